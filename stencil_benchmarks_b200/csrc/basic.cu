// Basic stencils (empty, copy, one-sided / symmetric average, Laplacian) for sm_100a.
//
// Replaces the bodies of stencil_benchmarks/benchmarks_collection/stencils/cuda_hip/basic.py:101-132
// and the loop templates cuda_hip/templates/basic_1d.j2:37-64 / basic_3d.j2:37-66.
// Arithmetic follows the oracle's operation order (stencils/base.py:180-254).
//
// Design: pure HBM streaming with neighbour reuse through L1.  One thread owns
// one 128-bit vector of consecutive i (2 doubles / 4 floats) and ROWS
// consecutive j rows, so ROWS (+2 for a j-Laplacian) independent 16-byte loads
// are in flight per thread and j-neighbours are reused from registers.
// Unit-offset i-neighbours are scalar L1 hits on lines the warp already
// fetched.  The scalar instantiation (VEC = 1) covers fields whose interior
// origin or strides are not 16-byte aligned (alignment=0 in the reference
// allocator, tools/array.py:139-174).
#include "common.cuh"

namespace sb200 {
namespace {

constexpr int kRows = 4;

// SHAPE: for averages the axis (0, 1, 2); for the Laplacian the axis mask.
template <class T, int KIND, int SHAPE, int VEC>
__global__ void __launch_bounds__(256)
    basic_kernel(const T* __restrict__ inp, T* __restrict__ out, int nx, int ny, int nz,
                 int64_t sy, int64_t sz) {
  if constexpr (KIND == SB200_BASIC_EMPTY) return;

  const int i0 = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  const int jb = (blockIdx.y * blockDim.y + threadIdx.y) * kRows;
  const int k = blockIdx.z;
  if (i0 >= nx || jb >= ny) return;

  const bool full = i0 + VEC <= nx;  // partial vectors at the i end go scalar
  const int nvalid = full ? VEC : nx - i0;
  const int64_t base = int64_t(k) * sz + int64_t(jb) * sy + i0;

  auto load_row = [&](int64_t offset, T(&v)[VEC]) {
    if (full) {
      load_vec<VEC>(inp + offset, v);
    } else {
#pragma unroll
      for (int n = 0; n < VEC; ++n) v[n] = n < nvalid ? inp[offset + n] : T(0);
    }
  };
  auto store_row = [&](int64_t offset, const T(&v)[VEC]) {
    if (full) {
      store_vec<VEC, Cache::Streaming>(out + offset, v);
    } else {
#pragma unroll
      for (int n = 0; n < VEC; ++n)
        if (n < nvalid) out[offset + n] = v[n];
    }
  };

  constexpr bool LAP = KIND == SB200_BASIC_LAPLACIAN;
  constexpr bool NEED_X = (KIND == SB200_BASIC_ONESIDED_AVG && SHAPE == 0) ||
                          (KIND == SB200_BASIC_SYMMETRIC_AVG && SHAPE == 0) || (LAP && (SHAPE & 1));
  constexpr bool NEED_XM = (KIND == SB200_BASIC_SYMMETRIC_AVG && SHAPE == 0) || (LAP && (SHAPE & 1));
  constexpr bool NEED_YP = (KIND == SB200_BASIC_ONESIDED_AVG && SHAPE == 1) ||
                           (KIND == SB200_BASIC_SYMMETRIC_AVG && SHAPE == 1) || (LAP && (SHAPE & 2));
  constexpr bool NEED_YM = (KIND == SB200_BASIC_SYMMETRIC_AVG && SHAPE == 1) || (LAP && (SHAPE & 2));
  constexpr bool NEED_ZP = (KIND == SB200_BASIC_ONESIDED_AVG && SHAPE == 2) ||
                           (KIND == SB200_BASIC_SYMMETRIC_AVG && SHAPE == 2) || (LAP && (SHAPE & 4));
  constexpr bool NEED_ZM = (KIND == SB200_BASIC_SYMMETRIC_AVG && SHAPE == 2) || (LAP && (SHAPE & 4));
  // the centre value is not used by a symmetric average along y/z
  constexpr bool NEED_C = !(KIND == SB200_BASIC_SYMMETRIC_AVG && SHAPE != 0);

  // rows jb-1 .. jb+kRows of the centre plane (index r+1 <-> row jb+r)
  T c[kRows + 2][VEC];
  T xm[kRows], xp[kRows];  // inp[i0-1], inp[i0+VEC] per row
  T zp[kRows][VEC], zm[kRows][VEC];

#pragma unroll
  for (int r = -1; r <= kRows; ++r) {
    const bool edge = r == -1 || r == kRows;
    if (edge && !((r == -1 && NEED_YM) || (r == kRows && NEED_YP))) continue;
    if (!edge && !(NEED_C || NEED_YP || NEED_YM)) continue;
    // rows beyond ny are only needed as +-1 neighbours of valid rows
    if (jb + r > ny || (jb + r == ny && !NEED_YP)) continue;
    load_row(base + int64_t(r) * sy, c[r + 1]);
  }
#pragma unroll
  for (int r = 0; r < kRows; ++r) {
    if (jb + r >= ny) continue;
    const int64_t o = base + int64_t(r) * sy;
    if constexpr (NEED_XM) xm[r] = inp[o - 1];
    if constexpr (NEED_X) xp[r] = inp[o + nvalid];
    if constexpr (NEED_ZP) load_row(o + sz, zp[r]);
    if constexpr (NEED_ZM) load_row(o - sz, zm[r]);
  }

#pragma unroll
  for (int r = 0; r < kRows; ++r) {
    if (jb + r >= ny) continue;
    T res[VEC];
#pragma unroll
    for (int n = 0; n < VEC; ++n) {
      const T centre = c[r + 1][n];
      // neighbours along i come from the vector itself or the two edge scalars;
      // for a partial vector the right neighbour of the last valid element is xp
      T left = T(0), right = T(0);
      if constexpr (NEED_XM) left = n == 0 ? xm[r] : c[r + 1][n - 1];
      if constexpr (NEED_X) right = (n == VEC - 1 || n == nvalid - 1) ? xp[r] : c[r + 1][n + 1];

      if constexpr (KIND == SB200_BASIC_COPY) {
        res[n] = centre;
      } else if constexpr (KIND == SB200_BASIC_ONESIDED_AVG) {
        const T plus = SHAPE == 0 ? right : (SHAPE == 1 ? c[r + 2][n] : zp[r][n]);
        res[n] = (plus + centre) / 2;
      } else if constexpr (KIND == SB200_BASIC_SYMMETRIC_AVG) {
        const T plus = SHAPE == 0 ? right : (SHAPE == 1 ? c[r + 2][n] : zp[r][n]);
        const T minus = SHAPE == 0 ? left : (SHAPE == 1 ? c[r][n] : zm[r][n]);
        res[n] = (plus + minus) / 2;
      } else {
        T acc = 0;
        if constexpr (SHAPE & 1) acc += 2 * centre - right - left;
        if constexpr (SHAPE & 2) acc += 2 * centre - c[r + 2][n] - c[r][n];
        if constexpr (SHAPE & 4) acc += 2 * centre - zp[r][n] - zm[r][n];
        res[n] = acc;
      }
    }
    store_row(base + int64_t(r) * sy, res);
  }
}

template <class T, int KIND, int SHAPE>
int launch_basic(const T* inp, T* out, int64_t nx, int64_t ny, int64_t nz, int64_t sy, int64_t sz,
                 int dry_runs, double* time, cudaStream_t stream) {
  constexpr int V = VecN<T>::value;
  const bool vector_ok = aligned_to(inp, 16) && aligned_to(out, 16) && sy % V == 0 && sz % V == 0;
  const int vec = vector_ok ? V : 1;
  const int64_t nvec = ceil_div(nx, vec);
  // 128 threads per block, as wide in i as the domain allows
  int bx = 128;
  while (bx > 32 && bx / 2 >= nvec) bx /= 2;
  const int by = 128 / bx;
  const dim3 block(bx, by, 1);
  const dim3 grid(unsigned(ceil_div(nvec, bx)), unsigned(ceil_div(ny, int64_t(by) * kRows)),
                  unsigned(nz));
  if (grid.y > 65535u || grid.z > 65535u) return fail("sb200_basic: domain too large for the launch grid");
  auto launch = [&] {
    if (vector_ok)
      basic_kernel<T, KIND, SHAPE, V><<<grid, block, 0, stream>>>(inp, out, int(nx), int(ny), int(nz), sy, sz);
    else
      basic_kernel<T, KIND, SHAPE, 1><<<grid, block, 0, stream>>>(inp, out, int(nx), int(ny), int(nz), sy, sz);
    count_launch();
  };
  return timed(launch, dry_runs, time, stream);
}

template <class T>
int dispatch_basic(int kind, const T* inp, T* out, int64_t nx, int64_t ny, int64_t nz, int64_t sy,
                   int64_t sz, int axis, int along, int dry_runs, double* time, cudaStream_t s) {
#define SB200_LAUNCH(KIND, SHAPE) \
  return launch_basic<T, KIND, SHAPE>(inp, out, nx, ny, nz, sy, sz, dry_runs, time, s)
  switch (kind) {
    case SB200_BASIC_EMPTY: SB200_LAUNCH(SB200_BASIC_EMPTY, 0);
    case SB200_BASIC_COPY: SB200_LAUNCH(SB200_BASIC_COPY, 0);
    case SB200_BASIC_ONESIDED_AVG:
      switch (axis) {
        case 0: SB200_LAUNCH(SB200_BASIC_ONESIDED_AVG, 0);
        case 1: SB200_LAUNCH(SB200_BASIC_ONESIDED_AVG, 1);
        case 2: SB200_LAUNCH(SB200_BASIC_ONESIDED_AVG, 2);
      }
      return fail("sb200_basic: axis must be 0, 1 or 2");
    case SB200_BASIC_SYMMETRIC_AVG:
      switch (axis) {
        case 0: SB200_LAUNCH(SB200_BASIC_SYMMETRIC_AVG, 0);
        case 1: SB200_LAUNCH(SB200_BASIC_SYMMETRIC_AVG, 1);
        case 2: SB200_LAUNCH(SB200_BASIC_SYMMETRIC_AVG, 2);
      }
      return fail("sb200_basic: axis must be 0, 1 or 2");
    case SB200_BASIC_LAPLACIAN:
      switch (along) {
        case 1: SB200_LAUNCH(SB200_BASIC_LAPLACIAN, 1);
        case 2: SB200_LAUNCH(SB200_BASIC_LAPLACIAN, 2);
        case 3: SB200_LAUNCH(SB200_BASIC_LAPLACIAN, 3);
        case 4: SB200_LAUNCH(SB200_BASIC_LAPLACIAN, 4);
        case 5: SB200_LAUNCH(SB200_BASIC_LAPLACIAN, 5);
        case 6: SB200_LAUNCH(SB200_BASIC_LAPLACIAN, 6);
        case 7: SB200_LAUNCH(SB200_BASIC_LAPLACIAN, 7);
      }
      return fail("sb200_basic: Laplacian needs at least one axis (along in 1..7)");
  }
#undef SB200_LAUNCH
  return fail("sb200_basic: unknown stencil kind");
}

}  // namespace
}  // namespace sb200

using namespace sb200;

extern "C" int sb200_basic(int kind, int dtype, const void* inp, void* out, int64_t nx, int64_t ny,
                           int64_t nz, int64_t sx, int64_t sy, int64_t sz, int axis, int along,
                           int dry_runs, double* time, void* stream) {
  if (nx <= 0 || ny <= 0 || nz <= 0) return fail("sb200_basic: domain must be positive");
  if (sx != 1) return fail("sb200_basic: only layout (2,1,0) is supported (unit stride along i)");
  if (nx > (int64_t(1) << 30) || ny > (int64_t(1) << 30)) return fail("sb200_basic: domain too large");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == SB200_F64)
    return dispatch_basic<double>(kind, static_cast<const double*>(inp), static_cast<double*>(out),
                                  nx, ny, nz, sy, sz, axis, along, dry_runs, time, s);
  if (dtype == SB200_F32)
    return dispatch_basic<float>(kind, static_cast<const float*>(inp), static_cast<float*>(out), nx,
                                 ny, nz, sy, sz, axis, along, dry_runs, time, s);
  return fail("sb200_basic: unsupported dtype");
}
