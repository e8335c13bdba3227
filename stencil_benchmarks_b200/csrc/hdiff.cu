// Horizontal diffusion for sm_100a: Laplacian -> flx/fly -> limiter -> update in ONE kernel.
//
// Replaces the nine GPU variants of
// stencil_benchmarks/benchmarks_collection/stencils/cuda_hip/horizontal_diffusion.py:50-128
// (templates cuda_hip/templates/horizontal_diffusion_*.j2); arithmetic and
// operation order follow the oracle, stencils/base.py:284-307:
//   lap = 4*inp - (inp[i+1] + inp[i-1] + inp[j+1] + inp[j-1])
//   flx = lap[i+1] - lap;  flx = flx*(inp[i+1]-inp) > 0 ? 0 : flx      (fly alike in j)
//   out = inp - coeff*(flx - flx[i-1] + fly - fly[j-1])
//
// Kernel "jmarch": a thread owns VEC consecutive i (one 128-bit vector: 2
// doubles / 4 floats) and marches over JT rows of j.  Per row it loads the
// VEC+4 wide inp strip [i0-2, i0+VEC+2) as aligned vectors, the coeff vector,
// and keeps in registers the rolling state that the oracle holds in whole
// temporaries: two Laplacian rows (VEC+2 wide), the previous fly row and two
// inp rows.  Every Laplacian / flux is computed once per thread; only the two
// i-edge Laplacians and one i-edge flx are recomputed by the neighbouring
// thread, which keeps the FP64 work at ~24 instructions per point (the
// on-the-fly form needs ~46 and would be FP64-bound on B200).  No shared
// memory, no barriers.  HBM traffic is the algorithmic minimum: inp halo
// columns/rows are L1/L2 hits on lines a neighbouring thread or block fetched.
#include "common.cuh"

namespace sb200 {
namespace {

template <class T>
__device__ __forceinline__ T limited(T flux, T delta) {
  // strict product-then-compare as in base.py:291-299
  return flux * delta > T(0) ? T(0) : flux;
}

// Strip of W = VEC + 4 values: index s <-> i = i0 - 2 + s.
template <class T, int VEC>
struct Strip {
  T v[VEC + 4];
};

template <class T, int VEC>
__device__ __forceinline__ void load_strip(const T* __restrict__ row, int i0, int nx, bool full,
                                           Strip<T, VEC>& s) {
  // row points at i = 0 of the row; i0 is a multiple of VEC, so for VEC > 1
  // i0-2 is 16-byte (double) / 8-byte (float) aligned.  The partial vector at
  // the i end of a row must not read beyond the halo of width 2.
  if (VEC > 1 && !full) {
#pragma unroll
    for (int n = 0; n < VEC + 4; ++n) s.v[n] = i0 - 2 + n <= nx + 1 ? row[i0 - 2 + n] : T(0);
  } else if constexpr (VEC == 2) {  // double: three LDG.E.128
    T a[2], b[2], c[2];
    load_vec<2>(row + i0 - 2, a);
    load_vec<2>(row + i0, b);
    load_vec<2>(row + i0 + 2, c);
    s.v[0] = a[0]; s.v[1] = a[1]; s.v[2] = b[0]; s.v[3] = b[1]; s.v[4] = c[0]; s.v[5] = c[1];
  } else if constexpr (VEC == 4) {  // float: LDG.E.64 + LDG.E.128 + LDG.E.64
    T a[2], b[4], c[2];
    load_vec<2>(row + i0 - 2, a);
    load_vec<4>(row + i0, b);
    load_vec<2>(row + i0 + 4, c);
    s.v[0] = a[0]; s.v[1] = a[1];
    s.v[2] = b[0]; s.v[3] = b[1]; s.v[4] = b[2]; s.v[5] = b[3];
    s.v[6] = c[0]; s.v[7] = c[1];
  } else {
#pragma unroll
    for (int n = 0; n < VEC + 4; ++n) s.v[n] = row[i0 - 2 + n];
  }
}

// Laplacian on row `c` for i in [i0-1, i0+VEC] (VEC+2 values; index m <-> i0-1+m)
template <class T, int VEC>
__device__ __forceinline__ void laplacian(const Strip<T, VEC>& below, const Strip<T, VEC>& c,
                                          const Strip<T, VEC>& above, T (&lap)[VEC + 2]) {
#pragma unroll
  for (int m = 0; m < VEC + 2; ++m) {
    const int s = m + 1;
    lap[m] = T(4) * c.v[s] - (c.v[s + 1] + c.v[s - 1] + above.v[s] + below.v[s]);
  }
}

template <class T, int VEC, int JT>
__global__ void __launch_bounds__(128)
    hdiff_jmarch_kernel(const T* __restrict__ inp, const T* __restrict__ coeff, T* __restrict__ out,
                        int nx, int ny, int64_t sy, int64_t sz) {
  const int i0 = (blockIdx.x * blockDim.x + threadIdx.x) * VEC;
  if (i0 >= nx) return;
  const int jb = blockIdx.y * JT;
  const int je = min(jb + JT, ny);
  const int64_t plane = int64_t(blockIdx.z) * sz;
  const T* __restrict__ ip = inp + plane;
  const T* __restrict__ cp = coeff + plane;
  T* __restrict__ op = out + plane;
  const bool full = i0 + VEC <= nx;

  // rolling state for output row j:
  //   rm = inp row j-1 ... only needed while building lap; kept implicitly
  //   rc = inp row j, rn = inp row j+1, lc = lap row j, ln = lap row j+1,
  //   fym = fly(j-1)
  Strip<T, VEC> r0, rc, rn, rnn;
  T lc[VEC + 2], ln[VEC + 2], fym[VEC];

  // warm-up: rows jb-2 .. jb+1  ->  lap(jb-1), lap(jb), fly(jb-1)
  load_strip<T, VEC>(ip + int64_t(jb - 2) * sy, i0, nx, full, r0);
  load_strip<T, VEC>(ip + int64_t(jb - 1) * sy, i0, nx, full, rc);
  load_strip<T, VEC>(ip + int64_t(jb) * sy, i0, nx, full, rn);
  load_strip<T, VEC>(ip + int64_t(jb + 1) * sy, i0, nx, full, rnn);
  laplacian<T, VEC>(r0, rc, rn, lc);   // lap(jb-1)
  laplacian<T, VEC>(rc, rn, rnn, ln);  // lap(jb)
#pragma unroll
  for (int n = 0; n < VEC; ++n)
    fym[n] = limited(ln[n + 1] - lc[n + 1], rn.v[n + 2] - rc.v[n + 2]);  // fly(jb-1)
  rc = rn;
  rn = rnn;
#pragma unroll
  for (int m = 0; m < VEC + 2; ++m) lc[m] = ln[m];

#pragma unroll 2
  for (int j = jb; j < je; ++j) {
    // newest row needed: j+2 (above the Laplacian row j+1)
    load_strip<T, VEC>(ip + int64_t(j + 2) * sy, i0, nx, full, rnn);
    T cf[VEC];
    if (full) {
      load_vec<VEC, Cache::Streaming>(cp + int64_t(j) * sy + i0, cf);
    } else {
#pragma unroll
      for (int n = 0; n < VEC; ++n) cf[n] = i0 + n < nx ? cp[int64_t(j) * sy + i0 + n] : T(0);
    }

    laplacian<T, VEC>(rc, rn, rnn, ln);  // lap(j+1)

    // flx(i, j) for i in [i0-1, i0+VEC-1]: index m <-> i0-1+m
    T flx[VEC + 1];
#pragma unroll
    for (int m = 0; m < VEC + 1; ++m)
      flx[m] = limited(lc[m + 1] - lc[m], rc.v[m + 2] - rc.v[m + 1]);

    T res[VEC];
#pragma unroll
    for (int n = 0; n < VEC; ++n) {
      const T fy = limited(ln[n + 1] - lc[n + 1], rn.v[n + 2] - rc.v[n + 2]);  // fly(i, j)
      res[n] = rc.v[n + 2] - cf[n] * (flx[n + 1] - flx[n] + fy - fym[n]);
      fym[n] = fy;
    }
    if (full) {
      store_vec<VEC, Cache::Streaming>(op + int64_t(j) * sy + i0, res);
    } else {
#pragma unroll
      for (int n = 0; n < VEC; ++n)
        if (i0 + n < nx) op[int64_t(j) * sy + i0 + n] = res[n];
    }

    rc = rn;
    rn = rnn;
#pragma unroll
    for (int m = 0; m < VEC + 2; ++m) lc[m] = ln[m];
  }
}

template <class T>
int launch_hdiff(const T* inp, const T* coeff, T* out, int64_t nx, int64_t ny, int64_t nz,
                 int64_t sy, int64_t sz, int dry_runs, double* time, cudaStream_t stream) {
  constexpr int V = VecN<T>::value;
  constexpr int JT = 64;
  // vector path: 16-byte aligned interior origin and strides that keep every row aligned
  const bool vector_ok = aligned_to(inp, 16) && aligned_to(coeff, 16) && aligned_to(out, 16) &&
                         sy % V == 0 && sz % V == 0;
  const int vec = vector_ok ? V : 1;
  const int64_t nvec = ceil_div(nx, vec);
  int bx = 128;
  while (bx > 32 && bx / 2 >= nvec) bx /= 2;
  const dim3 block(bx, 1, 1);
  const dim3 grid(unsigned(ceil_div(nvec, bx)), unsigned(ceil_div(ny, JT)), unsigned(nz));
  if (grid.y > 65535u || grid.z > 65535u) return fail("sb200_hdiff: domain too large for the launch grid");
  auto launch = [&] {
    if (vector_ok)
      hdiff_jmarch_kernel<T, V, JT><<<grid, block, 0, stream>>>(inp, coeff, out, int(nx), int(ny), sy, sz);
    else
      hdiff_jmarch_kernel<T, 1, JT><<<grid, block, 0, stream>>>(inp, coeff, out, int(nx), int(ny), sy, sz);
    count_launch();
  };
  return timed(launch, dry_runs, time, stream);
}

}  // namespace
}  // namespace sb200

using namespace sb200;

extern "C" int sb200_hdiff(int dtype, const void* inp, const void* coeff, void* out, int64_t nx,
                           int64_t ny, int64_t nz, int64_t sx, int64_t sy, int64_t sz,
                           int dry_runs, double* time, void* stream) {
  if (nx <= 0 || ny <= 0 || nz <= 0) return fail("sb200_hdiff: domain must be positive");
  if (sx != 1) return fail("sb200_hdiff: only layout (2,1,0) is supported (unit stride along i)");
  if (nx > (int64_t(1) << 30) || ny > (int64_t(1) << 30)) return fail("sb200_hdiff: domain too large");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == SB200_F64)
    return launch_hdiff<double>(static_cast<const double*>(inp), static_cast<const double*>(coeff),
                                static_cast<double*>(out), nx, ny, nz, sy, sz, dry_runs, time, s);
  if (dtype == SB200_F32)
    return launch_hdiff<float>(static_cast<const float*>(inp), static_cast<const float*>(coeff),
                               static_cast<float*>(out), nx, ny, nz, sy, sz, dry_runs, time, s);
  return fail("sb200_hdiff: unsupported dtype");
}
