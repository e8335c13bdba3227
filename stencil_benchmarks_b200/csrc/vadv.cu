// Vertical advection (implicit, per-column Thomas solve) for sm_100a.
//
// Replaces the four GPU variants of
// stencil_benchmarks/benchmarks_collection/stencils/cuda_hip/vertical_advection.py:60-73
// (templates cuda_hip/templates/vertical_advection_*.j2).  Arithmetic and
// operation order follow the oracle, stencils/base.py:410-473, including its
// reciprocal-then-multiply on interior levels and true division on the first
// and last level.
//
// Variant GLOBAL ("classic" data flow): one thread per (i, j) column,
// consecutive lanes on consecutive i so every level is one coalesced row
// segment per field; forward sweep k = 0..nz-1 keeps wcon / ustage of the
// neighbouring levels in registers (each value is loaded once) and writes the
// eliminated c and d to the ccol / dcol scratch fields; the backward sweep
// reads them back.  HBM traffic: 5 reads + 2 writes forward, 3 reads + 1 write
// backward.
#include "common.cuh"

namespace sb200 {
namespace {

template <class T>
struct VadvConst {
  static constexpr T dtr_stage = T(3) / T(20);
  static constexpr T beta_v = T(0);
  static constexpr T bet_m = T(0.5) * (T(1) - beta_v);
  static constexpr T bet_p = T(0.5) * (T(1) + beta_v);
};

template <class T>
__global__ void __launch_bounds__(128)
    vadv_global_kernel(const T* __restrict__ stage, const T* __restrict__ pos,
                       const T* __restrict__ tens, T* __restrict__ tensstage,
                       const T* __restrict__ wcon, T* __restrict__ ccol, T* __restrict__ dcol,
                       int nx, int ny, int nz, int64_t sy, int64_t sz, int64_t wshift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = blockIdx.y * blockDim.y + threadIdx.y;
  if (i >= nx || j >= ny) return;
  using C = VadvConst<T>;
  const T dtr_stage = C::dtr_stage, bet_m = C::bet_m, bet_p = C::bet_p;

  int64_t idx = int64_t(j) * sy + i;  // level 0

  // ---- forward sweep -----------------------------------------------------
  // rolling registers: wcon pair sums and stage values of neighbouring levels
  T wsum_next = wcon[idx + wshift + sz] + wcon[idx + sz];  // level 1 (shifted + unshifted)
  T stage_prev, stage_cur = stage[idx], stage_next = stage[idx + sz];
  T cprev, dprev;
  {
    const T gcv = T(0.25) * wsum_next;
    const T cs = gcv * bet_m;
    T c = gcv * bet_p;
    const T b = dtr_stage - c;
    const T correction = -cs * (stage_next - stage_cur);
    T d = dtr_stage * pos[idx] + tens[idx] + tensstage[idx] + correction;
    c /= b;
    d /= b;
    ccol[idx] = c;
    dcol[idx] = d;
    cprev = c;
    dprev = d;
  }

#pragma unroll 4
  for (int k = 1; k < nz - 1; ++k) {
    idx += sz;
    const T wsum_cur = wsum_next;
    wsum_next = wcon[idx + wshift + sz] + wcon[idx + sz];
    stage_prev = stage_cur;
    stage_cur = stage_next;
    stage_next = stage[idx + sz];

    const T gav = T(-0.25) * wsum_cur;
    const T gcv = T(0.25) * wsum_next;
    const T as = gav * bet_m;
    const T cs = gcv * bet_m;
    const T a = gav * bet_p;
    T c = gcv * bet_p;
    const T b = dtr_stage - a - c;
    const T correction = -as * (stage_prev - stage_cur) - cs * (stage_next - stage_cur);
    T d = dtr_stage * pos[idx] + tens[idx] + tensstage[idx] + correction;
    const T divided = T(1) / (b - cprev * a);
    c *= divided;
    d = (d - dprev * a) * divided;
    ccol[idx] = c;
    dcol[idx] = d;
    cprev = c;
    dprev = d;
  }

  T x;
  {
    idx += sz;  // level nz-1
    const T gav = T(-0.25) * wsum_next;
    const T as = gav * bet_m;
    const T a = gav * bet_p;
    const T b = dtr_stage - a;
    const T correction = -as * (stage_cur - stage_next);
    T d = dtr_stage * pos[idx] + tens[idx] + tensstage[idx] + correction;
    d = (d - dprev * a) / (b - cprev * a);
    // ---- backward sweep starts here (base.py:464-468) ----
    x = d;
    tensstage[idx] = dtr_stage * (x - pos[idx]);
  }

#pragma unroll 4
  for (int k = nz - 2; k >= 0; --k) {
    idx -= sz;
    x = dcol[idx] - ccol[idx] * x;
    tensstage[idx] = dtr_stage * (x - pos[idx]);
  }
}

template <class T>
int launch_vadv(const T* stage, const T* pos, const T* tens, T* tensstage, const T* wcon, T* ccol,
                T* dcol, int64_t nx, int64_t ny, int64_t nz, int64_t sy, int64_t sz, int ishift,
                int jshift, int variant, int dry_runs, double* time, cudaStream_t stream) {
  if (variant == SB200_VADV_ONCHIP) return fail("sb200_vadv: on-chip variant not available for this configuration");
  if (ccol == nullptr || dcol == nullptr) return fail("sb200_vadv: ccol/dcol scratch fields are required by the global variant");
  int bx = 64;
  while (bx > 32 && bx / 2 >= nx) bx /= 2;
  const dim3 block(bx, 1, 1);
  const dim3 grid(unsigned(ceil_div(nx, bx)), unsigned(ny), 1);
  if (grid.y > 65535u) return fail("sb200_vadv: domain too large for the launch grid");
  const int64_t wshift = int64_t(ishift) + int64_t(jshift) * sy;
  auto launch = [&] {
    vadv_global_kernel<T><<<grid, block, 0, stream>>>(stage, pos, tens, tensstage, wcon, ccol, dcol,
                                                      int(nx), int(ny), int(nz), sy, sz, wshift);
    count_launch();
  };
  return timed(launch, dry_runs, time, stream);
}

}  // namespace
}  // namespace sb200

using namespace sb200;

extern "C" int sb200_vadv(int dtype, const void* ustage, const void* upos, const void* utens,
                          void* utensstage, const void* wcon, void* ccol, void* dcol, void* datacol,
                          int64_t nx, int64_t ny, int64_t nz, int64_t sx, int64_t sy, int64_t sz,
                          int ishift, int jshift, int variant, int dry_runs, double* time,
                          void* stream) {
  (void)datacol;
  if (nx <= 0 || ny <= 0 || nz <= 0) return fail("sb200_vadv: domain must be positive");
  if (nz < 2) return fail("sb200_vadv: at least two vertical levels are required");
  if (sx != 1) return fail("sb200_vadv: only layout (2,1,0) is supported (unit stride along i)");
  if (ishift < 0 || ishift > 1 || jshift < 0 || jshift > 1) return fail("sb200_vadv: shifts must be 0 or 1");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == SB200_F64)
    return launch_vadv<double>(static_cast<const double*>(ustage), static_cast<const double*>(upos),
                               static_cast<const double*>(utens), static_cast<double*>(utensstage),
                               static_cast<const double*>(wcon), static_cast<double*>(ccol),
                               static_cast<double*>(dcol), nx, ny, nz, sy, sz, ishift, jshift,
                               variant, dry_runs, time, s);
  if (dtype == SB200_F32)
    return launch_vadv<float>(static_cast<const float*>(ustage), static_cast<const float*>(upos),
                              static_cast<const float*>(utens), static_cast<float*>(utensstage),
                              static_cast<const float*>(wcon), static_cast<float*>(ccol),
                              static_cast<float*>(dcol), nx, ny, nz, sy, sz, ishift, jshift, variant,
                              dry_runs, time, s);
  return fail("sb200_vadv: unsupported dtype");
}
