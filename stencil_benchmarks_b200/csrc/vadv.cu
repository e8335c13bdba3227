// Vertical advection (implicit, per-column Thomas solve) for sm_100a.
//
// Replaces the four GPU variants of
// stencil_benchmarks/benchmarks_collection/stencils/cuda_hip/vertical_advection.py:60-73
// (templates cuda_hip/templates/vertical_advection_*.j2).  Arithmetic and
// operation order follow the oracle, stencils/base.py:410-473, including its
// reciprocal-then-multiply on interior levels and true division on the first
// and last level.
//
// Variant ONCHIP (fast path, see the block comment above vadv_onchip_kernel): the eliminated
// coefficients never leave the SM -- c lives in tensor memory (TMEM), the folded right-hand side
// in shared memory -- so HBM traffic is the algorithmic minimum of 5 reads + 1 write.  One launch
// solves one component (u) or all three (u, v, w: the reference's merged kernel,
// templates/vertical_advection_localmemmerged.j2:392-462), which then share wcon through the L2.
//
// Variant GLOBAL ("classic" data flow): one thread per (i, j) column,
// consecutive lanes on consecutive i so every level is one coalesced row
// segment per field; forward sweep k = 0..nz-1 keeps wcon / ustage of the
// neighbouring levels in registers (each value is loaded once) and writes the
// eliminated c and d to the ccol / dcol scratch fields; the backward sweep
// reads them back.  HBM traffic: 5 reads + 2 writes forward, 3 reads + 1 write
// backward.
#include <algorithm>
#include <cstdlib>
#include <mutex>
#include <type_traits>

#include "common.cuh"
#include "tma.cuh"

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

// ---------------------------------------------------------------------------------
// ONCHIP variant
// ---------------------------------------------------------------------------------
// Persistent CTAs (one per SM): 4 compute warps (float32: 8) + 1 TMA producer warp.  A batch is
// 128 (float32: 256) consecutive-i columns of one j row; thread t owns column i_t + t for all levels.
// Work items (batch, component) are drawn from a global counter by the producer lane (see the
// comment above the kernel): SMs do not progress at the same speed, static round robin is 6 % slower.
//
//  * Input levels arrive through a TMA ring (boxes of one batch x KD levels for ustage, upos,
//    utens, utensstage and wcon; the wcon tile carries 16 more bytes per level so that wcon(i+1)
//    of the last column is there -- TMA needs 16-byte aligned box origins, so the i+1 neighbour
//    cannot be a second, shifted box; the j+1 neighbour of the v component can), so the
//    k-sequential compute never waits on a global load it issued itself.  Up to 8 stages of
//    4 levels are in flight (~160 KB per SM).
//  * The Thomas coefficients of a column are private to its thread and live on chip: c[k] of
//    every level in TMEM (tcgen05.st/ld, lane = thread), next to it the right-hand side of as
//    many levels as the thread's TMEM columns hold, the remaining right-hand sides in shared
//    memory.  The right-hand side is stored folded for the backward sweep:
//        e[k] = d[k] - pos[k] - c[k]*pos[k+1]   =>   z[k] = x[k] - pos[k] = e[k] - c[k]*z[k+1],
//        utensstage[k] = dtr * z[k]
//    so the backward sweep reads nothing from HBM.
//  * The backward sweep of batch n runs in lock-step with the forward sweep of batch n+1:
//    at step s the thread consumes slot p of the old column (level nz-1-s) and then stores the
//    new column's level s-1 into the same slot.  One set of nz-1 slots per thread suffices and
//    the two dependency chains (forward recurrence, backward FMA) overlap.
//  * The forward recurrence is division free (homogeneous triple p, r, q; see VadvForward).
// HBM traffic = 5 reads + 1 write per point.  Storage limits: (nz-1) * sizeof(T)/4 TMEM columns
// per thread (512, or 256 where two warps share a lane quarter) and shared memory for the
// right-hand sides that do not fit TMEM next to at least two ring stages.
namespace vcfg {
// Columns per batch = compute threads per CTA.  float64: 128 (4 warps, one per scheduler; the
// per-column store of a 160-level column fills TMEM + shared memory).  float32 values are half
// the size, so 256 columns fit: 8 warps, two per scheduler, which hides the per-level issue
// latency that bounds the float64 kernel.  Two warps on one scheduler share a TMEM lane quarter
// and split its 512 columns.
template <class T>
__host__ __device__ constexpr int cols() { return sizeof(T) == 8 ? 128 : 256; }
template <class T>
__host__ __device__ constexpr int threads() { return cols<T>() + 32; }
constexpr int kTmemCols = 512;
template <class T>
__host__ __device__ constexpr int tmem_cols_per_thread() { return kTmemCols / (cols<T>() / 128); }
// ring stage = 4 tiles of KD x cols values + 2 wcon tiles.  A wcon tile carries 16 more bytes per
// level (wcon(i+1) of the last column): inside the box where the box stays <= 256 elements
// (float64), as a second, 16-byte wide box otherwise (float32).  Tiles are padded to 128 bytes so
// every TMA destination stays 128-byte aligned.
template <class T>
__host__ __device__ constexpr int edge_elems() { return 16 / int(sizeof(T)); }
template <class T>
__host__ __device__ constexpr bool split_wcon() { return cols<T>() + edge_elems<T>() > 256; }
template <class T>
__host__ __device__ constexpr int wcon_width() { return split_wcon<T>() ? cols<T>() : cols<T>() + edge_elems<T>(); }
template <class T>
__host__ __device__ constexpr int tile_bytes(int kd) { return kd * cols<T>() * int(sizeof(T)); }
template <class T>
__host__ __device__ constexpr int wcon_bytes(int kd) { return kd * wcon_width<T>() * int(sizeof(T)); }
template <class T>
__host__ __device__ constexpr int wcon_main_bytes(int kd) { return (wcon_bytes<T>(kd) + 127) / 128 * 128; }
template <class T>
__host__ __device__ constexpr int wcon_edge_bytes(int kd) { return split_wcon<T>() ? (kd * 16 + 127) / 128 * 128 : 0; }
template <class T>
__host__ __device__ constexpr int wcon_tile_bytes(int kd) { return wcon_main_bytes<T>(kd) + wcon_edge_bytes<T>(kd); }
// bytes TMA delivers for one wcon tile with (first tile) / without (j+1 tile) the i+1 edge
template <class T>
__host__ __device__ constexpr int wcon_tx_bytes(int kd, bool with_edge) {
  return wcon_bytes<T>(kd) + (split_wcon<T>() && with_edge ? kd * 16 : 0);
}
}  // namespace vcfg

__device__ __forceinline__ void tmem_store(uint32_t taddr, double v) {
  const unsigned long long bits = (unsigned long long)__double_as_longlong(v);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr),
               "r"((uint32_t)(bits & 0xffffffffull)), "r"((uint32_t)(bits >> 32)));
}
__device__ __forceinline__ void tmem_store(uint32_t taddr, float v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(__float_as_uint(v)));
}
__device__ __forceinline__ void tmem_load(uint32_t taddr, double& v) {
  uint32_t lo, hi;
  // volatile without a memory clobber: TMEM accesses keep their mutual order, ordinary loads,
  // stores and arithmetic are free to move around them.  The wait ties the registers.
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(lo), "=r"(hi) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(lo), "+r"(hi));
  v = __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
}
__device__ __forceinline__ void tmem_load(uint32_t taddr, float& v) {
  uint32_t bits;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(bits) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(bits));
  v = __uint_as_float(bits);
}

// (c, e) pairs in adjacent TMEM columns: one LDTM / STTM per level
__device__ __forceinline__ void tmem_store_pair(uint32_t taddr, double c, double e) {
  const unsigned long long cb = (unsigned long long)__double_as_longlong(c);
  const unsigned long long eb = (unsigned long long)__double_as_longlong(e);
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr),
               "r"((uint32_t)(cb & 0xffffffffull)), "r"((uint32_t)(cb >> 32)),
               "r"((uint32_t)(eb & 0xffffffffull)), "r"((uint32_t)(eb >> 32)));
}
__device__ __forceinline__ void tmem_store_pair(uint32_t taddr, float c, float e) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr),
               "r"(__float_as_uint(c)), "r"(__float_as_uint(e)));
}
__device__ __forceinline__ void tmem_load_pair(uint32_t taddr, double& c, double& e) {
  uint32_t c0, c1, e0, e1;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(c0), "=r"(c1), "=r"(e0), "=r"(e1)
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(c0), "+r"(c1), "+r"(e0), "+r"(e1));
  c = __longlong_as_double((long long)(((unsigned long long)c1 << 32) | c0));
  e = __longlong_as_double((long long)(((unsigned long long)e1 << 32) | e0));
}
__device__ __forceinline__ void tmem_load_pair(uint32_t taddr, float& c, float& e) {
  uint32_t cb, eb;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(cb), "=r"(eb) : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" : "+r"(cb), "+r"(eb));
  c = __uint_as_float(cb);
  e = __uint_as_float(eb);
}

// Per-thread store of the eliminated coefficients (c[k], e[k]) by slot.  The first `paired`
// slots keep both values in TMEM (adjacent columns); the remaining slots keep c in TMEM and e in
// shared memory.  TMEM holds 512 32-bit columns per thread, so for float64 and nz = 160 that is
// 97 paired slots + 62 split slots -- which leaves ~160 KB of shared memory to the TMA ring.
template <class T>
struct VadvSlots {
  static constexpr int TCOLS = int(sizeof(T)) / 4;
  uint32_t tmem;  // TMEM address of this warp's lane quarter, column 0 of the allocation
  T* smem;        // this thread's column of the shared-memory part: slot q at smem[q * cols]
  int paired;

  // KIND: 0 = slot known to be split, 1 = known to be paired, 2 = decide at run time
  template <int KIND>
  __device__ __forceinline__ void load(int p, T& c, T& e) const {
    if (KIND == 1 || (KIND == 2 && p < paired)) {
      tmem_load_pair(tmem + uint32_t(2 * TCOLS * p), c, e);
    } else {
      tmem_load(tmem + uint32_t(2 * TCOLS * paired + TCOLS * (p - paired)), c);
      e = smem[(p - paired) * vcfg::cols<T>()];
    }
  }
  template <int KIND>
  __device__ __forceinline__ void store(int p, T c, T e) const {
    if (KIND == 1 || (KIND == 2 && p < paired)) {
      tmem_store_pair(tmem + uint32_t(2 * TCOLS * p), c, e);
    } else {
      tmem_store(tmem + uint32_t(2 * TCOLS * paired + TCOLS * (p - paired)), c);
      smem[(p - paired) * vcfg::cols<T>()] = e;
    }
  }
};

// The slots of one chunk: level r of the chunk uses slot p + r*dp.  For the two pure kinds the
// TMEM column and the shared-memory offset are affine in r with uniform coefficients, so the
// unrolled body addresses them without per-thread integer work.
template <class T, int KIND>
struct VadvCursor {
  static constexpr int TCOLS = VadvSlots<T>::TCOLS;
  const VadvSlots<T>& slots;
  int p, dp;
  __device__ __forceinline__ VadvCursor(const VadvSlots<T>& s, int p_, int dp_) : slots(s), p(p_), dp(dp_) {}
  __device__ __forceinline__ uint32_t column(int r) const {
    const int q = p + r * dp;
    return KIND == 1 ? uint32_t(2 * TCOLS * q) : uint32_t(2 * TCOLS * slots.paired + TCOLS * (q - slots.paired));
  }
  __device__ __forceinline__ T* shared(int r) const {
    return slots.smem + (p + r * dp - slots.paired) * vcfg::cols<T>();
  }
  __device__ __forceinline__ void load(int r, T& c, T& e) const {
    if (KIND == 1) {
      tmem_load_pair(slots.tmem + column(r), c, e);
    } else if (KIND == 0) {
      tmem_load(slots.tmem + column(r), c);
      e = *shared(r);
    } else {
      slots.template load<2>(p + r * dp, c, e);
    }
  }
  __device__ __forceinline__ void store(int r, T c, T e) const {
    if (KIND == 1) {
      tmem_store_pair(slots.tmem + column(r), c, e);
    } else if (KIND == 0) {
      tmem_store(slots.tmem + column(r), c);
      *shared(r) = e;
    } else {
      slots.template store<2>(p + r * dp, c, e);
    }
  }
};

// Reciprocal from the hardware seed (MUFU.RCP64H, ~2^-23) and two Newton steps: full double
// precision up to the last ulp, no special-case branches, 5 FP64 instructions on the critical path
// of nothing (the recurrence below is division free).
__device__ __forceinline__ double fast_rcp(double x) {
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  double t = fma(-x, r, 1.0);
  r = fma(r, t, r);
  t = fma(-x, r, 1.0);
  return fma(r, t, r);
}
__device__ __forceinline__ float fast_rcp(float x) { return __frcp_rn(x); }

// Exact power-of-two renormalisation of the homogeneous triple (p, r, q) so that q is in [1, 2).
__device__ __forceinline__ void rescale(double& p, double& r, double& q) {
  const int ex = (__double2hiint(q) >> 20) & 0x7ff;
  const double f = __hiloint2double((2046 - ex) << 20, 0);
  p *= f;
  r *= f;
  q *= f;
}
__device__ __forceinline__ void rescale(float& p, float& r, float& q) {
  const unsigned ex = (__float_as_uint(q) >> 23) & 0xffu;
  const float f = __uint_as_float((254u - ex) << 23);
  p *= f;
  r *= f;
  q *= f;
}

// Store of one output value.  MODE 1: guarded by a predicate instead of a branch (the k-sequential
// loop has one warp per scheduler, so every BSSY / BRA / BSYNC around a store is issue latency on
// the critical path); no "memory" clobber: the kernel reads utensstage through TMA only, and a
// compiler barrier per level would pin the shared-memory loads of the next level behind it.
// MODE 0: the plain conditional store.
template <int MODE>
__device__ __forceinline__ void store_if(bool valid, double* p, double v) {
  if (MODE == 1) {
    asm volatile("{ .reg .pred q; setp.ne.b32 q, %2, 0; @q st.global.f64 [%0], %1; }" ::"l"(p), "d"(v),
                 "r"(int(valid)));
  } else if (valid) {
    *p = v;
  }
}
template <int MODE>
__device__ __forceinline__ void store_if(bool valid, float* p, float v) {
  if (MODE == 1) {
    asm volatile("{ .reg .pred q; setp.ne.b32 q, %2, 0; @q st.global.f32 [%0], %1; }" ::"l"(p), "f"(v),
                 "r"(int(valid)));
  } else if (valid) {
    *p = v;
  }
}

// Forward state of the column a thread is eliminating.  c[k] = p/q and d[k] = r/q are kept as a
// homogeneous triple: with a, b, cc, d0 the coefficients of level k (base.py:432-445),
//     c[k] = cc / (b - c[k-1] a)            =>  p' = cc q,  q' = b q - a p
//     d[k] = (d0 - d[k-1] a) / (b - c[k-1] a)  =>  r' = d0 q - a r
// so the level-to-level dependency is two FMAs instead of a division; the division needed to
// store c[k], d[k] is off the critical path and pipelines across levels.
//
// With bet_m = bet_p = 1/2 the coefficients collapse onto h[k] = wsum[k] / 8 (exactly the oracle's
// 0.25 * wsum * 0.5): a = as = -h[k], cc = cs = h[k+1], b = dtr + h[k] - h[k+1], and with
// t[k] = h[k+1] * (stage[k+1] - stage[k]) the correction term is -t[k-1] - t[k].  Level 0 is the
// same formula with h[0] := 0, t[-1] := 0 and (p, r, q) = (0, 0, 1); wcon[0] is never used.
template <class T>
struct VadvForward {
  T h_cur = 0, h_next = 0, st_cur = 0, t_prev = 0;
  T pos_cur = 0, tens_cur = 0, tss_cur = 0;
  T p = 0, r = 0, q = 1;
};

// One lock-step iteration: level s of the new column has arrived (values v_*), level k = s-1 is
// eliminated and stored, after the old column's level nz-1-s has been back-substituted from the
// same slot.  FIRST: s may be 0 (resolved at compile time in the peeled first chunk).
template <class T, bool FIRST, int KIND, int STORE = 0>
__device__ __forceinline__ void vadv_step(int s, VadvForward<T>& f, T v_stage, T v_pos, T v_tens,
                                          T v_tss, T v_wsum, T& z, const VadvCursor<T, KIND>& slot, int r,
                                          T* old_out, bool old_valid) {
  const T dtr_stage = VadvConst<T>::dtr_stage;
  if (FIRST && s == 0) {
    f.h_next = 0;
    f.t_prev = 0;
    f.p = 0; f.r = 0; f.q = 1;
    f.st_cur = v_stage;
    f.pos_cur = v_pos; f.tens_cur = v_tens; f.tss_cur = v_tss;
    return;
  }
  f.h_cur = f.h_next;
  f.h_next = T(0.125) * v_wsum;
  const T t = f.h_next * (v_stage - f.st_cur);
  f.st_cur = v_stage;
  const T b = dtr_stage + f.h_cur - f.h_next;
  const T d0 = dtr_stage * f.pos_cur + f.tens_cur + f.tss_cur - f.t_prev - t;
  f.t_prev = t;
  const T pn = f.h_next * f.q;
  const T qn = b * f.q + f.h_cur * f.p;
  const T rn = d0 * f.q + f.h_cur * f.r;
  f.p = pn; f.r = rn; f.q = qn;
  const T rq = fast_rcp(qn);
  const T c = pn * rq;
  const T e = rq * (rn - pn * v_pos) - f.pos_cur;  // d - pos[k] - c * pos[k+1], folded
  // backward step of the old column on the slot that is about to be overwritten
  T c_old, e_old;
  slot.load(r, c_old, e_old);
  z = e_old - c_old * z;
  store_if<STORE>(old_valid, old_out, dtr_stage * z);
  slot.store(r, c, e);
  f.pos_cur = v_pos; f.tens_cur = v_tens; f.tss_cur = v_tss;
}


// Everything one launch needs about its components.  A launch solves `ncomp` = 1 or 3 systems:
// component c reads its own ustage / upos / utens / utensstage fields and the shared wcon field
// with the neighbour shift (ishift, jshift) = shifts[c] (u: i+1, v: j+1, w: none; base.py:475-483).
struct VadvMaps {
  CUtensorMap stage[3], pos[3], tens[3], tensstage[3];
  CUtensorMap wcon, wcon_edge;
};
template <class T>
struct VadvComponents {
  T* tensstage[3];
  int ishift[3], jshift[3];
  int ncomp;
  int two_wcon_tiles;  // some component has a j shift: ring stages carry a second wcon tile
};

// Work items are (batch, component) pairs, handed out dynamically: the producer lane draws the
// next item from a global counter and passes it to the compute warps in the header of the ring
// stage that carries the item's first chunk.  With three components the items of one batch are
// drawn back to back, so three CTAs sweep the same columns at the same time and the wcon tiles
// (and the j+1 tiles of the v component, which are the next row's own tiles) are fetched from HBM
// once and hit the L2 twice: one launch moves 16 fields' worth of HBM traffic, not 18.
template <class T, int KD, int STORE>
__global__ void __launch_bounds__(vcfg::threads<T>(), 1)
    vadv_onchip_kernel(const __grid_constant__ VadvMaps maps,
                       const __grid_constant__ VadvComponents<T> comps, unsigned int* __restrict__ counter,
                       int nx, int ny, int nz, int64_t sy, int64_t sz, int stages, int paired) {
  using C = VadvConst<T>;
  constexpr int COLS = vcfg::cols<T>();
  constexpr bool SPLIT = vcfg::split_wcon<T>();
  constexpr int EDGE = vcfg::edge_elems<T>();
  constexpr int WMAIN = vcfg::wcon_main_bytes<T>(KD);
  constexpr int TILE = vcfg::tile_bytes<T>(KD);           // one field, KD levels
  constexpr int WB = vcfg::wcon_width<T>();                // wcon tile width in elements
  constexpr int WTILE = vcfg::wcon_tile_bytes<T>(KD);
  const int STAGE = 4 * TILE + (comps.two_wcon_tiles ? 2 : 1) * WTILE;
  extern __shared__ __align__(128) unsigned char smem[];
  // layout: [ring: stages x STAGE][e store: (nz-1-paired) x COLS][barriers][headers][tmem base]
  unsigned char* ring = smem;
  T* estore = reinterpret_cast<T*>(smem + stages * STAGE);
  uint64_t* full =
      reinterpret_cast<uint64_t*>(smem + stages * STAGE + size_t(nz - 1 - paired) * COLS * sizeof(T));
  uint64_t* empty = full + stages;
  int* header = reinterpret_cast<int*>(empty + stages);  // item carried by a stage (first chunk only)
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(header + stages);

  const int warp = threadIdx.x >> 5;
  const int nbx = (nx + COLS - 1) / COLS;
  const int nitems = nbx * ny * comps.ncomp;
  const int nchunks = (nz + KD - 1) / KD;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      tma::mbar_init(&full[s], 1);
      tma::mbar_init(&empty[s], COLS / 32);
    }
    tma::fence_barrier_init();
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     tma::smem_u32(tmem_base_smem)),
                 "r"(vcfg::kTmemCols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  if (warp == COLS / 32) {
    // ===== producer warp =====
    if ((threadIdx.x & 31) == 0) {
      for (int c = 0; c < comps.ncomp; ++c) {
        tma::prefetch_tensormap(&maps.stage[c]);
        tma::prefetch_tensormap(&maps.pos[c]);
        tma::prefetch_tensormap(&maps.tens[c]);
        tma::prefetch_tensormap(&maps.tensstage[c]);
      }
      tma::prefetch_tensormap(&maps.wcon);
      if (SPLIT) tma::prefetch_tensormap(&maps.wcon_edge);
      int slot = 0, round = 0;  // ring position of the running chunk counter
      for (;;) {
        const int item = int(atomicAdd(counter, 1u));
        if (item >= nitems) {
          // end marker: a stage without data whose header says so
          if (round > 0) tma::mbar_wait(&empty[slot], (round - 1) & 1);
          header[slot] = -1;
          tma::mbar_arrive(&full[slot]);
          break;
        }
        const int comp = item % comps.ncomp;
        const int b = item / comps.ncomp;
        const int j = b / nbx;
        const int it = (b - j * nbx) * COLS;
        const int jshift = comps.jshift[comp];
        const uint32_t tx_bytes = 4 * TILE + vcfg::wcon_tx_bytes<T>(KD, true) +
                                  (jshift ? vcfg::wcon_tx_bytes<T>(KD, false) : 0);
        for (int c = 0; c < nchunks; ++c) {
          if (round > 0) tma::mbar_wait(&empty[slot], (round - 1) & 1);
          unsigned char* stage = ring + slot * STAGE;
          if (c == 0) header[slot] = item;
          tma::mbar_arrive_expect_tx(&full[slot], tx_bytes);
          const int k0 = c * KD;
          tma::load_3d(stage + 0 * TILE, &maps.stage[comp], it, j, k0, &full[slot]);
          tma::load_3d(stage + 1 * TILE, &maps.pos[comp], it, j, k0, &full[slot]);
          tma::load_3d(stage + 2 * TILE, &maps.tens[comp], it, j, k0, &full[slot]);
          tma::load_3d(stage + 3 * TILE, &maps.tensstage[comp], it, j, k0, &full[slot]);
          tma::load_3d(stage + 4 * TILE, &maps.wcon, it, j, k0, &full[slot]);
          if (SPLIT) tma::load_3d(stage + 4 * TILE + WMAIN, &maps.wcon_edge, it + COLS, j, k0, &full[slot]);
          if (jshift) tma::load_3d(stage + 4 * TILE + WTILE, &maps.wcon, it, j + 1, k0, &full[slot]);
          if (++slot == stages) {
            slot = 0;
            ++round;
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ===== compute warps =====
    const int t = threadIdx.x;
    const T dtr_stage = C::dtr_stage;
    // TMEM: lane quarter of the warp's scheduler; two warps on one scheduler split the columns
    const VadvSlots<T> slots{*tmem_base_smem + (uint32_t((warp & 3) * 32) << 16) +
                                 uint32_t((warp >> 2) * vcfg::tmem_cols_per_thread<T>()),
                             estore + t, paired};

    VadvForward<T> f;
    T z = 0;                // backward state of the old column: x - pos of the level above
    T* old_base = comps.tensstage[0];
    bool old_valid = false;
    bool any_batch = false;
    int slot = 0, round = 0;
    int dir = 0;  // slot direction of the NEW column: level k -> slot (dir ? nz-2-k : k)
    int ishift = 0, jshift = 0;  // neighbour shift of the running item's component

    // values of level r of the chunk in ring slot `slot`
    auto fetch = [&](const unsigned char* stage, int r, T& v_stage, T& v_pos, T& v_tens, T& v_tss,
                     T& v_wsum) {
      auto tile = [&](int field) { return reinterpret_cast<const T*>(stage + field * TILE)[r * COLS + t]; };
      v_stage = tile(0);
      v_pos = tile(1);
      v_tens = tile(2);
      v_tss = tile(3);
      // shifted + unshifted wcon, in the oracle's order: i+1 from the same (wider) tile,
      // j+1 from the second tile
      const T* w0 = reinterpret_cast<const T*>(stage + 4 * TILE) + r * WB + t;
      const T* w1 = reinterpret_cast<const T*>(stage + 4 * TILE + WTILE) + r * WB + t;
      T shifted;
      if (jshift) {
        shifted = w1[0];
      } else if (SPLIT && t + ishift >= COLS) {
        shifted = reinterpret_cast<const T*>(stage + 4 * TILE + WMAIN)[r * EDGE];  // wcon(i+1) of the last column
      } else {
        shifted = w0[ishift];
      }
      v_wsum = shifted + w0[0];
    };
    auto release = [&]() {
      __syncwarp();
      if ((t & 31) == 0) tma::mbar_arrive(&empty[slot]);
      if (++slot == stages) {
        slot = 0;
        ++round;
      }
    };

    for (;;) {
      // the stage that carries an item's first chunk names the item (the producer's end marker
      // is a stage of its own)
      tma::mbar_wait(&full[slot], round & 1);
      const int item = header[slot];
      if (item < 0) break;
      any_batch = true;
      const int comp = item % comps.ncomp;
      const int b = item / comps.ncomp;
      const int j = b / nbx;
      const int i = (b - j * nbx) * COLS + t;
      const bool new_valid = i < nx;
      T* new_base = comps.tensstage[comp] + int64_t(j) * sy + i;
      ishift = comps.ishift[comp];
      jshift = comps.jshift[comp];
      // slot of the new column's level s-1 / the old column's level nz-1-s at step s
      const int p0 = dir ? nz - 1 : -1;
      const int dp = dir ? -1 : 1;  // slot(s) = p0 + dp * s

      // ---- first chunk, peeled: levels 0 and 1 are special ----
      {
        const unsigned char* stage = ring + slot * STAGE;
        T v_stage[KD], v_pos[KD], v_tens[KD], v_tss[KD], v_wsum[KD];
#pragma unroll
        for (int r = 0; r < KD; ++r) fetch(stage, r, v_stage[r], v_pos[r], v_tens[r], v_tss[r], v_wsum[r]);
        release();
        VadvCursor<T, 2> slot(slots, p0, dp);
#pragma unroll
        for (int r = 0; r < KD; ++r) {
          if (r < nz)
            vadv_step<T, true, 2, STORE>(r, f, v_stage[r], v_pos[r], v_tens[r], v_tss[r], v_wsum[r], z, slot, r,
                                  old_base + int64_t(nz - 1 - r) * sz, old_valid);
        }
      }
      // ---- full chunks ----
      int s0 = KD;
      for (; s0 + KD <= nz; s0 += KD) {
        tma::mbar_wait(&full[slot], round & 1);
        const unsigned char* stage = ring + slot * STAGE;
        T v_stage[KD], v_pos[KD], v_tens[KD], v_tss[KD], v_wsum[KD];
#pragma unroll
        for (int r = 0; r < KD; ++r) fetch(stage, r, v_stage[r], v_pos[r], v_tens[r], v_tss[r], v_wsum[r]);
        release();
        rescale(f.p, f.r, f.q);
        // the slots of a chunk are consecutive: all paired, all split, or (once per sweep) mixed;
        // the two common cases run branch-free bodies
        const int pa = p0 + dp * s0, pb = p0 + dp * (s0 + KD - 1);
        const int kind = max(pa, pb) < paired ? 1 : (min(pa, pb) >= paired ? 0 : 2);
        auto body = [&](auto kind_tag) {
          constexpr int KIND = decltype(kind_tag)::value;
          VadvCursor<T, KIND> slot(slots, pa, dp);
          // output pointer of the old column.  STORE 0 (round 1): float64 schedules best when it is
          // recomputed per level, float32 (issue bound) when it is carried
          // (profiles/vadv_tuning_r01.log).  STORE 1: one 64-bit product per chunk, carried by
          // subtraction, next to the predicated store (a product per level costs 8 integer
          // instructions).  Computing all four pointers ahead of the chunk is slower in every
          // combination (profiles/vadv_storepath_r02.log).
          constexpr bool kRecompute = sizeof(T) == 8 && STORE == 0;
          T* out = old_base + int64_t(nz - 1 - s0) * sz;
#pragma unroll
          for (int r = 0; r < KD; ++r) {
            if (kRecompute) out = old_base + int64_t(nz - 1 - (s0 + r)) * sz;
            vadv_step<T, false, KIND, STORE>(s0 + r, f, v_stage[r], v_pos[r], v_tens[r], v_tss[r], v_wsum[r], z,
                                             slot, r, out, old_valid);
            if (!kRecompute) out -= sz;
          }
        };
        if (kind == 1)
          body(std::integral_constant<int, 1>{});
        else if (kind == 0)
          body(std::integral_constant<int, 0>{});
        else
          body(std::integral_constant<int, 2>{});
      }
      // ---- last, partial chunk ----
      if (s0 < nz) {
        tma::mbar_wait(&full[slot], round & 1);
        const unsigned char* stage = ring + slot * STAGE;
        rescale(f.p, f.r, f.q);
        VadvCursor<T, 2> slot(slots, p0 + dp * s0, dp);
        for (int s = s0; s < nz; ++s) {
          T v_stage, v_pos, v_tens, v_tss, v_wsum;
          fetch(stage, s - s0, v_stage, v_pos, v_tens, v_tss, v_wsum);
          vadv_step<T, false, 2, STORE>(s, f, v_stage, v_pos, v_tens, v_tss, v_wsum, z, slot, s - s0,
                                 old_base + int64_t(nz - 1 - s) * sz, old_valid);
        }
        release();
      }
      // ---- step nz: last level k = nz-1 of the new column (base.py:451-462) starts its
      //      backward sweep; the old column finished at step nz-1 ----
      {
        const T h = f.h_next;  // a = as = -h[nz-1], no c on the last level
        const T bb = dtr_stage + h;
        const T d0 = dtr_stage * f.pos_cur + f.tens_cur + f.tss_cur - f.t_prev;
        const T x_top = (d0 * f.q + h * f.r) / (bb * f.q + h * f.p);
        z = x_top - f.pos_cur;
        if (new_valid) new_base[int64_t(nz - 1) * sz] = dtr_stage * z;
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      old_base = new_base;
      old_valid = new_valid;
      dir ^= 1;
    }
    // ---- drain: backward sweep of the last column ----
    if (any_batch) {
      const int p0 = dir ? nz - 1 : -1;
      const int dp = dir ? -1 : 1;
      for (int s = 1; s <= nz - 1; ++s) {
        const int pslot = p0 + dp * s;
        T c_old, e_old;
        slots.template load<2>(pslot, c_old, e_old);
        z = e_old - c_old * z;
        if (old_valid) old_base[int64_t(nz - 1 - s) * sz] = dtr_stage * z;
      }
    }
  }

  // ---- teardown: the allocating warp frees TMEM after everybody is done ----
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(*tmem_base_smem),
                 "r"(vcfg::kTmemCols)
                 : "memory");
  }
}

// SB200_VADV_CFG="variant": 0 = auto, 1 = global, 2 = onchip (overrides the `variant` argument)
inline int vadv_variant_override() {
  if (const char* env = std::getenv("SB200_VADV_CFG")) return std::atoi(env);
  return 0;
}

template <class T>
struct VadvSystem {
  const T *stage, *pos, *tens;
  T* tensstage;
  int ishift, jshift;
};

template <class T, int KD>
int launch_vadv_onchip_kd(const VadvSystem<T>* systems, int ncomp, const T* wcon, int64_t nx, int64_t ny,
                          int64_t nz, int64_t sy, int64_t sz, int dry_runs, double* time, cudaStream_t stream,
                          bool* used) {
  constexpr int COLS = vcfg::cols<T>();
  *used = false;
  // TMEM: c of every slot, plus e of as many slots as the thread's remaining columns hold
  constexpr int TCOLS = int(sizeof(T)) / 4;
  constexpr int BUDGET = vcfg::tmem_cols_per_thread<T>();
  const int64_t slots = nz - 1;
  if (slots * TCOLS > BUDGET) return 0;
  int paired = int(std::min<int64_t>(slots, (BUDGET - slots * TCOLS) / TCOLS));
  if (const char* env = std::getenv("SB200_VADV_PAIRED")) paired = std::min(paired, std::atoi(env));
  int max_stages = 8;
  if (const char* env = std::getenv("SB200_VADV_STAGES")) max_stages = std::max(2, std::atoi(env));
  int ishift = 0, jshift = 0;
  for (int c = 0; c < ncomp; ++c) {
    ishift = std::max(ishift, systems[c].ishift);
    jshift = std::max(jshift, systems[c].jshift);
  }
  // as many ring stages as fit next to the shared-memory part of the store (2 ... 8)
  const size_t stage_size = 4 * vcfg::tile_bytes<T>(KD) + (jshift ? 2 : 1) * vcfg::wcon_tile_bytes<T>(KD);
  const size_t fixed = size_t(slots - paired) * COLS * sizeof(T) + 256;
  if (fixed + 2 * stage_size > 227 * 1024) return 0;
  const int stages = int(std::min<size_t>(size_t(max_stages), (227 * 1024 - fixed) / stage_size));
  const size_t smem = stages * stage_size + fixed;
  const auto type = tma::tensor_type<T>();
  const uint64_t s1 = uint64_t(sy) * sizeof(T), s2 = uint64_t(sz) * sizeof(T);
  VadvMaps maps;
  VadvComponents<T> comps;
  comps.ncomp = ncomp;
  comps.two_wcon_tiles = jshift;
  for (int c = 0; c < 3; ++c) {
    const VadvSystem<T>& sys = systems[c < ncomp ? c : 0];
    if (!tma::encode_3d(&maps.stage[c], type, sys.stage, nx, ny, nz, s1, s2, COLS, 1, KD) ||
        !tma::encode_3d(&maps.pos[c], type, sys.pos, nx, ny, nz, s1, s2, COLS, 1, KD) ||
        !tma::encode_3d(&maps.tens[c], type, sys.tens, nx, ny, nz, s1, s2, COLS, 1, KD) ||
        !tma::encode_3d(&maps.tensstage[c], type, sys.tensstage, nx, ny, nz, s1, s2, COLS, 1, KD))
      return 0;
    comps.tensstage[c] = sys.tensstage;
    comps.ishift[c] = sys.ishift;
    comps.jshift[c] = sys.jshift;
  }
  // the wcon extent includes the neighbour column / row some component reads
  if (!tma::encode_3d(&maps.wcon, type, wcon, nx + ishift, ny + jshift, nz, s1, s2, vcfg::wcon_width<T>(), 1, KD) ||
      !tma::encode_3d(&maps.wcon_edge, type, wcon, nx + ishift, ny + jshift, nz, s1, s2, vcfg::edge_elems<T>(), 1,
                      KD))
    return 0;
  // SB200_VADV_STORE (tuning aid): 0 = conditional stores, 1 = predicated stores + one pointer product per chunk
  int store_mode = sizeof(T) == 8 ? 1 : 0;
  if (const char* env = std::getenv("SB200_VADV_STORE")) store_mode = std::atoi(env) != 0;
  static std::atomic<uint64_t> attr_done{0};
  {
    int device = 0;
    SB200_CHECK(cudaGetDevice(&device));
    const uint64_t bit = uint64_t(1) << (device & 63);
    if (!(attr_done.load(std::memory_order_acquire) & bit)) {
      SB200_CHECK(cudaFuncSetAttribute(vadv_onchip_kernel<T, KD, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       227 * 1024));
      SB200_CHECK(cudaFuncSetAttribute(vadv_onchip_kernel<T, KD, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       227 * 1024));
      attr_done.fetch_or(bit, std::memory_order_release);
    }
  }
  const int64_t nitems = ceil_div(nx, COLS) * ny * ncomp;
  if (nitems > 0x7fffff00) return fail("sb200_vadv: domain too large for the on-chip variant");
  const unsigned grid = unsigned(std::min<int64_t>(nitems, sm_count()));
  *used = true;
  bool counter_ok = true;
  auto launch = [&] {
    unsigned int* counter = work_counter();
    if (counter == nullptr || cudaMemsetAsync(counter, 0, sizeof(unsigned int), stream) != cudaSuccess) {
      counter_ok = false;
      return;
    }
    if (store_mode)
      vadv_onchip_kernel<T, KD, 1><<<grid, vcfg::threads<T>(), smem, stream>>>(
          maps, comps, counter, int(nx), int(ny), int(nz), sy, sz, stages, paired);
    else
      vadv_onchip_kernel<T, KD, 0><<<grid, vcfg::threads<T>(), smem, stream>>>(
          maps, comps, counter, int(nx), int(ny), int(nz), sy, sz, stages, paired);
    count_launch();
  };
  const int rc = timed(launch, dry_runs, time, stream);
  if (!counter_ok) return fail("sb200_vadv: cannot allocate or reset the work counter");
  return rc;
}

// Levels per ring stage: 4 for float64, 8 for float32 (SB200_VADV_KD overrides, a tuning aid).
// A stage costs a fixed ~40 instructions per warp (barrier wait, release, slot bookkeeping, the
// renormalisation of the forward triple); the float32 kernel -- two warps per scheduler, issue
// bound -- gains 10 % from paying them half as often (0.625 against 0.696 ms, u/v/w 1.72 against
// 2.02 ms), the float64 kernel, which waits for HBM as often as for its own instructions, gains
// nothing (1.200 against 1.202 ms) and would give up four of its seven stages of prefetch depth
// (profiles/vadv_kd8_r02.log).
template <class T>
int launch_vadv_onchip(const VadvSystem<T>* systems, int ncomp, const T* wcon, int64_t nx, int64_t ny,
                       int64_t nz, int64_t sy, int64_t sz, int dry_runs, double* time, cudaStream_t stream,
                       bool* used) {
  int kd = sizeof(T) == 4 ? 8 : 4;
  if (const char* env = std::getenv("SB200_VADV_KD")) kd = std::atoi(env);
  if (kd == 8)
    return launch_vadv_onchip_kd<T, 8>(systems, ncomp, wcon, nx, ny, nz, sy, sz, dry_runs, time, stream, used);
  return launch_vadv_onchip_kd<T, 4>(systems, ncomp, wcon, nx, ny, nz, sy, sz, dry_runs, time, stream, used);
}

template <class T>
int launch_vadv(const VadvSystem<T>* systems, int ncomp, const T* wcon, T* ccol, T* dcol, int64_t nx,
                int64_t ny, int64_t nz, int64_t sy, int64_t sz, int variant, int dry_runs, double* time,
                cudaStream_t stream) {
  if (const int forced = vadv_variant_override()) variant = forced;
  constexpr int V = VecN<T>::value;
  bool aligned = aligned_to(wcon, 16) && sy % V == 0 && sz % V == 0;
  for (int c = 0; c < ncomp; ++c)
    aligned = aligned && aligned_to(systems[c].stage, 16) && aligned_to(systems[c].pos, 16) &&
              aligned_to(systems[c].tens, 16) && aligned_to(systems[c].tensstage, 16);
  if (variant != SB200_VADV_GLOBAL) {
    // the on-chip variant pays off once most of a 128-column batch is populated
    const bool wanted = variant == SB200_VADV_ONCHIP || nx >= 64;
    bool used = false;
    if (aligned && wanted) {
      const int rc = launch_vadv_onchip<T>(systems, ncomp, wcon, nx, ny, nz, sy, sz, dry_runs, time, stream, &used);
      if (used || rc != 0) return rc;
    }
    if (variant == SB200_VADV_ONCHIP)
      return fail("sb200_vadv: on-chip variant not available for this configuration "
                  "(needs 16-byte aligned fields and (nz-1)*128*sizeof(T) bytes of shared memory)");
  }
  if (ccol == nullptr || dcol == nullptr) return fail("sb200_vadv: ccol/dcol scratch fields are required by the global variant");
  int bx = 64;
  while (bx > 32 && bx / 2 >= nx) bx /= 2;
  const dim3 block(bx, 1, 1);
  const dim3 grid(unsigned(ceil_div(nx, bx)), unsigned(ny), 1);
  if (grid.y > 65535u) return fail("sb200_vadv: domain too large for the launch grid");
  // the components share the ccol / dcol scratch fields: one sweep after the other
  auto launch = [&] {
    for (int c = 0; c < ncomp; ++c) {
      const VadvSystem<T>& sys = systems[c];
      const int64_t wshift = int64_t(sys.ishift) + int64_t(sys.jshift) * sy;
      vadv_global_kernel<T><<<grid, block, 0, stream>>>(sys.stage, sys.pos, sys.tens, sys.tensstage, wcon, ccol,
                                                        dcol, int(nx), int(ny), int(nz), sy, sz, wshift);
      count_launch();
    }
  };
  return timed(launch, dry_runs, time, stream);
}

template <class T>
int vadv_entry(int ncomp, const void* const* ustage, const void* const* upos, const void* const* utens,
               void* const* utensstage, const int* ishift, const int* jshift, const void* wcon, void* ccol,
               void* dcol, int64_t nx, int64_t ny, int64_t nz, int64_t sy, int64_t sz, int variant, int dry_runs,
               double* time, cudaStream_t stream) {
  VadvSystem<T> systems[3];
  for (int c = 0; c < ncomp; ++c)
    systems[c] = {static_cast<const T*>(ustage[c]), static_cast<const T*>(upos[c]), static_cast<const T*>(utens[c]),
                  static_cast<T*>(utensstage[c]), ishift[c], jshift[c]};
  return launch_vadv<T>(systems, ncomp, static_cast<const T*>(wcon), static_cast<T*>(ccol), static_cast<T*>(dcol),
                        nx, ny, nz, sy, sz, variant, dry_runs, time, stream);
}

}  // namespace
}  // namespace sb200

using namespace sb200;

extern "C" int sb200_vadv_components(int dtype, int ncomp, const void* const* ustage,
                                     const void* const* upos, const void* const* utens,
                                     void* const* utensstage, const int* ishift, const int* jshift,
                                     const void* wcon, void* ccol, void* dcol, int64_t nx, int64_t ny,
                                     int64_t nz, int64_t sx, int64_t sy, int64_t sz, int variant,
                                     int dry_runs, double* time, void* stream) {
  if (ncomp < 1 || ncomp > 3) return fail("sb200_vadv: between one and three components per launch");
  if (ustage == nullptr || upos == nullptr || utens == nullptr || utensstage == nullptr || ishift == nullptr ||
      jshift == nullptr)
    return fail("sb200_vadv: null component table");
  if (nx <= 0 || ny <= 0 || nz <= 0) return fail("sb200_vadv: domain must be positive");
  if (nz < 2) return fail("sb200_vadv: at least two vertical levels are required");
  if (sx != 1) return fail("sb200_vadv: only layout (2,1,0) is supported (unit stride along i)");
  for (int c = 0; c < ncomp; ++c)
    if (ishift[c] < 0 || ishift[c] > 1 || jshift[c] < 0 || jshift[c] > 1)
      return fail("sb200_vadv: shifts must be 0 or 1");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == SB200_F64)
    return vadv_entry<double>(ncomp, ustage, upos, utens, utensstage, ishift, jshift, wcon, ccol, dcol, nx, ny, nz,
                              sy, sz, variant, dry_runs, time, s);
  if (dtype == SB200_F32)
    return vadv_entry<float>(ncomp, ustage, upos, utens, utensstage, ishift, jshift, wcon, ccol, dcol, nx, ny, nz,
                             sy, sz, variant, dry_runs, time, s);
  return fail("sb200_vadv: unsupported dtype");
}

extern "C" int sb200_vadv(int dtype, const void* ustage, const void* upos, const void* utens,
                          void* utensstage, const void* wcon, void* ccol, void* dcol, void* datacol,
                          int64_t nx, int64_t ny, int64_t nz, int64_t sx, int64_t sy, int64_t sz,
                          int ishift, int jshift, int variant, int dry_runs, double* time,
                          void* stream) {
  (void)datacol;
  return sb200_vadv_components(dtype, 1, &ustage, &upos, &utens, &utensstage, &ishift, &jshift, wcon, ccol, dcol,
                               nx, ny, nz, sx, sy, sz, variant, dry_runs, time, stream);
}
