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
// Two kernels share the per-thread arithmetic below:
//   * "tma"    (fast path, 16-byte aligned fields): inp halo tiles and coeff tiles are staged
//              in shared memory by TMA (cp.async.bulk.tensor) through an mbarrier ring fed by a
//              producer warp; four consumer warps march over j reading the tiles with LDS.128.
//   * "jmarch" (any alignment, tiny domains): the same march with direct global loads.
//
// Per-thread scheme: a thread owns VEC consecutive i (one 128-bit vector: 2
// doubles / 4 floats) and marches over JT rows of j.  Per row it loads the
// VEC+4 wide inp strip [i0-2, i0+VEC+2) as aligned vectors, the coeff vector,
// and keeps in registers the rolling state that the oracle holds in whole
// temporaries: two Laplacian rows (VEC+2 wide), the previous fly row and two
// inp rows.  Every Laplacian / flux is computed once per thread; only the two
// i-edge Laplacians and one i-edge flx are recomputed by the neighbouring
// thread, which keeps the FP64 work at ~24 instructions per point (the
// on-the-fly form needs ~46 and would be FP64-bound on B200).  The march needs
// no block-wide barrier in either kernel.  HBM traffic is the algorithmic
// minimum plus the 4 rows per march segment that two segments both read.
#include <cstdlib>
#include <mutex>
#include <utility>
#include <vector>

#include "common.cuh"
#include "tma.cuh"

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

// ---------------------------------------------------------------------------------
// TMA kernel
// ---------------------------------------------------------------------------------
// CTA = 4 consumer warps + 1 producer warp.  Tile: 128 threads x 16 bytes = 2 KB of i
// (256 doubles / 512 floats), marched over `jt` rows of one k level.  Per pipeline
// stage the producer issues three TMA boxes of R rows each (8-byte elements):
//   A: [i_t - 16 B, +2048 B)   inp, main part        R x 2048 B
//   B: [i_t + 2032 B, +32 B)   inp, right halo       R x   32 B   (TMA boxes are <= 256 elements)
//   C: coeff tile                                     R x 2048 B
// so row q of inp is the byte range [-16, 2064) around the tile = 130 16-byte vectors;
// thread t reads vectors t, t+1, t+2 (its own values and the two i-halo values on each
// side).  inp row q completes output row q-2, whose coeff row travels in the same stage.
namespace tmacfg {
constexpr int kConsumers = 128;
constexpr int kThreads = kConsumers + 32;
constexpr int kRowBytes = 2048;
constexpr int kHaloBytes = 32;
// the halo region of a stage reserves 128 bytes per row: one R-row box packs its rows at 32 bytes,
// the per-row boxes of the fused halo exchange need 128-byte aligned destinations
constexpr int kHaloPitchEdge = 128;
__host__ __device__ constexpr int stage_bytes(int rows) { return rows * (2 * kRowBytes + kHaloPitchEdge); }
__host__ __device__ constexpr int stage_tx_bytes(int rows) { return rows * (2 * kRowBytes + kHaloBytes); }
__host__ __device__ constexpr int smem_bytes(int rows, int stages) { return stages * stage_bytes(rows) + 2 * stages * 8; }
}  // namespace tmacfg

template <class T, int VEC>
__device__ __forceinline__ void read_strip(const unsigned char* a_row, const unsigned char* b_row,
                                           int t, Strip<T, VEC>& s) {
  auto vec = [&](int v) -> const unsigned char* {
    return v < tmacfg::kConsumers ? a_row + 16 * v : b_row + 16 * (v - tmacfg::kConsumers);
  };
  if constexpr (sizeof(T) == 8) {
    const double2 v0 = *reinterpret_cast<const double2*>(vec(t));
    const double2 v1 = *reinterpret_cast<const double2*>(vec(t + 1));
    const double2 v2 = *reinterpret_cast<const double2*>(vec(t + 2));
    s.v[0] = v0.x; s.v[1] = v0.y; s.v[2] = v1.x; s.v[3] = v1.y; s.v[4] = v2.x; s.v[5] = v2.y;
  } else {
    const float4 v0 = *reinterpret_cast<const float4*>(vec(t));
    const float4 v1 = *reinterpret_cast<const float4*>(vec(t + 1));
    const float4 v2 = *reinterpret_cast<const float4*>(vec(t + 2));
    s.v[0] = v0.z; s.v[1] = v0.w;
    s.v[2] = v1.x; s.v[3] = v1.y; s.v[4] = v1.z; s.v[5] = v1.w;
    s.v[6] = v2.x; s.v[7] = v2.y;
  }
}

// Fused halo exchange (multi-GPU): the j-halo rows of a slab are the edge rows of the neighbouring
// GPUs' slabs.  Instead of copying them into the local halo first, the producer of an edge
// segment fetches those rows straight from the neighbour's HBM over NVLink -- a TMA load on a
// tensor map whose base is the peer allocation (CUDA IPC mapping) -- one row per box, so local and
// remote rows of a stage can be mixed freely.  Index 0 = this GPU, 1 = lower, 2 = upper neighbour.
struct PeerMaps {
  CUtensorMap row_inp[3];   // box: 256 x 1 x 1 (8-byte elements)
  CUtensorMap row_halo[3];  // box:   4 x 1 x 1
  int ny_lower;             // rows of the lower neighbour's slab
  int has_lower, has_upper;
};

// Time loop over J slabs (inp and out swap roles every step): what orders the sweeps of
// neighbouring GPUs.  "Edge CTAs" are the CTAs that read a neighbour's rows and whose output rows
// a neighbour reads: the first segment of every (i tile, level) towards the lower neighbour, and
// towards the upper one every segment that ends at row ny-1 or ny (the last one, and the one
// before it when the last segment is a single row).  When an edge CTA has stored its rows, each of
// its four consumer warps adds 1 to the counter of its LEVEL in the neighbour's memory
// (red.release.sys over NVLink) -- after sweep m-1 a level's counter stands at m * (edge warps per
// level and sweep).  The producer of an edge CTA of sweep m spins on the LOCAL counter of its level
// (ld.acquire.sys) until it has reached that value before it issues its first load: then the
// neighbour's rows of the field it is about to read are complete (they were written in the
// neighbour's sweep m-1), and the neighbour no longer reads the rows this CTA is about to overwrite
// (it read them in its sweep m-1).  Counters are per level, so a CTA waits for the handful of
// neighbour CTAs it really depends on -- which finished most of a sweep ago -- and not for the
// neighbour's whole sweep.  No host involvement, no extra launch; CTAs that are not on an edge never
// look at a counter.  Sweeps of one GPU are ordered by the stream, so a waiting CTA only ever
// depends on sweeps that do not depend on it.
struct StepFlags {
  const unsigned int* arrived;   // local counters: [k] pushed by the lower, [levels + k] by the upper neighbour
  unsigned int* notify[2];       // the neighbours' counters this slab pushes: [0] lower (its upper half), [1] upper
  unsigned int wait_value[2];    // what a level's counter shows once the neighbour's previous sweep is through it
  int levels;
};

__device__ __forceinline__ void wait_for_neighbour(const unsigned int* counter, unsigned int value) {
  unsigned int seen;
  do {
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(counter) : "memory");
  } while (int(seen - value) < 0);
}

// Work decomposition of one sweep: CTA = (256-wide i tile, segment of jt rows, level k), numbered
// i tile fastest, then segment, then level: the CTAs resident at any time sweep one compact band of
// the field, and the eight i tiles of a segment, which together read whole rows, run side by side.
// (Segments graded from long to short towards the end of the launch, persistent CTAs walking a
// segment list through one continuous TMA ring -- with static and with dynamic assignment -- and L2
// eviction hints on the TMA loads were all built and measured; none beats this:
// profiles/hdiff_segments_r01.log, profiles/hdiff_persist_*_r02.log, profiles/hdiff_variants_r01.log.)
struct HdiffTiling {
  int xtiles;
  int segments;  // per (i tile, level)
  int jt;        // rows per segment
};

template <class T, int R, int S, bool PEER>
__global__ void __launch_bounds__(tmacfg::kThreads, 4)
    hdiff_tma_kernel(const __grid_constant__ CUtensorMap map_inp,
                     const __grid_constant__ CUtensorMap map_halo,
                     const __grid_constant__ CUtensorMap map_coeff,
                     const __grid_constant__ PeerMaps peer, const StepFlags flags,
                     T* __restrict__ out, int nx, int ny, const HdiffTiling tiling, int64_t sy,
                     int64_t sz) {
  constexpr int VEC = VecN<T>::value;
  constexpr int TW = tmacfg::kConsumers * VEC;
  constexpr int STAGE = tmacfg::stage_bytes(R);
  extern __shared__ __align__(128) unsigned char smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + S * STAGE);
  uint64_t* empty = full + S;

  int b = int(blockIdx.x);
  const int xt = b % tiling.xtiles;
  b /= tiling.xtiles;
  const int k = b / tiling.segments;
  const int jt = tiling.jt;
  const int it = xt * TW;  // first i of the tile
  const int jb = (b - k * tiling.segments) * jt;
  const int je = min(jb + jt, ny);
  const int nstages = (je - jb + 4 + R - 1) / R;  // rows jb-2 .. je+1
  const int warp = threadIdx.x >> 5;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      tma::mbar_init(&full[s], 1);
      tma::mbar_init(&empty[s], tmacfg::kConsumers / 32);
    }
    tma::fence_barrier_init();
  }
  __syncthreads();

  if (warp == tmacfg::kConsumers / 32) {
    // ===== producer warp: one lane drives the TMA ring =====
    if ((threadIdx.x & 31) == 0) {
      tma::prefetch_tensormap(&map_inp);
      tma::prefetch_tensormap(&map_halo);
      tma::prefetch_tensormap(&map_coeff);
      if (PEER) {
        for (int w = 0; w < 3; ++w) {
          tma::prefetch_tensormap(&peer.row_inp[w]);
          tma::prefetch_tensormap(&peer.row_halo[w]);
        }
      }
      const int c0 = it / (8 / int(sizeof(T)));  // tile origin in 8-byte elements
      if (PEER && (flags.wait_value[0] | flags.wait_value[1]) != 0) {
        // time loop: an edge CTA starts once the neighbour's previous sweep has finished with the
        // rows both touch; the async proxy (TMA) must not run ahead of the acquire
        const bool lower_edge = peer.has_lower && jb == 0;
        const bool upper_edge = peer.has_upper && je >= ny - 1;
        if (lower_edge) wait_for_neighbour(flags.arrived + k, flags.wait_value[0]);
        if (upper_edge) wait_for_neighbour(flags.arrived + flags.levels + k, flags.wait_value[1]);
        if (lower_edge || upper_edge) asm volatile("fence.proxy.async;" ::: "memory");
      }
      for (int n = 0; n < nstages; ++n) {
        const int slot = n % S;
        if (n >= S) tma::mbar_wait(&empty[slot], ((n / S) - 1) & 1);
        unsigned char* stage = smem + slot * STAGE;
        // the coeff rows of a stage belong to the output rows it completes: j = jb-4+nR+r; the
        // first R = 4 rows of a segment complete nothing, so stage 0 carries no coeff tile
        const bool with_coeff = (n + 1) * R > 4;
        tma::mbar_arrive_expect_tx(&full[slot], tmacfg::stage_tx_bytes(R) - (with_coeff ? 0 : R * tmacfg::kRowBytes));
        // tensor origins: inp at (i = -16 B, j = -2), coeff at (i = 0, j = 0)
        bool remote_rows = false;
        if (PEER) {
          const int q0 = jb - 2 + n * R;  // first inp row of the stage
          remote_rows = (peer.has_lower && q0 < 0) || (peer.has_upper && q0 + R > ny);
        }
        if (PEER && remote_rows) {
          // edge stage: row by row, each from the GPU that owns it
          for (int r = 0; r < R; ++r) {
            const int q = jb - 2 + n * R + r;
            int who = 0, row = q;
            if (peer.has_lower && q < 0) {
              who = 1;
              row = peer.ny_lower + q;
            } else if (peer.has_upper && q >= ny) {
              who = 2;
              row = q - ny;
            }
            tma::load_3d(stage + r * tmacfg::kRowBytes, &peer.row_inp[who], c0, row + 2, k, &full[slot]);
            tma::load_3d(stage + 2 * R * tmacfg::kRowBytes + r * tmacfg::kHaloPitchEdge, &peer.row_halo[who],
                         c0 + 256, row + 2, k, &full[slot]);
          }
          if (with_coeff)
            tma::load_3d(stage + R * tmacfg::kRowBytes, &map_coeff, c0, jb - 4 + n * R, k, &full[slot]);
        } else {
          tma::load_3d(stage, &map_inp, c0, jb + n * R, k, &full[slot]);
          tma::load_3d(stage + 2 * R * tmacfg::kRowBytes, &map_halo, c0 + 256, jb + n * R, k, &full[slot]);
          if (with_coeff)
            tma::load_3d(stage + R * tmacfg::kRowBytes, &map_coeff, c0, jb - 4 + n * R, k, &full[slot]);
        }
      }
    }
    return;
  }

  // ===== consumer warps =====
  const int t = threadIdx.x;
  const int i0 = it + t * VEC;
  const bool active = i0 < nx;
  const bool whole = i0 + VEC <= nx;
  T* __restrict__ op = out + int64_t(k) * sz + i0;

  Strip<T, VEC> rc, rn, rnn;
  T lc[VEC + 2], ln[VEC + 2], fym[VEC];
#pragma unroll
  for (int m = 0; m < VEC + 2; ++m) lc[m] = ln[m] = T(0);
#pragma unroll
  for (int n = 0; n < VEC; ++n) fym[n] = T(0);
#pragma unroll
  for (int n = 0; n < VEC + 4; ++n) rc.v[n] = rn.v[n] = T(0);

  for (int n = 0; n < nstages; ++n) {
    const int slot = n % S;
    tma::mbar_wait(&full[slot], (n / S) & 1);
    const unsigned char* stage = smem + slot * STAGE;
    int halo_pitch = tmacfg::kHaloBytes;
    if (PEER) {
      const int q0 = jb - 2 + n * R;
      if ((peer.has_lower && q0 < 0) || (peer.has_upper && q0 + R > ny)) halo_pitch = tmacfg::kHaloPitchEdge;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int j = jb + n * R + r - 4;  // output row completed by inp row j + 2
      read_strip<T, VEC>(stage + r * tmacfg::kRowBytes,
                         stage + 2 * R * tmacfg::kRowBytes + r * halo_pitch, t, rnn);
      T cf[VEC];
      {
        const unsigned char* c_row = stage + R * tmacfg::kRowBytes + r * tmacfg::kRowBytes + 16 * t;
        if constexpr (sizeof(T) == 8) {
          const double2 c = *reinterpret_cast<const double2*>(c_row);
          cf[0] = c.x; cf[1] = c.y;
        } else {
          const float4 c = *reinterpret_cast<const float4*>(c_row);
          cf[0] = c.x; cf[1] = c.y; cf[2] = c.z; cf[3] = c.w;
        }
      }
      // rows so far: rc = row j, rn = row j+1, rnn = row j+2; lc = lap(j), fym = fly(j-1).
      // The first four rows of a segment (n == 0) only fill this state; their results
      // are computed from zero-initialised registers and never stored (j < jb).
      laplacian<T, VEC>(rc, rn, rnn, ln);  // lap(j+1)
      T flx[VEC + 1];
#pragma unroll
      for (int m = 0; m < VEC + 1; ++m)
        flx[m] = limited(lc[m + 1] - lc[m], rc.v[m + 2] - rc.v[m + 1]);
      T res[VEC];
#pragma unroll
      for (int m = 0; m < VEC; ++m) {
        const T fy = limited(ln[m + 1] - lc[m + 1], rn.v[m + 2] - rc.v[m + 2]);
        res[m] = rc.v[m + 2] - cf[m] * (flx[m + 1] - flx[m] + fy - fym[m]);
        fym[m] = fy;
      }
      if (j >= jb && j < je && active) {
        if (whole) {
          store_vec<VEC, Cache::Streaming>(op + int64_t(j) * sy, res);
        } else {
#pragma unroll
          for (int m = 0; m < VEC; ++m)
            if (i0 + m < nx) op[int64_t(j) * sy + m] = res[m];
        }
      }
      rc = rn;
      rn = rnn;
#pragma unroll
      for (int m = 0; m < VEC + 2; ++m) lc[m] = ln[m];
    }
    __syncwarp();
    if ((threadIdx.x & 31) == 0) tma::mbar_arrive(&empty[slot]);
  }
  if (PEER) {
    // time loop: tell the neighbours that this edge CTA's rows are written (and that it no longer
    // reads theirs); the release at system scope covers the whole warp's stores
    const bool lower_edge = peer.has_lower && jb == 0 && flags.notify[0] != nullptr;
    const bool upper_edge = peer.has_upper && je >= ny - 1 && flags.notify[1] != nullptr;
    if (lower_edge || upper_edge) {
      // acq_rel is all a release needs (__threadfence_system() is the sequentially consistent
      // fence, MEMBAR.SC.SYS: measurably slower with 128 threads of every edge CTA issuing one)
      asm volatile("fence.acq_rel.sys;" ::: "memory");
      __syncwarp();
      if ((threadIdx.x & 31) == 0) {
        if (lower_edge)
          asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(flags.notify[0] + k) : "memory");
        if (upper_edge)
          asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(flags.notify[1] + k) : "memory");
      }
    }
  }
}

// SB200_HDIFF_CFG="variant,jt,pipeline" (tuning aid): variant 0 = auto, 1 = jmarch, 2 = tma;
// jt = rows per segment (0 = auto); pipeline = ring shape (see launch_hdiff_tma)
struct HdiffConfig {
  int variant = 0;
  int jt = 0;
  int pipeline = 0;
};

inline HdiffConfig hdiff_config() {
  HdiffConfig cfg;
  if (const char* env = std::getenv("SB200_HDIFF_CFG"))
    std::sscanf(env, "%d,%d,%d", &cfg.variant, &cfg.jt, &cfg.pipeline);
  return cfg;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is per device: set once per (kernel, device)
template <class Kernel>
int ensure_dynamic_smem(Kernel kernel, int smem, std::atomic<uint64_t>& done) {
  int device = 0;
  SB200_CHECK(cudaGetDevice(&device));
  const uint64_t bit = uint64_t(1) << (device & 63);
  if (done.load(std::memory_order_acquire) & bit) return 0;
  SB200_CHECK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  done.fetch_or(bit, std::memory_order_release);
  return 0;
}

// Segments of one sweep (host only).  R = rows per TMA stage.
inline bool hdiff_make_tiling(int R, int64_t xtiles, int64_t ny, int64_t nz, int jt_request,
                              HdiffTiling& tiling, int64_t& ctas_total) {
  int jt = jt_request;
  if (jt <= 0) {
    // enough CTAs for ~8 per SM; otherwise 32-row segments.  Longer marches re-read fewer rows
    // (4 per segment, and those mostly hit L2 because neighbouring segments run concurrently) but
    // measure slower: 128 rows 1.242 ms, 48: 1.229, 32: 1.196, 16: 1.167, 8: 1.519 for a single
    // sweep; in a loop 32 rows is the fastest on every box seen (profiles/hdiff_segments_r01.log,
    // profiles/hdiff_persist_dynamic_r02.log)
    const int64_t target = 148 * 8;
    int64_t segments = ceil_div(target, xtiles * nz);
    jt = int(std::min<int64_t>(32, std::max<int64_t>(16, ceil_div(ny, segments))));
  }
  jt = int(ceil_div(jt, R)) * R;
  tiling.xtiles = int(xtiles);
  tiling.jt = jt;
  tiling.segments = int(ceil_div(ny, jt));
  ctas_total = xtiles * tiling.segments * nz;
  return ctas_total <= 0x7fffffff;
}

// Tensor maps of one (fields, geometry) combination.  Encoding them is a handful of driver calls
// (3 maps, 9 with neighbours); a time loop sweeps the same fields again and again, so they are
// kept: the key is everything the maps depend on, the cache a small ring.
struct HdiffMapsKey {
  const void *inp, *coeff, *inp_lower, *inp_upper;
  int64_t nx, ny, nz, sy, sz, ny_lower, sz_lower, ny_upper, sz_upper;
  int element, rows, device;
  bool operator==(const HdiffMapsKey& o) const {
    return inp == o.inp && coeff == o.coeff && inp_lower == o.inp_lower && inp_upper == o.inp_upper &&
           nx == o.nx && ny == o.ny && nz == o.nz && sy == o.sy && sz == o.sz && ny_lower == o.ny_lower &&
           sz_lower == o.sz_lower && ny_upper == o.ny_upper && sz_upper == o.sz_upper &&
           element == o.element && rows == o.rows && device == o.device;
  }
};
struct HdiffMaps {
  CUtensorMap inp, halo, coeff;
  PeerMaps peer;
};

// 0 = encoded / found, 1 = error (reported), 2 = the TMA path does not apply to these fields
template <class T, int R>
int hdiff_maps(const T* inp, const T* coeff, int64_t nx, int64_t ny, int64_t nz, int64_t sy, int64_t sz,
               const T* inp_lower, int64_t ny_lower, int64_t sz_lower, const T* inp_upper, int64_t ny_upper,
               int64_t sz_upper, HdiffMaps& maps) {
  constexpr int E = 8 / int(sizeof(T));  // data elements per 8-byte TMA element
  constexpr int kEntries = 32;
  static std::mutex mutex;
  static std::vector<std::pair<HdiffMapsKey, HdiffMaps>> cache;
  static size_t next = 0;
  HdiffMapsKey key{inp, coeff, inp_lower, inp_upper, nx, ny, nz, sy, sz, ny_lower, sz_lower, ny_upper, sz_upper,
                   int(sizeof(T)), R, 0};
  if (cudaGetDevice(&key.device) != cudaSuccess) return fail("sb200_hdiff: no current device");
  {
    std::lock_guard<std::mutex> lock(mutex);
    for (const auto& entry : cache)
      if (entry.first == key) {
        maps = entry.second;
        return 0;
      }
  }
  // inp tensor: origin 16 bytes left of i = 0 and two rows below j = 0; extent covers
  // i in [-16 B, nx + 2), j in [-2, ny + 2)
  const T* inp_origin = inp - 16 / int(sizeof(T)) - 2 * sy;
  const uint64_t inp_d0 = uint64_t(ceil_div(16 + (nx + 2) * int64_t(sizeof(T)), 8));
  if (sizeof(T) == 4) {
    // float32: the origin lies 2 elements outside a halo of width 2 and an odd nx rounds the
    // extent up by one element; only take the TMA path if that memory belongs to the allocation
    const T* last = inp_origin + (nz - 1) * sz + (ny + 3) * sy + inp_d0 * E;
    if (!tma::range_is_allocated(inp_origin, last)) return 2;
  }
  const uint64_t bsy = uint64_t(sy) * sizeof(T), bsz = uint64_t(sz) * sizeof(T);
  // the right-halo box (4 x 8 bytes per row) needs its own descriptor: the box shape is part of it
  if (!tma::encode_3d_u64(&maps.inp, inp_origin, inp_d0, uint64_t(ny + 4), uint64_t(nz), bsy, bsz, 256, R, 1) ||
      !tma::encode_3d_u64(&maps.coeff, coeff, uint64_t(ceil_div(nx, E)), uint64_t(ny), uint64_t(nz), bsy, bsz,
                          256, R, 1) ||
      !tma::encode_3d_u64(&maps.halo, inp_origin, inp_d0, uint64_t(ny + 4), uint64_t(nz), bsy, bsz, 4, R, 1))
    return 2;
  // fused halo exchange: per-row maps on this slab and on the neighbours' slabs
  maps.peer.ny_lower = int(ny_lower);
  maps.peer.has_lower = inp_lower != nullptr;
  maps.peer.has_upper = inp_upper != nullptr;
  if (inp_lower != nullptr || inp_upper != nullptr) {
    const T* bases[3] = {inp, inp_lower ? inp_lower : inp, inp_upper ? inp_upper : inp};
    const int64_t rows[3] = {ny, inp_lower ? ny_lower : ny, inp_upper ? ny_upper : ny};
    // slabs with different row counts have different k strides
    const int64_t kstride[3] = {sz, inp_lower ? sz_lower : sz, inp_upper ? sz_upper : sz};
    for (int w = 0; w < 3; ++w) {
      const T* origin = bases[w] - 16 / int(sizeof(T)) - 2 * sy;
      if (!tma::encode_3d_u64(&maps.peer.row_inp[w], origin, inp_d0, uint64_t(rows[w] + 4), uint64_t(nz), bsy,
                              uint64_t(kstride[w]) * sizeof(T), 256, 1, 1) ||
          !tma::encode_3d_u64(&maps.peer.row_halo[w], origin, inp_d0, uint64_t(rows[w] + 4), uint64_t(nz), bsy,
                              uint64_t(kstride[w]) * sizeof(T), 4, 1, 1))
        return fail("sb200_hdiff_peer: cannot encode the tensor maps of the neighbouring slabs");
    }
  }
  std::lock_guard<std::mutex> lock(mutex);
  if (cache.size() < kEntries) {
    cache.emplace_back(key, maps);
  } else {
    cache[next] = {key, maps};
    next = (next + 1) % kEntries;
  }
  return 0;
}

template <class T, int R, int S>
int launch_hdiff_tma_rs(const T* inp, const T* coeff, T* out, int64_t nx, int64_t ny, int64_t nz,
                     int64_t sy, int64_t sz, int jt_request, int dry_runs, double* time,
                     cudaStream_t stream, bool* used, const T* inp_lower = nullptr, int64_t ny_lower = 0,
                     int64_t sz_lower = 0, const T* inp_upper = nullptr, int64_t ny_upper = 0,
                     int64_t sz_upper = 0, const StepFlags* step_flags = nullptr, unsigned step = 0) {
  constexpr int VEC = VecN<T>::value;
  constexpr int TW = tmacfg::kConsumers * VEC;
  *used = false;
  HdiffMaps maps;
  const int status = hdiff_maps<T, R>(inp, coeff, nx, ny, nz, sy, sz, inp_lower, ny_lower, sz_lower, inp_upper,
                                      ny_upper, sz_upper, maps);
  if (status == 2) return 0;
  if (status != 0) return status;
  const bool with_peers = inp_lower != nullptr || inp_upper != nullptr;
  const int64_t xtiles = ceil_div(nx, TW);
  constexpr int smem = tmacfg::smem_bytes(R, S);
  HdiffTiling tiling;
  int64_t ctas_total = 0;
  if (!hdiff_make_tiling(R, xtiles, ny, nz, jt_request, tiling, ctas_total))
    return fail("sb200_hdiff: domain too large for the launch grid");
  const dim3 grid{unsigned(ctas_total), 1, 1};
  static std::atomic<uint64_t> attr_local{0}, attr_peer{0};
  if (ensure_dynamic_smem(hdiff_tma_kernel<T, R, S, false>, smem, attr_local) ||
      ensure_dynamic_smem(hdiff_tma_kernel<T, R, S, true>, smem, attr_peer))
    return 1;
  *used = true;
  StepFlags flags{nullptr, {nullptr, nullptr}, {0, 0}, int(nz)};
  if (step_flags != nullptr) {
    flags = *step_flags;
    flags.levels = int(nz);
    // every edge CTA notifies once per consumer warp.  Towards its upper neighbour a slab has
    // one edge segment per (i tile, level), or two when its last segment is a single row -- what
    // this slab waits for from BELOW is therefore a property of the lower neighbour's tiling.
    const unsigned per_segment = unsigned(xtiles) * unsigned(tmacfg::kConsumers / 32);
    unsigned lower_segments = 1;
    if (inp_lower != nullptr) {
      HdiffTiling theirs;
      int64_t unused = 0;
      if (!hdiff_make_tiling(R, xtiles, ny_lower, nz, jt_request, theirs, unused))
        return fail("sb200_hdiff_step: the lower neighbour's slab is too large for the launch grid");
      if (theirs.segments > 1 && (ny_lower - 1) % theirs.jt == 0) lower_segments = 2;
    }
    flags.wait_value[0] = step * per_segment * lower_segments;
    flags.wait_value[1] = step * per_segment;
  }
  auto launch = [&] {
    if (with_peers)
      hdiff_tma_kernel<T, R, S, true><<<grid, tmacfg::kThreads, smem, stream>>>(
          maps.inp, maps.halo, maps.coeff, maps.peer, flags, out, int(nx), int(ny), tiling, sy, sz);
    else
      hdiff_tma_kernel<T, R, S, false><<<grid, tmacfg::kThreads, smem, stream>>>(
          maps.inp, maps.halo, maps.coeff, maps.peer, flags, out, int(nx), int(ny), tiling, sy, sz);
    count_launch();
  };
  return timed(launch, dry_runs, time, stream);
}

// rows per stage x stages of the TMA ring; SB200_HDIFF_CFG's third field selects the alternative
template <class T>
int launch_hdiff_tma(const T* inp, const T* coeff, T* out, int64_t nx, int64_t ny, int64_t nz, int64_t sy,
                     int64_t sz, int jt_request, int dry_runs, double* time,
                     cudaStream_t stream, bool* used, const T* inp_lower = nullptr, int64_t ny_lower = 0,
                     int64_t sz_lower = 0, const T* inp_upper = nullptr, int64_t ny_upper = 0,
                     int64_t sz_upper = 0, const StepFlags* step_flags = nullptr, unsigned step = 0) {
#define SB200_TMA_ARGS                                                                              \
  inp, coeff, out, nx, ny, nz, sy, sz, jt_request, dry_runs, time, stream, used, inp_lower, \
      ny_lower, sz_lower, inp_upper, ny_upper, sz_upper, step_flags, step
  // 4 rows x 4 stages is the best of the shapes measured (4x3, 8x2, 8x3, 2x6 are within 2.5 %:
  // profiles/hdiff_variants_r01.log); 4x3 is kept as the alternative for tuning runs
  if (hdiff_config().pipeline == 1) return launch_hdiff_tma_rs<T, 4, 3>(SB200_TMA_ARGS);
  return launch_hdiff_tma_rs<T, 4, 4>(SB200_TMA_ARGS);
#undef SB200_TMA_ARGS
}

template <class T>
int launch_hdiff(const T* inp, const T* coeff, T* out, int64_t nx, int64_t ny, int64_t nz,
                 int64_t sy, int64_t sz, int dry_runs, double* time, cudaStream_t stream) {
  constexpr int V = VecN<T>::value;
  constexpr int JT = 64;
  const HdiffConfig cfg = hdiff_config();
  // vector paths: 16-byte aligned interior origin and strides that keep every row aligned
  const bool vector_ok = aligned_to(inp, 16) && aligned_to(coeff, 16) && aligned_to(out, 16) &&
                         sy % V == 0 && sz % V == 0;
  // TMA path: faster than the generic march even when the 2 KB tile is only a quarter full
  // (profiles/size_sweep_r01.log)
  if (vector_ok && cfg.variant != 1 && (cfg.variant == 2 || nx * int64_t(sizeof(T)) >= 512)) {
    bool used = false;
    const int rc = launch_hdiff_tma<T>(inp, coeff, out, nx, ny, nz, sy, sz, cfg.jt, dry_runs,
                                       time, stream, &used);
    if (used || rc != 0) return rc;
  }
  const int vec = vector_ok ? V : 1;
  const int64_t nvec = ceil_div(nx, vec);
  int bx = 128;
  while (bx > 32 && bx / 2 >= nvec) bx /= 2;
  const dim3 block(bx, 1, 1);
  // long marches amortise the 4 warm-up rows; small domains need short ones to fill the GPU
  constexpr int JT_SHORT = 16;
  const bool short_march = ceil_div(nvec, bx) * ceil_div(ny, JT) * nz < 148 * 8;
  const int jt = short_march ? JT_SHORT : JT;
  const dim3 grid(unsigned(ceil_div(nvec, bx)), unsigned(ceil_div(ny, jt)), unsigned(nz));
  if (grid.y > 65535u || grid.z > 65535u) return fail("sb200_hdiff: domain too large for the launch grid");
  auto launch = [&] {
    if (vector_ok && short_march)
      hdiff_jmarch_kernel<T, V, JT_SHORT><<<grid, block, 0, stream>>>(inp, coeff, out, int(nx), int(ny), sy, sz);
    else if (vector_ok)
      hdiff_jmarch_kernel<T, V, JT><<<grid, block, 0, stream>>>(inp, coeff, out, int(nx), int(ny), sy, sz);
    else if (short_march)
      hdiff_jmarch_kernel<T, 1, JT_SHORT><<<grid, block, 0, stream>>>(inp, coeff, out, int(nx), int(ny), sy, sz);
    else
      hdiff_jmarch_kernel<T, 1, JT><<<grid, block, 0, stream>>>(inp, coeff, out, int(nx), int(ny), sy, sz);
    count_launch();
  };
  return timed(launch, dry_runs, time, stream);
}

}  // namespace
}  // namespace sb200

using namespace sb200;

namespace {
int hdiff_peer_entry(const char* who, int dtype, const void* inp, const void* coeff, void* out,
                     const void* inp_lower, int64_t ny_lower, int64_t sz_lower, const void* inp_upper,
                     int64_t ny_upper, int64_t sz_upper, const StepFlags* flags, unsigned step, int64_t nx,
                     int64_t ny, int64_t nz, int64_t sx, int64_t sy, int64_t sz, int dry_runs, double* time,
                     void* stream) {
  auto failed = [who](const char* what) {
    std::fprintf(stderr, "%s: %s\n", who, what);
    std::fflush(stderr);
    return 1;
  };
  if (nx <= 0 || ny <= 0 || nz <= 0) return failed("domain must be positive");
  if (sx != 1) return failed("only layout (2,1,0) is supported (unit stride along i)");
  if ((inp_lower != nullptr && ny_lower < 2) || (inp_upper != nullptr && ny_upper < 2))
    return failed("neighbouring slabs must hold at least two rows");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const HdiffConfig cfg = hdiff_config();
  bool used = false;
  int rc = 0;
  const bool aligned = aligned_to(inp, 16) && aligned_to(coeff, 16) && aligned_to(out, 16) &&
                       (inp_lower == nullptr || aligned_to(inp_lower, 16)) &&
                       (inp_upper == nullptr || aligned_to(inp_upper, 16));
  if (!aligned) return failed("fields must be 16-byte aligned");
  if (dtype == SB200_F64) {
    if (sy % 2 || sz % 2 || sz_lower % 2 || sz_upper % 2) return failed("strides must keep rows 16-byte aligned");
    rc = launch_hdiff_tma<double>(static_cast<const double*>(inp), static_cast<const double*>(coeff),
                                  static_cast<double*>(out), nx, ny, nz, sy, sz, cfg.jt, dry_runs, time, s,
                                  &used, static_cast<const double*>(inp_lower), ny_lower, sz_lower,
                                  static_cast<const double*>(inp_upper), ny_upper, sz_upper, flags, step);
  } else if (dtype == SB200_F32) {
    if (sy % 4 || sz % 4 || sz_lower % 4 || sz_upper % 4) return failed("strides must keep rows 16-byte aligned");
    rc = launch_hdiff_tma<float>(static_cast<const float*>(inp), static_cast<const float*>(coeff),
                                 static_cast<float*>(out), nx, ny, nz, sy, sz, cfg.jt, dry_runs, time, s,
                                 &used, static_cast<const float*>(inp_lower), ny_lower, sz_lower,
                                 static_cast<const float*>(inp_upper), ny_upper, sz_upper, flags, step);
  } else {
    return failed("unsupported dtype");
  }
  if (rc == 0 && !used) return failed("the TMA path is not available for these fields");
  return rc;
}
}  // namespace

extern "C" int sb200_hdiff_peer(int dtype, const void* inp, const void* coeff, void* out,
                                const void* inp_lower, int64_t ny_lower, int64_t sz_lower,
                                const void* inp_upper, int64_t ny_upper, int64_t sz_upper, int64_t nx,
                                int64_t ny, int64_t nz, int64_t sx, int64_t sy, int64_t sz, int dry_runs,
                                double* time, void* stream) {
  return hdiff_peer_entry("sb200_hdiff_peer", dtype, inp, coeff, out, inp_lower, ny_lower, sz_lower, inp_upper,
                          ny_upper, sz_upper, nullptr, 0, nx, ny, nz, sx, sy, sz, dry_runs, time, stream);
}

extern "C" int sb200_hdiff_step(int dtype, const void* inp, const void* coeff, void* out,
                                const void* inp_lower, int64_t ny_lower, int64_t sz_lower,
                                const void* inp_upper, int64_t ny_upper, int64_t sz_upper,
                                const uint32_t* arrived, uint32_t* notify_lower, uint32_t* notify_upper,
                                uint32_t step, int64_t nx, int64_t ny, int64_t nz, int64_t sx, int64_t sy,
                                int64_t sz, double* time, void* stream) {
  if (inp_lower == nullptr && inp_upper == nullptr) {
    // a slab without neighbours: nothing to order, the plain sweep
    return sb200_hdiff(dtype, inp, coeff, out, nx, ny, nz, sx, sy, sz, 0, time, stream);
  }
  if (arrived == nullptr || (inp_lower != nullptr && notify_lower == nullptr) ||
      (inp_upper != nullptr && notify_upper == nullptr))
    return fail("sb200_hdiff_step: a slab with neighbours needs its own counters and theirs");
  StepFlags flags;
  flags.arrived = arrived;
  flags.notify[0] = inp_lower != nullptr ? notify_lower : nullptr;
  flags.notify[1] = inp_upper != nullptr ? notify_upper : nullptr;
  flags.wait_value[0] = flags.wait_value[1] = 0;
  flags.levels = int(nz);
  return hdiff_peer_entry("sb200_hdiff_step", dtype, inp, coeff, out, inp_lower, ny_lower, sz_lower, inp_upper,
                          ny_upper, sz_upper, &flags, step, nx, ny, nz, sx, sy, sz, 0, time, stream);
}

extern "C" int sb200_hdiff_tiling(int dtype, int64_t nx, int64_t ny, int64_t nz, int* xtiles, int* segments,
                                  int* jt, int64_t* ctas) {
  if (nx <= 0 || ny <= 0 || nz <= 0) return fail("sb200_hdiff_tiling: domain must be positive");
  if (dtype != SB200_F64 && dtype != SB200_F32) return fail("sb200_hdiff_tiling: unsupported dtype");
  if (xtiles == nullptr || segments == nullptr || jt == nullptr || ctas == nullptr)
    return fail("sb200_hdiff_tiling: null output pointer");
  constexpr int R = 4;  // rows per stage of the default ring (launch_hdiff_tma)
  const int tile_width = tmacfg::kConsumers * (dtype == SB200_F64 ? VecN<double>::value : VecN<float>::value);
  HdiffTiling tiling;
  int64_t total = 0;
  if (!hdiff_make_tiling(R, ceil_div(nx, tile_width), ny, nz, hdiff_config().jt, tiling, total))
    return fail("sb200_hdiff_tiling: domain too large for the launch grid");
  *xtiles = tiling.xtiles;
  *segments = tiling.segments;
  *jt = tiling.jt;
  *ctas = total;
  return 0;
}

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
