/*
 * sbench_b200.h -- C ABI of the B200-native (sm_100a) backend for GridTools'
 * stencil_benchmarks ("sbench").
 *
 * This header is the drop-in boundary.  Every entry point below replaces one
 * FFI surface of the reference's GPU backend (paths relative to the reference
 * tree, `sb/` = stencil_benchmarks/, `sb/bc/` = sb/benchmarks_collection/):
 *
 *   - the JIT-compiled `extern "C" int kernel(double* time, T* f0, ..., T* fn)`
 *     of sb/bc/stencils/cuda_hip/templates/base.j2:68-193, called through
 *     ctypes from sb/bc/stencils/cuda_hip/mixin.py:162-175;
 *   - the `extern "C" int run()` of sb/bc/stream/cuda_hip.j2:179-288, called
 *     from sb/bc/stream/cuda_hip.py:121-144;
 *   - the libcudart calls the reference makes through ctypes for device
 *     memory (sb/bc/stencils/cuda_hip/api.py:39-104).
 *
 * Conventions kept from the reference (sb/tools/compilation.py:155-196):
 *   - every function returns int, 0 = success; on failure a one-line reason
 *     is written to stderr, the sticky CUDA error is cleared and 1 is returned;
 *   - field pointers are DEVICE pointers to the FIRST INTERIOR element
 *     (sb/tools/compilation.py:273-282 with offset = halo); the halo lies at
 *     negative offsets; strides are in ELEMENTS, in (i, j, k) order
 *     (sb/bc/stencils/base.py:119-121);
 *   - `time` (seconds) is measured with CUDA events around ONE sweep after
 *     `dry_runs` untimed sweeps (base.j2:137-184).  If `time` is NULL the
 *     sweep is only enqueued on `stream` (no events, no synchronisation), so
 *     a caller can time a batch itself or capture it in a CUDA graph.
 *
 * Differences from the reference ABI, on purpose:
 *   - geometry, dtype and dry_runs are run-time arguments of a pre-built
 *     library instead of literals baked into a JIT-compiled source;
 *   - an explicit `stream` (a cudaStream_t passed as void*, NULL = default
 *     stream) so halo exchange can overlap the interior sweep.
 *
 * No torch types, no C++ types: plain pointers and sizes only.
 */
#ifndef SBENCH_B200_H
#define SBENCH_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* dtype codes (NumPy names used by the `dtype` Parameter, sb/bc/stencils/base.py:53) */
#define SB200_F32 0
#define SB200_F64 1

/* basic stencil kinds (sb/bc/stencils/base.py:175-254) */
#define SB200_BASIC_EMPTY 0
#define SB200_BASIC_COPY 1
#define SB200_BASIC_ONESIDED_AVG 2
#define SB200_BASIC_SYMMETRIC_AVG 3
#define SB200_BASIC_LAPLACIAN 4

/* STREAM operations (sb/bc/stream/cuda_hip.j2:132-173) */
#define SB200_STREAM_COPY 0  /* c = a          */
#define SB200_STREAM_SCALE 1 /* b = s * c      */
#define SB200_STREAM_ADD 2   /* c = a + b      */
#define SB200_STREAM_TRIAD 3 /* a = b + s * c  */
#define SB200_STREAM_INIT 4  /* a = 1, b = 2, c = 0 */

/* vadv storage variants for the Thomas coefficients */
#define SB200_VADV_AUTO 0
#define SB200_VADV_GLOBAL 1 /* ccol/dcol round trip through HBM (reference "classic") */
#define SB200_VADV_ONCHIP 2 /* coefficients stay in shared memory / tensor memory    */

/* ------------------------------------------------------------------ */
/* Runtime (replaces sb/bc/stencils/cuda_hip/api.py:39-104)            */
/* ------------------------------------------------------------------ */

/* Library version as major*10000 + minor*100 + patch. Never fails. */
int sb200_version(void);

/* cudaGetDeviceCount; 0 devices is not an error here. */
int sb200_device_count(int* count);

/* cudaSetDevice / cudaGetDevice (base.j2:75-78 uses the current device). */
int sb200_set_device(int device);
int sb200_get_device(int* device);

/* Name, SM count, global memory bytes and L2 bytes of the current device. */
int sb200_device_info(char* name, int name_len, int* sm_count,
                      uint64_t* global_mem_bytes, uint64_t* l2_bytes);

/* cudaMalloc / cudaFree (api.py:54-78). */
int sb200_malloc(void** dptr, size_t nbytes);
int sb200_free(void* dptr);

/* Pinned host memory (cudaHostAlloc / cudaFreeHost) for the host fields, and
 * registration of memory that is already allocated (cudaHostRegister). */
int sb200_host_alloc(void** hptr, size_t nbytes);
int sb200_host_free(void* hptr);
int sb200_host_register(void* hptr, size_t nbytes);
int sb200_host_unregister(void* hptr);

/* cudaMemcpyAsync on `stream` followed by a stream synchronise when `sync`
 * is non-zero (api.py:80-93 + device_synchronize :95-96). */
int sb200_memcpy_h2d(void* dptr, const void* hptr, size_t nbytes, void* stream, int sync);
int sb200_memcpy_d2h(void* hptr, const void* dptr, size_t nbytes, void* stream, int sync);
int sb200_memcpy_d2d(void* dst, const void* src, size_t nbytes, void* stream, int sync);
int sb200_memset(void* dptr, int value, size_t nbytes, void* stream, int sync);

/* Strided (2-D) copies: `height` rows of `width_bytes` bytes, row pitches in bytes
 * (cudaMemcpy2DAsync).  Used to move j-slabs of a field: one "row" is the slab of one k level. */
int sb200_memcpy2d_h2d(void* dptr, size_t dpitch, const void* hptr, size_t hpitch,
                       size_t width_bytes, size_t height, void* stream);
int sb200_memcpy2d_d2h(void* hptr, size_t hpitch, const void* dptr, size_t dpitch,
                       size_t width_bytes, size_t height, void* stream);

/* Streams and events for pipelining copies against sweeps (cudaStreamCreateWithFlags
 * non-blocking, cudaEventCreate, cudaEventRecord, cudaStreamWaitEvent, cudaEventElapsedTime). */
int sb200_stream_create(void** stream);
int sb200_stream_destroy(void* stream);
int sb200_event_create(void** event);
int sb200_event_destroy(void* event);
int sb200_event_record(void* event, void* stream);
int sb200_stream_wait_event(void* stream, void* event);
int sb200_event_elapsed(void* start, void* stop, double* seconds);

/* cudaStreamSynchronize(stream) (stream == NULL: cudaDeviceSynchronize). */
int sb200_synchronize(void* stream);

/* Evict the L2 by overwriting a scratch buffer larger than the L2. */
int sb200_flush_l2(void* stream);

/* Number of sb200 kernels launched by this process so far (for bench.py's
 * `gpu_launches` claim). */
uint64_t sb200_launch_count(void);

/* ------------------------------------------------------------------ */
/* STREAM (replaces `run()` of sb/bc/stream/cuda_hip.j2:179-288)       */
/* ------------------------------------------------------------------ */

/* Launch shape of the STREAM kernels for the calls that follow (the reference bakes these into
 * its template: `block_size`, `unroll_factor`, `vector_size`, `streaming_loads/stores`,
 * sb/bc/stream/cuda_hip.py:44-66).  0 (streaming: negative) keeps the measured default of that
 * setting: 512 threads, 4 vectors per thread, 16-byte vectors, streaming (.cs) accesses.
 * vector_bytes is 16 or 32 (one LDG.E.128 or LDG.E.ENL2.256 per vector). */
int sb200_stream_configure(int block_size, int unroll_factor, int vector_bytes, int streaming);

/* Full McCalpin-style run on device arrays owned by the library: init
 * (a=1, b=2, c=0), `ntimes` rounds of copy/scale/add/triad each timed with
 * events, iteration 0 discarded, table printed to stdout in the reference's
 * format ("Copy: <MB/s> <avg> <min> <max>", cuda_hip.j2:262-275), closed-form
 * verification (cuda_hip.j2:290-345) when `verify` != 0.  Returns non-zero on
 * a CUDA error or failed verification. */
int sb200_stream_run(int dtype, uint64_t array_size, int ntimes, int verify);

/* One STREAM operation on caller-owned device arrays of `n` elements, scalar
 * 3 as in the reference unless given.  Arrays must be 16-byte aligned.
 * `time` as described at the top. */
int sb200_stream_op(int op, int dtype, void* a, void* b, void* c, uint64_t n,
                    double scalar, int dry_runs, double* time, void* stream);

/* ------------------------------------------------------------------ */
/* Stencils (replace `kernel()` of cuda_hip/templates/base.j2:68-193)  */
/* ------------------------------------------------------------------ */

/* Geometry arguments shared by all stencil entry points:
 *   nx, ny, nz   interior domain (sb/bc/stencils/base.py:46)
 *   sx, sy, sz   element strides of the padded field (base.py:119-121);
 *                sx must be 1 (layout (2,1,0): i is the unit-stride axis)
 * All fields of one stencil share the strides (they come from the same
 * allocator call with the same parameters, base.py:87-109). */

/* Basic stencils (bodies: sb/bc/stencils/cuda_hip/basic.py:101-132; oracle:
 * sb/bc/stencils/base.py:180-254).
 *   kind    SB200_BASIC_*
 *   axis    averaged axis 0/1/2 (one-sided: +1 neighbour; symmetric: +-1)
 *   along   bit mask of Laplacian axes: bit0 = x, bit1 = y, bit2 = z
 * Only interior points of `out` are written. */
int sb200_basic(int kind, int dtype, const void* inp, void* out,
                int64_t nx, int64_t ny, int64_t nz,
                int64_t sx, int64_t sy, int64_t sz,
                int axis, int along, int dry_runs, double* time, void* stream);

/* Horizontal diffusion (oracle: sb/bc/stencils/base.py:276-311; reference GPU
 * variants: sb/bc/stencils/cuda_hip/horizontal_diffusion.py:50-128).  One
 * fused kernel: Laplacian -> flx/fly -> limiter -> update.  Needs a halo of at
 * least 2 in i and j around `inp` (base.py:261-264).  `inp` and `coeff` are
 * never written; only interior points of `out` are written. */
int sb200_hdiff(int dtype, const void* inp, const void* coeff, void* out,
                int64_t nx, int64_t ny, int64_t nz,
                int64_t sx, int64_t sy, int64_t sz,
                int dry_runs, double* time, void* stream);

/* Work decomposition sb200_hdiff uses for a domain on its TMA path (host only, no
 * device needed; no counterpart in the reference, whose block sizes are template
 * literals: cuda_hip/horizontal_diffusion.py:41).  CTA b of `*ctas` sweeps i tile
 * b % *xtiles, rows [s * *jt, min((s+1) * *jt, ny)) with s = (b / *xtiles) % *segments,
 * on level b / (*xtiles * *segments). */
int sb200_hdiff_tiling(int dtype, int64_t nx, int64_t ny, int64_t nz,
                       int* xtiles, int* segments, int* jt, int64_t* ctas);

/* Vertical advection, u component (oracle: sb/bc/stencils/base.py:349-501 with
 * all_components=False; reference GPU variants:
 * sb/bc/stencils/cuda_hip/vertical_advection.py:60-73).  Per-(i,j)-column Thomas
 * solve, k-sequential.  `utensstage` is read and overwritten; `ccol`/`dcol`
 * are scratch (contents unspecified afterwards, as in the reference, base.py:485-501)
 * and may be NULL for the on-chip variant; `datacol` is never touched and may
 * be NULL.  Needs a halo of at least 1 in i around `wcon` (base.py:324-327).
 *   ishift, jshift   which wcon neighbour enters gav/gcv: (1,0) for u,
 *                    (0,1) for v, (0,0) for w (base.py:475-483)
 *   variant          SB200_VADV_*  */
int sb200_vadv(int dtype, const void* ustage, const void* upos, const void* utens,
               void* utensstage, const void* wcon, void* ccol, void* dcol, void* datacol,
               int64_t nx, int64_t ny, int64_t nz,
               int64_t sx, int64_t sy, int64_t sz,
               int ishift, int jshift, int variant,
               int dry_runs, double* time, void* stream);

/* Vertical advection of `ncomp` = 1...3 components in ONE sweep (all_components=True, oracle
 * base.py:475-483; reference counterpart: the merged u/v/w kernel
 * sb/bc/stencils/cuda_hip/templates/vertical_advection_localmemmerged.j2:392-462).  Component c
 * solves its own system from ustage[c] / upos[c] / utens[c] / utensstage[c] with the wcon
 * neighbour (ishift[c], jshift[c]); all components share `wcon`, which the on-chip variant reads
 * from HBM once per sweep (the components of a column batch are swept by different SMs at the same
 * time and meet in the L2): 13 reads + 3 writes per point for u, v, w instead of 3 x (5 + 1).
 * The tables hold `ncomp` entries; everything else as for sb200_vadv. */
int sb200_vadv_components(int dtype, int ncomp,
                          const void* const* ustage, const void* const* upos,
                          const void* const* utens, void* const* utensstage,
                          const int* ishift, const int* jshift,
                          const void* wcon, void* ccol, void* dcol,
                          int64_t nx, int64_t ny, int64_t nz,
                          int64_t sx, int64_t sy, int64_t sz,
                          int variant, int dry_runs, double* time, void* stream);

/* ------------------------------------------------------------------ */
/* Multi-GPU halo plumbing for the J-partitioned horizontal diffusion  */
/* ------------------------------------------------------------------ */

/* Peer memory inside one process that drives several GPUs: after cudaDeviceEnablePeerAccess(peer)
 * on `device`, kernels (and TMA) running on `device` can address `peer`'s allocations directly.
 * Already enabled is not an error. */
int sb200_enable_peer_access(int device, int peer);

/* Peer memory across processes (one process per GPU): cudaIpcGetMemHandle on the BASE of a
 * sb200_malloc allocation (64-byte handle, sent to the neighbour by any host channel),
 * cudaIpcOpenMemHandle / cudaIpcCloseMemHandle on the receiving side.  The mapped pointer
 * addresses the neighbour's HBM over NVLink. */
int sb200_ipc_get_handle(const void* dptr, void* handle64);
int sb200_ipc_open_handle(const void* handle64, void** dptr);
int sb200_ipc_close_handle(void* dptr);

/* Horizontal diffusion of one J slab with the halo exchange fused into the sweep: the two j-halo
 * rows on each side are read by TMA directly from the neighbouring GPUs' slabs (`inp_lower` /
 * `inp_upper`: pointers to the first interior element of the neighbour's inp field, mapped with
 * sb200_ipc_open_handle; NULL = no neighbour, the local halo is used).  `ny_*` / `sz_*` are the
 * neighbours' row counts and k strides (same i extent and j stride as the local slab).  One kernel
 * launch per sweep, no pack / send / receive / unpack.  Requires 16-byte aligned fields (TMA). */
int sb200_hdiff_peer(int dtype, const void* inp, const void* coeff, void* out,
                     const void* inp_lower, int64_t ny_lower, int64_t sz_lower,
                     const void* inp_upper, int64_t ny_upper, int64_t sz_upper,
                     int64_t nx, int64_t ny, int64_t nz,
                     int64_t sx, int64_t sy, int64_t sz,
                     int dry_runs, double* time, void* stream);

/* One sweep of a TIME LOOP over J slabs: like sb200_hdiff_peer, plus the ordering a loop needs
 * when `inp` and `out` swap roles every step (sweep m reads what sweep m-1 wrote, and overwrites
 * what sweep m-1 read).  The ordering is inside the kernel: the CTAs that sweep the slab's first /
 * last rows wait (ld.acquire.sys) until the neighbour's sweep m-1 has finished with the rows both
 * touch, and tell the neighbours when they are done themselves (red.release.sys into the
 * neighbours' memory over NVLink); every other CTA runs unhindered, the host is not involved.
 *   arrived        this slab's counters, 2 * nz zero-initialised uint32 in its own device memory:
 *                  [k] is pushed by the lower, [nz + k] by the upper neighbour (one counter per
 *                  level: a CTA waits for the neighbour CTAs of its own level only)
 *   notify_lower   address of the LOWER neighbour's arrived[nz] (the half its upper neighbour
 *                  pushes), mapped into this process
 *   notify_upper   address of the UPPER neighbour's arrived[0]
 *   step           m = 0, 1, 2, ...: index of this sweep in the loop; every slab calls with the
 *                  same sequence.  The fields, their neighbours and the domain must stay the same
 *                  for the whole loop apart from the inp <-> out swap.
 * Without neighbours this is sb200_hdiff.  No dry runs (they would advance the counters). */
int sb200_hdiff_step(int dtype, const void* inp, const void* coeff, void* out,
                     const void* inp_lower, int64_t ny_lower, int64_t sz_lower,
                     const void* inp_upper, int64_t ny_upper, int64_t sz_upper,
                     const uint32_t* arrived, uint32_t* notify_lower, uint32_t* notify_upper,
                     uint32_t step,
                     int64_t nx, int64_t ny, int64_t nz,
                     int64_t sx, int64_t sy, int64_t sz,
                     double* time, void* stream);

/* Pack `nrows` consecutive j-rows (all nz levels, i range [-hx, nx+hx)) of a
 * field into a contiguous buffer / unpack them again, so that one
 * ncclSend/ncclRecv moves a whole halo face.  `field` points to the first
 * interior element, `j0` is the first row relative to it (negative = halo).
 * Buffer layout: [k][row][i] with nx + 2*hx elements per row. */
int sb200_pack_rows(int dtype, const void* field, void* buffer,
                    int64_t nx, int64_t nz, int64_t hx,
                    int64_t sy, int64_t sz, int64_t j0, int64_t nrows, void* stream);
int sb200_unpack_rows(int dtype, void* field, const void* buffer,
                      int64_t nx, int64_t nz, int64_t hx,
                      int64_t sy, int64_t sz, int64_t j0, int64_t nrows, void* stream);

#ifdef __cplusplus
}
#endif

#endif /* SBENCH_B200_H */
