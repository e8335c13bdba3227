// STREAM copy / scale / add / triad for sm_100a.
//
// Replaces the generated kernels + `run()` of the reference
// (stencil_benchmarks/benchmarks_collection/stream/cuda_hip.j2:132-173, :179-288)
// and its closed-form verification (:290-345).
//
// Design (HBM roofline kernel, no reuse): every thread moves UNROLL vectors of
// VB bytes (16 B = LDG.E.128, 32 B = LDG.E.ENL2.256), all loads issued before
// the first store so that UNROLL requests per input array are in flight per
// thread; consecutive threads touch consecutive vectors (full 128-byte lines
// per warp request); streaming cache policy because nothing is reused.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <limits>
#include <vector>

#include "common.cuh"

namespace sb200 {
namespace {

template <class T, int VB>
struct Chunk {
  static constexpr int N = VB / int(sizeof(T));
  T v[N];
};

// ---- VB-byte global accesses ---------------------------------------------------
template <class T, int VB, bool STREAMING>
__device__ __forceinline__ Chunk<T, VB> load_chunk(const T* p) {
  Chunk<T, VB> c;
  if constexpr (VB == 32) {
    unsigned long long w0, w1, w2, w3;
    if constexpr (STREAMING)
      asm volatile("ld.global.cs.v4.b64 {%0,%1,%2,%3}, [%4];"
                   : "=l"(w0), "=l"(w1), "=l"(w2), "=l"(w3)
                   : "l"(p));
    else
      asm volatile("ld.global.v4.b64 {%0,%1,%2,%3}, [%4];"
                   : "=l"(w0), "=l"(w1), "=l"(w2), "=l"(w3)
                   : "l"(p));
    unsigned long long w[4] = {w0, w1, w2, w3};
    if constexpr (sizeof(T) == 8) {
#pragma unroll
      for (int n = 0; n < 4; ++n) c.v[n] = __longlong_as_double((long long)w[n]);
    } else {
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        c.v[2 * n] = __uint_as_float((unsigned)(w[n] & 0xffffffffull));
        c.v[2 * n + 1] = __uint_as_float((unsigned)(w[n] >> 32));
      }
    }
  } else {
    load_vec<Chunk<T, VB>::N, STREAMING ? Cache::Streaming : Cache::Default>(p, c.v);
  }
  return c;
}

template <class T, int VB, bool STREAMING>
__device__ __forceinline__ void store_chunk(T* p, const Chunk<T, VB>& c) {
  if constexpr (VB == 32) {
    unsigned long long w[4];
    if constexpr (sizeof(T) == 8) {
#pragma unroll
      for (int n = 0; n < 4; ++n) w[n] = (unsigned long long)__double_as_longlong(c.v[n]);
    } else {
#pragma unroll
      for (int n = 0; n < 4; ++n)
        w[n] = (unsigned long long)__float_as_uint(c.v[2 * n]) |
               ((unsigned long long)__float_as_uint(c.v[2 * n + 1]) << 32);
    }
    if constexpr (STREAMING)
      asm volatile("st.global.cs.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(w[0]), "l"(w[1]),
                   "l"(w[2]), "l"(w[3])
                   : "memory");
    else
      asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(w[0]), "l"(w[1]),
                   "l"(w[2]), "l"(w[3])
                   : "memory");
  } else {
    store_vec<Chunk<T, VB>::N, STREAMING ? Cache::Streaming : Cache::Default>(p, c.v);
  }
}

// ---- the kernel ---------------------------------------------------------------
// dst = f(s1, s2):  COPY dst=s1; SCALE dst=q*s1; ADD dst=s1+s2; TRIAD dst=s1+q*s2
template <class T, int OP>
__device__ __forceinline__ T apply(T x, T y, T q) {
  if constexpr (OP == SB200_STREAM_COPY) return x;
  if constexpr (OP == SB200_STREAM_SCALE) return q * x;
  if constexpr (OP == SB200_STREAM_ADD) return x + y;
  return x + q * y;
}

template <class T, int OP, int VB, int UNROLL, bool STREAMING>
__global__ void __launch_bounds__(1024)
    stream_kernel(T* __restrict__ dst, const T* __restrict__ s1, const T* __restrict__ s2, T q,
                  size_t nchunks, size_t n) {
  constexpr int N = VB / int(sizeof(T));
  constexpr bool TWO = OP == SB200_STREAM_ADD || OP == SB200_STREAM_TRIAD;
  const size_t base = size_t(blockIdx.x) * (size_t(blockDim.x) * UNROLL) + threadIdx.x;

  if (base + size_t(UNROLL - 1) * blockDim.x < nchunks) {
    Chunk<T, VB> x[UNROLL], y[UNROLL];
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const size_t e = (base + size_t(u) * blockDim.x) * N;
      x[u] = load_chunk<T, VB, STREAMING>(s1 + e);
      if constexpr (TWO) y[u] = load_chunk<T, VB, STREAMING>(s2 + e);
    }
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const size_t e = (base + size_t(u) * blockDim.x) * N;
      Chunk<T, VB> r;
#pragma unroll
      for (int m = 0; m < N; ++m) r.v[m] = apply<T, OP>(x[u].v[m], TWO ? y[u].v[m] : T(0), q);
      store_chunk<T, VB, STREAMING>(dst + e, r);
    }
  } else {
    // last block: bounds-checked chunks, then the scalar remainder (n % N)
#pragma unroll
    for (int u = 0; u < UNROLL; ++u) {
      const size_t c = base + size_t(u) * blockDim.x;
      if (c < nchunks) {
        const size_t e = c * N;
        Chunk<T, VB> x = load_chunk<T, VB, STREAMING>(s1 + e), y, r;
        if constexpr (TWO) y = load_chunk<T, VB, STREAMING>(s2 + e);
#pragma unroll
        for (int m = 0; m < N; ++m) r.v[m] = apply<T, OP>(x.v[m], TWO ? y.v[m] : T(0), q);
        store_chunk<T, VB, STREAMING>(dst + e, r);
      }
    }
  }
  if (blockIdx.x == gridDim.x - 1) {
    const size_t e = nchunks * N + threadIdx.x;
    if (e < n) dst[e] = apply<T, OP>(s1[e], TWO ? s2[e] : T(0), q);
  }
}

template <class T, int VB, int UNROLL, bool STREAMING>
__global__ void __launch_bounds__(1024)
    init_kernel(T* __restrict__ a, T* __restrict__ b, T* __restrict__ c, size_t nchunks, size_t n) {
  constexpr int N = VB / int(sizeof(T));
  const size_t base = size_t(blockIdx.x) * (size_t(blockDim.x) * UNROLL) + threadIdx.x;
  Chunk<T, VB> one, two, zero;
#pragma unroll
  for (int m = 0; m < N; ++m) {
    one.v[m] = T(1);
    two.v[m] = T(2);
    zero.v[m] = T(0);
  }
#pragma unroll
  for (int u = 0; u < UNROLL; ++u) {
    const size_t ch = base + size_t(u) * blockDim.x;
    if (ch < nchunks) {
      store_chunk<T, VB, STREAMING>(a + ch * N, one);
      store_chunk<T, VB, STREAMING>(b + ch * N, two);
      store_chunk<T, VB, STREAMING>(c + ch * N, zero);
    }
  }
  if (blockIdx.x == gridDim.x - 1) {
    const size_t e = nchunks * N + threadIdx.x;
    if (e < n) {
      a[e] = T(1);
      b[e] = T(2);
      c[e] = T(0);
    }
  }
}

// ---- launch configuration -------------------------------------------------------
// Defaults chosen from the sweep in profiles/ (see DESIGN.md); the SB200_STREAM_CFG
// environment variable "block,unroll,vector_bytes,streaming" overrides them for
// tuning runs.
struct StreamConfig {
  int block = 512;
  int unroll = 4;
  int vector_bytes = 16;
  int streaming = 1;
};

// sb200_stream_configure: explicit settings of the calling benchmark (0 = keep the default)
StreamConfig g_stream_override{0, 0, 0, -1};

StreamConfig stream_config() {
  StreamConfig cfg;
  if (g_stream_override.block > 0) cfg.block = g_stream_override.block;
  if (g_stream_override.unroll > 0) cfg.unroll = g_stream_override.unroll;
  if (g_stream_override.vector_bytes > 0) cfg.vector_bytes = g_stream_override.vector_bytes;
  if (g_stream_override.streaming >= 0) cfg.streaming = g_stream_override.streaming;
  if (const char* env = std::getenv("SB200_STREAM_CFG")) {
    int b, u, v, s;
    if (std::sscanf(env, "%d,%d,%d,%d", &b, &u, &v, &s) == 4) {
      cfg.block = b;
      cfg.unroll = u;
      cfg.vector_bytes = v;
      cfg.streaming = s;
    }
  }
  return cfg;
}

template <class T, int OP, int VB, int UNROLL, bool STREAMING>
void launch_op(T* dst, const T* s1, const T* s2, T q, size_t n, int block, cudaStream_t stream) {
  constexpr int N = VB / int(sizeof(T));
  const size_t nchunks = n / N;
  const size_t per_block = size_t(block) * UNROLL;
  const size_t grid = std::max<size_t>(1, (nchunks + per_block - 1) / per_block);
  stream_kernel<T, OP, VB, UNROLL, STREAMING>
      <<<unsigned(grid), block, 0, stream>>>(dst, s1, s2, q, nchunks, n);
  count_launch();
}

template <class T, int OP, int VB, int UNROLL>
void launch_op_s(T* dst, const T* s1, const T* s2, T q, size_t n, const StreamConfig& cfg,
                 cudaStream_t stream) {
  if (cfg.streaming)
    launch_op<T, OP, VB, UNROLL, true>(dst, s1, s2, q, n, cfg.block, stream);
  else
    launch_op<T, OP, VB, UNROLL, false>(dst, s1, s2, q, n, cfg.block, stream);
}

template <class T, int OP, int VB>
int launch_op_u(T* dst, const T* s1, const T* s2, T q, size_t n, const StreamConfig& cfg,
                cudaStream_t stream) {
  switch (cfg.unroll) {
    case 1: launch_op_s<T, OP, VB, 1>(dst, s1, s2, q, n, cfg, stream); return 0;
    case 2: launch_op_s<T, OP, VB, 2>(dst, s1, s2, q, n, cfg, stream); return 0;
    case 4: launch_op_s<T, OP, VB, 4>(dst, s1, s2, q, n, cfg, stream); return 0;
    case 8: launch_op_s<T, OP, VB, 8>(dst, s1, s2, q, n, cfg, stream); return 0;
  }
  return fail("sb200 stream: unroll must be 1, 2, 4 or 8");
}

template <class T, int OP>
int launch_op_v(T* dst, const T* s1, const T* s2, T q, size_t n, const StreamConfig& cfg,
                cudaStream_t stream) {
  if (cfg.block < 32 || cfg.block > 1024 || cfg.block % 32)
    return fail("sb200 stream: block must be a multiple of 32 in [32, 1024]");
  const bool ok32 = aligned_to(dst, 32) && aligned_to(s1, 32) && (s2 == nullptr || aligned_to(s2, 32));
  if (cfg.vector_bytes == 32 && ok32) return launch_op_u<T, OP, 32>(dst, s1, s2, q, n, cfg, stream);
  if (cfg.vector_bytes != 16 && cfg.vector_bytes != 32)
    return fail("sb200 stream: vector_bytes must be 16 or 32");
  return launch_op_u<T, OP, 16>(dst, s1, s2, q, n, cfg, stream);
}

// a, b, c as in McCalpin's STREAM; which array plays dst/s1/s2 follows
// cuda_hip.j2:132-173.
template <class T>
int stream_op(int op, T* a, T* b, T* c, size_t n, T q, cudaStream_t stream) {
  const StreamConfig cfg = stream_config();
  switch (op) {
    case SB200_STREAM_COPY: return launch_op_v<T, SB200_STREAM_COPY>(c, a, (const T*)nullptr, q, n, cfg, stream);
    case SB200_STREAM_SCALE: return launch_op_v<T, SB200_STREAM_SCALE>(b, c, (const T*)nullptr, q, n, cfg, stream);
    case SB200_STREAM_ADD: return launch_op_v<T, SB200_STREAM_ADD>(c, a, b, q, n, cfg, stream);
    case SB200_STREAM_TRIAD: return launch_op_v<T, SB200_STREAM_TRIAD>(a, b, c, q, n, cfg, stream);
    case SB200_STREAM_INIT: {
      const size_t nchunks = n / (16 / sizeof(T));
      const size_t per_block = size_t(512) * 4;
      const size_t grid = std::max<size_t>(1, (nchunks + per_block - 1) / per_block);
      init_kernel<T, 16, 4, false><<<unsigned(grid), 512, 0, stream>>>(a, b, c, nchunks, n);
      count_launch();
      return 0;
    }
  }
  return fail("sb200 stream: unknown operation");
}

template <class T>
int stream_op_timed(int op, void* a, void* b, void* c, uint64_t n, double scalar, int dry_runs,
                    double* time, cudaStream_t stream) {
  if (!aligned_to(a, 16) || !aligned_to(b, 16) || !aligned_to(c, 16))
    return fail("sb200_stream_op: arrays must be 16-byte aligned");
  int status = 0;
  auto launch = [&] {
    status |= stream_op<T>(op, static_cast<T*>(a), static_cast<T*>(b), static_cast<T*>(c),
                           size_t(n), T(scalar), stream);
  };
  const int rc = timed(launch, dry_runs, time, stream);
  return rc | status;
}

// ---- verification (cuda_hip.j2:290-345) --------------------------------------------
template <class T>
__global__ void __launch_bounds__(256)
    abs_error_kernel(const T* __restrict__ x, size_t n, T expected, double* __restrict__ sum,
                     unsigned long long* __restrict__ bad, double epsilon) {
  double local = 0;
  unsigned long long local_bad = 0;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += size_t(gridDim.x) * blockDim.x) {
    const T v = x[i];
    local += fabs(double(v) - double(expected));
    if (fabs(double(v) / double(expected) - 1) > epsilon) ++local_bad;
  }
  for (int o = 16; o > 0; o >>= 1) {
    local += __shfl_xor_sync(0xffffffffu, local, o);
    local_bad += __shfl_xor_sync(0xffffffffu, local_bad, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicAdd(sum, local);
    atomicAdd(bad, local_bad);
  }
}

template <class T>
int check_array(const char* name, const T* x, size_t n, T expected, bool* ok) {
  if (std::isinf(double(expected))) {
    std::fprintf(stderr, "expected value is infinite, ntimes too large for verication\n");
    *ok = false;
    return 0;
  }
  const double epsilon = sizeof(T) == 4 ? 1e-6 : 1e-13;
  double* dsum = nullptr;
  unsigned long long* dbad = nullptr;
  struct Scratch {  // freed on every path out of here
    double*& sum;
    unsigned long long*& bad;
    ~Scratch() {
      cudaFree(sum);
      cudaFree(bad);
    }
  } scratch{dsum, dbad};
  SB200_CHECK(cudaMalloc(&dsum, sizeof(double)));
  SB200_CHECK(cudaMalloc(&dbad, sizeof(unsigned long long)));
  SB200_CHECK(cudaMemset(dsum, 0, sizeof(double)));
  SB200_CHECK(cudaMemset(dbad, 0, sizeof(unsigned long long)));
  abs_error_kernel<T><<<148 * 8, 256>>>(x, n, expected, dsum, dbad, epsilon);
  count_launch();
  SB200_CHECK(cudaGetLastError());
  double sum;
  unsigned long long bad;
  SB200_CHECK(cudaMemcpy(&sum, dsum, sizeof(double), cudaMemcpyDeviceToHost));
  SB200_CHECK(cudaMemcpy(&bad, dbad, sizeof(bad), cudaMemcpyDeviceToHost));
  const double avg_err = sum / double(n);
  *ok = true;
  if (std::fabs(avg_err / double(expected)) > epsilon) {
    std::fprintf(stderr,
                 "failed validation on array %s[]\nexpected value: %g avg. abs. error: %g "
                 "avg. rel. error: %g\nfor array %s[], %llu errors were found\n",
                 name, double(expected), avg_err, std::fabs(avg_err / double(expected)), name, bad);
    *ok = bad == 0;
  }
  return 0;
}

template <class T>
int stream_run(uint64_t n, int ntimes, int verify) {
  if (n == 0) return fail("sb200_stream_run: array_size must be positive");
  if (ntimes < 2) return fail("sb200_stream_run: ntimes must be at least 2");
  T *a, *b, *c;
  SB200_CHECK(cudaMalloc(&a, sizeof(T) * n));
  SB200_CHECK(cudaMalloc(&b, sizeof(T) * n));
  SB200_CHECK(cudaMalloc(&c, sizeof(T) * n));
  struct Guard {
    T *a, *b, *c;
    ~Guard() {
      cudaFree(a);
      cudaFree(b);
      cudaFree(c);
    }
  } guard{a, b, c};

  if (stream_op<T>(SB200_STREAM_INIT, a, b, c, n, T(0), nullptr)) return 1;
  SB200_CHECK(cudaGetLastError());
  SB200_CHECK(cudaDeviceSynchronize());

  const T scalar = 3;
  EventPair events;
  SB200_CHECK(cudaEventCreate(&events.start));
  SB200_CHECK(cudaEventCreate(&events.stop));
  const cudaEvent_t start = events.start, stop = events.stop;
  std::vector<double> times[4];
  for (auto& t : times) t.resize(size_t(ntimes));
  const int ops[4] = {SB200_STREAM_COPY, SB200_STREAM_SCALE, SB200_STREAM_ADD, SB200_STREAM_TRIAD};
  for (int k = 0; k < ntimes; ++k) {
    for (int j = 0; j < 4; ++j) {
      float ms;
      SB200_CHECK(cudaEventRecord(start));
      if (stream_op<T>(ops[j], a, b, c, n, scalar, nullptr)) return 1;
      SB200_CHECK(cudaGetLastError());
      SB200_CHECK(cudaEventRecord(stop));
      SB200_CHECK(cudaEventSynchronize(stop));
      SB200_CHECK(cudaEventElapsedTime(&ms, start, stop));
      times[j][size_t(k)] = double(ms) / 1000.0;
    }
  }

  const char* label[4] = {"Copy:      ", "Scale:     ", "Add:       ", "Triad:     "};
  const double bytes[4] = {2.0 * sizeof(T) * n, 2.0 * sizeof(T) * n, 3.0 * sizeof(T) * n,
                           3.0 * sizeof(T) * n};
  std::printf("Function    Best Rate MB/s  Avg time     Min time     Max time\n");
  for (int j = 0; j < 4; ++j) {
    double avg = 0, mn = std::numeric_limits<double>::max(), mx = 0;
    for (int k = 1; k < ntimes; ++k) {  // iteration 0 is discarded (cuda_hip.j2:250)
      avg += times[j][size_t(k)];
      mn = std::min(mn, times[j][size_t(k)]);
      mx = std::max(mx, times[j][size_t(k)]);
    }
    avg /= double(ntimes - 1);
    std::printf("%s%12.1f  %11.6f  %11.6f  %11.6f\n", label[j], 1.0e-6 * bytes[j] / mn, avg, mn, mx);
  }
  std::fflush(stdout);

  bool verifies = true;
  if (verify) {
    T aj = 1, bj = 2, cj = 0;
    for (int k = 0; k < ntimes; ++k) {
      cj = aj;
      bj = scalar * cj;
      cj = aj + bj;
      aj = bj + scalar * cj;
    }
    bool ok;
    if (check_array<T>("a", a, n, aj, &ok)) return 1;
    verifies = verifies && ok;
    if (verifies) {
      if (check_array<T>("b", b, n, bj, &ok)) return 1;
      verifies = verifies && ok;
    }
    if (verifies) {
      if (check_array<T>("c", c, n, cj, &ok)) return 1;
      verifies = verifies && ok;
    }
    std::fflush(stderr);
  }
  return verifies ? 0 : 1;
}

}  // namespace
}  // namespace sb200

using namespace sb200;

extern "C" {

int sb200_stream_configure(int block_size, int unroll_factor, int vector_bytes, int streaming) {
  if (block_size != 0 && (block_size < 32 || block_size > 1024 || block_size % 32))
    return fail("sb200_stream_configure: block size must be a multiple of 32 in [32, 1024]");
  if (unroll_factor != 0 && unroll_factor != 1 && unroll_factor != 2 && unroll_factor != 4 && unroll_factor != 8)
    return fail("sb200_stream_configure: unroll factor must be 1, 2, 4 or 8");
  if (vector_bytes != 0 && vector_bytes != 16 && vector_bytes != 32)
    return fail("sb200_stream_configure: vectors are 16 or 32 bytes wide");
  g_stream_override = StreamConfig{block_size, unroll_factor, vector_bytes, streaming < 0 ? -1 : (streaming != 0)};
  return 0;
}

int sb200_stream_run(int dtype, uint64_t array_size, int ntimes, int verify) {
  if (dtype == SB200_F64) return stream_run<double>(array_size, ntimes, verify);
  if (dtype == SB200_F32) return stream_run<float>(array_size, ntimes, verify);
  return fail("sb200_stream_run: unsupported dtype");
}

int sb200_stream_op(int op, int dtype, void* a, void* b, void* c, uint64_t n, double scalar,
                    int dry_runs, double* time, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == SB200_F64) return stream_op_timed<double>(op, a, b, c, n, scalar, dry_runs, time, s);
  if (dtype == SB200_F32) return stream_op_timed<float>(op, a, b, c, n, scalar, dry_runs, time, s);
  return fail("sb200_stream_op: unsupported dtype");
}

}  // extern "C"
