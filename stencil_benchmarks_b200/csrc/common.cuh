// Shared host/device helpers of libsbench_b200 (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>

#include "../../include/sbench_b200.h"

namespace sb200 {

// --- error convention -------------------------------------------------------
// Same contract as the reference's CHECK macro (cuda_hip/templates/base.j2:48-55):
// message on stderr, sticky error drained, non-zero return.
inline int report_cuda_error(cudaError_t error, const char* what) {
  std::fprintf(stderr, "%s failed: %s\n", what, cudaGetErrorString(error));
  std::fflush(stderr);
  while (cudaGetLastError() != cudaSuccess) {
  }
  return 1;
}

inline int fail(const char* message) {
  std::fprintf(stderr, "%s\n", message);
  std::fflush(stderr);
  return 1;
}

#define SB200_CHECK(call)                                            \
  do {                                                               \
    cudaError_t sb200_error_ = (call);                               \
    if (sb200_error_ != cudaSuccess)                                 \
      return ::sb200::report_cuda_error(sb200_error_, #call);        \
  } while (0)

// --- launch accounting --------------------------------------------------------
extern std::atomic<uint64_t> g_launches;
inline void count_launch(uint64_t n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// device counter for a persistent kernel that hands out its work items dynamically (runtime.cu);
// the caller zeroes it on the launch's stream.  nullptr: no device / out of memory.
unsigned int* work_counter();

// --- timing wrapper -----------------------------------------------------------
// dry_runs untimed launches, then one launch bracketed by events on `stream`
// (base.j2:137-184).  time == nullptr: enqueue only.
struct EventPair {
  cudaEvent_t start = nullptr, stop = nullptr;
  ~EventPair() {
    if (start != nullptr) cudaEventDestroy(start);
    if (stop != nullptr) cudaEventDestroy(stop);
  }
};

template <class Launch>
int timed(Launch&& launch, int dry_runs, double* time, cudaStream_t stream) {
  for (int i = 0; i < dry_runs; ++i) launch();
  if (time == nullptr) {
    launch();
    SB200_CHECK(cudaGetLastError());
    return 0;
  }
  EventPair events;  // destroyed on every path out of here
  SB200_CHECK(cudaEventCreate(&events.start));
  SB200_CHECK(cudaEventCreate(&events.stop));
  SB200_CHECK(cudaEventRecord(events.start, stream));
  launch();
  SB200_CHECK(cudaGetLastError());
  SB200_CHECK(cudaEventRecord(events.stop, stream));
  SB200_CHECK(cudaEventSynchronize(events.stop));
  float ms = 0.f;
  SB200_CHECK(cudaEventElapsedTime(&ms, events.start, events.stop));
  *time = double(ms) / 1000.0;
  return 0;
}

// --- 128-bit vectors ------------------------------------------------------------
// VecN<T> = number of elements of T in one 128-bit access.
template <class T>
struct VecN;
template <>
struct VecN<double> {
  static constexpr int value = 2;
};
template <>
struct VecN<float> {
  static constexpr int value = 4;
};

enum class Cache { Default, Streaming, ReadOnly };

// Load N consecutive elements starting at p with the widest access that N
// allows: 16 B (LDG.E.128), 8 B or scalar.  p must be aligned to N*sizeof(T)
// (at most 16 B).
template <int N, Cache C = Cache::Default, class T>
__device__ __forceinline__ void load_vec(const T* __restrict__ p, T (&v)[N]) {
  constexpr int bytes = N * int(sizeof(T));
  if constexpr (bytes == 16 && sizeof(T) == 8) {
    const double2* q = reinterpret_cast<const double2*>(p);
    double2 t = C == Cache::Streaming ? __ldcs(q) : (C == Cache::ReadOnly ? __ldg(q) : *q);
    v[0] = t.x;
    v[1] = t.y;
  } else if constexpr (bytes == 16 && sizeof(T) == 4) {
    const float4* q = reinterpret_cast<const float4*>(p);
    float4 t = C == Cache::Streaming ? __ldcs(q) : (C == Cache::ReadOnly ? __ldg(q) : *q);
    v[0] = t.x;
    v[1] = t.y;
    v[2] = t.z;
    v[3] = t.w;
  } else if constexpr (bytes == 8 && sizeof(T) == 4) {
    const float2* q = reinterpret_cast<const float2*>(p);
    float2 t = C == Cache::Streaming ? __ldcs(q) : (C == Cache::ReadOnly ? __ldg(q) : *q);
    v[0] = t.x;
    v[1] = t.y;
  } else {
#pragma unroll
    for (int n = 0; n < N; ++n)
      v[n] = C == Cache::Streaming ? __ldcs(p + n) : (C == Cache::ReadOnly ? __ldg(p + n) : p[n]);
  }
}

template <int N, Cache C = Cache::Default, class T>
__device__ __forceinline__ void store_vec(T* __restrict__ p, const T (&v)[N]) {
  constexpr int bytes = N * int(sizeof(T));
  if constexpr (bytes == 16 && sizeof(T) == 8) {
    double2 t = make_double2(v[0], v[1]);
    if (C == Cache::Streaming)
      __stcs(reinterpret_cast<double2*>(p), t);
    else
      *reinterpret_cast<double2*>(p) = t;
  } else if constexpr (bytes == 16 && sizeof(T) == 4) {
    float4 t = make_float4(v[0], v[1], v[2], v[3]);
    if (C == Cache::Streaming)
      __stcs(reinterpret_cast<float4*>(p), t);
    else
      *reinterpret_cast<float4*>(p) = t;
  } else if constexpr (bytes == 8 && sizeof(T) == 4) {
    float2 t = make_float2(v[0], v[1]);
    if (C == Cache::Streaming)
      __stcs(reinterpret_cast<float2*>(p), t);
    else
      *reinterpret_cast<float2*>(p) = t;
  } else {
#pragma unroll
    for (int n = 0; n < N; ++n) {
      if (C == Cache::Streaming)
        __stcs(p + n, v[n]);
      else
        p[n] = v[n];
    }
  }
}

inline bool aligned_to(const void* p, size_t bytes) {
  return (reinterpret_cast<uintptr_t>(p) % bytes) == 0;
}

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

// SMs of the current device (every GPU of a box is the same part: cached once per process)
inline int sm_count() {
  static int count = [] {
    int device = 0, n = 148;
    if (cudaGetDevice(&device) == cudaSuccess)
      cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, device);
    return n;
  }();
  return count;
}

}  // namespace sb200
