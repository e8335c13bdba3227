// Device runtime entry points of libsbench_b200: what the reference does from
// Python through ctypes on libcudart (cuda_hip/api.py:39-104).
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace sb200 {
std::atomic<uint64_t> g_launches{0};

// Work counters of the persistent kernels (dynamic scheduling): a small per-device pool, one
// counter per launch in turn, zeroed by the caller on the launch's stream right before the kernel
// (launches on different streams of one device may overlap; 64 of them in flight at once is more
// than any caller here does).
unsigned int* work_counter() {
  constexpr int kDevices = 64, kPool = 64;
  static std::mutex mutex;
  static unsigned int* pool[kDevices] = {};
  static unsigned next[kDevices] = {};
  int device = 0;
  if (cudaGetDevice(&device) != cudaSuccess || device < 0 || device >= kDevices) return nullptr;
  std::lock_guard<std::mutex> lock(mutex);
  if (pool[device] == nullptr &&
      cudaMalloc(reinterpret_cast<void**>(&pool[device]), kPool * sizeof(unsigned int)) != cudaSuccess) {
    while (cudaGetLastError() != cudaSuccess) {
    }
    return nullptr;
  }
  return pool[device] + (next[device]++ % kPool);
}

namespace {
__global__ void __launch_bounds__(256) flush_kernel(uint4* __restrict__ buffer, size_t n) {
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (; i < n; i += stride) buffer[i] = make_uint4(0u, 0u, 0u, 0u);
}

struct FlushBuffer {
  int device = -1;
  void* ptr = nullptr;
  size_t bytes = 0;
};
std::mutex g_flush_mutex;
FlushBuffer g_flush[16];
}  // namespace
}  // namespace sb200

using namespace sb200;

extern "C" {

int sb200_version(void) { return 100; }

int sb200_device_count(int* count) {
  if (count == nullptr) return fail("sb200_device_count: count is NULL");
  cudaError_t error = cudaGetDeviceCount(count);
  if (error == cudaErrorNoDevice || error == cudaErrorInsufficientDriver) {
    *count = 0;
    while (cudaGetLastError() != cudaSuccess) {
    }
    return 0;
  }
  SB200_CHECK(error);
  return 0;
}

int sb200_set_device(int device) {
  SB200_CHECK(cudaSetDevice(device));
  return 0;
}

int sb200_get_device(int* device) {
  if (device == nullptr) return fail("sb200_get_device: device is NULL");
  SB200_CHECK(cudaGetDevice(device));
  return 0;
}

int sb200_device_info(char* name, int name_len, int* sm_count, uint64_t* global_mem_bytes,
                      uint64_t* l2_bytes) {
  int device;
  cudaDeviceProp properties;
  SB200_CHECK(cudaGetDevice(&device));
  SB200_CHECK(cudaGetDeviceProperties(&properties, device));
  if (name != nullptr && name_len > 0) {
    std::strncpy(name, properties.name, size_t(name_len) - 1);
    name[name_len - 1] = '\0';
  }
  if (sm_count != nullptr) *sm_count = properties.multiProcessorCount;
  if (global_mem_bytes != nullptr) *global_mem_bytes = properties.totalGlobalMem;
  if (l2_bytes != nullptr) *l2_bytes = uint64_t(properties.l2CacheSize);
  return 0;
}

int sb200_malloc(void** dptr, size_t nbytes) {
  if (dptr == nullptr) return fail("sb200_malloc: dptr is NULL");
  SB200_CHECK(cudaMalloc(dptr, nbytes));
  return 0;
}

int sb200_free(void* dptr) {
  SB200_CHECK(cudaFree(dptr));
  return 0;
}

int sb200_host_alloc(void** hptr, size_t nbytes) {
  if (hptr == nullptr) return fail("sb200_host_alloc: hptr is NULL");
  // portable: one process may drive several devices (the partitioned benchmarks)
  SB200_CHECK(cudaHostAlloc(hptr, nbytes, cudaHostAllocPortable));
  return 0;
}

int sb200_host_free(void* hptr) {
  SB200_CHECK(cudaFreeHost(hptr));
  return 0;
}

int sb200_host_register(void* hptr, size_t nbytes) {
  SB200_CHECK(cudaHostRegister(hptr, nbytes, cudaHostRegisterDefault));
  return 0;
}

int sb200_host_unregister(void* hptr) {
  SB200_CHECK(cudaHostUnregister(hptr));
  return 0;
}

int sb200_memcpy_h2d(void* dptr, const void* hptr, size_t nbytes, void* stream, int sync) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SB200_CHECK(cudaMemcpyAsync(dptr, hptr, nbytes, cudaMemcpyHostToDevice, s));
  if (sync) SB200_CHECK(cudaStreamSynchronize(s));
  return 0;
}

int sb200_memcpy_d2h(void* hptr, const void* dptr, size_t nbytes, void* stream, int sync) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SB200_CHECK(cudaMemcpyAsync(hptr, dptr, nbytes, cudaMemcpyDeviceToHost, s));
  if (sync) SB200_CHECK(cudaStreamSynchronize(s));
  return 0;
}

int sb200_memcpy_d2d(void* dst, const void* src, size_t nbytes, void* stream, int sync) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SB200_CHECK(cudaMemcpyAsync(dst, src, nbytes, cudaMemcpyDeviceToDevice, s));
  if (sync) SB200_CHECK(cudaStreamSynchronize(s));
  return 0;
}

int sb200_memset(void* dptr, int value, size_t nbytes, void* stream, int sync) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  SB200_CHECK(cudaMemsetAsync(dptr, value, nbytes, s));
  if (sync) SB200_CHECK(cudaStreamSynchronize(s));
  return 0;
}

int sb200_memcpy2d_h2d(void* dptr, size_t dpitch, const void* hptr, size_t hpitch, size_t width_bytes,
                       size_t height, void* stream) {
  SB200_CHECK(cudaMemcpy2DAsync(dptr, dpitch, hptr, hpitch, width_bytes, height, cudaMemcpyHostToDevice,
                                static_cast<cudaStream_t>(stream)));
  return 0;
}

int sb200_memcpy2d_d2h(void* hptr, size_t hpitch, const void* dptr, size_t dpitch, size_t width_bytes,
                       size_t height, void* stream) {
  SB200_CHECK(cudaMemcpy2DAsync(hptr, hpitch, dptr, dpitch, width_bytes, height, cudaMemcpyDeviceToHost,
                                static_cast<cudaStream_t>(stream)));
  return 0;
}

int sb200_stream_create(void** stream) {
  if (stream == nullptr) return fail("sb200_stream_create: stream is NULL");
  cudaStream_t s;
  SB200_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
  *stream = s;
  return 0;
}

int sb200_stream_destroy(void* stream) {
  SB200_CHECK(cudaStreamDestroy(static_cast<cudaStream_t>(stream)));
  return 0;
}

int sb200_event_create(void** event) {
  if (event == nullptr) return fail("sb200_event_create: event is NULL");
  cudaEvent_t e;
  SB200_CHECK(cudaEventCreate(&e));
  *event = e;
  return 0;
}

int sb200_event_destroy(void* event) {
  SB200_CHECK(cudaEventDestroy(static_cast<cudaEvent_t>(event)));
  return 0;
}

int sb200_event_record(void* event, void* stream) {
  SB200_CHECK(cudaEventRecord(static_cast<cudaEvent_t>(event), static_cast<cudaStream_t>(stream)));
  return 0;
}

int sb200_stream_wait_event(void* stream, void* event) {
  SB200_CHECK(cudaStreamWaitEvent(static_cast<cudaStream_t>(stream), static_cast<cudaEvent_t>(event), 0));
  return 0;
}

int sb200_event_elapsed(void* start, void* stop, double* seconds) {
  if (seconds == nullptr) return fail("sb200_event_elapsed: seconds is NULL");
  float ms = 0.f;
  SB200_CHECK(cudaEventSynchronize(static_cast<cudaEvent_t>(stop)));
  SB200_CHECK(cudaEventElapsedTime(&ms, static_cast<cudaEvent_t>(start), static_cast<cudaEvent_t>(stop)));
  *seconds = double(ms) / 1000.0;
  return 0;
}

int sb200_enable_peer_access(int device, int peer) {
  int previous = 0, possible = 0;
  SB200_CHECK(cudaGetDevice(&previous));
  SB200_CHECK(cudaDeviceCanAccessPeer(&possible, device, peer));
  if (!possible) return fail("sb200_enable_peer_access: the devices cannot access each other's memory");
  SB200_CHECK(cudaSetDevice(device));
  const cudaError_t status = cudaDeviceEnablePeerAccess(peer, 0);
  if (status != cudaSuccess && status != cudaErrorPeerAccessAlreadyEnabled) {
    cudaSetDevice(previous);
    return report_cuda_error(status, "cudaDeviceEnablePeerAccess");
  }
  while (cudaGetLastError() != cudaSuccess) {
  }
  SB200_CHECK(cudaSetDevice(previous));
  return 0;
}

int sb200_ipc_get_handle(const void* dptr, void* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handles are exchanged as 64 bytes");
  if (handle64 == nullptr) return fail("sb200_ipc_get_handle: handle is NULL");
  cudaIpcMemHandle_t handle;
  SB200_CHECK(cudaIpcGetMemHandle(&handle, const_cast<void*>(dptr)));
  std::memcpy(handle64, &handle, sizeof(handle));
  return 0;
}

int sb200_ipc_open_handle(const void* handle64, void** dptr) {
  if (handle64 == nullptr || dptr == nullptr) return fail("sb200_ipc_open_handle: NULL argument");
  cudaIpcMemHandle_t handle;
  std::memcpy(&handle, handle64, sizeof(handle));
  SB200_CHECK(cudaIpcOpenMemHandle(dptr, handle, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}

int sb200_ipc_close_handle(void* dptr) {
  SB200_CHECK(cudaIpcCloseMemHandle(dptr));
  return 0;
}

int sb200_synchronize(void* stream) {
  if (stream == nullptr)
    SB200_CHECK(cudaDeviceSynchronize());
  else
    SB200_CHECK(cudaStreamSynchronize(static_cast<cudaStream_t>(stream)));
  return 0;
}

int sb200_flush_l2(void* stream) {
  int device;
  SB200_CHECK(cudaGetDevice(&device));
  if (device < 0 || device >= 16) return fail("sb200_flush_l2: unsupported device index");
  std::lock_guard<std::mutex> lock(g_flush_mutex);
  FlushBuffer& fb = g_flush[device];
  if (fb.ptr == nullptr) {
    cudaDeviceProp properties;
    SB200_CHECK(cudaGetDeviceProperties(&properties, device));
    // twice the L2 so that every set is overwritten
    fb.bytes = size_t(properties.l2CacheSize) * 2;
    if (fb.bytes < (size_t(64) << 20)) fb.bytes = size_t(64) << 20;
    SB200_CHECK(cudaMalloc(&fb.ptr, fb.bytes));
    fb.device = device;
  }
  flush_kernel<<<148 * 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<uint4*>(fb.ptr), fb.bytes / sizeof(uint4));
  SB200_CHECK(cudaGetLastError());
  return 0;
}

uint64_t sb200_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

}  // extern "C"
