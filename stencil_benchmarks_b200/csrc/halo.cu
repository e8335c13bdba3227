// Halo-face pack / unpack for the J-partitioned multi-GPU horizontal diffusion.
//
// New functionality (the reference is single-GPU, SURVEY.md §2 "Parallelism
// strategies"): rows [j0, j0+nrows) of every level are gathered into one
// contiguous message so a single ncclSend / ncclRecv moves a whole face over
// NVLink.  Rows are contiguous in i, so the copy is row-wise coalesced.
#include "common.cuh"

namespace sb200 {
namespace {

template <class T, bool PACK>
__global__ void __launch_bounds__(256)
    rows_kernel(T* __restrict__ field, T* __restrict__ buffer, int64_t row_len, int64_t nrows,
                int64_t nz, int64_t hx, int64_t sy, int64_t sz, int64_t j0) {
  const int64_t total = row_len * nrows * nz;
  for (int64_t e = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; e < total;
       e += int64_t(gridDim.x) * blockDim.x) {
    const int64_t i = e % row_len;
    const int64_t r = (e / row_len) % nrows;
    const int64_t k = e / (row_len * nrows);
    const int64_t f = k * sz + (j0 + r) * sy + (i - hx);
    if (PACK)
      buffer[e] = field[f];
    else
      field[f] = buffer[e];
  }
}

template <class T, bool PACK>
int launch_rows(T* field, T* buffer, int64_t nx, int64_t nz, int64_t hx, int64_t sy, int64_t sz,
                int64_t j0, int64_t nrows, cudaStream_t stream) {
  if (nx <= 0 || nz <= 0 || nrows <= 0 || hx < 0) return fail("sb200 pack/unpack: invalid extent");
  const int64_t total = (nx + 2 * hx) * nrows * nz;
  const unsigned grid = unsigned(std::min<int64_t>(ceil_div(total, 256), 148 * 16));
  rows_kernel<T, PACK><<<grid, 256, 0, stream>>>(field, buffer, nx + 2 * hx, nrows, nz, hx, sy, sz, j0);
  count_launch();
  SB200_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace
}  // namespace sb200

using namespace sb200;

extern "C" {

int sb200_pack_rows(int dtype, const void* field, void* buffer, int64_t nx, int64_t nz, int64_t hx,
                    int64_t sy, int64_t sz, int64_t j0, int64_t nrows, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == SB200_F64)
    return launch_rows<double, true>(static_cast<double*>(const_cast<void*>(field)),
                                     static_cast<double*>(buffer), nx, nz, hx, sy, sz, j0, nrows, s);
  if (dtype == SB200_F32)
    return launch_rows<float, true>(static_cast<float*>(const_cast<void*>(field)),
                                    static_cast<float*>(buffer), nx, nz, hx, sy, sz, j0, nrows, s);
  return fail("sb200_pack_rows: unsupported dtype");
}

int sb200_unpack_rows(int dtype, void* field, const void* buffer, int64_t nx, int64_t nz, int64_t hx,
                      int64_t sy, int64_t sz, int64_t j0, int64_t nrows, void* stream) {
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (dtype == SB200_F64)
    return launch_rows<double, false>(static_cast<double*>(field),
                                      static_cast<double*>(const_cast<void*>(buffer)), nx, nz, hx,
                                      sy, sz, j0, nrows, s);
  if (dtype == SB200_F32)
    return launch_rows<float, false>(static_cast<float*>(field),
                                     static_cast<float*>(const_cast<void*>(buffer)), nx, nz, hx, sy,
                                     sz, j0, nrows, s);
  return fail("sb200_unpack_rows: unsupported dtype");
}

}  // extern "C"
