// nnk_shard.cu -- row-segment copy used by the sharded (multi-GPU) MLPG path: after the all-gather
// the result lives in shard order (bucket, rank, utterance); this kernel moves whole utterances
// (row segments) between two row-major matrices, e.g. back into the caller's utterance order.
// Pure data movement, HBM bound: every segment row is read once and written once with 16-byte
// accesses when the row pitch allows, 4-byte words otherwise.
#include <string.h>

#include "nnk_common.cuh"

namespace nnk {

struct SegCopyParams {
  const unsigned char* src;
  unsigned char* dst;
  int64_t row_bytes, src_pitch, dst_pitch;  // bytes
  const int64_t* src_row;
  const int64_t* dst_row;
  const int32_t* len;
  int rows_per_block;
};

template <typename W>
__global__ void __launch_bounds__(256) segment_copy_kernel(const SegCopyParams p) {
  const int seg = blockIdx.y;
  const int n = p.len[seg];
  const int r0 = blockIdx.x * p.rows_per_block;
  if (r0 >= n) return;
  const int r1 = min(n, r0 + p.rows_per_block);
  const int64_t wpr = p.row_bytes / (int64_t)sizeof(W);  // words per row
  const unsigned char* s = p.src + (p.src_row[seg] + r0) * p.src_pitch;
  unsigned char* d = p.dst + (p.dst_row[seg] + r0) * p.dst_pitch;
  const int64_t total = (int64_t)(r1 - r0) * wpr;
  for (int64_t i = threadIdx.x; i < total; i += blockDim.x) {
    const int64_t r = i / wpr, c = i - r * wpr;
    const W v = __ldcs(reinterpret_cast<const W*>(s + r * p.src_pitch) + c);
    __stcs(reinterpret_cast<W*>(d + r * p.dst_pitch) + c, v);
  }
}

}  // namespace nnk

using namespace nnk;

extern "C" int nnk_segment_copy(const void* src, void* dst, int32_t elem_bytes, int64_t cols, int64_t src_ld,
                                int64_t dst_ld, const int64_t* src_row, const int64_t* dst_row, const int32_t* len,
                                int32_t n_seg, int32_t max_len, void* stream) {
  NNK_REQUIRE(src && dst && src_row && dst_row && len, NNK_ERR_ARG, "NULL pointer");
  NNK_REQUIRE((elem_bytes == 4 || elem_bytes == 8) && cols > 0 && src_ld >= cols && dst_ld >= cols, NNK_ERR_ARG, "bad size");
  NNK_REQUIRE(n_seg >= 0 && max_len >= 0 && n_seg <= 65535, NNK_ERR_ARG, "bad segment count (max 65535 per call)");
  if (n_seg == 0 || max_len == 0) return NNK_OK;
  DeviceGuard guard(src);
  SegCopyParams p;
  p.src = (const unsigned char*)src; p.dst = (unsigned char*)dst;
  p.row_bytes = cols * elem_bytes; p.src_pitch = src_ld * elem_bytes; p.dst_pitch = dst_ld * elem_bytes;
  p.src_row = src_row; p.dst_row = dst_row; p.len = len;
  p.rows_per_block = 64;
  dim3 grid((unsigned)((max_len + p.rows_per_block - 1) / p.rows_per_block), (unsigned)n_seg);
  cudaStream_t st = (cudaStream_t)stream;
  const bool vec = (p.row_bytes % 16 == 0) && (p.src_pitch % 16 == 0) && (p.dst_pitch % 16 == 0) &&
                   ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0);
  if (vec) segment_copy_kernel<uint4><<<grid, 256, 0, st>>>(p);
  else segment_copy_kernel<uint32_t><<<grid, 256, 0, st>>>(p);
  count_launch();
  NNK_CUDA_CHECK(cudaGetLastError());
  return NNK_OK;
}

// ---- peer-memory transport of the sharded result (one process per GPU, NVLink / NVSwitch) -----------------
// The all-gather of the trajectories does not need SMs at all: every rank PUSHES its slot of the result
// buffer into the same slot of every peer's buffer with copy-engine DMA over NVLink (cudaMemcpyAsync
// between peer-mapped allocations).  Unlike an NCCL kernel, which has to wait for SM slots the solve
// kernel holds, the copies run while the next bucket is being solved.  The buffers are plain cudaMalloc
// allocations shared between the per-GPU processes through CUDA IPC handles.
extern "C" int nnk_peer_alloc(size_t bytes, void** ptr) {
  NNK_REQUIRE(ptr != nullptr && bytes > 0, NNK_ERR_ARG, "bad argument");
  NNK_CUDA_CHECK(cudaMalloc(ptr, bytes));
  NNK_CUDA_CHECK(cudaMemset(*ptr, 0, bytes));
  return NNK_OK;
}

extern "C" int nnk_peer_free(void* ptr) {
  if (ptr) NNK_CUDA_CHECK(cudaFree(ptr));
  return NNK_OK;
}

extern "C" int nnk_peer_export(const void* ptr, unsigned char* handle64) {
  NNK_REQUIRE(ptr && handle64, NNK_ERR_ARG, "NULL pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  cudaIpcMemHandle_t h;
  NNK_CUDA_CHECK(cudaIpcGetMemHandle(&h, const_cast<void*>(ptr)));
  memcpy(handle64, &h, 64);
  return NNK_OK;
}

extern "C" int nnk_peer_open(const unsigned char* handle64, void** ptr) {
  NNK_REQUIRE(ptr && handle64, NNK_ERR_ARG, "NULL pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, 64);
  NNK_CUDA_CHECK(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return NNK_OK;
}

extern "C" int nnk_peer_close(void* ptr) {
  if (ptr) NNK_CUDA_CHECK(cudaIpcCloseMemHandle(ptr));
  return NNK_OK;
}

extern "C" int nnk_peer_copy(void* dst_peer, const void* src_local, size_t bytes, void* stream) {
  NNK_REQUIRE(dst_peer && src_local, NNK_ERR_ARG, "NULL pointer");
  if (bytes == 0) return NNK_OK;
  DeviceGuard guard(src_local);
  NNK_CUDA_CHECK(cudaMemcpyAsync(dst_peer, src_local, bytes, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return NNK_OK;
}
