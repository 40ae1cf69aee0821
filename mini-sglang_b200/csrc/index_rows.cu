// Row gather: out[t] = weights[indices[t]] (optionally masked to a vocabulary shard).
//
// Replaces `indexing(weights, indices, output=, vocab_range=)` (python/minisgl/kernel/index.py:32-53
// -> kernel/csrc/jit/index.cu:34-96), the embedding lookup of VocabParallelEmbedding.forward
// (python/minisgl/layers/embedding.py:31-41), and -- same access pattern -- the last-token gather
// `x[indices].contiguous()` in front of the LM head (layers/embedding.py:92-94).
//
// Pure byte movement, HBM-bound: algorithmic bytes = 2 * n * row_bytes (+ n index reads).  The
// reference gives a warp (or 2 / 4 warps) one row; here the work item is a 16-byte vector and the
// grid is sized from the byte count, so short rows (a 256 B kv-head row) and long rows (16 KB hidden
// states) both spread over all SMs: each thread keeps four independent 16-byte loads in flight
// (the table is read through the non-coherent path, the output is written with streaming stores).
#include "b200attn.h"
#include "common.cuh"

namespace b200 {

template <typename IdxT, bool kMasked>
__global__ void __launch_bounds__(256)
index_rows_kernel(const uint8_t* __restrict__ weights, int64_t w_stride, const IdxT* __restrict__ indices,
                  int64_t n, int vecs_per_row, uint8_t* __restrict__ out, int64_t o_stride,
                  uint64_t vocab_start, uint64_t vocab_len) {
  pdl_wait();
  pdl_launch_dependents();
  const int64_t total = n * (int64_t)vecs_per_row;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  constexpr int kUnroll = 4;
  for (int64_t base = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; base < total; base += stride * kUnroll) {
    Vec8 val[kUnroll];
    int64_t dst[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int64_t i = base + u * stride;
      dst[u] = -1;
      val[u] = Vec8{{0u, 0u, 0u, 0u}};
      if (i < total) {
        const int64_t t = i / vecs_per_row;
        const int c = (int)(i - t * vecs_per_row);
        dst[u] = t * o_stride + (int64_t)c * 16;
        // unsigned compare = the reference's `pos < length` on size_t (negative positions wrap)
        const uint64_t pos = (uint64_t)(int64_t)indices[t] - vocab_start;
        if (!kMasked || pos < vocab_len) val[u] = ldg_stream(weights + (int64_t)pos * w_stride + (int64_t)c * 16);
      }
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u)
      if (dst[u] >= 0) stg_stream(out + dst[u], val[u]);
  }
}

template <typename IdxT>
static int launch_index(const uint8_t* w, int64_t w_stride, const void* idx, int64_t n, int vecs,
                        uint8_t* out, int64_t o_stride, int64_t start, int64_t len, cudaStream_t st) {
  const int64_t total = n * (int64_t)vecs;
  // enough CTAs for four vectors per thread, at most eight resident CTAs per SM
  int64_t blocks = ceil_div<int64_t>(total, 256 * 4);
  const int64_t cap = (int64_t)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  if (len >= 0) {
    B200_CHECK_CUDA(launch_pdl(index_rows_kernel<IdxT, true>, dim3((unsigned)blocks), dim3(256), 0, st, w,
                               w_stride, (const IdxT*)idx, n, vecs, out, o_stride, (uint64_t)start,
                               (uint64_t)len));
  } else {
    B200_CHECK_CUDA(launch_pdl(index_rows_kernel<IdxT, false>, dim3((unsigned)blocks), dim3(256), 0, st, w,
                               w_stride, (const IdxT*)idx, n, vecs, out, o_stride, (uint64_t)0,
                               (uint64_t)0));
  }
  B200_POST_LAUNCH();
  return 0;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_index_rows(const void* weights, int64_t weight_row_stride_bytes, const void* indices,
                               int idx64, int64_t num_indices, int64_t row_bytes, void* out,
                               int64_t out_row_stride_bytes, int64_t vocab_start, int64_t vocab_length,
                               void* stream) {
  B200_CHECK_ARG(row_bytes > 0 && row_bytes % 16 == 0, "index_rows: row_bytes %lld not a multiple of 16",
                 (long long)row_bytes);
  B200_CHECK_ARG(weight_row_stride_bytes % 16 == 0 && out_row_stride_bytes % 16 == 0 &&
                     weight_row_stride_bytes >= row_bytes && out_row_stride_bytes >= row_bytes,
                 "index_rows: row strides must be multiples of 16 bytes and cover a row");
  B200_CHECK_ARG(((uintptr_t)weights % 16) == 0 && ((uintptr_t)out % 16) == 0,
                 "index_rows: pointers must be 16-byte aligned");
  B200_CHECK_ARG(num_indices >= 0, "index_rows: negative length");
  B200_CHECK_ARG(vocab_length < 0 || vocab_start >= 0, "index_rows: negative vocab_start");
  B200_CHECK_ARG(row_bytes / 16 < (1ll << 30), "index_rows: row too long");
  if (num_indices == 0) return 0;
  auto st = (cudaStream_t)stream;
  const int vecs = (int)(row_bytes / 16);
  if (idx64)
    return launch_index<int64_t>((const uint8_t*)weights, weight_row_stride_bytes, indices, num_indices, vecs,
                                 (uint8_t*)out, out_row_stride_bytes, vocab_start, vocab_length, st);
  return launch_index<int32_t>((const uint8_t*)weights, weight_row_stride_bytes, indices, num_indices, vecs,
                               (uint8_t*)out, out_row_stride_bytes, vocab_start, vocab_length, st);
}
