// Split-KV combine shared by the decode kernels: merges the per-chunk partial (o, m, l) written by
// the attention kernel into the final output (flash-decoding reduction).
#pragma once
#include "common.cuh"

namespace b200 {

constexpr int kPlanHeader = 4;   // decode_plan = {chunk_tokens, total_chunks, bs, 0, chunk_start[bs+1]}
constexpr int kMaxSplits = 16;   // chunks per request (metadata.cu guarantees the bound)
constexpr int kHeadDim = 128;

// ------------------------------------------------------------------------------- combine
// grid (bs, hq), 128 threads (one per output dim). Requests with a single chunk were already
// written by the attention kernel.
template <typename T>
__global__ void __launch_bounds__(kHeadDim) attn_combine_kernel(const float* __restrict__ part_o,
                                                          const float* __restrict__ part_ml,
                                                          const int32_t* __restrict__ plan, int hq,
                                                          T* __restrict__ out) {
  const int r = blockIdx.x, hh = blockIdx.y, d = threadIdx.x;
  pdl_wait();                // the decode kernel's partials
  pdl_launch_dependents();   // the next kernel may start its prologue
  const int32_t* chunk_start = plan + kPlanHeader;
  const int n = chunk_start[r + 1] - chunk_start[r];
  if (n <= 1) return;
  float m = -INFINITY;
  for (int c = 0; c < n; ++c)
    m = fmaxf(m, part_ml[(((int64_t)r * kMaxSplits + c) * hq + hh) * 2]);
  float acc = 0.f, l = 0.f;
  for (int c = 0; c < n; ++c) {
    const int64_t idx = ((int64_t)r * kMaxSplits + c) * hq + hh;
    const float w = fast_exp2(part_ml[idx * 2] - m);
    l += w * part_ml[idx * 2 + 1];
    acc += w * part_o[idx * kHeadDim + d];
  }
  out[((int64_t)r * hq + hh) * kHeadDim + d] = DTypeTraits<T>::from_float(acc / l);
}


}  // namespace b200
