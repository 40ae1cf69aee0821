// Per-head RMSNorm + neox RoPE on one lane group -- the arithmetic shared by the stand-alone
// qknorm_rope_kernel (elementwise.cu) and the Q loader of the fused decode kernel
// (attn_decode_tc.cu), so that both produce the same bits (reference sequence:
// python/minisgl/layers/attention.py:50-54 = flashinfer.rmsnorm x2 + apply_rope_with_cos_sin_cache_inplace).
//
// A group of G = D/8 consecutive lanes owns one (token, head) row; lane j holds elements [8j, 8j+8).
// Lanes j < G/2 hold the first half of the row, their rotation partner is lane j ^ (G/2).
// EVERY lane of the warp must call this (full-mask shuffles), active or not.
#pragma once
#include "common.cuh"

namespace b200 {

template <typename T, int G, bool kNorm>
__device__ __forceinline__ Vec8 qknorm_rope_lanes(const Vec8 xv, const int j, const T* __restrict__ w /* nullable, uniform per group */,
                                                  const float eps, const float* __restrict__ cs_row /* cos|sin row of the token's position */,
                                                  const bool active) {
  constexpr int kDim = G * 8;
  constexpr int kHalf = kDim / 2;
  float f[8];
  unpack8<T>(xv, f);
  if constexpr (kNorm) {
    // the group reduction runs unconditionally: groups with and without a weight can share a warp, so a
    // full-mask shuffle must not sit behind the per-group `w != nullptr` branch
    float ss = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) ss += f[i] * f[i];
#pragma unroll
    for (int o = G / 2; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (w != nullptr) {  // uniform per group
      const float rcp = rsqrtf(ss / (float)kDim + eps);
      float wf[8];
      Vec8 wv = *reinterpret_cast<const Vec8*>(w + j * 8);
      unpack8<T>(wv, wf);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = f[i] * rcp * wf[i];
      // the unfused reference stores the normed value (one rounding) before RoPE reads it
      Vec8 rounded = pack8<T>(f);
      unpack8<T>(rounded, f);
    }
  }
  // rotation partner values
  float p[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) p[i] = __shfl_xor_sync(0xffffffffu, f[i], G / 2);
  Vec8 out = xv;
  if (active) {
    const int ci = (j * 8) % kHalf;
    const float4* cp = reinterpret_cast<const float4*>(cs_row + ci);
    const float4* sp = reinterpret_cast<const float4*>(cs_row + kHalf + ci);
    float4 c0 = __ldg(cp), c1 = __ldg(cp + 1), s0 = __ldg(sp), s1 = __ldg(sp + 1);
    const float c[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
    const float s[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
    const bool first = j < G / 2;
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = f[i] * c[i] + (first ? -p[i] : p[i]) * s[i];
    out = pack8<T>(o);
  }
  return out;
}

}  // namespace b200
