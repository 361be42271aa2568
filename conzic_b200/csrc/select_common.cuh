// Device helpers shared by the selection kernels (select_ops.cu) and the certified-argmax kernels (cert_ops.cu).
// The cosine / logit arithmetic lives here once so that a candidate re-scored by the certification pass gets
// bit-for-bit the value the plain score_select_kernel computes from the same embedding.
#pragma once
#include "kernels.h"

namespace conzic {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// Block-wide reductions with a fixed tree (deterministic run to run).  scratch: >= 33 floats.
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = lane < nw ? scratch[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}
__device__ __forceinline__ float block_max(float v, float* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[w] = v;
  __syncthreads();
  if (w == 0) {
    float t = lane < nw ? scratch[lane] : -INFINITY;
    t = warp_max(t);
    if (lane == 0) scratch[32] = t;
  }
  __syncthreads();
  return scratch[32];
}

// scale * <e / ||e||, v / ||v||> (clip/clip.py:91-96), one warp per text embedding row e (image row v); every lane
// returns the value.  D % 4 == 0, rows 16-byte aligned.
__device__ __forceinline__ float sel_logit(const float* __restrict__ e, const float* __restrict__ v, int D, float scale,
                                           int lane) {
  float s2 = 0.f, v2 = 0.f;
  for (int i = lane * 4; i < D; i += 128) {
    const float4 x = *reinterpret_cast<const float4*>(e + i);
    const float4 y = __ldg(reinterpret_cast<const float4*>(v + i));
    s2 += (x.x * x.x + x.y * x.y) + (x.z * x.z + x.w * x.w);
    v2 += (y.x * y.x + y.y * y.y) + (y.z * y.z + y.w * y.w);
  }
  const float en = sqrtf(warp_sum(s2)), vn = sqrtf(warp_sum(v2));
  float dot = 0.f;
  for (int i = lane * 4; i < D; i += 128) {
    const float4 x = *reinterpret_cast<const float4*>(e + i);
    const float4 y = __ldg(reinterpret_cast<const float4*>(v + i));
    dot += ((x.x / en) * (y.x / vn) + (x.y / en) * (y.y / vn)) + ((x.z / en) * (y.z / vn) + (x.w / en) * (y.w / vn));
  }
  return warp_sum(dot) * scale;
}

// out[k] = softmax over k of x[0..K) (x in shared or global memory, out in shared memory); whole block, fixed
// reduction tree; ends with a barrier.
__device__ __forceinline__ void sel_softmax(const float* x, int K, float* out, float* scratch) {
  float mx = -INFINITY;
  for (int k = threadIdx.x; k < K; k += blockDim.x) mx = fmaxf(mx, x[k]);
  mx = block_max(mx, scratch);
  float sum = 0.f;
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    const float e = expf(x[k] - mx);
    out[k] = e;
    sum += e;
  }
  sum = block_sum(sum, scratch);
  for (int k = threadIdx.x; k < K; k += blockDim.x) out[k] = out[k] / sum;
  __syncthreads();
}

// alpha * p + beta * clip_score (+ gamma * senti_prob + 0.1 * (1 - exp(repeats)))   (gen_utils.py:77,
// control_gen_utils.py:59); `o` = b * K + k.
__device__ __forceinline__ float sel_fuse(const SelectArgs& a, size_t o, float cscore, float sprob) {
  float f = a.alpha * a.probs[o] + a.beta * cscore;
  if (a.senti) {
    f = f + a.gamma * sprob;
    f = f + 0.1f * (1.0f - expf(a.repeats ? a.repeats[o] : 0.f));
  }
  return f;
}
// the part of sel_fuse that does not depend on the CLIP tower
__device__ __forceinline__ float sel_fuse_exact_terms(const SelectArgs& a, size_t o, float sprob) {
  float f = a.alpha * a.probs[o];
  if (a.senti) {
    f = f + a.gamma * sprob;
    f = f + 0.1f * (1.0f - expf(a.repeats ? a.repeats[o] : 0.f));
  }
  return f;
}

// Block argmax of per-thread (value, index) candidates, lowest index on ties; every thread returns the index.
// scratch: >= 40 floats.
__device__ __forceinline__ int sel_block_argmax(float bestv, int besti, float* scratch) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bestv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, besti, o);
    if (ov > bestv || (ov == bestv && oi < besti)) { bestv = ov; besti = oi; }
  }
  float* rv = scratch;
  int* ri = reinterpret_cast<int*>(scratch + 16);
  __syncthreads();
  if (lane == 0) { rv[w] = bestv; ri[w] = besti; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float bv = rv[0]; int bi = ri[0];
    for (int ww = 1; ww < nw; ++ww)
      if (rv[ww] > bv || (rv[ww] == bv && ri[ww] < bi)) { bv = rv[ww]; bi = ri[ww]; }
    ri[15] = bi;
  }
  __syncthreads();
  return ri[15];
}

// gen_utils.py:79-80, control_gen_utils.py:63: the winner's token, cosine and control score.
__device__ __forceinline__ void sel_write_winner(const SelectArgs& a, int b, int bi, float logit_bi) {
  const size_t o = static_cast<size_t>(b) * a.K + bi;
  if (a.inp) a.inp[static_cast<size_t>(b) * a.L + a.pos] = a.ids_masked[o];
  if (a.out_clip_ref) a.out_clip_ref[b] = logit_bi / a.scale;
  if (a.out_senti && a.senti) a.out_senti[b] = a.senti[o];
  if (a.tr_best) a.tr_best[b] = bi;
}

}  // namespace conzic
