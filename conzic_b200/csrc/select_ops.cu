// The reduction-shaped part of the Gibbs step: temperature softmax * stop-word mask -> top-K,
// candidate -> CLIP id assembly, cosine / softmax / score fuse / argmax.  HBM-bandwidth or latency bound:
// coalesced vector loads, warp-shuffle reductions, everything stays on the device.
#include "kernels.h"

namespace conzic {

namespace {

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
__device__ float block_sum(float v, float* scratch) {
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
__device__ float block_max(float v, float* scratch) {
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

// ---------------------------------------------------------------------------------------------------
// generate_caption_step (gen_utils.py:33-49): one CTA per image row.  The whole vocabulary row lives in
// shared memory (V * 4 B = 122 KB): softmax(logits / T) * mask, exact K-th value by 4-pass radix select,
// ties at the threshold taken in ascending vocabulary index, bitonic sort of the K survivors.
// ---------------------------------------------------------------------------------------------------
constexpr int TOPK_THREADS = 1024;

__global__ void __launch_bounds__(TOPK_THREADS) topk_kernel(const float* __restrict__ logits, int ldl, int V,
                                                            const float* __restrict__ mask, float temperature, int K,
                                                            int K2, float* __restrict__ probs,
                                                            int64_t* __restrict__ ids) {
  PDL_ENTRY();
  extern __shared__ __align__(16) unsigned char topk_smem[];
  unsigned long long* sortbuf = reinterpret_cast<unsigned long long*>(topk_smem);          // [K2]
  uint32_t* keys = reinterpret_cast<uint32_t*>(sortbuf + K2);                                // [V]
  int* hist = reinterpret_cast<int*>(keys + ((V + 3) & ~3));                                 // [32][256]
  float* scratch = reinterpret_cast<float*>(hist + 32 * 256);                                // [40]
  int* ctl = reinterpret_cast<int*>(scratch + 40);                                           // [8]

  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const float* x = logits + static_cast<size_t>(blockIdx.x) * ldl;
  float* kf = reinterpret_cast<float*>(keys);

  float mx = -INFINITY;
  for (int i = tid; i < V; i += TOPK_THREADS) {
    const float t = x[i] / temperature;
    kf[i] = t;
    mx = fmaxf(mx, t);
  }
  mx = block_max(mx, scratch);
  float sum = 0.f;
  for (int i = tid; i < V; i += TOPK_THREADS) {
    const float e = expf(kf[i] - mx);
    kf[i] = e;
    sum += e;
  }
  sum = block_sum(sum, scratch);
  for (int i = tid; i < V; i += TOPK_THREADS) kf[i] = (kf[i] / sum) * __ldg(mask + i);
  __syncthreads();

  // ---- radix select of the K-th largest key (non-negative floats order like their bit patterns)
  uint32_t prefix = 0, pmask = 0;
  int remaining = K;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int i = tid; i < 32 * 256; i += TOPK_THREADS) hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < V; i += TOPK_THREADS) {
      const uint32_t k = keys[i];
      if ((k & pmask) == prefix) atomicAdd(&hist[w * 256 + ((k >> shift) & 255)], 1);
    }
    __syncthreads();
    if (tid < 256) {
      int c = 0;
      for (int ww = 0; ww < 32; ++ww) c += hist[ww * 256 + tid];
      hist[tid] = c;  // warp 0's histogram row now holds the totals (each thread only touches column tid)
    }
    __syncthreads();
    if (tid == 0) {
      int cum = 0, digit = 0, rem = remaining;
      for (int bin = 255; bin >= 0; --bin) {
        const int h = hist[bin];
        if (cum + h >= rem) { digit = bin; rem -= cum; break; }
        cum += h;
      }
      ctl[0] = digit;
      ctl[1] = rem;
    }
    __syncthreads();
    prefix |= static_cast<uint32_t>(ctl[0]) << shift;
    pmask |= 255u << shift;
    remaining = ctl[1];
    __syncthreads();
  }
  const uint32_t T = prefix;          // K-th largest value
  const int n_gt = K - remaining;     // strictly greater than T; `remaining` ties are still needed

  if (tid == 0) { ctl[2] = 0; ctl[3] = 0; }
  for (int i = tid; i < K2; i += TOPK_THREADS) sortbuf[i] = 0ull;
  __syncthreads();
  for (int i = tid; i < V; i += TOPK_THREADS) {
    const uint32_t k = keys[i];
    if (k > T) {
      const int slot = atomicAdd(&ctl[2], 1);
      sortbuf[slot] = (static_cast<unsigned long long>(k) << 32) | (0xFFFFFFFFu - static_cast<uint32_t>(i));
    }
  }
  // ties at T: lowest vocabulary indices first (ordered block scan over the row)
  int tie_base = 0;
  int* wtot = reinterpret_cast<int*>(scratch);  // reuse [0,32)
  for (int c0 = 0; c0 < V && tie_base < remaining; c0 += TOPK_THREADS) {
    const int i = c0 + tid;
    const bool is = (i < V) && (keys[i] == T);
    const unsigned bal = __ballot_sync(0xffffffffu, is);
    __syncthreads();
    if (lane == 0) wtot[w] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
    for (int ww = 0; ww < 32; ++ww) {
      const int c = wtot[ww];
      if (ww < w) before += c;
      total += c;
    }
    if (is) {
      const int rank = tie_base + before + __popc(bal & ((1u << lane) - 1u));
      if (rank < remaining)
        sortbuf[n_gt + rank] = (static_cast<unsigned long long>(T) << 32) | (0xFFFFFFFFu - static_cast<uint32_t>(i));
    }
    tie_base += total;
  }
  __syncthreads();
  // ---- bitonic sort, descending, K2 a power of two
  for (int size = 2; size <= K2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < (K2 >> 1); t += TOPK_THREADS) {
        const int i = ((t / stride) * (stride << 1)) + (t % stride);
        const int j = i + stride;
        const bool desc = (i & size) == 0;
        const unsigned long long a = sortbuf[i], b = sortbuf[j];
        if (desc ? (a < b) : (a > b)) { sortbuf[i] = b; sortbuf[j] = a; }
      }
      __syncthreads();
    }
  }
  for (int k = tid; k < K; k += TOPK_THREADS) {
    const unsigned long long c = sortbuf[k];
    probs[static_cast<size_t>(blockIdx.x) * K + k] = __uint_as_float(static_cast<uint32_t>(c >> 32));
    ids[static_cast<size_t>(blockIdx.x) * K + k] = static_cast<int64_t>(0xFFFFFFFFu - static_cast<uint32_t>(c));
  }
}

size_t topk_smem_bytes(int V, int K2) {
  return static_cast<size_t>(K2) * 8 + static_cast<size_t>((V + 3) & ~3) * 4 + 32 * 256 * 4 + 40 * 4 + 8 * 4;
}

// ---------------------------------------------------------------------------------------------------
// Candidate captions -> CLIP ids (gen_utils.py:71-75, clip/clip.py:71-77) through the BERT-id -> BPE CSR
// table.  One CTA per image; thread 0 walks the caption once (prefix before `pos`, tail after), then each
// thread builds the suffix of one candidate.
// ---------------------------------------------------------------------------------------------------
constexpr int MAX_BODY = 96;

__device__ __forceinline__ bool is_special(const AssembleArgs& a, int64_t id) {
  return id == a.special[0] || id == a.special[1] || id == a.special[2] || id == a.special[3] || id == a.special[4];
}

__global__ void assemble_kernel(AssembleArgs a) {
  PDL_ENTRY();
  __shared__ int pre[MAX_BODY];
  __shared__ int tail[MAX_BODY];
  __shared__ int s_np, s_nt;
  const int b = blockIdx.x;
  const int64_t* row = a.inp + static_cast<size_t>(b) * a.L;
  const int body_max = a.maxlen - 2;
  if (threadIdx.x == 0) {
    int np = 0, nt = 0;
    if (a.ov_mask && a.ov_mask[b]) {  // host-tokenised prefix / tail (a merged '##' word sits in this caption)
      for (int t = a.ov_off[2 * b]; t < a.ov_off[2 * b + 1]; ++t) if (np < MAX_BODY) pre[np++] = a.ov_tok[t];
      for (int t = a.ov_off[2 * b + 1]; t < a.ov_off[2 * b + 2]; ++t) if (nt < MAX_BODY) tail[nt++] = a.ov_tok[t];
    } else
    for (int j = 0; j < a.L; ++j) {
      if (j == a.pos) continue;
      const int64_t id = row[j];
      if (is_special(a, id) || id < 0 || id >= a.V) continue;
      for (int t = a.off[id]; t < a.off[id + 1]; ++t) {
        if (j < a.pos) { if (np < MAX_BODY) pre[np++] = a.tok[t]; }
        else           { if (nt < MAX_BODY) tail[nt++] = a.tok[t]; }
      }
    }
    if (np > body_max) np = body_max;
    s_np = np; s_nt = nt;
    if (a.P > 0) {
      int32_t* o = a.ids_prefix + static_cast<size_t>(b) * a.P;
      int n = 0;
      if (n < a.P) o[n++] = a.bos;
      for (int t = 0; t < np && n < a.P; ++t) o[n++] = pre[t];
      const int p0v = n;
      while (n < a.P) o[n++] = a.eos;
      a.p0[b] = p0v;
    } else if (a.p0) {
      a.p0[b] = 0;
    }
  }
  __syncthreads();
  const int np = s_np, nt = s_nt;
  for (int k = threadIdx.x; k < a.K; k += blockDim.x) {
    const size_t bk = static_cast<size_t>(b) * a.K + k;
    const int64_t c = a.ids[bk];
    const float m = a.token_mask[c];
    const int64_t cm = static_cast<int64_t>(static_cast<float>(c) * m);   // idxs * token_mask[0][idxs] (gen_utils.py:72)
    a.ids_masked[bk] = cm;
    int32_t* o = a.ids_suffix + bk * a.S;
    int n = 0;
    if (a.P == 0) {
      if (n < a.S) o[n++] = a.bos;
      for (int t = 0; t < np && n < a.S; ++t) o[n++] = pre[t];
    }
    int budget = body_max - np;
    if (!(is_special(a, cm) || cm < 0 || cm >= a.V)) {
      for (int t = a.off[cm]; t < a.off[cm + 1] && budget > 0 && n < a.S; ++t, --budget) o[n++] = a.tok[t];
    }
    for (int t = 0; t < nt && budget > 0 && n < a.S; ++t, --budget) o[n++] = tail[t];
    if (n >= a.S) n = a.S - 1;  // host sizing guarantees this never triggers; keeps memory safe
    a.eos_idx[bk] = n;
    while (n < a.S) o[n++] = a.eos;
    if (a.repeats) {
      // control_gen_utils.py:53: (idxs_ == topk_inp).sum - 1, the candidate column itself included in the sum
      int rep = 0;
      for (int j = 0; j < a.L; ++j)
        if (j != a.pos && row[j] == cm) ++rep;
      a.repeats[bk] = static_cast<float>(rep);
    }
    if (a.senti) {
      // per-word control score summed over the visible words of the candidate caption, in caption order
      double s = 0.0;
      for (int j = 0; j < a.L; ++j) {
        const int64_t id = (j == a.pos) ? cm : row[j];
        if (is_special(a, id) || id < 0 || id >= a.V) continue;
        s += static_cast<double>(a.senti_table[id]);
      }
      a.senti[bk] = static_cast<float>(s);
    }
  }
}

__global__ void step_prologue_kernel(int64_t* inp, int B, int L, int pos, int mask_id, float* token_mask, int dot_id,
                                     int dot_allowed) {
  PDL_ENTRY();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) inp[static_cast<size_t>(i) * L + pos] = mask_id;       // gen_utils.py:67
  if (i == 0 && token_mask && dot_id >= 0) token_mask[dot_id] = dot_allowed ? 1.0f : 0.0f;  // utils.py:53-59
}

__global__ void gather_rows_index_kernel(int32_t* rows, int B, int L, int pos) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) rows[i] = i * L + pos;
}

__global__ void pool_index_kernel(int32_t* rows, const int32_t* eos_idx, int B, int P, int K, int S) {
  PDL_ENTRY();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * K) rows[i] = B * P + i * S + eos_idx[i];
}

// ---------------------------------------------------------------------------------------------------
// clip/clip.py:86-98 + gen_utils.py:77-81 + control_gen_utils.py:59-65.  One CTA per image.
// ---------------------------------------------------------------------------------------------------
constexpr int SEL_THREADS = 256;

__global__ void __launch_bounds__(SEL_THREADS) score_select_kernel(SelectArgs a) {
  PDL_ENTRY();
  extern __shared__ float sel_smem[];
  float* vhat = sel_smem;            // [D]
  float* logit = vhat + a.D;         // [K]
  float* cscore = logit + a.K;       // [K]
  float* sprob = cscore + a.K;       // [K]
  float* scratch = sprob + a.K;      // [40]
  __shared__ float s_bestv;
  __shared__ int s_besti;

  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, w = tid >> 5, nw = SEL_THREADS >> 5;
  const float* v = a.image + static_cast<size_t>(b) * a.D;
  float ss = 0.f;
  for (int i = tid; i < a.D; i += SEL_THREADS) ss += v[i] * v[i];
  const float vn = sqrtf(block_sum(ss, scratch));
  for (int i = tid; i < a.D; i += SEL_THREADS) vhat[i] = v[i] / vn;
  __syncthreads();

  for (int k = w; k < a.K; k += nw) {
    const float* e = a.text + (static_cast<size_t>(b) * a.K + k) * a.D;
    float s2 = 0.f;
    for (int i = lane * 4; i < a.D; i += 128) {
      float4 x = *reinterpret_cast<const float4*>(e + i);
      s2 += (x.x * x.x + x.y * x.y) + (x.z * x.z + x.w * x.w);
    }
    const float en = sqrtf(warp_sum(s2));
    float dot = 0.f;
    for (int i = lane * 4; i < a.D; i += 128) {
      float4 x = *reinterpret_cast<const float4*>(e + i);
      float4 y = *reinterpret_cast<const float4*>(vhat + i);
      dot += ((x.x / en) * y.x + (x.y / en) * y.y) + ((x.z / en) * y.z + (x.w / en) * y.w);
    }
    dot = warp_sum(dot);
    if (lane == 0) logit[k] = dot * a.scale;
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int k = tid; k < a.K; k += SEL_THREADS) mx = fmaxf(mx, logit[k]);
  mx = block_max(mx, scratch);
  float sum = 0.f;
  for (int k = tid; k < a.K; k += SEL_THREADS) {
    const float e = expf(logit[k] - mx);
    cscore[k] = e;
    sum += e;
  }
  sum = block_sum(sum, scratch);
  for (int k = tid; k < a.K; k += SEL_THREADS) {
    const float c = cscore[k] / sum;
    cscore[k] = c;
    const size_t o = static_cast<size_t>(b) * a.K + k;
    if (a.tr_clip_score) a.tr_clip_score[o] = c;
    if (a.tr_clip_ref) a.tr_clip_ref[o] = logit[k] / a.scale;
  }
  if (!a.probs) return;

  if (a.senti) {  // softmax over K of the raw control scores (sentiments_classifer.py:46-47, temperature 1)
    float smx = -INFINITY;
    for (int k = tid; k < a.K; k += SEL_THREADS) smx = fmaxf(smx, a.senti[static_cast<size_t>(b) * a.K + k]);
    smx = block_max(smx, scratch);
    float ssum = 0.f;
    for (int k = tid; k < a.K; k += SEL_THREADS) {
      const float e = expf(a.senti[static_cast<size_t>(b) * a.K + k] - smx);
      sprob[k] = e;
      ssum += e;
    }
    ssum = block_sum(ssum, scratch);
    for (int k = tid; k < a.K; k += SEL_THREADS) sprob[k] = sprob[k] / ssum;
  }
  __syncthreads();

  float bestv = -INFINITY;
  int besti = 0x7fffffff;
  for (int k = tid; k < a.K; k += SEL_THREADS) {
    const size_t o = static_cast<size_t>(b) * a.K + k;
    float f = a.alpha * a.probs[o] + a.beta * cscore[k];
    if (a.senti) {
      f = f + a.gamma * sprob[k];
      f = f + 0.1f * (1.0f - expf(a.repeats ? a.repeats[o] : 0.f));
    }
    if (a.tr_final) a.tr_final[o] = f;
    if (f > bestv || (f == bestv && k < besti)) { bestv = f; besti = k; }
  }
  // argmax, first index on ties (torch.argmax on CPU)
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
  if (tid == 0) {
    float bv = rv[0]; int bi = ri[0];
    for (int ww = 1; ww < nw; ++ww)
      if (rv[ww] > bv || (rv[ww] == bv && ri[ww] < bi)) { bv = rv[ww]; bi = ri[ww]; }
    s_bestv = bv; s_besti = bi;
    if (bi < 0 || bi >= a.K) bi = 0;  // all-NaN row: fall back to candidate 0 like argmax of NaNs is undefined
    const size_t o = static_cast<size_t>(b) * a.K + bi;
    if (a.inp) a.inp[static_cast<size_t>(b) * a.L + a.pos] = a.ids_masked[o];    // gen_utils.py:79
    if (a.out_clip_ref) a.out_clip_ref[b] = logit[bi] / a.scale;                  // gen_utils.py:80
    if (a.out_senti && a.senti) a.out_senti[b] = a.senti[o];                      // control_gen_utils.py:63
    if (a.tr_best) a.tr_best[b] = bi;
  }
}

}  // namespace

bool topk_configure() {
  return cuda_ok(cudaFuncSetAttribute(topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024),
                 "cudaFuncSetAttribute(topk)");
}

bool launch_topk(const float* logits, int ldl, int B, int V, const float* mask, float temperature, int K, float* probs,
                 int64_t* ids, cudaStream_t st) {
  ++g_launches;
  ProfScope prof_(CAT_TOPK, static_cast<double>(B) * V, st);
  if (K < 1 || K > 1024 || K > V) {
    set_error("topk: K must be in [1, min(1024, V)]");
    return false;
  }
  int K2 = 2;
  while (K2 < K) K2 <<= 1;
  const size_t smem = topk_smem_bytes(V, K2);
  if (smem > 200 * 1024) {
    set_error("topk: vocabulary too large for the shared-memory row buffer");
    return false;
  }
  launch_k(topk_kernel, dim3(B), dim3(TOPK_THREADS), smem, st, logits, ldl, V, mask, temperature, K, K2, probs, ids);
  return cuda_ok(cudaGetLastError(), "topk launch");
}

void launch_assemble(const AssembleArgs& a, cudaStream_t st) {
  ++g_launches;
  ProfScope prof_(CAT_ASSEMBLE, 0, st);
  launch_k(assemble_kernel, dim3(a.B), dim3(256), 0, st, a);
}

void launch_step_prologue(int64_t* inp, int B, int L, int pos, int mask_id, float* token_mask, int dot_id,
                          int dot_allowed, cudaStream_t st) {
  ++g_launches;
  ProfScope prof_(CAT_MISC, 0, st);
  launch_k(step_prologue_kernel, dim3((B + 255) / 256), dim3(256), 0, st, inp, B, L, pos, mask_id, token_mask, dot_id, dot_allowed);
}

void launch_gather_rows_index(int32_t* rows, int B, int L, int pos, cudaStream_t st) {
  ++g_launches;
  ProfScope prof_(CAT_MISC, 0, st);
  gather_rows_index_kernel<<<(B + 255) / 256, 256, 0, st>>>(rows, B, L, pos);
}

void launch_pool_index(int32_t* rows, const int32_t* eos_idx, int B, int P, int K, int S, cudaStream_t st) {
  ++g_launches;
  ProfScope prof_(CAT_MISC, 0, st);
  launch_k(pool_index_kernel, dim3((B * K + 255) / 256), dim3(256), 0, st, rows, eos_idx, B, P, K, S);
}

void launch_score_select(const SelectArgs& a, cudaStream_t st) {
  ++g_launches;
  ProfScope prof_(CAT_SELECT, 0, st);
  const size_t smem = static_cast<size_t>(a.D + 3 * a.K + 40) * sizeof(float);
  launch_k(score_select_kernel, dim3(a.B), dim3(SEL_THREADS), smem, st, a);
}

}  // namespace conzic
