// The reduction-shaped part of the Gibbs step: temperature softmax * stop-word mask -> top-K,
// candidate -> CLIP id assembly, cosine / softmax / score fuse / argmax.  HBM-bandwidth or latency bound:
// coalesced vector loads, warp-shuffle reductions, everything stays on the device.
#include <cooperative_groups.h>

#include "kernels.h"
#include "select_common.cuh"

namespace conzic {

namespace {

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

// ---------------------------------------------------------------------------------------------------
// The same selection with one image row spread over a thread-block cluster of TOPK_CL CTAs (B = 64 rows would
// otherwise occupy 64 of 148 SMs, each streaming a 122 KB row on its own): CTA r of the cluster holds the slice
// [r * per, (r + 1) * per) of the vocabulary in its shared memory; the row maximum, the softmax sum, the four radix
// histograms and the candidate counts are exchanged through distributed shared memory (cluster.map_shared_rank), the
// K survivors are written into CTA 0's sort buffer, which sorts and stores them.  Same results as topk_kernel up to
// the rounding of the softmax sum (partial sums per slice, added in rank order); ties still in ascending index order.
// ---------------------------------------------------------------------------------------------------
constexpr int TOPK_CL = 4;
constexpr int TOPK_CL_THREADS = 512;

struct TopkClSmem {  // dynamic shared memory layout of one CTA
  unsigned long long* sortbuf;  // [K2]   (used in CTA 0 only, allocated everywhere so offsets agree)
  uint32_t* keys;               // [per]
  int* hist;                    // [16][256] per-warp histograms
  int* mine;                    // [2][256] this CTA's merged histogram, double buffered over the radix passes
  int* tot;                     // [256] cluster-wide histogram
  float* scratch;               // [40]
  float* red;                   // [4]  values other CTAs read: max, sum
  int* cnt;                     // [4]  values other CTAs read: > T count, == T count
  int* ctl;                     // [8]
  __device__ TopkClSmem(unsigned char* base, int K2, int per) {
    sortbuf = reinterpret_cast<unsigned long long*>(base);
    keys = reinterpret_cast<uint32_t*>(sortbuf + K2);
    hist = reinterpret_cast<int*>(keys + ((per + 3) & ~3));
    mine = hist + 16 * 256;
    tot = mine + 2 * 256;
    scratch = reinterpret_cast<float*>(tot + 256);
    red = scratch + 40;
    cnt = reinterpret_cast<int*>(red + 4);
    ctl = cnt + 4;
  }
};
size_t topk_cl_smem_bytes(int K2, int per) {
  return static_cast<size_t>(K2) * 8 + static_cast<size_t>((per + 3) & ~3) * 4 + (16 * 256 + 3 * 256) * 4 + (40 + 4) * 4 + (4 + 8) * 4;
}

__global__ void __launch_bounds__(TOPK_CL_THREADS) topk_cluster_kernel(const float* __restrict__ logits, int ldl, int V,
                                                                       const float* __restrict__ mask, float temperature,
                                                                       int K, int K2, int per, float* __restrict__ probs,
                                                                       int64_t* __restrict__ ids) {
  PDL_ENTRY();
  namespace cg = cooperative_groups;
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ __align__(16) unsigned char topk_cl_smem[];
  TopkClSmem sm(topk_cl_smem, K2, per);
  const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
  const int rank = static_cast<int>(cluster.block_rank());
  const int row = blockIdx.x / TOPK_CL;
  const int lo = rank * per, n = max(0, min(V, lo + per) - lo);
  const float* x = logits + static_cast<size_t>(row) * ldl + lo;
  float* kf = reinterpret_cast<float*>(sm.keys);

  if (rank == 0)
    for (int i = tid; i < K2; i += TOPK_CL_THREADS) sm.sortbuf[i] = 0ull;
  float mx = -INFINITY;
  for (int i = tid; i < n; i += TOPK_CL_THREADS) {
    const float t = x[i] / temperature;
    kf[i] = t;
    mx = fmaxf(mx, t);
  }
  mx = block_max(mx, sm.scratch);
  if (tid == 0) sm.red[0] = mx;
  cluster.sync();
  for (int r = 0; r < TOPK_CL; ++r) mx = fmaxf(mx, *cluster.map_shared_rank(&sm.red[0], r));
  float sum = 0.f;
  for (int i = tid; i < n; i += TOPK_CL_THREADS) {
    const float e = expf(kf[i] - mx);
    kf[i] = e;
    sum += e;
  }
  sum = block_sum(sum, sm.scratch);
  if (tid == 0) sm.red[1] = sum;
  cluster.sync();
  sum = 0.f;
  for (int r = 0; r < TOPK_CL; ++r) sum += *cluster.map_shared_rank(&sm.red[1], r);
  for (int i = tid; i < n; i += TOPK_CL_THREADS) kf[i] = (kf[i] / sum) * __ldg(mask + lo + i);
  __syncthreads();

  // ---- radix select of the K-th largest key over the whole row (non-negative floats order like their bit patterns).
  // One cluster barrier per pass: the merged histograms are double buffered, so a CTA can only overwrite buffer p & 1
  // after the barrier of pass p + 1, which every CTA reaches after it has read the pass-p histograms.
  uint32_t prefix = 0, pmask = 0;
  int remaining = K;
  for (int shift = 24, pass = 0; shift >= 0; shift -= 8, ++pass) {
    int* mine = sm.mine + (pass & 1) * 256;
    for (int i = tid; i < 16 * 256; i += TOPK_CL_THREADS) sm.hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += TOPK_CL_THREADS) {
      const uint32_t k = sm.keys[i];
      if ((k & pmask) == prefix) atomicAdd(&sm.hist[w * 256 + ((k >> shift) & 255)], 1);
    }
    __syncthreads();
    if (tid < 256) {
      int c = 0;
      for (int ww = 0; ww < 16; ++ww) c += sm.hist[ww * 256 + tid];
      mine[tid] = c;
    }
    cluster.sync();
    if (tid < 256) {
      int c = 0;
      for (int r = 0; r < TOPK_CL; ++r) c += *cluster.map_shared_rank(&mine[tid], r);
      sm.tot[tid] = c;
    }
    __syncthreads();
    if (w == 0) {
      // the highest bin whose count, added to everything above it, reaches `remaining`: lane l scans bins
      // 255 - 8 l ... 248 - 8 l, a warp scan orders the lanes (every CTA computes the same digit)
      int loc[8], ssum = 0;
#pragma unroll
      for (int u = 0; u < 8; ++u) { loc[u] = sm.tot[255 - (8 * lane + u)]; ssum += loc[u]; }
      int incl = ssum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      const unsigned hit = __ballot_sync(0xffffffffu, incl >= remaining);
      const int first = __ffs(hit) - 1;  // some lane always hits: the row holds at least `remaining` matching keys
      if (lane == first) {
        int cum = incl - ssum;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (cum + loc[u] >= remaining) { sm.ctl[0] = 255 - (8 * lane + u); sm.ctl[1] = remaining - cum; break; }
          cum += loc[u];
        }
      }
    }
    __syncthreads();
    prefix |= static_cast<uint32_t>(sm.ctl[0]) << shift;
    pmask |= 255u << shift;
    remaining = sm.ctl[1];
  }
  const uint32_t T = prefix;       // K-th largest value of the row
  const int n_gt = K - remaining;  // strictly greater than T; `remaining` ties are still needed

  // ---- how many keys > T and == T each CTA holds -> where its survivors go in CTA 0's sort buffer
  float c_gt = 0.f, c_eq = 0.f;
  for (int i = tid; i < n; i += TOPK_CL_THREADS) {
    const uint32_t k = sm.keys[i];
    c_gt += k > T ? 1.f : 0.f;
    c_eq += k == T ? 1.f : 0.f;
  }
  const int my_gt = static_cast<int>(block_sum(c_gt, sm.scratch));
  const int my_eq = static_cast<int>(block_sum(c_eq, sm.scratch));
  if (tid == 0) { sm.cnt[0] = my_gt; sm.cnt[1] = my_eq; sm.ctl[2] = 0; }
  cluster.sync();
  int gt_off = 0, eq_off = 0, eq_total = 0;
  for (int r = 0; r < TOPK_CL; ++r) {
    const int g = *cluster.map_shared_rank(&sm.cnt[0], r), e = *cluster.map_shared_rank(&sm.cnt[1], r);
    if (r < rank) { gt_off += g; eq_off += e; }
    eq_total += e;
  }
  unsigned long long* dst = cluster.map_shared_rank(sm.sortbuf, 0);
  // the usual case: every key equal to T is needed (the K-th value is unique): all keys >= T go in, in any order
  const bool all_ties = (eq_total == remaining);
  const int base = all_ties ? gt_off + eq_off : gt_off;
  for (int i = tid; i < n; i += TOPK_CL_THREADS) {
    const uint32_t k = sm.keys[i];
    if (k > T || (all_ties && k == T)) {
      const int slot = base + atomicAdd(&sm.ctl[2], 1);
      dst[slot] = (static_cast<unsigned long long>(k) << 32) | (0xFFFFFFFFu - static_cast<uint32_t>(lo + i));
    }
  }
  // more ties than needed (e.g. underflowed zeros): lowest vocabulary indices first -- CTAs in rank order, inside a CTA
  // an ordered block scan
  int tie_base = eq_off;
  int* wtot = reinterpret_cast<int*>(sm.scratch);  // reuse [0,16)
  for (int c0 = 0; !all_ties && c0 < n && tie_base < remaining; c0 += TOPK_CL_THREADS) {
    const int i = c0 + tid;
    const bool is = (i < n) && (sm.keys[i] == T);
    const unsigned bal = __ballot_sync(0xffffffffu, is);
    __syncthreads();
    if (lane == 0) wtot[w] = __popc(bal);
    __syncthreads();
    int before = 0, total = 0;
    for (int ww = 0; ww < TOPK_CL_THREADS / 32; ++ww) {
      const int c = wtot[ww];
      if (ww < w) before += c;
      total += c;
    }
    if (is) {
      const int rk = tie_base + before + __popc(bal & ((1u << lane) - 1u));
      if (rk < remaining)
        dst[n_gt + rk] = (static_cast<unsigned long long>(T) << 32) | (0xFFFFFFFFu - static_cast<uint32_t>(lo + i));
    }
    tie_base += total;
  }
  cluster.sync();  // all survivors are in CTA 0's buffer; the other CTAs are done
  if (rank != 0) return;
  // ---- bitonic sort, descending, K2 a power of two
  for (int size = 2; size <= K2; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int t = tid; t < (K2 >> 1); t += TOPK_CL_THREADS) {
        const int i = ((t / stride) * (stride << 1)) + (t % stride);
        const int j = i + stride;
        const bool desc = (i & size) == 0;
        const unsigned long long a = sm.sortbuf[i], b = sm.sortbuf[j];
        if (desc ? (a < b) : (a > b)) { sm.sortbuf[i] = b; sm.sortbuf[j] = a; }
      }
      __syncthreads();
    }
  }
  for (int k = tid; k < K; k += TOPK_CL_THREADS) {
    const unsigned long long c = sm.sortbuf[k];
    probs[static_cast<size_t>(row) * K + k] = __uint_as_float(static_cast<uint32_t>(c >> 32));
    ids[static_cast<size_t>(row) * K + k] = static_cast<int64_t>(0xFFFFFFFFu - static_cast<uint32_t>(c));
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

__global__ void pool_index_kernel(int32_t* rows, const int32_t* eos_idx, int n_pre_rows, int n_cand, int S) {
  PDL_ENTRY();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_cand) rows[i] = n_pre_rows + i * S + eos_idx[i];
}

// ---------------------------------------------------------------------------------------------------
// clip/clip.py:86-98 + gen_utils.py:77-81 + control_gen_utils.py:59-65, in two launches:
//   clip_logits_kernel   one WARP per candidate caption, grid over all B*K of them: scale * cos(text, image) --
//                        the only part that touches the [B*K, D] embeddings (HBM bound: B*K*D*4 bytes);
//   score_select_kernel  one CTA per image over its K logits: softmax_K, score fuse, argmax, write-back.
// ---------------------------------------------------------------------------------------------------
constexpr int SEL_THREADS = 256;

__global__ void __launch_bounds__(256) clip_logits_kernel(const float* __restrict__ text, const float* __restrict__ image,
                                                          const int32_t* __restrict__ row_bk, int n_rows, int K, int D,
                                                          float scale, float* __restrict__ logit) {
  PDL_ENTRY();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= n_rows) return;
  const int b = (row_bk ? row_bk[r] : r) / K;
  const float l = sel_logit(text + static_cast<size_t>(r) * D, image + static_cast<size_t>(b) * D, D, scale, lane);
  if (lane == 0) logit[r] = l;
}

__global__ void __launch_bounds__(SEL_THREADS) score_select_kernel(SelectArgs a) {
  PDL_ENTRY();
  extern __shared__ float sel_smem[];
  float* logit = sel_smem;           // [K]
  float* cscore = logit + a.K;       // [K]
  float* sprob = cscore + a.K;       // [K]
  float* scratch = sprob + a.K;      // [40]

  const int b = blockIdx.x, tid = threadIdx.x;
  for (int k = tid; k < a.K; k += SEL_THREADS) logit[k] = a.logit[static_cast<size_t>(b) * a.K + k];
  __syncthreads();
  sel_softmax(logit, a.K, cscore, scratch);
  for (int k = tid; k < a.K; k += SEL_THREADS) {
    const size_t o = static_cast<size_t>(b) * a.K + k;
    if (a.tr_clip_score) a.tr_clip_score[o] = cscore[k];
    if (a.tr_clip_ref) a.tr_clip_ref[o] = logit[k] / a.scale;
  }
  if (!a.probs) return;
  if (a.senti) sel_softmax(a.senti + static_cast<size_t>(b) * a.K, a.K, sprob, scratch);
  __syncthreads();

  float bestv = -INFINITY;
  int besti = 0x7fffffff;
  for (int k = tid; k < a.K; k += SEL_THREADS) {
    const size_t o = static_cast<size_t>(b) * a.K + k;
    const float f = sel_fuse(a, o, cscore[k], a.senti ? sprob[k] : 0.f);
    if (a.tr_final) a.tr_final[o] = f;
    if (f > bestv || (f == bestv && k < besti)) { bestv = f; besti = k; }
  }
  int bi = sel_block_argmax(bestv, besti, scratch);  // first index on ties (torch.argmax on CPU)
  if (tid == 0) {
    if (bi < 0 || bi >= a.K) bi = 0;  // all-NaN row: fall back to candidate 0 like argmax of NaNs is undefined
    sel_write_winner(a, b, bi, logit[bi]);
  }
}

}  // namespace

bool topk_configure() {
  return cuda_ok(cudaFuncSetAttribute(topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024),
                 "cudaFuncSetAttribute(topk)") &&
         cuda_ok(cudaFuncSetAttribute(topk_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024),
                 "cudaFuncSetAttribute(topk_cluster)");
}

bool launch_topk(const float* logits, int ldl, int B, int V, const float* mask, float temperature, int K, float* probs,
                 int64_t* ids, cudaStream_t st) {
  count_launch();
  ProfScope prof_(CAT_TOPK, static_cast<double>(B) * V, st);
  if (K < 1 || K > 1024 || K > V) {
    set_error("topk: K must be in [1, min(1024, V)]");
    return false;
  }
  int K2 = 2;
  while (K2 < K) K2 <<= 1;
  if (V >= 4096) {  // one row per cluster of TOPK_CL CTAs
    const int per = (V + TOPK_CL - 1) / TOPK_CL;
    const size_t smem_cl = topk_cl_smem_bytes(K2, per);
    if (smem_cl <= 100 * 1024) {
      cudaLaunchConfig_t cfg{};
      cfg.gridDim = dim3(static_cast<unsigned>(B) * TOPK_CL);
      cfg.blockDim = dim3(TOPK_CL_THREADS);
      cfg.dynamicSmemBytes = smem_cl;
      cfg.stream = st;
      cudaLaunchAttribute attr[2];
      attr[0].id = cudaLaunchAttributeClusterDimension;
      attr[0].val.clusterDim.x = TOPK_CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
      attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[1].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = pdl_enabled() ? 2 : 1;
      return cuda_ok(cudaLaunchKernelEx(&cfg, topk_cluster_kernel, logits, ldl, V, mask, temperature, K, K2, per, probs, ids),
                     "topk_cluster launch");
    }
  }
  const size_t smem = topk_smem_bytes(V, K2);
  if (smem > 200 * 1024) {
    set_error("topk: vocabulary too large for the shared-memory row buffer");
    return false;
  }
  launch_k(topk_kernel, dim3(B), dim3(TOPK_THREADS), smem, st, logits, ldl, V, mask, temperature, K, K2, probs, ids);
  return cuda_ok(cudaGetLastError(), "topk launch");
}

void launch_assemble(const AssembleArgs& a, cudaStream_t st) {
  count_launch();
  ProfScope prof_(CAT_ASSEMBLE, 0, st);
  launch_k(assemble_kernel, dim3(a.B), dim3(256), 0, st, a);
}

void launch_step_prologue(int64_t* inp, int B, int L, int pos, int mask_id, float* token_mask, int dot_id,
                          int dot_allowed, cudaStream_t st) {
  count_launch();
  ProfScope prof_(CAT_MISC, 0, st);
  launch_k(step_prologue_kernel, dim3((B + 255) / 256), dim3(256), 0, st, inp, B, L, pos, mask_id, token_mask, dot_id, dot_allowed);
}

void launch_pool_index(int32_t* rows, const int32_t* eos_idx, int n_pre_rows, int n_cand, int S, cudaStream_t st) {
  count_launch();
  ProfScope prof_(CAT_MISC, 0, st);
  if (n_cand <= 0) return;
  launch_k(pool_index_kernel, dim3((n_cand + 255) / 256), dim3(256), 0, st, rows, eos_idx, n_pre_rows, n_cand, S);
}

void launch_clip_logits(const float* text, const float* image, const int32_t* row_bk, int n_rows, int K, int D,
                        float scale, float* logit, cudaStream_t st) {
  count_launch();
  ProfScope prof_(CAT_SELECT, static_cast<double>(n_rows) * D * 4, st);
  if (n_rows <= 0) return;
  launch_k(clip_logits_kernel, dim3((n_rows + 7) / 8), dim3(256), 0, st, text, image, row_bk, n_rows, K, D, scale, logit);
}

void launch_score_select(const SelectArgs& a, cudaStream_t st) {
  count_launch();
  ProfScope prof_(CAT_SELECT, 0, st);
  const size_t smem = static_cast<size_t>(3 * a.K + 40) * sizeof(float);
  launch_k(score_select_kernel, dim3(a.B), dim3(SEL_THREADS), smem, st, a);
}

}  // namespace conzic
