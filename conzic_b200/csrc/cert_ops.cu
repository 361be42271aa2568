// Certified argmax: the winner of gen_utils.py:77-79 (control_gen_utils.py:59-61) as the exact (bf16x3, fp32-grade)
// CLIP tower would pick it, computed from the fast bf16 tower plus an exact re-score of a handful of candidates.
//
// The fused score of candidate k is  f_k = E_k + beta * exp(t_k) / Z,  Z = sum_j exp(t_j),  where t_k is the exact
// logit (scale * cosine) and E_k = alpha * p_k (+ gamma * senti_k + 0.1 (1 - exp(rep_k))) does not depend on the
// tower.  The bf16 tower gives a_k with -lo <= a_k - t_k <= hi (scale * measured bounds on the cosine error with a
// safety factor: the bf16 tower over-estimates on average, so the two sides differ; conzic_config.cert_dcos = hi,
// cert_dcos_lo = lo).  For the bf16 winner w and any other candidate k
//     f_w - f_k = (E_w - E_k) + beta * (exp(t_w) - exp(t_k)) / Z
//              >= (E_w - E_k) + beta * D / (D >= 0 ? Zhi : Zlo),   D = exp(a_w - hi) - exp(a_k + lo),
// with Zlo = r_lo Z_a <= Z <= r_hi Z_a = Zhi (r_lo = exp(-hi), r_hi = exp(lo) follow from the per-candidate bounds; the
// sum over ~200 candidates is far better behaved than its worst term, and cert_zratio_lo / _hi carry measured bounds on
// the ratio exact / bf16 of the part of the denominator that comes from bf16 logits).  If that lower bound is > tau, k cannot win in exact
// arithmetic and is dropped (round 1).  The survivors (always including w, whose exact cosine is also what the
// caller reports) are re-encoded by the exact tower; round 2 repeats the test among them with their exact logits
// (no error for them, [-lo, hi] for the rest of Z).  An image whose survivors still cannot be ordered -- or that has more
// than `fcap` of them -- is re-encoded in full by the exact tower and decided by the plain score_select_kernel, so
// every decision equals the bf16x3 mode's.  Candidates that are the same caption (masked ids, gen_utils.py:72) have
// identical logits in any arithmetic and are ordered by their exact terms alone.
// Round 2's only uncertainty is the part of Z that still comes from bf16 logits.  Listing the "heavy" candidates
// as well (cert_heavy > 0: softmax weight >= cert_heavy, doubled until the list fits fcap) narrows it, but with the
// flat softmax of this workload (the top weight is ~0.05) it costs more exact rows than the occasional full re-encode
// of an image (measured: 1 748 instead of 305 re-scored candidates per step, profiles/r02c), so it is off by default.
#include "kernels.h"
#include "select_common.cuh"

namespace conzic {

namespace {

constexpr int CERT_THREADS = 256;
constexpr int CERT_HEAVY_BIT = 1 << 30;  // img_k entry: listed for its weight in Z only, already ruled out as a winner

// lower bound of f_w - f_k given logits relative to a common maximum (ew = exp(a_w - m), ek likewise), their error
// bounds, and bounds on the softmax denominator (same reference m)
// (hi, lo): a - t lies in [-lo, hi] for both candidates (0, 0 when their logits are the exact ones)
__device__ __forceinline__ float cert_lower_bound(float dE, float beta, float ew, float ek, float hi, float lo,
                                                  float zlo, float zhi) {
  if (beta >= 0.f) {
    const float d = ew * expf(-hi) - ek * expf(lo);  // lower bound of exp(t_w) - exp(t_k)
    return dE + beta * (d >= 0.f ? d / zhi : d / zlo);
  }
  const float d = ew * expf(lo) - ek * expf(-hi);    // upper bound of exp(t_w) - exp(t_k)
  return dE + beta * (d >= 0.f ? d / zlo : d / zhi);
}

// Shared layout of both rounds: logit[K] | e[K] (exp, then clip score) | sprob[K] | exact[K] | scratch[40] | flag bytes[K]
struct CertSmem {
  float *logit, *e, *sprob, *exact, *scratch;
  unsigned char* flag;
  __device__ CertSmem(float* base, int K) {
    logit = base; e = logit + K; sprob = e + K; exact = sprob + K; scratch = exact + K;
    flag = reinterpret_cast<unsigned char*>(scratch + 40);
  }
};
size_t cert_smem_bytes(int K) { return static_cast<size_t>(4 * K + 40) * sizeof(float) + static_cast<size_t>((K + 15) & ~15); }

__global__ void __launch_bounds__(CERT_THREADS) cert_round1_kernel(CertArgs c) {
  PDL_ENTRY();
  extern __shared__ float cert_smem[];
  const SelectArgs& a = c.q;
  const int K = a.K;
  CertSmem sm(cert_smem, K);
  const int b = blockIdx.x, tid = threadIdx.x;
  const size_t o0 = static_cast<size_t>(b) * K;

  for (int k = tid; k < K; k += CERT_THREADS) sm.logit[k] = a.logit[o0 + k];
  __syncthreads();
  float mx = -INFINITY;
  for (int k = tid; k < K; k += CERT_THREADS) mx = fmaxf(mx, sm.logit[k]);
  mx = block_max(mx, sm.scratch);
  float sum = 0.f;
  for (int k = tid; k < K; k += CERT_THREADS) {
    const float e = expf(sm.logit[k] - mx);
    sm.e[k] = e;
    sum += e;
  }
  const float Z = block_sum(sum, sm.scratch);
  if (a.senti) sel_softmax(a.senti + o0, K, sm.sprob, sm.scratch);
  __syncthreads();

  float bestv = -INFINITY;
  int besti = 0x7fffffff;
  for (int k = tid; k < K; k += CERT_THREADS) {
    const float sp = a.senti ? sm.sprob[k] : 0.f;
    const float f = sel_fuse(a, o0 + k, sm.e[k] / Z, sp);
    sm.exact[k] = sel_fuse_exact_terms(a, o0 + k, sp);
    if (f > bestv || (f == bestv && k < besti)) { bestv = f; besti = k; }
  }
  int w = sel_block_argmax(bestv, besti, sm.scratch);
  if (w < 0 || w >= K) w = 0;
  __syncthreads();

  const float zlo = Z * c.zr_lo, zhi = Z * c.zr_hi;
  const float Ew = sm.exact[w], ew = sm.e[w];
  const int64_t idw = a.ids_masked[o0 + w];
  const float pw = a.probs[o0 + w];
  for (int k = tid; k < K; k += CERT_THREADS) {
    bool alive = true;
    if (k != w) {
      const float dE = Ew - sm.exact[k];
      if (a.ids_masked[o0 + k] == idw) {
        // the same caption: identical logits whatever the arithmetic
        alive = !(dE > c.tau || (a.probs[o0 + k] == pw && k > w));
      } else {
        alive = !(cert_lower_bound(dE, a.beta, ew, sm.e[k], c.eps_hi, c.eps_lo, zlo, zhi) > c.tau);
      }
    }
    sm.flag[k] = alive ? 1 : 0;
  }
  __syncthreads();
  float cnt = 0.f;
  for (int k = tid; k < K; k += CERT_THREADS) cnt += sm.flag[k] ? 1.f : 0.f;
  const int n_alive = static_cast<int>(block_sum(cnt, sm.scratch));
  if (n_alive <= c.fcap && c.heavy > 0.f) {
    float theta = c.heavy * Z;
#pragma unroll 1
    for (int it = 0; it < 8; ++it, theta *= 2.0f) {
      float add = 0.f;
      for (int k = tid; k < K; k += CERT_THREADS) add += (!sm.flag[k] && sm.e[k] >= theta) ? 1.f : 0.f;
      const int n_add = static_cast<int>(block_sum(add, sm.scratch));
      if (n_alive + n_add <= c.fcap) {
        for (int k = tid; k < K; k += CERT_THREADS)
          if (!sm.flag[k] && sm.e[k] >= theta) sm.flag[k] = 2;
        break;
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    int n = 0;
    for (int k = 0; k < K; ++k) n += sm.flag[k] != 0;
    if (n_alive > 1) atomicAdd(&c.counters[2], 1);
    if (n_alive > c.fcap) {
      c.img_nflag[b] = -1;
      c.full_list[atomicAdd(&c.counters[1], 1)] = b;
      atomicAdd(&c.counters[3], 1);
    } else {
      const int slot0 = atomicAdd(&c.counters[0], n);
      c.img_nflag[b] = n;
      c.img_slot0[b] = slot0;
      int i = 0;
      for (int k = 0; k < K; ++k)
        if (sm.flag[k]) {
          c.img_k[b * c.fcap + i] = sm.flag[k] == 2 ? (k | CERT_HEAVY_BIT) : k;
          c.flag_list[slot0 + i] = b * K + k;
          ++i;
        }
    }
  }
}

__global__ void __launch_bounds__(CERT_THREADS) cert_round2_kernel(CertArgs c) {
  PDL_ENTRY();
  extern __shared__ float cert_smem[];
  const SelectArgs& a = c.q;
  const int K = a.K;
  CertSmem sm(cert_smem, K);
  const int b = blockIdx.x, tid = threadIdx.x;
  const int n = c.img_nflag[b];
  if (n < 0) return;  // full re-encode decides this image
  const size_t o0 = static_cast<size_t>(b) * K;
  const int* ks = c.img_k + b * c.fcap;
  const int slot0 = c.img_slot0[b];

  for (int k = tid; k < K; k += CERT_THREADS) { sm.logit[k] = a.logit[o0 + k]; sm.flag[k] = 0; }
  __syncthreads();
  for (int i = tid; i < n; i += CERT_THREADS) {
    const int k = ks[i] & ~CERT_HEAVY_BIT;
    sm.logit[k] = c.logit3[slot0 + i];
    sm.flag[k] = 1;
  }
  __syncthreads();
  float mx = -INFINITY;
  for (int k = tid; k < K; k += CERT_THREADS) mx = fmaxf(mx, sm.logit[k]);
  mx = block_max(mx, sm.scratch);
  float sum = 0.f, sum_rest = 0.f;
  for (int k = tid; k < K; k += CERT_THREADS) {
    const float e = expf(sm.logit[k] - mx);
    sm.e[k] = e;
    sum += e;
    if (!sm.flag[k]) sum_rest += e;
  }
  const float Z = block_sum(sum, sm.scratch);
  const float Zrest = block_sum(sum_rest, sm.scratch);
  if (a.senti) sel_softmax(a.senti + o0, K, sm.sprob, sm.scratch);
  __syncthreads();
  for (int k = tid; k < K; k += CERT_THREADS) {
    const float sp = a.senti ? sm.sprob[k] : 0.f;
    const float cs = sm.e[k] / Z;
    const float f = sel_fuse(a, o0 + k, cs, sp);
    sm.exact[k] = sel_fuse_exact_terms(a, o0 + k, sp);
    sm.sprob[k] = f;  // the senti softmax is folded into exact[] now; reuse the slot for the fused score
    if (a.tr_clip_score) a.tr_clip_score[o0 + k] = cs;
    if (a.tr_clip_ref) a.tr_clip_ref[o0 + k] = sm.logit[k] / a.scale;
    if (a.tr_final) a.tr_final[o0 + k] = f;
  }
  __syncthreads();
  if (tid == 0) {
    int w = -1;
    for (int i = 0; i < n; ++i) {
      if (ks[i] & CERT_HEAVY_BIT) continue;  // ruled out in round 1
      if (w < 0 || sm.sprob[ks[i]] > sm.sprob[w]) w = ks[i];  // ks ascending: the lowest index wins ties
    }
    const float zfl = Z - Zrest;  // exact part of the denominator
    const float zlo = zfl + Zrest * c.zr_lo, zhi = zfl + Zrest * c.zr_hi;
    bool ok = true;
    const int64_t idw = a.ids_masked[o0 + w];
    for (int i = 0; i < n && ok; ++i) {
      const int k = ks[i];
      if (k == w || (k & CERT_HEAVY_BIT)) continue;
      const float dE = sm.exact[w] - sm.exact[k];
      if (a.ids_masked[o0 + k] == idw) ok = dE > c.tau || (a.probs[o0 + k] == a.probs[o0 + w] && k > w);
      else ok = cert_lower_bound(dE, a.beta, sm.e[w], sm.e[k], 0.f, 0.f, zlo, zhi) > c.tau;
    }
    if (ok) sel_write_winner(a, b, w, sm.logit[w]);
    else c.full_list[atomicAdd(&c.counters[1], 1)] = b;
  }
}

__global__ void cert_gather_suffix_kernel(const int32_t* __restrict__ flag_list, int n, const int32_t* __restrict__ ids_suffix,
                                          const int32_t* __restrict__ eos_idx, int K, int S, int32_t* __restrict__ out_ids,
                                          int32_t* __restrict__ out_eos, int32_t* __restrict__ out_img) {
  PDL_ENTRY();
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (r >= n) return;
  const int bk = flag_list[r];
  const int32_t* suf = ids_suffix + static_cast<size_t>(bk) * S;
  int32_t* o = out_ids + static_cast<size_t>(r) * S;
  for (int t = lane; t < S; t += 32) o[t] = suf[t];
  if (lane == 0) { out_eos[r] = eos_idx[bk]; out_img[r] = bk / K; }
}

__global__ void cert_compact_kernel(CertCompact a) {
  PDL_ENTRY();
  const int i = blockIdx.x, b = a.full_list[i], tid = threadIdx.x;
  if (a.P > 0) {
    for (int t = tid; t < a.P; t += blockDim.x) a.c_ids_prefix[static_cast<size_t>(i) * a.P + t] = a.ids_prefix[static_cast<size_t>(b) * a.P + t];
    if (tid == 0) a.c_p0[i] = a.p0[b];
  }
  const size_t ks = static_cast<size_t>(a.K) * a.S;
  for (size_t t = tid; t < ks; t += blockDim.x) a.c_ids_suffix[i * ks + t] = a.ids_suffix[b * ks + t];
  for (int k = tid; k < a.K; k += blockDim.x) {
    const size_t s = static_cast<size_t>(b) * a.K + k, d = static_cast<size_t>(i) * a.K + k;
    a.c_eos_idx[d] = a.eos_idx[s];
    a.c_probs[d] = a.probs[s];
    a.c_ids_masked[d] = a.ids_masked[s];
    if (a.senti) a.c_senti[d] = a.senti[s];
    if (a.repeats) a.c_repeats[d] = a.repeats[s];
  }
  for (int t = tid; t < a.D; t += blockDim.x) a.c_image[static_cast<size_t>(i) * a.D + t] = a.image[static_cast<size_t>(b) * a.D + t];
}

__global__ void cert_scatter_kernel(CertScatter a) {
  PDL_ENTRY();
  const int i = blockIdx.x, b = a.full_list[i], tid = threadIdx.x;
  if (tid == 0) {
    a.inp[static_cast<size_t>(b) * a.L + a.pos] = a.c_inp[static_cast<size_t>(i) * a.L + a.pos];
    a.clip_ref[b] = a.c_clip_ref[i];
    if (a.senti && a.c_senti) a.senti[b] = a.c_senti[i];
    if (a.tr_best && a.c_tr_best) a.tr_best[b] = a.c_tr_best[i];
  }
  for (int k = tid; k < a.K; k += blockDim.x) {
    const size_t s = static_cast<size_t>(i) * a.K + k, d = static_cast<size_t>(b) * a.K + k;
    if (a.tr_score) a.tr_score[d] = a.c_tr_score[s];
    if (a.tr_ref) a.tr_ref[d] = a.c_tr_ref[s];
    if (a.tr_final) a.tr_final[d] = a.c_tr_final[s];
  }
}

}  // namespace

void launch_cert_round1(const CertArgs& a, cudaStream_t st) {
  count_launch();
  ProfScope prof_(CAT_SELECT, 0, st);
  launch_k(cert_round1_kernel, dim3(a.q.B), dim3(CERT_THREADS), cert_smem_bytes(a.q.K), st, a);
}

void launch_cert_round2(const CertArgs& a, cudaStream_t st) {
  count_launch();
  ProfScope prof_(CAT_SELECT, 0, st);
  launch_k(cert_round2_kernel, dim3(a.q.B), dim3(CERT_THREADS), cert_smem_bytes(a.q.K), st, a);
}

void launch_cert_gather_suffix(const int32_t* flag_list, int n, const int32_t* ids_suffix, const int32_t* eos_idx, int K,
                               int S, int32_t* out_ids, int32_t* out_eos, int32_t* out_img, cudaStream_t st) {
  count_launch();
  ProfScope prof_(CAT_ASSEMBLE, 0, st);
  if (n <= 0) return;
  launch_k(cert_gather_suffix_kernel, dim3((n + 7) / 8), dim3(256), 0, st, flag_list, n, ids_suffix, eos_idx, K, S, out_ids,
           out_eos, out_img);
}

void launch_cert_compact(const CertCompact& a, cudaStream_t st) {
  count_launch();
  ProfScope prof_(CAT_ASSEMBLE, 0, st);
  if (a.n <= 0) return;
  launch_k(cert_compact_kernel, dim3(a.n), dim3(256), 0, st, a);
}

void launch_cert_scatter(const CertScatter& a, cudaStream_t st) {
  count_launch();
  ProfScope prof_(CAT_SELECT, 0, st);
  if (a.n <= 0) return;
  launch_k(cert_scatter_kernel, dim3(a.n), dim3(256), 0, st, a);
}

}  // namespace conzic
