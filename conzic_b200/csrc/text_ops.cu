// Candidate captions -> CLIP ids for vocabularies with '##' word pieces (every real BERT vocabulary), on the device:
// the reference's per-step string round trip (gen_utils.py:75 tokenizer.batch_decode, clip/clip.py:71-72
// CLIPTokenizer) as two kernels around text_pipeline.cuh.
//   text_tokenize_kernel  one CTA per image, one thread per candidate: the whole candidate caption through
//                         WordPiece decode + CLIP pre-tokenise + byte-level BPE; BOS / EOS framing; then the longest
//                         prefix shared by all K sequences of the image (a piece candidate merges into the word
//                         before it and a piece after `pos` into the candidate, so the split point is found from
//                         the sequences themselves instead of from token positions);
//   text_layout_kernel    the shared-prefix / per-candidate-suffix row layout the CLIP tower consumes, for row
//                         capacities (P, S) the host picked after reading back the two maxima kernel 1 reports.
#include "kernels.h"
#include "select_common.cuh"
#include "text_pipeline.cuh"

namespace conzic {

namespace {

__global__ void __launch_bounds__(256) text_tokenize_kernel(TextAssembleArgs a) {
  PDL_ENTRY();
  __shared__ int s_min_lcp, s_min_len, s_max_len, s_err;
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) { s_min_lcp = 1 << 20; s_min_len = 1 << 20; s_max_len = 0; s_err = 0; }
  __syncthreads();
  const int64_t* row = a.inp + static_cast<size_t>(b) * a.L;
  const int T = a.maxlen;
  for (int k = tid; k < a.K; k += blockDim.x) {
    const size_t bk = static_cast<size_t>(b) * a.K + k;
    const int64_t c = a.ids[bk];
    const float m = a.token_mask[c];
    const int64_t cm = static_cast<int64_t>(static_cast<float>(c) * m);  // idxs * token_mask[0][idxs] (gen_utils.py:72)
    a.ids_masked[bk] = cm;
    int32_t* seq = a.seq + bk * T;
    int err = 0;
    seq[0] = a.bos;
    const int nb = txt_caption_to_clip(a.vocab, row, a.L, a.pos, cm, a.special, seq + 1, T - 2, &err);
    seq[1 + nb] = a.eos;
    // the tower pools at the FIRST id equal to EOS (HF:models/clip/modeling_clip.py:564-584) and is causal, so nothing
    // after it matters: a vocabulary whose unknown-symbol id is the EOS id (the CLIP default) ends the caption there
    int len = 2;
    while (seq[len - 1] != a.eos) ++len;
    a.len[bk] = len;
    if (err) atomicOr(&s_err, err);
    atomicMin(&s_min_len, len);
    atomicMax(&s_max_len, len);
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
        if (txt_is_special(a.special, id) || id < 0 || id >= a.vocab.V) continue;
        s += static_cast<double>(a.senti_table[id]);
      }
      a.senti[bk] = static_cast<float>(s);
    }
  }
  __syncthreads();
  // longest common prefix of the K sequences = min over k of LCP(seq_k, seq_0)
  const int32_t* s0 = a.seq + static_cast<size_t>(b) * a.K * T;
  const int len0 = a.len[static_cast<size_t>(b) * a.K];
  for (int k = tid; k < a.K; k += blockDim.x) {
    const int32_t* sk = s0 + static_cast<size_t>(k) * T;
    const int n = min(len0, a.len[static_cast<size_t>(b) * a.K + k]);
    int l = 0;
    while (l < n && sk[l] == s0[l]) ++l;
    atomicMin(&s_min_lcp, l);
  }
  __syncthreads();
  if (tid == 0) {
    // every candidate keeps at least its EOS as its own row; the prefix holds at least BOS
    int p0 = min(s_min_lcp, s_min_len - 1);
    if (p0 < 1) p0 = 1;
    a.p0[b] = p0;
    atomicMax(&a.dims[0], p0);
    atomicMax(&a.dims[1], s_max_len - p0);
    if (s_err) atomicOr(&a.dims[2], s_err);
  }
}

__global__ void __launch_bounds__(256) text_layout_kernel(TextAssembleArgs a) {
  PDL_ENTRY();
  const int b = blockIdx.x, tid = threadIdx.x;
  const int T = a.maxlen;
  int p0 = a.p0[b];
  if (p0 > a.P) {  // a shorter shared prefix is still a shared prefix
    p0 = a.P;
    __syncthreads();
    if (tid == 0) a.p0[b] = p0;
  }
  const int32_t* s0 = a.seq + static_cast<size_t>(b) * a.K * T;
  for (int t = tid; t < a.P; t += blockDim.x) a.ids_prefix[static_cast<size_t>(b) * a.P + t] = t < p0 ? s0[t] : a.eos;
  for (int k = tid; k < a.K; k += blockDim.x) {
    const size_t bk = static_cast<size_t>(b) * a.K + k;
    const int32_t* sk = s0 + static_cast<size_t>(k) * T;
    int ns = a.len[bk] - p0;
    if (ns > a.S) { ns = a.S; atomicOr(&a.dims[2], TXT_ERR_ROWS); }
    int32_t* o = a.ids_suffix + bk * a.S;
    for (int t = 0; t < a.S; ++t) o[t] = t < ns ? sk[p0 + t] : a.eos;
    a.eos_idx[bk] = ns - 1;
  }
}

}  // namespace

void launch_text_tokenize(const TextAssembleArgs& a, cudaStream_t st) {
  count_launch();
  ProfScope prof_(CAT_ASSEMBLE, 0, st);
  launch_k(text_tokenize_kernel, dim3(a.B), dim3(256), 0, st, a);
}

void launch_text_layout(const TextAssembleArgs& a, cudaStream_t st) {
  count_launch();
  ProfScope prof_(CAT_ASSEMBLE, 0, st);
  launch_k(text_layout_kernel, dim3(a.B), dim3(256), 0, st, a);
}

}  // namespace conzic
