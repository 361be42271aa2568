/*
 * conzic.h -- C ABI of libconzic.so: the B200 (sm_100a) Gibbs-BERT caption-polishing step.
 *
 * The reference (joeyz0z/ConZIC) is pure Python and has no FFI; the functions below are what a
 * binding for its hot path would attach to.  Each entry point cites the reference lines it replaces
 * (paths relative to the reference repo; "HF:" = the transformers package the reference calls into).
 *
 * Conventions
 *   - every pointer marked "dev" is a DEVICE pointer owned by the caller (PyTorch tensors in practice);
 *     "host" pointers are ordinary host memory; nothing here allocates or frees caller memory;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no call synchronises the
 *     device or copies device->host -- with ONE exception: in CONZIC_PREC_CERTIFIED mode conzic_gibbs_step and
 *     conzic_score_select read two 64-byte counter blocks back per call (how many candidates need the exact
 *     re-score) and wait for the stream each time;
 *   - scratch memory is one caller-owned device buffer `ws` of at least conzic_workspace_bytes();
 *   - return value: 0 = ok, negative = error; conzic_last_error() gives the message (thread local);
 *   - integer ids are int64 where the reference holds torch.long tensors, int32 for CLIP ids.
 *   - there is NO CPU fallback: on a machine without an sm_100a device every compute call fails.
 */
#ifndef CONZIC_H_
#define CONZIC_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CONZIC_ABI_VERSION 6

typedef struct conzic_ctx conzic_ctx;

/* arithmetic mode of the GEMM operands (accumulation, LayerNorm, softmax, scores are always fp32) */
enum {
  CONZIC_PREC_BF16 = 0,      /* bf16 operands everywhere, one tcgen05.mma per k-step (tolerance-only parity)      */
  CONZIC_PREC_BF16X3 = 1,    /* every fp32 operand split hi+lo bf16, 3 MMAs per k-step: fp32-grade, the reference's
                                token ids (the exact mode the other two are judged against)                        */
  CONZIC_PREC_CERTIFIED = 2  /* the default of the Python layer: BERT and the image tower in bf16x3 (so the top-K set,
                                its probabilities and the image embedding are the exact ones), the CLIP text tower in
                                bf16 over all K candidates, then a certified argmax: every candidate the bf16 scores
                                cannot rule out (error bound cert_dcos) is re-encoded by a bf16x3 copy of the tower and
                                the decision is taken on exact scores -- same winners and same reported cosines as
                                CONZIC_PREC_BF16X3 (gen_utils.py:77-80), at close to bf16 speed                     */
};
/* d = cos(bf16 tower) - cos(bf16x3 tower) of one candidate is taken to lie in [-CERT_DCOS_LO, CERT_DCOS]: 1.3 x the
 * extremes over 3.0 M candidates of the config-2 / config-3 workloads (-5.45e-4 / +1.26e-3; the bf16 tower
 * over-estimates by 3.3e-4 on average; tools/cert_bound.py, profiles/r02b_cert_bound.md, profiles/r02j_cert_bound.md,
 * DESIGN.md section 2) */
#define CONZIC_CERT_DCOS_DEFAULT 1.6e-3f
#define CONZIC_CERT_DCOS_LO_DEFAULT 7.1e-4f
/* R = sum_j exp(exact logit_j) / sum_j exp(bf16 logit_j) over the candidates of an image that are NOT re-scored (the part
 * of the softmax denominator that stays approximate) is taken to lie in [ZRATIO_LO, ZRATIO_HI]; the per-candidate bounds
 * alone only give [exp(-100 hi), exp(100 lo)] = [0.852, 1.067]; measured over 5 120 image-steps R stays within
 * [0.9406, 1.0003] (a sum over ~190 candidates averages the per-candidate errors: profiles/r02j_cert_bound.md); the
 * defaults leave the same kind of margin as the per-candidate bounds (3.3e-4 / 3.0e-4 in mean-cosine terms).
 * Used only with the default error bounds; conzic_config.cert_zratio_* < 0 switches them off. */
#define CONZIC_CERT_ZRATIO_LO_DEFAULT 0.91f
#define CONZIC_CERT_ZRATIO_HI_DEFAULT 1.03f
#define CONZIC_CERT_STATS 8

/* conzic_config.flags: each selects the slower form of a kernel choice, for same-box A/B measurements */
enum {
  CONZIC_FLAG_NO_PDL = 1,         /* launch without programmatic dependent launch                                       */
  CONZIC_FLAG_LN_STANDALONE = 2,  /* CLIP bf16 tower: LayerNorm as its own kernel instead of in the O-proj / fc2
                                     epilogues (HF:models/clip/modeling_clip.py:369-384)                                */
  CONZIC_FLAG_WIDE_LSU = 4,       /* N = 512 GEMM kernel: per-lane epilogue accesses (16 warps) instead of the TMA
                                     reduce-add epilogue                                                                */
  CONZIC_FLAG_LSU_OUT = 16        /* persistent GEMM kernel: bf16 outputs by per-lane stores instead of TMA boxes       */
};

/* GEMM implementation: 0 is the product path; 1 is a slow SIMT kernel kept for cross-checking in tests */
enum { CONZIC_GEMM_TCGEN05 = 0, CONZIC_GEMM_SIMT_DEBUG = 1 };

typedef struct conzic_config {
  /* BERT masked LM (HF:models/bert/configuration_bert.py defaults = bert-base-uncased) */
  int32_t bert_layers, bert_hidden, bert_heads, bert_ffn, bert_vocab, bert_maxpos;
  float bert_ln_eps;
  /* CLIP text tower (HF:models/clip/configuration_clip.py defaults = ViT-B/32 text) */
  int32_t clip_layers, clip_hidden, clip_heads, clip_ffn, clip_vocab, clip_maxpos, clip_proj;
  float clip_ln_eps;
  /* ids that tokenizer.batch_decode(skip_special_tokens=True) drops (gen_utils.py:75) */
  int32_t pad_id, unk_id, cls_id, sep_id, mask_id;
  int32_t dot_id;              /* tokenizer.vocab['.'] (utils.py:53-59) */
  int32_t clip_bos, clip_eos;  /* 49406 / 49407; pad id == eos id (clip/clip.py:71-72) */
  int32_t precision;           /* CONZIC_PREC_* */
  int32_t gemm_impl;           /* CONZIC_GEMM_* */
  int32_t clip_chunk_rows;     /* CLIP token rows processed per pass; 0 = default (303104 = 16 waves of 148 x 128-row tiles) */
  float cert_dcos;             /* CERTIFIED: upper bound of cos(bf16) - cos(exact); 0 = CONZIC_CERT_DCOS_DEFAULT */
  float cert_dcos_lo;          /* CERTIFIED: -(lower bound); 0 = cert_dcos when that is given, else the default */
  float cert_zratio_lo;        /* CERTIFIED: bounds on the denominator ratio R (see CONZIC_CERT_ZRATIO_*); 0 = default when the */
  float cert_zratio_hi;        /*   error bounds are the defaults too, < 0 = only what the error bounds imply */
  int32_t cert_fcap;           /* CERTIFIED: an image with more unbeaten candidates than this is re-encoded in full; 0 = 64 */
  int32_t flags;               /* CONZIC_FLAG_* */
} conzic_config;

/* ---- weight tables: arrays of fp32 DEVICE pointers in this order (HF state-dict tensors) -------------
 * BERT (n = 10 + 16*layers):
 *   0 word_embeddings[V,H] 1 position_embeddings[P,H] 2 token_type_embeddings[2,H] 3 emb LN gamma 4 emb LN beta
 *   5 cls.predictions.bias[V] 6 cls.predictions.transform.dense.weight[H,H] 7 .bias 8 transform LN gamma 9 beta
 *   per layer l at 10+16*l: q.w q.b k.w k.b v.w v.b attn.out.w attn.out.b attnLN.g attnLN.b
 *                           intermediate.w[F,H] intermediate.b output.w[H,F] output.b outLN.g outLN.b
 * CLIP text (n = 5 + 16*layers):
 *   0 token_embedding[V,H] 1 position_embedding[77,H] 2 final_layer_norm gamma 3 beta 4 text_projection.weight[P,H]
 *   per layer l at 5+16*l: ln1.g ln1.b q.w q.b k.w k.b v.w v.b out.w out.b ln2.g ln2.b fc1.w[F,H] fc1.b fc2.w[H,F] fc2.b
 * The context makes its own bf16 copies; the caller may free the fp32 tensors after create returns and
 * the stream has drained.
 */
#define CONZIC_BERT_GLOBALS 10
#define CONZIC_CLIP_GLOBALS 5
#define CONZIC_PER_LAYER 16

/* Replaces model/clip construction + .to(device) (run.py:134-141, clip/clip.py:7-18) for this path. */
int conzic_ctx_create(const conzic_config* cfg, const void* const* bert_weights_dev, int n_bert,
                      const void* const* clip_weights_dev, int n_clip, void* stream, conzic_ctx** out);
void conzic_ctx_destroy(conzic_ctx* ctx);
const char* conzic_last_error(void);
int conzic_abi_version(void);

/* BERT id -> CLIP BPE ids, CSR (off[V+1], tok[off[V]]), int32 DEVICE arrays copied into the context.
 * Device-side replacement of tokenizer.batch_decode + CLIPTokenizer for vocabularies without '##'
 * merges (gen_utils.py:75, clip/clip.py:71-72).  Special ids must have empty rows. */
int conzic_set_bert2clip(conzic_ctx* ctx, const int32_t* off_dev, const int32_t* tok_dev, int n_tok,
                         int max_tok_per_word, void* stream);

/* Device text pipeline for vocabularies with '##' word pieces -- every real BERT vocabulary (replaces
 * tokenizer.batch_decode + CLIPTokenizer, gen_utils.py:75 + clip/clip.py:71-72, for the Hugging Face fast
 * BertTokenizer / CLIPTokenizer pair; csrc/text_pipeline.cuh states the algorithm).  All pointers are DEVICE arrays,
 * copied into the context.  conzic_set_bert2clip must have been called: whole-word tokens keep their table rows.
 * With a text vocabulary set, conzic_gibbs_step builds every candidate's CLIP ids from the candidate caption's
 * bytes (a piece merges into its neighbour word and changes that word's BPE) and reads the two row capacities of
 * the step back (16 bytes, one stream synchronisation per step). */
typedef struct conzic_text_vocab {
  const int32_t* tok_off;      /* [V+1] byte offsets of each BERT token's text ('##' stripped, NFC, lowercase) */
  const uint8_t* tok_bytes;    /* UTF-8 */
  const uint8_t* tok_cls;      /* per byte: 0 other / 1 letter / 2 number / 3 space, | 4 on a character's first byte */
  const uint8_t* tok_flags;    /* [V]: 1 '##' piece, 2 no space before it (WordPiece clean-up), 4 dropped, 8 simple */
  const int32_t* byte_sym;     /* [512] CLIP id of byte b inside a word / [256+b] as the last byte of a word */
  const uint64_t* merge_keys;  /* open-addressing table of 1 << merge_bits slots: (id_a << 32 | id_b) + 1, 0 = empty */
  const uint32_t* merge_vals;  /* rank << 16 | merged id */
  int32_t n_bytes, merge_bits;
} conzic_text_vocab;
int conzic_set_text_vocab(conzic_ctx* ctx, const conzic_text_vocab* v, void* stream);

/* Scratch bytes needed by any call below with at most B images, L BERT tokens, K candidates. */
size_t conzic_workspace_bytes(const conzic_ctx* ctx, int B, int L, int K);

/* model(inp).logits[:, pos] (gen_utils.py:69 + :42; HF:models/bert/modeling_bert.py:944-987) --
 * the MLM head is evaluated on row `pos` only.  inp int64[B,L] dev (already holding [MASK] where the
 * caller wants it); logits f32[B,ldl] dev, ldl >= V and ldl % 4 == 0. */
int conzic_bert_mlm_row(conzic_ctx* ctx, const int64_t* inp_dev, int B, int L, int pos, float* logits_dev,
                        int ldl, void* ws_dev, size_t ws_bytes, void* stream);

/* generate_caption_step (gen_utils.py:33-49): softmax(logits/temperature) * token_mask, top-K.
 * Ties (equal probabilities, e.g. underflowed zeros) are ordered by ascending vocabulary index.
 * probs f32[B,K], ids int64[B,K], sorted by descending probability.  K <= 1024. */
int conzic_topk_mask(conzic_ctx* ctx, const float* logits_dev, int ldl, int B, const float* token_mask_dev,
                     float temperature, int K, float* probs_dev, int64_t* ids_dev, void* stream);

/* Candidate captions as CLIP ids (gen_utils.py:71-75 + clip/clip.py:71-77), dense form:
 * clip_ids int32[B*K,T] (BOS .. EOS, right padded with EOS, truncated to 77), clip_len int32[B*K]
 * (tokens up to and including the first EOS), ids_masked int64[B,K] = ids * token_mask[ids]. */
int conzic_build_clip_ids(conzic_ctx* ctx, const int64_t* inp_dev, int B, int L, int pos, const int64_t* ids_dev,
                          const float* token_mask_dev, int K, int32_t* clip_ids_dev, int T, int32_t* clip_len_dev,
                          int64_t* ids_masked_dev, void* stream);

/* CLIP.compute_text_representation after tokenisation (clip/clip.py:78-83;
 * HF:models/clip/modeling_clip.py:531-589): clip_ids int32[N,T] right padded with EOS ->
 * text_embeds f32[N,proj]. */
int conzic_clip_text_encode(conzic_ctx* ctx, const int32_t* clip_ids_dev, int N, int T, float* text_embeds_dev,
                            void* ws_dev, size_t ws_bytes, void* stream);

/* compute_image_text_similarity_via_embeddings (clip/clip.py:86-98): text f32[B*K,D], image f32[B,D]
 * -> clip_score f32[B,K] (softmax over K of scale*cos) and clip_ref f32[B,K] (cos). */
int conzic_image_text_similarity(conzic_ctx* ctx, const float* text_embeds_dev, const float* image_embeds_dev,
                                 int B, int K, float logit_scale_exp, float* clip_score_dev, float* clip_ref_dev,
                                 void* stream);

/* Score fuse + argmax + write-back on its own (gen_utils.py:77-81, control_gen_utils.py:59-65), for callers that
 * build the candidate texts themselves (host string path: POS-template control needs every caption as a string):
 *   final = alpha*probs + beta*softmax_K(scale*cos) (+ gamma*softmax_K(senti_raw) + 0.1*(1-exp(repeats)));
 *   inp[b,pos] = ids_masked[b, argmax]; out_clip_ref[b] = cos of the winner; out_senti[b] = senti_raw of the winner;
 *   out_best[b] = argmax (int64, may be NULL) -- the POS-template path needs it to pick the winner's tag
 *   sequence and raw score on the host (control_gen_utils.py:171-178).
 * A control term with its own softmax temperature t (POS: 0.1, control_gen_utils.py:166) passes senti_raw = raw / t
 * and repeats = NULL (no repeat penalty in that formula).
 * text f32[B*K,D] (from conzic_clip_text_encode), image f32[B,D], probs f32[B,K], ids_masked int64[B,K],
 * senti_raw / repeats f32[B,K] or NULL.  clip_ids int32[B*K,T] = the rows text was encoded from: required in
 * CERTIFIED mode (unbeaten candidates are re-encoded from them by the exact tower), ignored otherwise.
 * ws: conzic_workspace_bytes(ctx, B, L, K). */
int conzic_score_select(conzic_ctx* ctx, const float* text_embeds_dev, const float* image_embeds_dev,
                        const int32_t* clip_ids_dev, int T, int B, int K, float logit_scale_exp,
                        const float* probs_dev, const int64_t* ids_masked_dev, const float* senti_raw_dev,
                        const float* repeats_dev, float alpha, float beta, float gamma, int64_t* inp_dev, int L,
                        int pos, float* out_clip_ref_dev, float* out_senti_dev, int64_t* out_best_dev, void* ws_dev,
                        size_t ws_bytes, void* stream);

/* One whole Gibbs step (gen_utils.py:66-81, control_gen_utils.py:45-67) with no host round trip:
 *   token_mask[dot_id] = dot_allowed; inp[:,pos] = [MASK]; BERT row logits; top-K; candidates -> CLIP ids
 *   (shared caption prefix encoded once per image, candidate suffixes per candidate); CLIP text tower;
 *   cosine / softmax; alpha*p + beta*c (+ gamma*softmax_K(senti) + 0.1*(1-exp(repeats))); argmax;
 *   inp[:,pos] = winner.
 * In/out: inp int64[B,L] dev, token_mask f32[V] dev (mutated like utils.py:53-59).
 * In: image_embeds f32[B,proj] dev; senti_table f32[V] dev or NULL (NULL = caption mode; the table is
 *     the per-word control score, sign already applied for "negative"); visited_after = upper bound on the
 *     number of caption positions > pos that hold a word (host bookkeeping, sizes the suffix tile);
 *     visited_before likewise for positions < pos (prompt words included).
 * Out: out_clip_ref f32[B] (cosine of the winner, gen_utils.py:80), out_senti f32[B] or NULL.
 * Optional trace (any may be NULL): tr_probs f32[B,K], tr_ids int64[B,K], tr_clip_score f32[B,K],
 *     tr_clip_ref f32[B,K], tr_final f32[B,K], tr_best int64[B]. */
typedef struct conzic_step_args {
  int64_t* inp;
  float* token_mask;
  const float* image_embeds;
  const float* senti_table;
  int32_t B, L, K, pos;
  int32_t dot_allowed;
  int32_t visited_before, visited_after;
  float temperature, alpha, beta, gamma, logit_scale_exp;
  float* out_clip_ref;
  float* out_senti;
  float* tr_probs;
  int64_t* tr_ids;
  float* tr_clip_score;
  float* tr_clip_ref;
  float* tr_final;
  int64_t* tr_best;
  float* tr_logits;   /* f32[B, ldl = roundup(V,4)] or NULL */
  /* Span order (gen_utils.py:148-195): logits of row `pos` from an EARLIER forward, f32[B, ldl] from
   * conzic_bert_mlm_row; when non-NULL the step skips its own BERT forward and scores these. */
  const float* logits_in;
} conzic_step_args;

int conzic_gibbs_step(conzic_ctx* ctx, const conzic_step_args* args, void* ws_dev, size_t ws_bytes, void* stream);

/* Counters: kernels launched for this context since it was created (bench.py "gpu_launches"). */
uint64_t conzic_launch_count(const conzic_ctx* ctx);

/* CERTIFIED mode bookkeeping since the context was created, out[0..n): 0 selection calls, 1 images decided,
 * 2 candidates re-encoded exactly (one per image at least: the winner's reported cosine is always the exact one),
 * 3 images with more than one unbeaten candidate after the bf16 pass, 4 images sent to the full exact re-encode
 * because more than cert_fcap candidates were unbeaten, 5 images fully re-encoded in total (4 + those the exact
 * re-score of the unbeaten candidates still could not order).  Zeros in the other modes. */
int conzic_cert_stats(const conzic_ctx* ctx, uint64_t* out, int n);

/* Optional device timing by kernel category (CUDA events around each launch, on the launching stream).
 * Categories: 0 persistent pair GEMM (CLIP / vision towers; work = executed FLOPs), 1 attention, 2 LayerNorm,
 * 3 embeddings, 4 top-k, 5 CLIP-id assembly, 6 score/select, 7 misc, 8 gridded GEMM (BERT, bf16x3 mode).  conzic_profile(ctx, 1) clears and enables, (ctx, 0) disables;
 * conzic_profile_read waits for the recorded events and returns summed milliseconds, work and launches.
 * Categories 100 + p sum the same records by step phase p instead: 0 other, 1 BERT, 2 top-k + candidate assembly,
 * 3 CLIP text tower over all candidates, 4 logits + selection kernels, 5 certified mode: exact re-score of the listed
 * candidates, 6 certified mode: full exact re-encode of undecided images, 7 image tower. */
int conzic_profile(conzic_ctx* ctx, int enable);
int conzic_profile_read(conzic_ctx* ctx, int category, double* ms, double* work, int* launches);

/* Plain GEMM entry used by tests and the roofline microbench:
 *   out[M,N] = act(A[M,K] * W[N,K]^T + bias) (+ resid), A/W fp32 dev, converted internally to the
 *   context's operand format; exercises exactly the kernel the towers use.  act: 0 none, 1 quick_gelu,
 *   2 erf-gelu; act | 16 routes the result through the kernel's bf16 activation output (bf16 operands, no
 *   residual) before it is widened into out; act | 32 (CERTIFIED contexts) uses the exact tower's operand format and
 *   kernel instead of the bf16 one; act | 64 forces the gridded 128 x 128 kernel, act | 128 the CTA-pair kernel (bf16x3 operands).  out fp32[M,N]. */
int conzic_debug_linear(conzic_ctx* ctx, const float* A_dev, const float* W_dev, const float* bias_dev,
                        const float* resid_dev, int M, int N, int K, int act, float* out_dev, void* ws_dev,
                        size_t ws_bytes, void* stream);

/* ---- CLIP image tower (clip/clip.py:48-62; HF:models/clip/modeling_clip.py:676-686): once per call, before the
 * Gibbs loop.  Weight table (fp32 device pointers, n = 8 + 16*layers):
 *   0 patch_embedding.weight[H,3,p,p] 1 class_embedding[H] 2 position_embedding[T,H] 3 pre_layrnorm g 4 b
 *   5 post_layernorm g 6 b 7 visual_projection.weight[proj,H]
 *   per layer l at 8+16*l: same 16 entries as the text tower (ln1.g ln1.b q.w q.b k.w k.b v.w v.b out.w out.b
 *   ln2.g ln2.b fc1.w fc1.b fc2.w fc2.b). */
typedef struct conzic_vision_config {
  int32_t layers, hidden, heads, ffn, image_size, patch, proj;
  float ln_eps;
} conzic_vision_config;
int conzic_set_vision(conzic_ctx* ctx, const conzic_vision_config* vc, const void* const* weights_dev, int n,
                      void* stream);
size_t conzic_vision_workspace_bytes(const conzic_ctx* ctx, int B);
/* pixel_values f32[B,3,S,S] dev (already resized / normalised by the caller's processor) -> image_embeds
 * f32[B,proj] dev (not L2-normalised, like compute_image_representation_from_image_instance). */
int conzic_clip_image_encode(conzic_ctx* ctx, const float* pixel_values_dev, int B, float* image_embeds_dev,
                             void* ws_dev, size_t ws_bytes, void* stream);

/* ---- CLIPImageProcessor on the device (clip/clip.py:55-58 -> HF CLIPImageProcessor, torchvision backend): n uint8
 * HWC images of ONE size -> pixel_values f32[n, 3, S, S].  Antialiased bicubic resize of the shortest edge on uint8
 * in ATen's fixed-point arithmetic (horizontal pass, vertical pass, a rounded uint8 after each), centre crop, fused
 * (x - mean) / std with mean / std in 0..255 units.  The caller supplies, per axis and for the output positions inside
 * the crop only, the integer filter taps ATen computes (conzic_b200/imageproc.py restates that computation; identity
 * != 0: the axis is not resized, `first` just crops).  All pointers are DEVICE pointers.
 * ws: n * (row_hi - row_lo) * horiz.n_out * 3 bytes. */
typedef struct conzic_resize_axis {
  const int16_t* weights;  /* [n_out, taps] */
  const int32_t* first;    /* [n_out] first input index read by each output position */
  const int32_t* count;    /* [n_out] taps used */
  int32_t taps, precision, n_out, identity;
} conzic_resize_axis;
int conzic_image_preprocess(conzic_ctx* ctx, const uint8_t* images_dev, int n, int H, int W,
                            const conzic_resize_axis* horiz, const conzic_resize_axis* vert, int row_lo, int row_hi,
                            const float* mean3_host, const float* std3_host, float* pixel_values_dev, void* ws_dev,
                            size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CONZIC_H_ */
