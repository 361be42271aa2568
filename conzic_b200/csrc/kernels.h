// Internal interfaces between the translation units of libconzic.so (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

namespace conzic {

typedef __nv_bfloat16 bf16;

// Per-context launch state.  Every extern "C" entry point binds the calling thread to its context's state for the
// duration of the call (StateScope in engine.cu), so two contexts -- or two threads with a context each -- do not
// share counters, the PDL switch or profiling records.
struct ProfRec { cudaEvent_t a, b; int cat; int phase; double work; };
struct CtxState {
  uint64_t launches = 0;  // kernels launched for this context (bench.py "gpu_launches")
  int pdl = 1;            // launches may carry the programmatic-dependent-launch attribute (conzic_config.no_pdl clears it)
  int pdl_now = 1;        // set by the engine per section: measured on B200, PDL gains ~3.5 % on steps made of short
                          // kernels (BERT, CLIP passes under ~50 k rows) and costs 1-2 % when the kernels are long
  bool prof_on = false;
  int phase = 0;          // which part of a step the engine is in (PHASE_*), recorded with every profiled launch
  std::vector<ProfRec> prof;
};
enum { PHASE_OTHER = 0, PHASE_BERT = 1, PHASE_CANDIDATES = 2, PHASE_TOWER = 3, PHASE_SELECT = 4, PHASE_CERT_RESCORE = 5,
       PHASE_CERT_FULL = 6, PHASE_IMAGE = 7, PHASE_COUNT = 8 };
extern thread_local CtxState* t_state;
inline void count_launch() { if (t_state) ++t_state->launches; }
inline bool pdl_enabled() { return t_state && t_state->pdl && t_state->pdl_now; }
inline void set_pdl_now(int v) { if (t_state) t_state->pdl_now = v; }
inline void set_phase(int p) { if (t_state) t_state->phase = p; }

// Programmatic dependent launch: every hot-path kernel starts with PDL_ENTRY() -- it lets the NEXT kernel in the
// stream begin launching right away (griddepcontrol.launch_dependents) and then waits until everything the
// PREVIOUS kernels wrote is complete and visible (griddepcontrol.wait), before touching any global memory.
// The launch latency and block scheduling of kernel N+1 thereby overlap the execution of kernel N; a step is
// ~180 dependent launches, many of them 5-20 us long.
#define PDL_ENTRY()                                                   \
  do {                                                                \
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   \
    asm volatile("griddepcontrol.wait;" ::: "memory");                \
  } while (0)

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}
void set_error(const std::string& msg);
bool cuda_ok(cudaError_t e, const char* what);

// Optional per-category device timing (CUDA events around each launch on the launching stream); off by default.
enum { CAT_GEMM = 0, CAT_ATTN = 1, CAT_LN = 2, CAT_EMBED = 3, CAT_TOPK = 4, CAT_ASSEMBLE = 5, CAT_SELECT = 6,
       CAT_MISC = 7, CAT_GEMM_SMALL = 8, CAT_COUNT = 9 };
struct ProfScope {
  int cat; cudaStream_t st; bool on;
  ProfScope(int cat, double work, cudaStream_t st);
  ~ProfScope();
};
void prof_enable(CtxState* s, bool on);
bool prof_read(CtxState* s, int cat, double* ms, double* work, int* n);

// A GEMM-input activation matrix.  bf16 mode: [rows, K].  bf16x3 mode: [rows, 2K], the bf16 "hi" plane in
// columns [0,K) and the residual "lo" plane (x - float(hi)) in [K,2K).
struct Act {
  bf16* p;
  int ld;  // elements per row: K or 2K
  int K;
};

// A linear layer's weight in operand format ([N,K] or [N,2K] like Act) plus its TMA descriptors.
struct LinearW {
  bf16* w = nullptr;
  const float* bias = nullptr;
  int N = 0, K = 0;
  CUtensorMap tmap128;  // box 128 rows x 64 cols, 128B swizzle
  CUtensorMap tmap256;  // box 256 rows x 64 cols
  CUtensorMap tmap64;   // box 64 rows x 64 cols (narrow-N bf16x3 GEMMs on a few thousand rows)
};

enum { ACT_NONE = 0, ACT_QUICK_GELU = 1, ACT_ERF_GELU = 2 };

// Epilogue: v = acc + bias[n] ; v = act(v) ; v += resid[m,n] ; then written as fp32 and/or as an Act.
struct Epi {
  const float* bias = nullptr;
  const float* resid = nullptr;
  int ldr = 0;
  float* out_f32 = nullptr;
  int ldo_f32 = 0;
  bf16* out_act = nullptr;
  int ldo_act = 0;
  int out_K = 0;  // column offset of the lo plane when writing a split Act
  int act = ACT_NONE;
  // LayerNorm of the rows this GEMM writes, applied by the SAME kernel (gemm_wide_kernel only: N == 512 == H, a
  // CTA owns whole rows): after x = acc + bias + resid has been stored, lnf_out[m, :] = bf16(LN(x[m, :]) * g + b).
  // Replaces the stand-alone LayerNorm launch that would re-read x from HBM (HF:models/clip/modeling_clip.py:369-384).
  bf16* lnf_out = nullptr;
  int lnf_ld = 0;
  const float* lnf_g = nullptr;
  const float* lnf_b = nullptr;
  float lnf_eps = 0.f;
};

struct GemmOpts {
  int split = 0;    // 1 = bf16x3
  int impl = 0;     // 0 tcgen05, 1 SIMT debug
  int bn = 128;     // N tile: 128 or 256
  int stages = 3;   // smem pipeline depth
  int persist = 0;  // 1 = persistent A-resident kernel with double-buffered TMEM accumulators (bf16 mode only)
  int cg = 1;       // persistent kernel: 2 = CTA pairs (tcgen05 cta_group::2), 1 = single CTAs
  int ksplit = 1;   // gridded kernel: split-K factor (raw fp32 partials, summed by the following LayerNorm)
  int force_pair = 0;  // bf16x3 + persist: the pair kernel whatever the tile count (tests)
  int lsu_out = 0;       // persistent kernel: bf16 outputs by per-lane stores instead of TMA boxes (A/B)
  int wide_lsu = 0;      // N = 512 wide kernel: per-lane epilogue accesses (16 warps) instead of the TMA reduce epilogue (A/B)
};

bool tma_init();  // resolves cuTensorMapEncodeTiled through the runtime (no link-time libcuda dependency)
bool make_tmap_bf16_2d(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                       uint32_t box_rows);
bool gemm_configure();  // opt-in to large dynamic shared memory for every instantiation
bool launch_linear(const Act& A, int M, const LinearW& W, const Epi& epi, const GemmOpts& o, cudaStream_t st);

// ---- transformer pieces (transformer_ops.cu) ---------------------------------------------------------
void launch_f32_to_act(const float* src, int rows, int K, int lds, bf16* dst, int ldd, int split, cudaStream_t st);

struct LNArgs {
  const float* x;      // [*, H] fp32
  const int32_t* rows; // optional gather: output row r reads input row rows[r]; null = identity
  int n_rows, H;
  const float* gamma;
  const float* beta;
  float eps;
  float* out_f32;      // optional [n_rows, H]
  bf16* out_act;       // optional Act [n_rows, ld]
  int ld_act, split;
  size_t x_row_stride = 0;  // elements between input rows; 0 = H (dense)
  // split-K consumer: the row to normalise is x + add_bias + sum_z partials[z * part_stride + row * H ...]
  const float* partials = nullptr;
  int n_parts = 0;
  size_t part_stride = 0;
  const float* add_bias = nullptr;
};
void launch_layernorm(const LNArgs& a, cudaStream_t st);

// Image patches as GEMM rows: pix f32[B,3,S,S] -> Act [B*g*g, 3*p*p] (channel, py, px order = conv weight layout)
void launch_im2col(const float* pix, int B, int S, int p, bf16* dst, int ldd, int split, cudaStream_t st);
// x[b,0,:] = cls + pos[0]; x[b,1+i,:] = patch[b*g2+i,:] + pos[1+i]    (HF:models/clip/modeling_clip.py:190-200)
void launch_vision_embed(const float* patch, const float* cls, const float* pos, int B, int T, int H, float* x,
                         cudaStream_t st);
void launch_gather_rows(const void* src, size_t row_bytes, const int32_t* rows, int n, void* dst, cudaStream_t st);

void launch_bert_embed_ln(const int64_t* inp, int rows, int L, const float* word, const float* pos, const float* type,
                          const float* gamma, const float* beta, float eps, int H, float* x_f32, bf16* act, int ld_act,
                          int split, cudaStream_t st);
// ln_out optional (H == 512, bf16 operands): LayerNorm 1 of the first block written by the same kernel.
// cand_img / n_cand: see AttnArgs (null / 0 = K candidates per image)
void launch_clip_embed(const int32_t* ids_prefix, const int32_t* ids_suffix, const int32_t* p0, int B, int P, int K,
                       int S, int maxpos, const float* tok, const float* pos, int H, float* x_f32, cudaStream_t st,
                       const float* ln_g = nullptr, const float* ln_b = nullptr, float ln_eps = 0.f,
                       bf16* ln_out = nullptr, int ln_ld = 0, const int32_t* cand_img = nullptr, int n_cand = 0);

// Attention over packed token rows.  Row layout: B*P "prefix" rows (image b, position t) followed by
// B*K*S "suffix" rows (image b, candidate k, offset s).  A suffix row attends to its image's first p0[b]
// prefix rows and to suffix rows 0..s of its own candidate (causal) -- or to all S rows when !causal.
// qkv is either bf16 [rows, 3H] (split=0) or fp32 [rows, 3H] (split=1).
struct AttnArgs {
  const void* qkv;
  int ld_qkv;      // elements per row
  int qkv_f32;     // 1 = fp32 qkv
  const int32_t* p0;  // [B] or null (=> P rows all valid for prefix rows; suffix sees min(P, .))
  int B, P, K, S, H, heads;
  int causal;
  float scale;
  bf16* out_act;
  int ld_act, split;
  // optional: the suffix blocks are n_cand candidates of arbitrary images, candidate i belonging to image cand_img[i]
  // (certified re-score: a few candidates per image); null = K candidates per image in image order
  const int32_t* cand_img = nullptr;
  int n_cand = 0;
  int cpt = 1;            // filled by launch_attention: candidates packed into one 16-row query tile
  int cand_per_task = 8;  // filled by launch_attention: candidates per warp task
};
bool attention_configure();  // opt-in to large dynamic shared memory for every instantiation (current device)
bool launch_attention(const AttnArgs& a, cudaStream_t st);

// ---- image pre-processing (image_ops.cu) ---------------------------------------------------------------
struct ResizeAxis {
  const int16_t* w;      // [n_out, taps] fixed-point filter taps
  const int32_t* first;  // [n_out] first input index
  const int32_t* count;  // [n_out] taps used
  int taps, precision, n_out, identity;
};
struct ImagePreArgs {
  const uint8_t* src;  // [n, H, W, 3]
  int n, H, W;
  ResizeAxis hz, vt;
  int row_lo, row_hi;  // input rows the vertical pass reads
  float mean[3], std[3];
  uint8_t* tmp;        // [n, row_hi - row_lo, hz.n_out, 3]
  float* out;          // [n, 3, vt.n_out, hz.n_out]
};
void launch_image_preprocess(const ImagePreArgs& a, cudaStream_t st);

// ---- selection pieces (select_ops.cu) -----------------------------------------------------------------
bool topk_configure();
bool launch_topk(const float* logits, int ldl, int B, int V, const float* mask, float temperature, int K, float* probs,
                 int64_t* ids, cudaStream_t st);

struct AssembleArgs {
  const int64_t* inp;       // [B,L], [MASK] at pos
  const int64_t* ids;       // [B,K] top-k ids
  const float* token_mask;  // [V]
  const int32_t* off;       // CSR bert id -> clip tokens
  const int32_t* tok;
  const float* senti_table; // [V] or null
  int B, L, K, pos, V;
  int special[5];
  int bos, eos, maxlen;     // maxlen = 77
  // outputs
  int32_t* ids_prefix;  // [B,P] or null when P == 0 (dense mode: everything goes to the suffix)
  int32_t* ids_suffix;  // [B,K,S]
  int32_t* p0;          // [B]
  int32_t* eos_idx;     // [B*K] suffix-relative index of the first EOS
  int P, S;
  int64_t* ids_masked;  // [B,K]
  float* repeats;       // [B,K] or null
  float* senti;         // [B,K] or null
};
void launch_assemble(const AssembleArgs& a, cudaStream_t st);

}  // namespace conzic
#include "text_pipeline.cuh"
namespace conzic {
// Device text pipeline (text_ops.cu): the candidates of one Gibbs step through WordPiece decode + CLIP BPE
struct TextAssembleArgs {
  TextVocab vocab;
  const int64_t* inp;       // [B,L], [MASK] at pos
  const int64_t* ids;       // [B,K] top-k ids
  const float* token_mask;  // [V]
  const float* senti_table; // [V] or null
  int B, L, K, pos;
  int special[5];
  int bos, eos, maxlen;     // maxlen = 77
  // kernel 1 outputs
  int32_t* seq;             // [B*K, maxlen] BOS ... EOS of every candidate caption
  int32_t* len;             // [B*K] tokens including BOS and EOS
  int32_t* p0;              // [B] rows of the image's shared prefix (>= 1: BOS)
  int32_t* dims;            // [0] max p0, [1] max suffix rows, [2] TXT_ERR_* bits (all atomically merged; zeroed by the caller)
  int64_t* ids_masked;      // [B,K]
  float* repeats;           // [B,K] or null
  float* senti;             // [B,K] or null
  // kernel 2: the tower's row layout for capacities P, S
  int P, S;
  int32_t* ids_prefix;      // [B,P]
  int32_t* ids_suffix;      // [B,K,S]
  int32_t* eos_idx;         // [B*K] suffix-relative index of the EOS
};
void launch_text_tokenize(const TextAssembleArgs& a, cudaStream_t st);
void launch_text_layout(const TextAssembleArgs& a, cudaStream_t st);

void launch_step_prologue(int64_t* inp, int B, int L, int pos, int mask_id, float* token_mask, int dot_id,
                          int dot_allowed, cudaStream_t st);
// rows[i] = n_pre_rows + i * S + eos_idx[i] for the n_cand candidate blocks that follow the prefix rows
void launch_pool_index(int32_t* rows, const int32_t* eos_idx, int n_pre_rows, int n_cand, int S, cudaStream_t st);

struct SelectArgs {
  const float* logit;  // [B,K] scale * cos(text, image) from launch_clip_logits
  int B, K;
  float scale;
  const float* probs;        // [B,K] or null (similarity only)
  const int64_t* ids_masked; // [B,K]
  const float* senti;        // [B,K] raw control scores or null
  const float* repeats;      // [B,K] or null
  float alpha, beta, gamma;
  int64_t* inp; int L, pos;  // winner written to inp[b,pos] when inp != null
  float* out_clip_ref;       // [B]
  float* out_senti;          // [B] or null
  float* tr_clip_score; float* tr_clip_ref; float* tr_final; int64_t* tr_best;
};
// logit[r] = scale * cos(text row r, image row b) with b = (row_bk ? row_bk[r] : r) / K
void launch_clip_logits(const float* text, const float* image, const int32_t* row_bk, int n_rows, int K, int D,
                        float scale, float* logit, cudaStream_t st);
void launch_score_select(const SelectArgs& a, cudaStream_t st);

// ---- certified argmax (cert_ops.cu): the winner the exact (bf16x3) tower would pick, from a bf16 tower plus an
// exact re-score of the few candidates that are not provably beaten.  See the file header for the bound.
struct CertArgs {
  SelectArgs q;        // q.logit = main-tower logits [B,K]; round 2 patches the listed candidates in place
  float eps_hi, eps_lo; // main-tower logit - exact logit of one candidate lies in [-eps_lo, eps_hi] (scale * cosine bounds)
  float zr_lo, zr_hi;   // sum exp(exact logit) / sum exp(main-tower logit) over the candidates not re-scored lies in [zr_lo, zr_hi]
  float tau;           // slack for the fp rounding of the comparisons themselves
  int fcap;            // an image with more than fcap unbeaten candidates goes straight to the full exact re-encode
  float heavy;         // > 0: also list candidates whose softmax weight is >= heavy (narrows round 2's bound on Z)
  int32_t* img_nflag;  // [B] unbeaten candidates of image b (its winner included); -1 = full re-encode
  int32_t* img_k;      // [B,fcap] their candidate indices, ascending
  int32_t* flag_list;  // [<= B*fcap] b * K + k of every listed candidate; image b's occupy img_slot0[b] ... + img_nflag[b]
  int32_t* img_slot0;  // [B]
  int32_t* full_list;  // [B] images that need the full exact re-encode
  int32_t* counters;   // [0] listed candidates, [1] images in full_list, [2] images with more than one unbeaten
                       // candidate, [3] images sent to full_list by round 1
  const float* logit3; // round 2: exact logits of the listed candidates, [counters[0]]
};
void launch_cert_round1(const CertArgs& a, cudaStream_t st);
void launch_cert_round2(const CertArgs& a, cudaStream_t st);
// suffix rows of the listed candidates: out_ids[n, S] = ids_suffix[flag_list[n]], out_eos[n] = eos_idx[flag_list[n]],
// out_img[n] = flag_list[n] / K (the image whose shared prefix rows the candidate attends to)
void launch_cert_gather_suffix(const int32_t* flag_list, int n, const int32_t* ids_suffix, const int32_t* eos_idx, int K,
                               int S, int32_t* out_ids, int32_t* out_eos, int32_t* out_img, cudaStream_t st);
// Full re-encode of n images: compact copies of their per-image inputs (image i of the copy = image full_list[i]) ...
struct CertCompact {
  const int32_t* full_list; int n;
  int B, K, P, S, D;
  const int32_t *ids_prefix, *ids_suffix, *p0, *eos_idx;
  const float* probs; const int64_t* ids_masked; const float* senti; const float* repeats; const float* image;
  int32_t *c_ids_prefix, *c_ids_suffix, *c_p0, *c_eos_idx;
  float* c_probs; int64_t* c_ids_masked; float* c_senti; float* c_repeats; float* c_image;
};
void launch_cert_compact(const CertCompact& a, cudaStream_t st);
// ... and the way back: winners / scores / trace rows of the compact run scattered to the images they belong to
struct CertScatter {
  const int32_t* full_list; int n;
  int K, L, pos;
  const int64_t* c_inp; int64_t* inp;                // [n,L] / [B,L], column pos
  const float* c_clip_ref; float* clip_ref;          // [n] / [B]
  const float* c_senti; float* senti;                // [n] / [B] or null
  const float *c_tr_score, *c_tr_ref, *c_tr_final;   // [n,K] or null
  float *tr_score, *tr_ref, *tr_final;               // [B,K] or null
  const int64_t* c_tr_best; int64_t* tr_best;        // [n] / [B] or null
};
void launch_cert_scatter(const CertScatter& a, cudaStream_t st);

}  // namespace conzic
