// Internal interfaces between the translation units of libconzic.so (not part of the C ABI).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

namespace conzic {

typedef __nv_bfloat16 bf16;

extern uint64_t g_launches;  // kernels launched by this library (all contexts)
extern int g_pdl;            // 1 = launches may carry the programmatic-dependent-launch attribute (CONZIC_PDL, default 1)
extern int g_pdl_now;        // set by the engine per section: measured on B200, PDL gains ~3.5 % on steps made of short
                             // kernels (BERT, CLIP passes under ~50 k rows) and costs 1-2 % when the kernels are long

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
  cfg.numAttrs = (g_pdl && g_pdl_now) ? 1 : 0;
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
void prof_enable(bool on);
bool prof_read(int cat, double* ms, double* work, int* n);

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
  // LayerNorm folded into the GEMM (persistent kernel, bf16 mode).  Consumer side: A holds bf16(x) (NOT
  // normalised), W holds bf16(gamma * W), `bias` holds b + W beta, ln_s[n] = sum_k W'[n,k]; with the row's
  // mean and rstd from the partial sums in ln_stats the epilogue forms rstd*(acc - mean*ln_s[n]) + bias[n],
  // which equals LN(x) W^T + b.  Producer side: stats_out receives per-row (sum, sum of squares) of the fp32
  // values this GEMM writes, one float2 per 128-column slice: stats_out[row * stats_parts + slice].
  const float* ln_s = nullptr;
  const float2* ln_stats = nullptr;
  int ln_parts = 0;
  int ln_width = 0;     // number of columns the statistics cover (H)
  float ln_eps = 0.f;
  float2* stats_out = nullptr;
  int stats_parts = 0;
  // LayerNorm of the rows this GEMM writes, applied by the SAME kernel (gemm_wide_kernel only: N == 512 == H, a
  // CTA owns whole rows): after x = acc + bias + resid has been stored, lnf_out[m, :] = bf16(LN(x[m, :]) * g + b).
  // Replaces the stand-alone LayerNorm launch that would re-read x from HBM (HF:models/clip/modeling_clip.py:369-384).
  bf16* lnf_out = nullptr;
  int lnf_ld = 0;
  const float* lnf_g = nullptr;
  const float* lnf_b = nullptr;
  float lnf_eps = 0.f;
  int lnf_mode = 1;  // 1: the LayerNorm pass re-reads the fp32 values the lane just stored (L2) and the accumulators are
                     // handed back as soon as they are drained; 2: the values are kept in TMEM until the pass is done
};

struct GemmOpts {
  int split = 0;    // 1 = bf16x3
  int impl = 0;     // 0 tcgen05, 1 SIMT debug
  int bn = 128;     // N tile: 128 or 256
  int stages = 3;   // smem pipeline depth
  int persist = 0;  // 1 = persistent A-resident kernel with double-buffered TMEM accumulators (bf16 mode only)
  int cg = 1;       // persistent kernel: 2 = CTA pairs (tcgen05 cta_group::2), 1 = single CTAs
  int ksplit = 1;   // gridded kernel: split-K factor (raw fp32 partials, summed by the following LayerNorm)
  int force_wide = 0;  // persistent pair path, N == 512, fp32 output: use gemm_wide_kernel whatever K is
};

bool tma_init();  // resolves cuTensorMapEncodeTiled through the runtime (no link-time libcuda dependency)
bool make_tmap_bf16_2d(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                       uint32_t box_rows);
bool gemm_configure();  // opt-in to large dynamic shared memory for every instantiation
bool launch_linear(const Act& A, int M, const LinearW& W, const Epi& epi, const GemmOpts& o, cudaStream_t st,
                   uint64_t* launches);

// Fused MLP (bf16 mode): x_out = resid + fc2(act(fc1(Hin))) in one persistent launch; `scratch` is a bf16
// buffer of mlp_scratch_rows() x W1.N elements that holds the per-CTA fc1 tiles (L2 resident, re-used per tile).
int mlp_scratch_rows();
bool launch_mlp_fused(const Act& Hin, int M, const LinearW& W1, const LinearW& W2, bf16* scratch, int act,
                      const float* resid, int ldr, float* out_f32, int ldo_f32, bf16* out_act, int ldo_act,
                      cudaStream_t st);

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
// xb / stats optional: bf16 copy of the rows and their (sum, sum of squares) in stats[row * stats_parts + 0]
// (the other slices are zeroed), for a consumer GEMM with folded LayerNorm.
void launch_clip_embed(const int32_t* ids_prefix, const int32_t* ids_suffix, const int32_t* p0, int B, int P, int K,
                       int S, int maxpos, const float* tok, const float* pos, int H, float* x_f32, bf16* xb,
                       float2* stats, int stats_parts, cudaStream_t st, const float* ln_g = nullptr,
                       const float* ln_b = nullptr, float ln_eps = 0.f, bf16* ln_out = nullptr, int ln_ld = 0);
// Folds a LayerNorm into the linear layer that consumes it: w_out[n,k] = bf16(gamma[k] * w[n,k]),
// s_out[n] = sum_k float(w_out[n,k]), bias_out[n] = bias[n] + sum_k beta[k] * w[n,k].
void launch_fold_ln(const float* w, const float* bias, const float* gamma, const float* beta, int N, int K, bf16* w_out,
                    float* s_out, float* bias_out, cudaStream_t st);

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
  int cpt = 1;            // filled by launch_attention: candidates packed into one 16-row query tile
  int cand_per_task = 8;  // filled by launch_attention: candidates per warp task
  int prefetch = 0;       // filled by launch_attention: fetch the next candidate tile's q/k/v rows into registers
                          // before computing the current one (tiles of <= 16 own rows)
};
bool launch_attention(const AttnArgs& a, cudaStream_t st);

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
  // optional per-image override of the table walk (captions that hold a merged '##' word outside `pos`): for images
  // with ov_mask[b] != 0 the prefix tokens are ov_tok[ov_off[2b] .. ov_off[2b+1]) and the tail tokens
  // ov_tok[ov_off[2b+1] .. ov_off[2b+2]) -- what the host's tokenizers made of the decoded strings
  const int32_t* ov_mask = nullptr;  // [B]
  const int32_t* ov_off = nullptr;   // [2B+1]
  const int32_t* ov_tok = nullptr;
};
void launch_assemble(const AssembleArgs& a, cudaStream_t st);

void launch_step_prologue(int64_t* inp, int B, int L, int pos, int mask_id, float* token_mask, int dot_id,
                          int dot_allowed, cudaStream_t st);
void launch_gather_rows_index(int32_t* rows, int B, int L, int pos, cudaStream_t st);
void launch_pool_index(int32_t* rows, const int32_t* eos_idx, int B, int P, int K, int S, cudaStream_t st);

struct SelectArgs {
  const float* text;   // [B*K, D]
  const float* image;  // [B, D]
  int B, K, D;
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
void launch_score_select(const SelectArgs& a, cudaStream_t st);

}  // namespace conzic
