// libconzic.so: context (weights in operand format + TMA descriptors), the two towers as sequences of
// kernel launches on one stream, and the extern "C" boundary declared in include/conzic.h.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/conzic.h"
#include "kernels.h"

namespace conzic {

static thread_local std::string g_err;
thread_local CtxState* t_state = nullptr;

void set_error(const std::string& msg) { g_err = msg; }
bool cuda_ok(cudaError_t e, const char* what) {
  if (e == cudaSuccess) return true;
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return false;
}

ProfScope::ProfScope(int c, double work, cudaStream_t s) : cat(c), st(s), on(t_state && t_state->prof_on) {
  if (!on) return;
  ProfRec r;
  cudaEventCreate(&r.a);
  cudaEventCreate(&r.b);
  r.cat = c;
  r.phase = t_state->phase;
  r.work = work;
  cudaEventRecord(r.a, st);
  t_state->prof.push_back(r);
}
ProfScope::~ProfScope() {
  if (on) cudaEventRecord(t_state->prof.back().b, st);
}
void prof_enable(CtxState* s, bool on) {
  for (auto& r : s->prof) { cudaEventDestroy(r.a); cudaEventDestroy(r.b); }
  s->prof.clear();
  s->prof_on = on;
}
bool prof_read(CtxState* s, int cat, double* ms, double* work, int* n) {
  double t = 0, w = 0;
  int c = 0;
  for (auto& r : s->prof) {
    if (cat >= 100 ? (r.phase != cat - 100) : (r.cat != cat)) continue;
    if (cudaEventSynchronize(r.b) != cudaSuccess) return false;
    float e = 0;
    if (cudaEventElapsedTime(&e, r.a, r.b) != cudaSuccess) return false;
    t += e; w += r.work; ++c;
  }
  *ms = t; *work = w; *n = c;
  return true;
}

namespace {

// binds the calling thread to a context's launch state for the duration of one API call
struct StateScope {
  CtxState* prev;
  explicit StateScope(CtxState* s) : prev(t_state) { t_state = s; }
  ~StateScope() { t_state = prev; }
};

__global__ void find_eos_kernel(const int32_t* ids, int N, int T, int eos, int32_t* eos_idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  int r = 0;  // HF:models/clip/modeling_clip.py:564-584: argmax of (ids == eos), 0 when absent
  for (int t = 0; t < T; ++t)
    if (ids[static_cast<size_t>(i) * T + t] == eos) { r = t; break; }
  eos_idx[i] = r;
}
__global__ void bf16_to_f32_kernel(const bf16* src, float* dst, size_t n) {
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n;
       i += static_cast<size_t>(gridDim.x) * blockDim.x)
    dst[i] = __bfloat162float(src[i]);
}
__global__ void pad_eos_kernel(int32_t* ids, const int32_t* len, int N, int T, int eos) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * T) return;
  if (i % T >= len[i / T]) ids[i] = eos;
}
__global__ void add_one_kernel(int32_t* v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] += 1;
}

struct Layer {
  LinearW qkv, o, f1, f2;
  float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
};

// The linear layers of one transformer tower in one operand format (split = 1: bf16 hi | lo planes, 3 MMAs per
// k-step).  LayerNorm parameters are fp32 and shared between the formats.
struct Tower {
  int split = 0;
  GemmOpts gopt;
  std::vector<Layer> layers;
  LinearW proj;
};

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Bump {
  char* base;
  size_t off = 0, cap;
  Bump(void* p, size_t c) : base(static_cast<char*>(p)), cap(c) {}
  template <typename T>
  T* take(size_t n) {
    off = align_up(off, 256);
    T* r = reinterpret_cast<T*>(base + off);
    off += n * sizeof(T);
    return r;
  }
};

}  // namespace
}  // namespace conzic

using namespace conzic;

struct conzic_ctx {
  conzic_config cfg;
  CtxState state;
  std::vector<void*> owned;
  // BERT
  int bert_split = 0;
  GemmOpts gopt_bert;
  float *b_word = nullptr, *b_pos = nullptr, *b_type = nullptr, *b_eln_g = nullptr, *b_eln_b = nullptr;
  float *b_hln_g = nullptr, *b_hln_b = nullptr;
  LinearW b_transform, b_decoder;
  std::vector<Layer> bert;
  // CLIP text: `clip` encodes every candidate; `clip3` (CERTIFIED mode only) is the exact tower that re-scores the few
  // candidates the bf16 result cannot rule out
  float *c_tok = nullptr, *c_pos = nullptr, *c_fln_g = nullptr, *c_fln_b = nullptr;
  Tower clip, clip3;
  bool certified = false;
  float cert_dcos = 0.f, cert_dcos_lo = 0.f, cert_zratio_lo = 0.f, cert_zratio_hi = 0.f;
  int cert_fcap = 64;
  int32_t* cert_host = nullptr;  // pinned, 16 ints: the counters read back twice per certified step
  uint64_t cert_stats[CONZIC_CERT_STATS] = {};
  // CLIP image tower (optional; conzic_set_vision)
  conzic_vision_config vcfg{};
  bool has_vision = false;
  Tower vis;  // exact operands unless the whole context is bf16
  LinearW v_patch;
  float *v_cls = nullptr, *v_pos = nullptr, *v_pre_g = nullptr, *v_pre_b = nullptr, *v_post_g = nullptr, *v_post_b = nullptr;
  // bert id -> clip ids
  int32_t *b2c_off = nullptr, *b2c_tok = nullptr;
  int max_tok_per_word = 1;
  // device text pipeline (conzic_set_text_vocab): vocabularies with '##' pieces
  TextVocab text{};
  bool has_text = false;
  int32_t* dims_host = nullptr;  // pinned, 4 ints: the row capacities a text-path step reads back
  int32_t* aux = nullptr;        // small device scratch for conzic_build_clip_ids on the text path
  int aux_n = 0;
  int chunk_rows = 303104;
  int chunk_rows3 = 16384;
  int wide_ln = 0;  // CLIP bf16 tower: LayerNorm written by the epilogue of the GEMM that produces the residual stream

  ~conzic_ctx() {
    prof_enable(&state, false);
    for (void* p : owned) cudaFree(p);
    if (cert_host) cudaFreeHost(cert_host);
    if (dims_host) cudaFreeHost(dims_host);
  }
  template <typename T>
  T* dalloc(size_t n) {
    void* p = nullptr;
    if (!cuda_ok(cudaMalloc(&p, n * sizeof(T)), "cudaMalloc(weights)")) return nullptr;
    owned.push_back(p);
    return reinterpret_cast<T*>(p);
  }
  float* copy_f32(const void* src, size_t n, cudaStream_t st) {
    float* d = dalloc<float>(n);
    if (d) cuda_ok(cudaMemcpyAsync(d, src, n * sizeof(float), cudaMemcpyDeviceToDevice, st), "copy weights");
    return d;
  }
  // Build a LinearW from `parts` row blocks of fp32 [n_i, K] (e.g. q, k, v stacked into one [3H, K]).
  bool make_linear(LinearW* L, const void* const* w_parts, const void* const* b_parts, const int* n_parts, int parts,
                   int K, int split, cudaStream_t st, const float* shared_bias = nullptr) {
    int N = 0;
    for (int i = 0; i < parts; ++i) N += n_parts[i];
    L->N = N;
    L->K = K;
    const int ld = split ? 2 * K : K;
    L->w = dalloc<bf16>(static_cast<size_t>(N) * ld);
    if (!L->w) return false;
    float* bias = nullptr;
    if (b_parts && !shared_bias) {
      bias = dalloc<float>(N);
      if (!bias) return false;
    }
    int r0 = 0;
    for (int i = 0; i < parts; ++i) {
      launch_f32_to_act(static_cast<const float*>(w_parts[i]), n_parts[i], K, K, L->w + static_cast<size_t>(r0) * ld, ld,
                        split, st);
      if (bias)
        cuda_ok(cudaMemcpyAsync(bias + r0, b_parts[i], n_parts[i] * sizeof(float), cudaMemcpyDeviceToDevice, st),
                "copy bias");
      r0 += n_parts[i];
    }
    L->bias = shared_bias ? shared_bias : bias;
    if (cfg.gemm_impl == CONZIC_GEMM_TCGEN05) {
      if (!make_tmap_bf16_2d(&L->tmap128, L->w, N, ld, ld, 128)) return false;
      if (!make_tmap_bf16_2d(&L->tmap256, L->w, N, ld, ld, 256)) return false;
      if (!make_tmap_bf16_2d(&L->tmap64, L->w, N, ld, ld, 64)) return false;
    }
    return true;
  }
  // One pre-LN block's linears (16 state-dict tensors `t`, CLIP order) in the given operand format.  `like` (optional)
  // is the same block of another tower whose fp32 LayerNorm parameters and biases are shared instead of copied.
  bool make_clip_layer(Layer* ly, const void* const* t, int H, int F, int split, cudaStream_t st, const Layer* like) {
    const void* wq[3] = {t[2], t[4], t[6]}; const void* bq[3] = {t[3], t[5], t[7]}; int nq[3] = {H, H, H};
    const void* wo[1] = {t[8]}; const void* bo[1] = {t[9]}; int no[1] = {H};
    const void* wf[1] = {t[12]}; const void* bf[1] = {t[13]}; int nf[1] = {F};
    const void* wg[1] = {t[14]}; const void* bg[1] = {t[15]}; int ng[1] = {H};
    bool ok = make_linear(&ly->qkv, wq, bq, nq, 3, H, split, st, like ? like->qkv.bias : nullptr) &&
              make_linear(&ly->o, wo, bo, no, 1, H, split, st, like ? like->o.bias : nullptr) &&
              make_linear(&ly->f1, wf, bf, nf, 1, H, split, st, like ? like->f1.bias : nullptr) &&
              make_linear(&ly->f2, wg, bg, ng, 1, F, split, st, like ? like->f2.bias : nullptr);
    if (!ok) return false;
    if (like) {
      ly->ln1_g = like->ln1_g; ly->ln1_b = like->ln1_b; ly->ln2_g = like->ln2_g; ly->ln2_b = like->ln2_b;
    } else {
      ly->ln1_g = copy_f32(t[0], H, st); ly->ln1_b = copy_f32(t[1], H, st);
      ly->ln2_g = copy_f32(t[10], H, st); ly->ln2_b = copy_f32(t[11], H, st);
    }
    return ly->ln1_g && ly->ln1_b && ly->ln2_g && ly->ln2_b;
  }
};

namespace {

// Buffers of one pass of a CLIP text tower over at most `rows_cap` token rows
struct ClipBufs {
  float *cx, *cxe, *text;
  bf16 *ch, *cattn, *cffn, *cpool;
  void* cqkv;
  int32_t* pool_rows;
  size_t rows_cap;
};

struct Plan {
  // BERT
  float *bx, *by, *bpart;
  bf16 *bh, *battn, *bffn;
  void* bqkv;
  float *bt, *logits;
  bf16* bt_act;
  int ldl;
  // top-k / assembly / selection
  float *probs, *repeats, *senti, *clogit;
  int64_t *ids, *ids_masked;
  int32_t *ids_prefix, *ids_suffix, *p0, *eos_idx;
  int32_t *seq, *seq_len, *dims;  // text path: every candidate's full CLIP id sequence
  ClipBufs main;
  // certified mode
  ClipBufs exact;
  int32_t *img_nflag, *img_k, *img_slot0, *flag_list, *full_list, *counters, *ids3, *eos3, *img3;
  float* logit3;
  int32_t *c_ids_prefix, *c_ids_suffix, *c_p0, *c_eos_idx;
  float *c_probs, *c_senti, *c_repeats, *c_image, *c_clip_ref, *c_senti_out, *c_tr_score, *c_tr_ref, *c_tr_final;
  int64_t *c_ids_masked, *c_inp, *c_tr_best;
  size_t bytes;
};

size_t tower_rows_cap(const conzic_ctx* c, int chunk_rows, int B, int K) {
  const size_t per_img = static_cast<size_t>(c->cfg.clip_maxpos) * (K + 1);
  size_t cap = static_cast<size_t>(chunk_rows) > per_img ? static_cast<size_t>(chunk_rows) : per_img;
  const size_t most = static_cast<size_t>(B > 0 ? B : 1) * per_img;  // never more than B images can produce
  return cap < most ? cap : most;
}

void take_clip_bufs(Bump& b, ClipBufs& cb, size_t rows, size_t BK, int Hc, int Fc, int proj, int s) {
  cb.rows_cap = rows;
  cb.cx = b.take<float>(rows * Hc);
  cb.ch = b.take<bf16>(rows * Hc * (1 + s));
  cb.cattn = b.take<bf16>(rows * Hc * (1 + s));
  cb.cffn = b.take<bf16>(rows * Fc * (1 + s));
  cb.cqkv = b.take<char>(rows * 3 * Hc * (s ? 4 : 2));
  cb.cpool = b.take<bf16>(BK * Hc * (1 + s));
  cb.text = b.take<float>(BK * proj);
  cb.cxe = b.take<float>(BK * Hc);
  cb.pool_rows = b.take<int32_t>(BK);
}

Plan make_plan(const conzic_ctx* c, void* ws, int B, int L, int K) {
  const conzic_config& g = c->cfg;
  const int s = c->bert_split;
  Bump b(ws, 0);
  Plan p{};
  const size_t Mb = static_cast<size_t>(B) * (L > 0 ? L : 1);
  const int Hb = g.bert_hidden, Fb = g.bert_ffn, Hc = g.clip_hidden, Fc = g.clip_ffn;
  p.bx = b.take<float>(Mb * Hb);
  p.by = b.take<float>(Mb * Hb);
  p.bpart = b.take<float>(Mb * Hb * 4);  // split-K partial sums of the O-proj / FFN-out GEMMs (<= 4 splits)
  p.bh = b.take<bf16>(Mb * Hb * (1 + s));
  p.battn = b.take<bf16>(Mb * Hb * (1 + s));
  p.bffn = b.take<bf16>(Mb * Fb * (1 + s));
  p.bqkv = b.take<char>(Mb * 3 * Hb * (s ? 4 : 2));
  p.bt = b.take<float>(static_cast<size_t>(B) * Hb);
  p.bt_act = b.take<bf16>(static_cast<size_t>(B) * Hb * (1 + s));
  p.ldl = (g.bert_vocab + 3) & ~3;
  p.logits = b.take<float>(static_cast<size_t>(B) * p.ldl);
  const size_t BK = static_cast<size_t>(B) * K;
  p.probs = b.take<float>(BK);
  p.repeats = b.take<float>(BK);
  p.senti = b.take<float>(BK);
  p.clogit = b.take<float>(BK);
  p.ids = b.take<int64_t>(BK);
  p.ids_masked = b.take<int64_t>(BK);
  p.ids_prefix = b.take<int32_t>(static_cast<size_t>(B) * g.clip_maxpos);
  p.ids_suffix = b.take<int32_t>(BK * g.clip_maxpos);
  p.p0 = b.take<int32_t>(B);
  p.eos_idx = b.take<int32_t>(BK);
  p.seq = b.take<int32_t>(c->has_text ? BK * g.clip_maxpos : 1);
  p.seq_len = b.take<int32_t>(c->has_text ? BK : 1);
  p.dims = b.take<int32_t>(4);
  take_clip_bufs(b, p.main, tower_rows_cap(c, c->chunk_rows, B, K), BK, Hc, Fc, g.clip_proj, c->clip.split);
  if (c->certified) {
    take_clip_bufs(b, p.exact, tower_rows_cap(c, c->chunk_rows3, B, K), BK, Hc, Fc, g.clip_proj, 1);
    const size_t BF = static_cast<size_t>(B) * c->cert_fcap;
    p.img_nflag = b.take<int32_t>(B);
    p.img_k = b.take<int32_t>(BF);
    p.img_slot0 = b.take<int32_t>(B);
    p.flag_list = b.take<int32_t>(BF);
    p.full_list = b.take<int32_t>(B);
    p.counters = b.take<int32_t>(16);
    p.ids3 = b.take<int32_t>(BF * g.clip_maxpos);
    p.eos3 = b.take<int32_t>(BF);
    p.img3 = b.take<int32_t>(BF);
    p.logit3 = b.take<float>(BF > BK ? BF : BK);
    p.c_ids_prefix = b.take<int32_t>(static_cast<size_t>(B) * g.clip_maxpos);
    p.c_ids_suffix = b.take<int32_t>(BK * g.clip_maxpos);
    p.c_p0 = b.take<int32_t>(B);
    p.c_eos_idx = b.take<int32_t>(BK);
    p.c_probs = b.take<float>(BK);
    p.c_senti = b.take<float>(BK);
    p.c_repeats = b.take<float>(BK);
    p.c_image = b.take<float>(static_cast<size_t>(B) * g.clip_proj);
    p.c_clip_ref = b.take<float>(B);
    p.c_senti_out = b.take<float>(B);
    p.c_tr_score = b.take<float>(BK);
    p.c_tr_ref = b.take<float>(BK);
    p.c_tr_final = b.take<float>(BK);
    p.c_ids_masked = b.take<int64_t>(BK);
    p.c_inp = b.take<int64_t>(static_cast<size_t>(B) * (L > 0 ? L : 1));
    p.c_tr_best = b.take<int64_t>(B);
  }
  p.bytes = align_up(b.off, 256);
  return p;
}

bool check_ws(const conzic_ctx* c, size_t ws_bytes, int B, int L, int K) {
  Plan p = make_plan(c, nullptr, B, L, K);
  if (ws_bytes < p.bytes) {
    set_error("workspace too small: need " + std::to_string(p.bytes) + " bytes, got " + std::to_string(ws_bytes));
    return false;
  }
  return true;
}

Epi epi_act_out(const LinearW& W, bf16* out, int ld, int outK, int act) {
  Epi e;
  e.bias = W.bias; e.out_act = out; e.ldo_act = ld; e.out_K = outK; e.act = act;
  return e;
}
Epi epi_f32_out(const LinearW& W, float* out, int ld, const float* resid, int ldr, int act) {
  Epi e;
  e.bias = W.bias; e.out_f32 = out; e.ldo_f32 = ld; e.resid = resid; e.ldr = ldr; e.act = act;
  return e;
}

// ---- BERT: logits of row `pos` for every image ------------------------------------------------------
bool bert_row_logits(conzic_ctx* c, const int64_t* inp, int B, int L, int pos, float* logits, int ldl, Plan& p,
                     cudaStream_t st) {
  const conzic_config& g = c->cfg;
  const int H = g.bert_hidden, F = g.bert_ffn, s = c->bert_split;
  const int M = B * L;
  const int ldh = H * (1 + s), ldf = F * (1 + s);
  if (L > g.bert_maxpos) { set_error("bert: sequence longer than position table"); return false; }
  set_pdl_now(1);  // ~85 short launches: overlap their launch latencies
  launch_bert_embed_ln(inp, M, L, c->b_word, c->b_pos, c->b_type, c->b_eln_g, c->b_eln_b, g.bert_ln_eps, H, p.bx, p.bh,
                       ldh, s, st);
  for (size_t l = 0; l < c->bert.size(); ++l) {
    const Layer& ly = c->bert[l];
    Act h{p.bh, ldh, H};
    Epi e;
    if (s) e = epi_f32_out(ly.qkv, static_cast<float*>(p.bqkv), 3 * H, nullptr, 0, ACT_NONE);
    else   e = epi_act_out(ly.qkv, static_cast<bf16*>(p.bqkv), 3 * H, 0, ACT_NONE);
    const GemmOpts o = c->gopt_bert;  // few token rows (B*L): the (n, m)-gridded kernel spreads them over more SMs
    // These GEMMs are latency bound (<= 150 CTAs, 12-48 k blocks each): a deep TMA ring (6 x 32 KB in flight per
    // CTA) matters more than a second resident CTA, except for fc1 whose grid exceeds one CTA per SM.
    GemmOpts od = o;
    od.stages = 6;
    const bool one_wave = static_cast<long long>((M + 127) / 128) * ((F + o.bn - 1) / o.bn) <= 148;
    // bf16x3: the wide-N GEMMs (QKV, fc1) run on the CTA-pair kernel from 512 rows on -- 256 x 256 tiles move half the
    // operand bytes per FLOP of the 128 x 128 gridded kernel, which the three passes make L2-bandwidth bound (measured
    // 22 / 72 us per launch at 960 rows).  The N = H GEMMs keep the split-K gridded form (3 pair tiles would idle the chip).
    GemmOpts opair = o;
    opair.persist = (s && M >= 512 && o.impl == 0) ? 1 : 0;
    if (!launch_linear(h, M, ly.qkv, e, opair.persist ? opair : od, st)) return false;
    AttnArgs at;
    at.qkv = p.bqkv; at.ld_qkv = 3 * H; at.qkv_f32 = s; at.p0 = nullptr;
    at.B = B; at.P = 0; at.K = 1; at.S = L; at.H = H; at.heads = g.bert_heads; at.causal = 0;
    at.scale = 0.125f; at.out_act = p.battn; at.ld_act = ldh; at.split = s;
    if (!launch_attention(at, st)) return false;
    Act a{p.battn, ldh, H};
    // The two N = H GEMMs have few tiles (48 at 64 images): split K so ~148 CTAs stream the operands, write raw
    // fp32 partial sums, and let the LayerNorm that follows add them to the residual and the bias.
    const int tiles = ((M + 127) / 128) * ((H + od.bn - 1) / od.bn);
    int ks = 148 / (tiles > 0 ? tiles : 1);
    if (ks > 4) ks = 4;
    if (ks < 1) ks = 1;
    GemmOpts osk = od;
    osk.ksplit = ks;
    auto nh_gemm = [&](const Act& in, const LinearW& W, const float* lg, const float* lb) -> bool {
      if (ks > 1) {
        Epi e2;
        e2.out_f32 = p.bpart; e2.ldo_f32 = H;
        if (!launch_linear(in, M, W, e2, osk, st)) return false;
        LNArgs ln{p.bx, nullptr, M, H, lg, lb, g.bert_ln_eps, p.bx, p.bh, ldh, s};
        ln.partials = p.bpart; ln.n_parts = ks; ln.part_stride = static_cast<size_t>(M) * H; ln.add_bias = W.bias;
        launch_layernorm(ln, st);
        return true;
      }
      if (!launch_linear(in, M, W, epi_f32_out(W, p.by, H, p.bx, H, ACT_NONE), od, st)) return false;
      LNArgs ln{p.by, nullptr, M, H, lg, lb, g.bert_ln_eps, p.bx, p.bh, ldh, s};
      launch_layernorm(ln, st);
      return true;
    };
    if (!nh_gemm(a, ly.o, ly.ln1_g, ly.ln1_b)) return false;
    if (!launch_linear(h, M, ly.f1, epi_act_out(ly.f1, p.bffn, ldf, F, ACT_ERF_GELU), opair.persist ? opair : (one_wave ? od : o), st))
      return false;
    Act f{p.bffn, ldf, F};
    if (!nh_gemm(f, ly.f2, ly.ln2_g, ly.ln2_b)) return false;
  }
  // MLM head on row `pos` only: a strided view of the hidden states (row stride L*ldh) needs no gather.
  Act hrow{p.bh + static_cast<size_t>(pos) * ldh, L * ldh, H};
  const GemmOpts o = c->gopt_bert;
  GemmOpts ot = o;
  ot.stages = 6;  // 6 CTAs, 36 k blocks each: latency bound, the deep TMA ring is all that helps
  if (!launch_linear(hrow, B, c->b_transform, epi_f32_out(c->b_transform, p.bt, H, nullptr, 0, ACT_ERF_GELU), ot, st))
    return false;
  LNArgs lnh{p.bt, nullptr, B, H, c->b_hln_g, c->b_hln_b, g.bert_ln_eps, nullptr, p.bt_act, ldh, s};
  launch_layernorm(lnh, st);
  Act t{p.bt_act, ldh, H};
  return launch_linear(t, B, c->b_decoder, epi_f32_out(c->b_decoder, logits, ldl, nullptr, 0, ACT_NONE), o, st);
}

// ---- a CLIP text tower over packed prefix/suffix rows ------------------------------------------------
// One pass: nb images' shared prefix rows (nb * P) followed by n_cand candidate blocks of S rows.  cand_img == null:
// the candidates are K per image in image order (n_cand == nb * K); otherwise candidate i belongs to image cand_img[i]
// (the certified re-score of a few candidates per image).  text receives one embedding per candidate.
bool encode_pass(conzic_ctx* c, const Tower& tw, ClipBufs& p, const int32_t* idp, const int32_t* ids, const int32_t* p0c,
                 const int32_t* cand_img, const int32_t* eos, int nb, int P, int K, int n_cand, int S, float* text,
                 float* xe, cudaStream_t st) {
  const conzic_config& g = c->cfg;
  const int H = g.clip_hidden, F = g.clip_ffn, s = tw.split;
  const int ldh = H * (1 + s), ldf = F * (1 + s);
  const float scale = 0.125f;  // head_dim^-0.5, HF:models/clip/modeling_clip.py:283
  const int wl = (&tw == &c->clip) ? c->wide_ln : 0;
  const int M = nb * P + n_cand * S;
  set_pdl_now((M < 50000) ? 1 : 0);
  // the embedding kernel holds each row in registers: it also writes LN1 of the first block (bit-identical to the
  // stand-alone launch it replaces)
  const bool embed_ln = !s && H == 512 && !tw.layers.empty();
  launch_clip_embed(idp, ids, p0c, nb, P, K, S, g.clip_maxpos, c->c_tok, c->c_pos, H, p.cx, st,
                    embed_ln ? tw.layers[0].ln1_g : nullptr, embed_ln ? tw.layers[0].ln1_b : nullptr, g.clip_ln_eps,
                    embed_ln ? p.ch : nullptr, ldh, cand_img, n_cand);
  const int NE = n_cand;  // rows that are pooled: one EOS row per candidate caption
  bool h_ready = embed_ln, pooled_ready = false;
  // LayerNorm written by the producing GEMM's epilogue (gemm_wide_kernel owns whole 512-column rows)
  auto with_ln = [&](Epi e, bf16* out, const float* gamma, const float* beta) {
    e.lnf_out = out; e.lnf_ld = ldh; e.lnf_g = gamma; e.lnf_b = beta; e.lnf_eps = g.clip_ln_eps;
    return e;
  };
  for (size_t l = 0; l < tw.layers.size(); ++l) {
    const Layer& ly = tw.layers[l];
    const bool last = (l + 1 == tw.layers.size());
    if (!h_ready) {  // otherwise the embedding kernel / the previous block's fc2 epilogue already wrote LN1(x) into ch
      LNArgs ln1{p.cx, nullptr, M, H, ly.ln1_g, ly.ln1_b, g.clip_ln_eps, nullptr, p.ch, ldh, s};
      launch_layernorm(ln1, st);
    }
    h_ready = false;  // consumed by this block's QKV
    Act h{p.ch, ldh, H};
    Epi e;
    if (s) e = epi_f32_out(ly.qkv, static_cast<float*>(p.cqkv), 3 * H, nullptr, 0, ACT_NONE);
    else   e = epi_act_out(ly.qkv, static_cast<bf16*>(p.cqkv), 3 * H, 0, ACT_NONE);
    if (!launch_linear(h, M, ly.qkv, e, tw.gopt, st)) return false;
    AttnArgs at;
    at.qkv = p.cqkv; at.ld_qkv = 3 * H; at.qkv_f32 = s; at.p0 = p0c;
    at.B = nb; at.P = P; at.K = K; at.S = S; at.H = H; at.heads = g.clip_heads; at.causal = 1;
    at.scale = scale; at.out_act = p.cattn; at.ld_act = ldh; at.split = s;
    at.cand_img = cand_img; at.n_cand = n_cand;
    if (!launch_attention(at, st)) return false;
    // The tower's output is read at ONE row per caption (its first EOS) and a row of the last block depends on
    // other rows only through this block's attention: after it, only the EOS rows go on (exact, not an
    // approximation).  Rows are compacted: attention out -> ch, residual -> xe; LN2 output re-uses cattn.
    int Mr = M;
    float* x = p.cx;
    bf16 *a_in = p.cattn, *h2 = p.ch;
    if (last) {
      launch_pool_index(p.pool_rows, eos, nb * P, NE, S, st);
      launch_gather_rows(p.cattn, static_cast<size_t>(ldh) * sizeof(bf16), p.pool_rows, NE, p.ch, st);
      launch_gather_rows(p.cx, static_cast<size_t>(H) * sizeof(float), p.pool_rows, NE, xe, st);
      Mr = NE; x = xe; a_in = p.ch; h2 = p.cattn;
    }
    Act a{a_in, ldh, H};
    if (wl) {  // O-proj through the wide pair kernel, LN2 written by its epilogue
      if (!launch_linear(a, Mr, ly.o, with_ln(epi_f32_out(ly.o, x, H, x, H, ACT_NONE), h2, ly.ln2_g, ly.ln2_b), tw.gopt, st))
        return false;
    } else {
      if (!launch_linear(a, Mr, ly.o, epi_f32_out(ly.o, x, H, x, H, ACT_NONE), tw.gopt, st)) return false;
      LNArgs ln2{x, nullptr, Mr, H, ly.ln2_g, ly.ln2_b, g.clip_ln_eps, nullptr, h2, ldh, s};
      launch_layernorm(ln2, st);
    }
    Act hh{h2, ldh, H};
    if (!launch_linear(hh, Mr, ly.f1, epi_act_out(ly.f1, p.cffn, ldf, F, ACT_QUICK_GELU), tw.gopt, st)) return false;
    Act f{p.cffn, ldf, F};
    Epi e2 = epi_f32_out(ly.f2, x, H, x, H, ACT_NONE);
    if (wl && !last) {  // LN1 of the next block, straight into its QKV operand
      e2 = with_ln(e2, p.ch, tw.layers[l + 1].ln1_g, tw.layers[l + 1].ln1_b);
      h_ready = true;
    } else if (wl) {    // last block (EOS rows only): the final LayerNorm, straight into the projection operand
      e2 = with_ln(e2, p.cpool, c->c_fln_g, c->c_fln_b);
      pooled_ready = true;
    }
    if (!launch_linear(f, Mr, ly.f2, e2, tw.gopt, st)) return false;
  }
  if (tw.layers.empty()) {  // degenerate 0-layer tower: pool straight from the embeddings
    launch_pool_index(p.pool_rows, eos, nb * P, NE, S, st);
    launch_gather_rows(p.cx, static_cast<size_t>(H) * sizeof(float), p.pool_rows, NE, xe, st);
  }
  // pooled = final LN of the hidden state at the first EOS (already compacted); text_projection without bias
  if (!pooled_ready) {
    LNArgs lnf{xe, nullptr, NE, H, c->c_fln_g, c->c_fln_b, g.clip_ln_eps, nullptr, p.cpool, ldh, s};
    launch_layernorm(lnf, st);
  }
  Act pooled{p.cpool, ldh, H};
  Epi e;
  e.out_f32 = text;
  e.ldo_f32 = g.clip_proj;
  return launch_linear(pooled, NE, tw.proj, e, tw.gopt, st);
}

// K candidates per image, in passes of whole images
bool clip_encode(conzic_ctx* c, const Tower& tw, ClipBufs& p, int chunk_rows, const int32_t* ids_prefix,
                 const int32_t* ids_suffix, const int32_t* p0, const int32_t* eos_idx, int B, int P, int K, int S,
                 float* text, cudaStream_t st) {
  const conzic_config& g = c->cfg;
  const int per_img = P + K * S;
  int Bc = chunk_rows / per_img;
  if (Bc < 1) Bc = 1;
  if (Bc > B) Bc = B;
  if (static_cast<size_t>(Bc) * per_img > p.rows_cap) Bc = static_cast<int>(p.rows_cap / per_img);
  if (Bc < 1) {
    set_error("clip_encode: one image's token rows exceed the workspace plan");
    return false;
  }
  for (int b0 = 0; b0 < B; b0 += Bc) {
    const int nb = (B - b0 < Bc) ? (B - b0) : Bc;
    if (!encode_pass(c, tw, p, P > 0 ? ids_prefix + static_cast<size_t>(b0) * P : nullptr,
                     ids_suffix + static_cast<size_t>(b0) * K * S, p0 ? p0 + b0 : nullptr, nullptr,
                     eos_idx + static_cast<size_t>(b0) * K, nb, P, K, nb * K, S,
                     text + static_cast<size_t>(b0) * K * g.clip_proj, p.cxe + static_cast<size_t>(b0) * K * g.clip_hidden, st))
      return false;
  }
  return cuda_ok(cudaGetLastError(), "clip_encode");
}

// n_cand candidates of arbitrary images (cand_img), every pass re-encoding the B shared prefixes (B * P rows: small next
// to what a dense re-encode of the candidates' whole captions costs)
bool clip_encode_mapped(conzic_ctx* c, const Tower& tw, ClipBufs& p, int chunk_rows, const int32_t* ids_prefix,
                        const int32_t* ids_suffix, const int32_t* p0, const int32_t* cand_img, const int32_t* eos_idx,
                        int B, int P, int n_cand, int S, float* text, cudaStream_t st) {
  const conzic_config& g = c->cfg;
  const size_t pre = static_cast<size_t>(B) * P;
  size_t cap = static_cast<size_t>(chunk_rows) < p.rows_cap ? static_cast<size_t>(chunk_rows) : p.rows_cap;
  if (cap <= pre + static_cast<size_t>(S)) cap = p.rows_cap;
  if (cap < pre + static_cast<size_t>(S)) {
    set_error("clip_encode: prefix rows exceed the workspace plan");
    return false;
  }
  const int Cc = static_cast<int>((cap - pre) / S);
  for (int c0 = 0; c0 < n_cand; c0 += Cc) {
    const int nc = (n_cand - c0 < Cc) ? (n_cand - c0) : Cc;
    if (!encode_pass(c, tw, p, P > 0 ? ids_prefix : nullptr, ids_suffix + static_cast<size_t>(c0) * S, P > 0 ? p0 : nullptr,
                     cand_img + c0, eos_idx + c0, P > 0 ? B : 0, P, 1, nc, S, text + static_cast<size_t>(c0) * g.clip_proj,
                     p.cxe + static_cast<size_t>(c0) * g.clip_hidden, st))
      return false;
  }
  return cuda_ok(cudaGetLastError(), "clip_encode (mapped)");
}

bool have_device() {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    set_error("no CUDA device: libconzic has no CPU fallback");
    return false;
  }
  int dev = 0, major = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10) {
    set_error("libconzic is built for sm_100a (B200) only; device compute capability major is " + std::to_string(major));
    return false;
  }
  return true;
}

// How the candidate captions' CLIP ids are laid out: P shared prefix rows per image + S rows per candidate
// (P == 0: every candidate is a whole dense row of S ids).
struct CandLayout {
  const int32_t *ids_prefix, *ids_suffix, *p0, *eos_idx;
  int P, S;
};

// gen_utils.py:77-81 from the main tower's embeddings `text` [B*K, D]: logits, then either the plain fused
// score / argmax or, in CERTIFIED mode, the certified one (cert_ops.cu) with its exact re-encodes.  q carries
// everything but q.logit.  Synchronises the stream twice per call in CERTIFIED mode (two 64-byte reads).
bool select_winner(conzic_ctx* c, Plan& p, const CandLayout& lay, const float* text, const float* image, SelectArgs q,
                   cudaStream_t st) {
  const conzic_config& g = c->cfg;
  const int B = q.B, K = q.K, D = g.clip_proj;
  set_pdl_now(1);
  set_phase(PHASE_SELECT);
  launch_clip_logits(text, image, nullptr, B * K, K, D, q.scale, p.clogit, st);
  q.logit = p.clogit;
  if (!c->certified) {
    launch_score_select(q, st);
    return cuda_ok(cudaGetLastError(), "score_select");
  }
  CertArgs ca{};
  ca.q = q;
  ca.eps_hi = q.scale * c->cert_dcos;
  ca.eps_lo = q.scale * c->cert_dcos_lo;
  // the denominator ratio: the measured bounds, never wider than what the per-candidate bounds imply
  ca.zr_lo = fmaxf(expf(-ca.eps_hi), c->cert_zratio_lo);
  ca.zr_hi = c->cert_zratio_hi > 0.f ? fminf(expf(ca.eps_lo), c->cert_zratio_hi) : expf(ca.eps_lo);
  ca.tau = 2e-6f * (fabsf(q.alpha) + fabsf(q.beta) + fabsf(q.gamma) + 1.0f);
  ca.fcap = c->cert_fcap;
  ca.heavy = 0.f;
  ca.img_nflag = p.img_nflag; ca.img_k = p.img_k; ca.img_slot0 = p.img_slot0; ca.flag_list = p.flag_list;
  ca.full_list = p.full_list; ca.counters = p.counters; ca.logit3 = p.logit3;
  if (!cuda_ok(cudaMemsetAsync(p.counters, 0, 16 * sizeof(int32_t), st), "memset(cert counters)")) return false;
  launch_cert_round1(ca, st);
  int32_t* h = c->cert_host;
  if (!cuda_ok(cudaMemcpyAsync(h, p.counters, 16 * sizeof(int32_t), cudaMemcpyDeviceToHost, st), "read cert counters") ||
      !cuda_ok(cudaStreamSynchronize(st), "sync(cert round 1)"))
    return false;
  const int n_list = h[0];
  c->cert_stats[0] += 1;
  c->cert_stats[1] += static_cast<uint64_t>(B);
  c->cert_stats[2] += static_cast<uint64_t>(n_list);
  c->cert_stats[3] += static_cast<uint64_t>(h[2]);
  c->cert_stats[4] += static_cast<uint64_t>(h[3]);
  if (n_list > 0) {
    set_phase(PHASE_CERT_RESCORE);
    // exact re-encode of the listed candidates: their suffix rows + the images' shared prefix rows (bit-identical to
    // what the exact tower produces for them inside a full pass: every kernel of that tower is row-wise deterministic)
    launch_cert_gather_suffix(p.flag_list, n_list, lay.ids_suffix, lay.eos_idx, K, lay.S, p.ids3, p.eos3, p.img3, st);
    if (!clip_encode_mapped(c, c->clip3, p.exact, c->chunk_rows3, lay.ids_prefix, p.ids3, lay.p0, p.img3, p.eos3, B, lay.P,
                            n_list, lay.S, p.exact.text, st))
      return false;
    set_pdl_now(1);
    launch_clip_logits(p.exact.text, image, p.flag_list, n_list, K, D, q.scale, p.logit3, st);
  }
  launch_cert_round2(ca, st);
  if (!cuda_ok(cudaMemcpyAsync(h, p.counters, 16 * sizeof(int32_t), cudaMemcpyDeviceToHost, st), "read cert counters") ||
      !cuda_ok(cudaStreamSynchronize(st), "sync(cert round 2)"))
    return false;
  const int n_full = h[1];
  c->cert_stats[5] += static_cast<uint64_t>(n_full);
  if (n_full > 0) {
    set_phase(PHASE_CERT_FULL);
    // images the bound could not decide: every candidate through the exact tower, then the plain kernel
    CertCompact cc{};
    cc.full_list = p.full_list; cc.n = n_full; cc.B = B; cc.K = K; cc.P = lay.P; cc.S = lay.S; cc.D = D;
    cc.ids_prefix = lay.ids_prefix; cc.ids_suffix = lay.ids_suffix; cc.p0 = lay.p0; cc.eos_idx = lay.eos_idx;
    cc.probs = q.probs; cc.ids_masked = q.ids_masked; cc.senti = q.senti; cc.repeats = q.repeats; cc.image = image;
    cc.c_ids_prefix = p.c_ids_prefix; cc.c_ids_suffix = p.c_ids_suffix; cc.c_p0 = p.c_p0; cc.c_eos_idx = p.c_eos_idx;
    cc.c_probs = p.c_probs; cc.c_ids_masked = p.c_ids_masked; cc.c_senti = p.c_senti; cc.c_repeats = p.c_repeats;
    cc.c_image = p.c_image;
    launch_cert_compact(cc, st);
    if (!clip_encode(c, c->clip3, p.exact, c->chunk_rows3, lay.P > 0 ? p.c_ids_prefix : nullptr, p.c_ids_suffix,
                     lay.P > 0 ? p.c_p0 : nullptr, p.c_eos_idx, n_full, lay.P, K, lay.S, p.exact.text, st))
      return false;
    set_pdl_now(1);
    launch_clip_logits(p.exact.text, p.c_image, nullptr, n_full * K, K, D, q.scale, p.logit3, st);
    SelectArgs q3 = q;
    q3.logit = p.logit3; q3.B = n_full;
    q3.probs = p.c_probs; q3.ids_masked = p.c_ids_masked;
    q3.senti = q.senti ? p.c_senti : nullptr; q3.repeats = q.repeats ? p.c_repeats : nullptr;
    q3.inp = p.c_inp; q3.out_clip_ref = p.c_clip_ref; q3.out_senti = q.out_senti ? p.c_senti_out : nullptr;
    q3.tr_clip_score = q.tr_clip_score ? p.c_tr_score : nullptr;
    q3.tr_clip_ref = q.tr_clip_ref ? p.c_tr_ref : nullptr;
    q3.tr_final = q.tr_final ? p.c_tr_final : nullptr;
    q3.tr_best = q.tr_best ? p.c_tr_best : nullptr;
    launch_score_select(q3, st);
    CertScatter cs{};
    cs.full_list = p.full_list; cs.n = n_full; cs.K = K; cs.L = q.L; cs.pos = q.pos;
    cs.c_inp = p.c_inp; cs.inp = q.inp;
    cs.c_clip_ref = p.c_clip_ref; cs.clip_ref = q.out_clip_ref;
    cs.c_senti = q3.out_senti; cs.senti = q.out_senti;
    cs.c_tr_score = q3.tr_clip_score; cs.c_tr_ref = q3.tr_clip_ref; cs.c_tr_final = q3.tr_final;
    cs.tr_score = q.tr_clip_score; cs.tr_ref = q.tr_clip_ref; cs.tr_final = q.tr_final;
    cs.c_tr_best = q3.tr_best; cs.tr_best = q.tr_best;
    launch_cert_scatter(cs, st);
  }
  return cuda_ok(cudaGetLastError(), "certified select");
}

void fill_assemble(const conzic_ctx* c, AssembleArgs& a) {
  const conzic_config& g = c->cfg;
  a.off = c->b2c_off; a.tok = c->b2c_tok; a.V = g.bert_vocab;
  a.special[0] = g.pad_id; a.special[1] = g.unk_id; a.special[2] = g.cls_id; a.special[3] = g.sep_id;
  a.special[4] = g.mask_id;
  a.bos = g.clip_bos; a.eos = g.clip_eos; a.maxlen = g.clip_maxpos;
}

}  // namespace

extern "C" {

int conzic_abi_version(void) { return CONZIC_ABI_VERSION; }
const char* conzic_last_error(void) { return g_err.c_str(); }

int conzic_ctx_create(const conzic_config* cfg, const void* const* bw, int n_bert, const void* const* cw, int n_clip,
                      void* stream, conzic_ctx** out) {
  if (!cfg || !bw || !cw || !out) { set_error("ctx_create: null argument"); return -1; }
  if (!have_device()) return -2;
  if (n_bert != CONZIC_BERT_GLOBALS + CONZIC_PER_LAYER * cfg->bert_layers ||
      n_clip != CONZIC_CLIP_GLOBALS + CONZIC_PER_LAYER * cfg->clip_layers) {
    set_error("ctx_create: weight table length does not match the layer counts");
    return -1;
  }
  if (cfg->bert_hidden != cfg->bert_heads * 64 || cfg->clip_hidden != cfg->clip_heads * 64) {
    set_error("ctx_create: head_dim must be 64");
    return -1;
  }
  if (cfg->precision < CONZIC_PREC_BF16 || cfg->precision > CONZIC_PREC_CERTIFIED) {
    set_error("ctx_create: unknown precision mode");
    return -1;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  conzic_ctx* c = new conzic_ctx();
  StateScope scope(&c->state);
  c->cfg = *cfg;
  c->state.pdl = (cfg->flags & CONZIC_FLAG_NO_PDL) ? 0 : 1;
  c->certified = cfg->precision == CONZIC_PREC_CERTIFIED;
  // BERT: exact operands unless the whole context is bf16 (the top-K set and its probabilities must be the
  // reference's: they decide which candidates exist at all)
  c->bert_split = cfg->precision == CONZIC_PREC_BF16 ? 0 : 1;
  c->clip.split = cfg->precision == CONZIC_PREC_BF16X3 ? 1 : 0;
  c->clip3.split = 1;
  c->vis.split = c->bert_split;
  auto opts = [&](int split, bool persist) {
    GemmOpts o;
    o.split = split; o.impl = cfg->gemm_impl; o.bn = 128; o.stages = 3;
    o.persist = persist ? 1 : 0;  // persistent CTA-pair kernel (tcgen05 cta_group::2): the CLIP towers, bf16 and bf16x3
    o.cg = 2;
    o.wide_lsu = (cfg->flags & CONZIC_FLAG_WIDE_LSU) ? 1 : 0;
    o.lsu_out = (cfg->flags & CONZIC_FLAG_LSU_OUT) ? 1 : 0;  // CONZIC_FLAG_WIDE_LSU16 / _LSU8
    return o;
  };
  c->gopt_bert = opts(c->bert_split, false);  // few token rows: gridded kernel (QKV / fc1 in bf16x3 opt into the pair kernel)
  c->clip.gopt = opts(c->clip.split, true);
  c->clip3.gopt = opts(1, true);
  c->vis.gopt = opts(c->vis.split, true);
  c->wide_ln = (!(cfg->flags & CONZIC_FLAG_LN_STANDALONE) && !c->clip.split && cfg->gemm_impl == CONZIC_GEMM_TCGEN05 &&
                cfg->clip_hidden == 512) ? 1 : 0;
  c->cert_dcos = cfg->cert_dcos > 0.f ? cfg->cert_dcos : CONZIC_CERT_DCOS_DEFAULT;
  // an explicit upper bound without a lower one is taken as symmetric
  c->cert_dcos_lo = cfg->cert_dcos_lo > 0.f ? cfg->cert_dcos_lo : (cfg->cert_dcos > 0.f ? cfg->cert_dcos : CONZIC_CERT_DCOS_LO_DEFAULT);
  c->cert_fcap = cfg->cert_fcap > 0 ? cfg->cert_fcap : 64;
  // measured denominator-ratio bounds go with the measured default error bounds; explicit error bounds without explicit
  // ratio bounds fall back to what the error bounds imply (negative = "none")
  const bool own_bounds = cfg->cert_dcos > 0.f || cfg->cert_dcos_lo > 0.f;
  c->cert_zratio_lo = cfg->cert_zratio_lo > 0.f ? cfg->cert_zratio_lo : (own_bounds || cfg->cert_zratio_lo < 0.f ? 0.f : CONZIC_CERT_ZRATIO_LO_DEFAULT);
  c->cert_zratio_hi = cfg->cert_zratio_hi > 0.f ? cfg->cert_zratio_hi : (own_bounds || cfg->cert_zratio_hi < 0.f ? 0.f : CONZIC_CERT_ZRATIO_HI_DEFAULT);
  // default: 16 x (148 SMs x 128 rows) token rows per pass; measured on B200: the larger the pass the better
  // (every kernel is a persistent or grid-stride launch; nothing stays L2 resident between kernels anyway)
  c->chunk_rows = cfg->clip_chunk_rows > 0 ? cfg->clip_chunk_rows : 303104;
  bool ok = true;
  if (cfg->gemm_impl == CONZIC_GEMM_TCGEN05) ok = tma_init() && gemm_configure();
  ok = ok && topk_configure() && attention_configure();
  if (ok && c->certified)
    ok = cuda_ok(cudaHostAlloc(reinterpret_cast<void**>(&c->cert_host), 16 * sizeof(int32_t), cudaHostAllocDefault),
                 "cudaHostAlloc(cert counters)");
  const int Hb = cfg->bert_hidden, Fb = cfg->bert_ffn, Vb = cfg->bert_vocab;
  const int Hc = cfg->clip_hidden, Fc = cfg->clip_ffn;
  const int sb = c->bert_split;
  if (ok) {
    c->b_word = c->copy_f32(bw[0], static_cast<size_t>(Vb) * Hb, st);
    c->b_pos = c->copy_f32(bw[1], static_cast<size_t>(cfg->bert_maxpos) * Hb, st);
    c->b_type = c->copy_f32(bw[2], static_cast<size_t>(2) * Hb, st);
    c->b_eln_g = c->copy_f32(bw[3], Hb, st);
    c->b_eln_b = c->copy_f32(bw[4], Hb, st);
    c->b_hln_g = c->copy_f32(bw[8], Hb, st);
    c->b_hln_b = c->copy_f32(bw[9], Hb, st);
    ok = c->b_word && c->b_pos && c->b_type && c->b_eln_g && c->b_eln_b && c->b_hln_g && c->b_hln_b;
  }
  if (ok) {
    const void* w1[1] = {bw[6]}; const void* b1[1] = {bw[7]}; int n1[1] = {Hb};
    ok = c->make_linear(&c->b_transform, w1, b1, n1, 1, Hb, sb, st);
    const void* w2[1] = {bw[0]}; const void* b2[1] = {bw[5]}; int n2[1] = {Vb};
    ok = ok && c->make_linear(&c->b_decoder, w2, b2, n2, 1, Hb, sb, st);  // decoder tied to word embeddings
  }
  for (int l = 0; ok && l < cfg->bert_layers; ++l) {
    const void* const* t = bw + CONZIC_BERT_GLOBALS + CONZIC_PER_LAYER * l;
    Layer ly;
    const void* wq[3] = {t[0], t[2], t[4]}; const void* bq[3] = {t[1], t[3], t[5]}; int nq[3] = {Hb, Hb, Hb};
    ok = c->make_linear(&ly.qkv, wq, bq, nq, 3, Hb, sb, st);
    const void* wo[1] = {t[6]}; const void* bo[1] = {t[7]}; int no[1] = {Hb};
    ok = ok && c->make_linear(&ly.o, wo, bo, no, 1, Hb, sb, st);
    ly.ln1_g = c->copy_f32(t[8], Hb, st); ly.ln1_b = c->copy_f32(t[9], Hb, st);
    const void* wf[1] = {t[10]}; const void* bf[1] = {t[11]}; int nf[1] = {Fb};
    ok = ok && c->make_linear(&ly.f1, wf, bf, nf, 1, Hb, sb, st);
    const void* wg[1] = {t[12]}; const void* bg[1] = {t[13]}; int ng[1] = {Hb};
    ok = ok && c->make_linear(&ly.f2, wg, bg, ng, 1, Fb, sb, st);
    ly.ln2_g = c->copy_f32(t[14], Hb, st); ly.ln2_b = c->copy_f32(t[15], Hb, st);
    ok = ok && ly.ln1_g && ly.ln1_b && ly.ln2_g && ly.ln2_b;
    c->bert.push_back(ly);
  }
  if (ok) {
    c->c_tok = c->copy_f32(cw[0], static_cast<size_t>(cfg->clip_vocab) * Hc, st);
    c->c_pos = c->copy_f32(cw[1], static_cast<size_t>(cfg->clip_maxpos) * Hc, st);
    c->c_fln_g = c->copy_f32(cw[2], Hc, st);
    c->c_fln_b = c->copy_f32(cw[3], Hc, st);
    const void* wp[1] = {cw[4]}; int np[1] = {cfg->clip_proj};
    ok = c->c_tok && c->c_pos && c->c_fln_g && c->c_fln_b &&
         c->make_linear(&c->clip.proj, wp, nullptr, np, 1, Hc, c->clip.split, st);
    if (ok && c->certified) ok = c->make_linear(&c->clip3.proj, wp, nullptr, np, 1, Hc, 1, st);
  }
  for (int l = 0; ok && l < cfg->clip_layers; ++l) {
    const void* const* t = cw + CONZIC_CLIP_GLOBALS + CONZIC_PER_LAYER * l;
    Layer ly;
    ok = c->make_clip_layer(&ly, t, Hc, Fc, c->clip.split, st, nullptr);
    c->clip.layers.push_back(ly);
    if (ok && c->certified) {
      Layer l3;
      ok = c->make_clip_layer(&l3, t, Hc, Fc, 1, st, &c->clip.layers.back());
      c->clip3.layers.push_back(l3);
    }
  }
  ok = ok && cuda_ok(cudaStreamSynchronize(st), "ctx_create sync");
  if (!ok) { delete c; return -3; }
  c->state.launches = 0;
  *out = c;
  return 0;
}

void conzic_ctx_destroy(conzic_ctx* ctx) { delete ctx; }

int conzic_set_bert2clip(conzic_ctx* c, const int32_t* off, const int32_t* tok, int n_tok, int max_tok_per_word,
                         void* stream) {
  if (!c || !off || (!tok && n_tok > 0)) { set_error("set_bert2clip: null argument"); return -1; }
  StateScope scope(&c->state);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  c->b2c_off = c->dalloc<int32_t>(c->cfg.bert_vocab + 1);
  c->b2c_tok = c->dalloc<int32_t>(n_tok > 0 ? n_tok : 1);
  if (!c->b2c_off || !c->b2c_tok) return -3;
  bool ok = cuda_ok(cudaMemcpyAsync(c->b2c_off, off, (c->cfg.bert_vocab + 1) * sizeof(int32_t), cudaMemcpyDeviceToDevice, st), "copy b2c off");
  if (n_tok > 0)
    ok = ok && cuda_ok(cudaMemcpyAsync(c->b2c_tok, tok, n_tok * sizeof(int32_t), cudaMemcpyDeviceToDevice, st), "copy b2c tok");
  c->max_tok_per_word = max_tok_per_word > 0 ? max_tok_per_word : 1;
  return ok ? 0 : -3;
}

int conzic_set_text_vocab(conzic_ctx* c, const conzic_text_vocab* tv, void* stream) {
  if (!c || !tv || !tv->tok_off || !tv->tok_bytes || !tv->tok_cls || !tv->tok_flags || !tv->byte_sym || !tv->merge_keys ||
      !tv->merge_vals) { set_error("set_text_vocab: null argument"); return -1; }
  if (!c->b2c_off) { set_error("set_text_vocab: conzic_set_bert2clip must come first (per-token ids of whole words)"); return -1; }
  if (tv->merge_bits < 1 || tv->merge_bits > 24 || tv->n_bytes < 0) { set_error("set_text_vocab: bad table sizes"); return -1; }
  StateScope scope(&c->state);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int V = c->cfg.bert_vocab;
  const size_t nb = tv->n_bytes > 0 ? tv->n_bytes : 1, nm = static_cast<size_t>(1) << tv->merge_bits;
  int32_t* off = c->dalloc<int32_t>(V + 1);
  uint8_t* bytes = c->dalloc<uint8_t>(nb);
  uint8_t* cls = c->dalloc<uint8_t>(nb);
  uint8_t* flags = c->dalloc<uint8_t>(V);
  int32_t* sym = c->dalloc<int32_t>(512);
  uint64_t* keys = c->dalloc<uint64_t>(nm);
  uint32_t* vals = c->dalloc<uint32_t>(nm);
  if (!off || !bytes || !cls || !flags || !sym || !keys || !vals) return -3;
  auto cp = [&](void* d, const void* s_, size_t n) {
    return cuda_ok(cudaMemcpyAsync(d, s_, n, cudaMemcpyDeviceToDevice, st), "copy text vocab");
  };
  bool ok = cp(off, tv->tok_off, (V + 1) * sizeof(int32_t)) && cp(bytes, tv->tok_bytes, nb) && cp(cls, tv->tok_cls, nb) &&
            cp(flags, tv->tok_flags, V) && cp(sym, tv->byte_sym, 512 * sizeof(int32_t)) &&
            cp(keys, tv->merge_keys, nm * sizeof(uint64_t)) && cp(vals, tv->merge_vals, nm * sizeof(uint32_t));
  if (ok && !c->dims_host)
    ok = cuda_ok(cudaHostAlloc(reinterpret_cast<void**>(&c->dims_host), 4 * sizeof(int32_t), cudaHostAllocDefault),
                 "cudaHostAlloc(text dims)");
  ok = ok && cuda_ok(cudaStreamSynchronize(st), "set_text_vocab sync");
  if (!ok) return -3;
  c->text = TextVocab{off, bytes, cls, flags, c->b2c_off, c->b2c_tok, sym, keys, vals, tv->merge_bits, V};
  c->has_text = true;
  return 0;
}

size_t conzic_workspace_bytes(const conzic_ctx* c, int B, int L, int K) {
  if (!c) return 0;
  return make_plan(c, nullptr, B, L, K).bytes;
}

int conzic_profile(conzic_ctx* c, int enable) {
  if (!c) { set_error("profile: null context"); return -1; }
  prof_enable(&c->state, enable != 0);
  return 0;
}
int conzic_profile_read(conzic_ctx* c, int category, double* ms, double* work, int* launches) {
  if (!c || category < 0 || (category >= CAT_COUNT && (category < 100 || category >= 100 + PHASE_COUNT)) || !ms || !work ||
      !launches) { set_error("profile_read: bad argument"); return -1; }
  return prof_read(&c->state, category, ms, work, launches) ? 0 : -4;
}

uint64_t conzic_launch_count(const conzic_ctx* c) { return c ? c->state.launches : 0; }

int conzic_cert_stats(const conzic_ctx* c, uint64_t* out, int n) {
  if (!c || !out || n < 0) { set_error("cert_stats: bad argument"); return -1; }
  for (int i = 0; i < n && i < CONZIC_CERT_STATS; ++i) out[i] = c->cert_stats[i];
  return 0;
}

int conzic_bert_mlm_row(conzic_ctx* c, const int64_t* inp, int B, int L, int pos, float* logits, int ldl, void* ws,
                        size_t ws_bytes, void* stream) {
  if (!c || !inp || !logits || !ws) { set_error("bert_mlm_row: null argument"); return -1; }
  if (pos < 0 || pos >= L || ldl < c->cfg.bert_vocab || (ldl & 3)) { set_error("bert_mlm_row: bad pos / ldl"); return -1; }
  if (!check_ws(c, ws_bytes, B, L, 1)) return -1;
  StateScope scope(&c->state);
  Plan p = make_plan(c, ws, B, L, 1);
  return bert_row_logits(c, inp, B, L, pos, logits, ldl, p, static_cast<cudaStream_t>(stream)) ? 0 : -4;
}

int conzic_topk_mask(conzic_ctx* c, const float* logits, int ldl, int B, const float* token_mask, float temperature,
                     int K, float* probs, int64_t* ids, void* stream) {
  if (!c || !logits || !token_mask || !probs || !ids) { set_error("topk_mask: null argument"); return -1; }
  StateScope scope(&c->state);
  return launch_topk(logits, ldl, B, c->cfg.bert_vocab, token_mask, temperature, K, probs, ids,
                     static_cast<cudaStream_t>(stream)) ? 0 : -4;
}

int conzic_build_clip_ids(conzic_ctx* c, const int64_t* inp, int B, int L, int pos, const int64_t* ids,
                          const float* token_mask, int K, int32_t* clip_ids, int T, int32_t* clip_len,
                          int64_t* ids_masked, void* stream) {
  if (!c || !inp || !ids || !token_mask || !clip_ids || !clip_len || !ids_masked) { set_error("build_clip_ids: null argument"); return -1; }
  if (!c->b2c_off) { set_error("build_clip_ids: conzic_set_bert2clip has not been called"); return -1; }
  StateScope scope(&c->state);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (c->has_text) {
    // device text pipeline: the sequences are written straight into clip_ids, which must have full-length rows
    if (T != c->cfg.clip_maxpos || L > TXT_MAX_TOK) { set_error("build_clip_ids: with a text vocabulary T must be 77 and L <= 96"); return -1; }
    if (c->aux_n < B + 4) {
      c->aux = c->dalloc<int32_t>(B + 4);
      c->aux_n = c->aux ? B + 4 : 0;
      if (!c->aux) return -3;
    }
    const conzic_config& g = c->cfg;
    TextAssembleArgs t{};
    t.vocab = c->text;
    t.inp = inp; t.ids = ids; t.token_mask = token_mask; t.senti_table = nullptr;
    t.B = B; t.L = L; t.K = K; t.pos = pos;
    t.special[0] = g.pad_id; t.special[1] = g.unk_id; t.special[2] = g.cls_id; t.special[3] = g.sep_id; t.special[4] = g.mask_id;
    t.bos = g.clip_bos; t.eos = g.clip_eos; t.maxlen = T;
    t.seq = clip_ids; t.len = clip_len; t.p0 = c->aux; t.dims = c->aux + B;
    t.ids_masked = ids_masked;
    if (!cuda_ok(cudaMemsetAsync(t.dims, 0, 4 * sizeof(int32_t), st), "memset(text dims)")) return -4;
    launch_text_tokenize(t, st);
    pad_eos_kernel<<<(B * K * T + 255) / 256, 256, 0, st>>>(clip_ids, clip_len, B * K, T, g.clip_eos);
    count_launch();
    return cuda_ok(cudaGetLastError(), "build_clip_ids") ? 0 : -4;
  }
  AssembleArgs a{};
  fill_assemble(c, a);
  a.inp = inp; a.ids = ids; a.token_mask = token_mask; a.senti_table = nullptr;
  a.B = B; a.L = L; a.K = K; a.pos = pos;
  a.ids_prefix = nullptr; a.ids_suffix = clip_ids; a.p0 = nullptr; a.eos_idx = clip_len; a.P = 0; a.S = T;
  a.ids_masked = ids_masked; a.repeats = nullptr; a.senti = nullptr;
  launch_assemble(a, st);
  add_one_kernel<<<(B * K + 255) / 256, 256, 0, st>>>(clip_len, B * K);
  count_launch();
  return cuda_ok(cudaGetLastError(), "build_clip_ids") ? 0 : -4;
}

int conzic_clip_text_encode(conzic_ctx* c, const int32_t* clip_ids, int N, int T, float* text, void* ws,
                            size_t ws_bytes, void* stream) {
  if (!c || !clip_ids || !text || !ws) { set_error("clip_text_encode: null argument"); return -1; }
  if (T < 1 || T > c->cfg.clip_maxpos) { set_error("clip_text_encode: T must be in [1, 77]"); return -1; }
  if (!check_ws(c, ws_bytes, N, 0, 1)) return -1;
  StateScope scope(&c->state);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Plan p = make_plan(c, ws, N, 0, 1);
  find_eos_kernel<<<(N + 255) / 256, 256, 0, st>>>(clip_ids, N, T, c->cfg.clip_eos, p.eos_idx);
  count_launch();
  if (!clip_encode(c, c->clip, p.main, c->chunk_rows, nullptr, clip_ids, nullptr, p.eos_idx, N, 0, 1, T, p.main.text, st))
    return -4;
  return cuda_ok(cudaMemcpyAsync(text, p.main.text, static_cast<size_t>(N) * c->cfg.clip_proj * sizeof(float),
                                 cudaMemcpyDeviceToDevice, st), "copy text embeds") ? 0 : -4;
}

int conzic_image_text_similarity(conzic_ctx* c, const float* text, const float* image, int B, int K, float scale,
                                 float* clip_score, float* clip_ref, void* stream) {
  if (!c || !text || !image || !clip_ref) { set_error("image_text_similarity: null argument"); return -1; }
  if (K > 1024) { set_error("image_text_similarity: K > 1024"); return -1; }
  StateScope scope(&c->state);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // clip_ref doubles as the logit buffer: the select kernel reads every logit of its row before it writes the cosines
  launch_clip_logits(text, image, nullptr, B * K, K, c->cfg.clip_proj, scale, clip_ref, st);
  SelectArgs a{};
  a.logit = clip_ref; a.B = B; a.K = K; a.scale = scale;
  a.tr_clip_score = clip_score; a.tr_clip_ref = clip_ref;
  launch_score_select(a, st);
  return cuda_ok(cudaGetLastError(), "image_text_similarity") ? 0 : -4;
}

int conzic_score_select(conzic_ctx* c, const float* text, const float* image, const int32_t* clip_ids, int T, int B,
                        int K, float scale, const float* probs, const int64_t* ids_masked, const float* senti_raw,
                        const float* repeats, float alpha, float beta, float gamma, int64_t* inp, int L, int pos,
                        float* out_clip_ref, float* out_senti, int64_t* out_best, void* ws, size_t ws_bytes,
                        void* stream) {
  if (!c || !text || !image || !probs || !ids_masked || !inp || !out_clip_ref || !ws) { set_error("score_select: null argument"); return -1; }
  if (K < 1 || K > 1024 || pos < 0 || pos >= L) { set_error("score_select: bad K / pos"); return -1; }
  if (c->certified && (!clip_ids || T < 1 || T > c->cfg.clip_maxpos)) {
    set_error("score_select: the certified mode needs the candidates' CLIP ids (clip_ids int32[B*K, T], T <= 77)");
    return -1;
  }
  if (!check_ws(c, ws_bytes, B, L, K)) return -1;
  StateScope scope(&c->state);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Plan p = make_plan(c, ws, B, L, K);
  SelectArgs q{};
  q.B = B; q.K = K; q.scale = scale;
  q.probs = probs; q.ids_masked = ids_masked; q.senti = senti_raw; q.repeats = repeats;
  q.alpha = alpha; q.beta = beta; q.gamma = gamma;
  q.inp = inp; q.L = L; q.pos = pos; q.out_clip_ref = out_clip_ref; q.out_senti = out_senti; q.tr_best = out_best;
  CandLayout lay{nullptr, clip_ids, nullptr, p.eos_idx, 0, T};
  if (c->certified) {
    find_eos_kernel<<<(B * K + 255) / 256, 256, 0, st>>>(clip_ids, B * K, T, c->cfg.clip_eos, p.eos_idx);
    count_launch();
  }
  return select_winner(c, p, lay, text, image, q, st) ? 0 : -4;
}

int conzic_gibbs_step(conzic_ctx* c, const conzic_step_args* s, void* ws, size_t ws_bytes, void* stream) {
  if (!c || !s || !ws || !s->inp || !s->token_mask || !s->image_embeds) { set_error("gibbs_step: null argument"); return -1; }
  if (!c->b2c_off) { set_error("gibbs_step: conzic_set_bert2clip has not been called"); return -1; }
  const int B = s->B, L = s->L, K = s->K, pos = s->pos;
  if (B < 1 || K < 1 || K > 1024 || pos < 1 || pos >= L - 1) { set_error("gibbs_step: bad B / K / pos"); return -1; }
  if (!check_ws(c, ws_bytes, B, L, K)) return -1;
  StateScope scope(&c->state);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  Plan p = make_plan(c, ws, B, L, K);
  const conzic_config& g = c->cfg;
  set_pdl_now(1);
  set_phase(PHASE_BERT);
  launch_step_prologue(s->inp, B, L, pos, g.mask_id, s->token_mask, g.dot_id, s->dot_allowed, st);
  float* logits = s->tr_logits ? s->tr_logits : p.logits;
  if (s->logits_in) {
    logits = const_cast<float*>(s->logits_in);  // span order: scores from an earlier forward, no BERT here
  } else if (!bert_row_logits(c, s->inp, B, L, pos, logits, p.ldl, p, st)) {
    return -4;
  }
  float* probs = s->tr_probs ? s->tr_probs : p.probs;
  int64_t* ids = s->tr_ids ? s->tr_ids : p.ids;
  set_phase(PHASE_CANDIDATES);
  if (!launch_topk(logits, p.ldl, B, g.bert_vocab, s->token_mask, s->temperature, K, probs, ids, st)) return -4;
  const int W = c->max_tok_per_word;
  const int cap = g.clip_maxpos - 1;
  int P = 1 + W * s->visited_before; if (P > cap) P = cap;
  int S = W * (1 + s->visited_after) + 1; if (S > cap) S = cap;
  const bool ctl = s->senti_table != nullptr;
  if (c->has_text) {
    // vocabularies with '##' pieces: every candidate caption through the device text pipeline; the row capacities
    // are data dependent (a piece merges into its neighbour), so the two maxima are read back -- 16 bytes
    if (L > TXT_MAX_TOK) { set_error("gibbs_step: caption longer than the text pipeline's token buffer"); return -1; }
    TextAssembleArgs t{};
    t.vocab = c->text;
    t.inp = s->inp; t.ids = ids; t.token_mask = s->token_mask; t.senti_table = s->senti_table;
    t.B = B; t.L = L; t.K = K; t.pos = pos;
    t.special[0] = g.pad_id; t.special[1] = g.unk_id; t.special[2] = g.cls_id; t.special[3] = g.sep_id; t.special[4] = g.mask_id;
    t.bos = g.clip_bos; t.eos = g.clip_eos; t.maxlen = g.clip_maxpos;
    t.seq = p.seq; t.len = p.seq_len; t.p0 = p.p0; t.dims = p.dims;
    t.ids_masked = p.ids_masked; t.repeats = ctl ? p.repeats : nullptr; t.senti = ctl ? p.senti : nullptr;
    if (!cuda_ok(cudaMemsetAsync(p.dims, 0, 4 * sizeof(int32_t), st), "memset(text dims)")) return -4;
    launch_text_tokenize(t, st);
    int32_t* h = c->dims_host;
    if (!cuda_ok(cudaMemcpyAsync(h, p.dims, 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, st), "read text dims") ||
        !cuda_ok(cudaStreamSynchronize(st), "sync(text dims)"))
      return -4;
    if (h[2]) {
      set_error("gibbs_step: a caption overflows the device text pipeline's buffers (TXT_ERR bits " + std::to_string(h[2]) + ")");
      return -4;
    }
    P = h[0] < 1 ? 1 : h[0];
    S = h[1] < 1 ? 1 : h[1];
    if (P + S > 96) {  // one attention tile holds prefix + suffix keys: shorten the shared prefix, lengthen the suffixes
      const int cut = P + S - 96;
      P -= cut; S += cut;
      if (P < 1 || S > cap) { set_error("gibbs_step: captions too long for the attention tile"); return -4; }
    }
    t.P = P; t.S = S; t.ids_prefix = p.ids_prefix; t.ids_suffix = p.ids_suffix; t.eos_idx = p.eos_idx;
    launch_text_layout(t, st);
  } else {
    AssembleArgs a{};
    fill_assemble(c, a);
    a.inp = s->inp; a.ids = ids; a.token_mask = s->token_mask; a.senti_table = s->senti_table;
    a.B = B; a.L = L; a.K = K; a.pos = pos;
    a.ids_prefix = p.ids_prefix; a.ids_suffix = p.ids_suffix; a.p0 = p.p0; a.eos_idx = p.eos_idx; a.P = P; a.S = S;
    a.ids_masked = p.ids_masked; a.repeats = ctl ? p.repeats : nullptr; a.senti = ctl ? p.senti : nullptr;
    launch_assemble(a, st);
  }
  set_phase(PHASE_TOWER);
  if (!clip_encode(c, c->clip, p.main, c->chunk_rows, p.ids_prefix, p.ids_suffix, p.p0, p.eos_idx, B, P, K, S, p.main.text, st))
    return -4;
  SelectArgs q{};
  q.B = B; q.K = K; q.scale = s->logit_scale_exp;
  q.probs = probs; q.ids_masked = p.ids_masked; q.senti = ctl ? p.senti : nullptr; q.repeats = ctl ? p.repeats : nullptr;
  q.alpha = s->alpha; q.beta = s->beta; q.gamma = s->gamma;
  q.inp = s->inp; q.L = L; q.pos = pos;
  q.out_clip_ref = s->out_clip_ref; q.out_senti = s->out_senti;
  q.tr_clip_score = s->tr_clip_score; q.tr_clip_ref = s->tr_clip_ref; q.tr_final = s->tr_final; q.tr_best = s->tr_best;
  CandLayout lay{p.ids_prefix, p.ids_suffix, p.p0, p.eos_idx, P, S};
  const bool ok = select_winner(c, p, lay, p.main.text, s->image_embeds, q, st);
  set_phase(PHASE_OTHER);
  return ok ? 0 : -4;
}

int conzic_debug_linear(conzic_ctx* c, const float* A, const float* Wf, const float* bias, const float* resid, int M,
                        int N, int K, int act, float* out, void* ws, size_t ws_bytes, void* stream) {
  if (!c || !A || !Wf || !out || !ws) { set_error("debug_linear: null argument"); return -1; }
  StateScope scope(&c->state);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool bf16_out = (act & 16) != 0;  // exercise the bf16 activation output path (bf16 operands only)
  const bool exact = (act & 32) != 0;     // CERTIFIED contexts: use the exact (bf16x3) operand format
  const bool gridded = (act & 64) != 0;   // force the gridded 128 x 128 kernel (bit-identity checks against the pair kernel)
  const int act_flags = act;              // | 128: force the pair kernel
  act &= 15;
  const Tower& tw = (exact && c->certified) ? c->clip3 : c->clip;
  const int s = tw.split;
  if (bf16_out && (s || resid)) { set_error("debug_linear: bf16 output needs bf16 operands and no residual"); return -1; }
  const int ld = K * (1 + s);
  const size_t need = (static_cast<size_t>(M) + N) * ld * sizeof(bf16) + (bf16_out ? static_cast<size_t>(M) * N * 2 : 0) + 2048;
  if (ws_bytes < need) { set_error("debug_linear: workspace too small, need " + std::to_string(need)); return -1; }
  Bump b(ws, ws_bytes);
  bf16* a_act = b.take<bf16>(static_cast<size_t>(M) * ld);
  bf16* w_act = b.take<bf16>(static_cast<size_t>(N) * ld);
  launch_f32_to_act(A, M, K, K, a_act, ld, s, st);
  launch_f32_to_act(Wf, N, K, K, w_act, ld, s, st);
  LinearW W;
  W.w = w_act; W.bias = bias; W.N = N; W.K = K;
  if (c->cfg.gemm_impl == CONZIC_GEMM_TCGEN05) {
    if (!make_tmap_bf16_2d(&W.tmap128, w_act, N, ld, ld, 128)) return -4;
    if (!make_tmap_bf16_2d(&W.tmap256, w_act, N, ld, ld, 256)) return -4;
    if (!make_tmap_bf16_2d(&W.tmap64, w_act, N, ld, ld, 64)) return -4;
  }
  Epi e;
  e.bias = bias; e.resid = resid; e.ldr = N; e.act = act;
  bf16* o16 = nullptr;
  if (bf16_out) {
    o16 = b.take<bf16>(static_cast<size_t>(M) * N);
    e.out_act = o16; e.ldo_act = N; e.out_K = 0;
  } else {
    e.out_f32 = out; e.ldo_f32 = N;
  }
  Act a{a_act, ld, K};
  GemmOpts go = tw.gopt;
  if (gridded) { go.persist = 0; go.stages = 6; }
  if (act_flags & 128) go.force_pair = 1;
  if (!launch_linear(a, M, W, e, go, st)) return -4;
  if (bf16_out) {
    bf16_to_f32_kernel<<<1184, 256, 0, st>>>(o16, out, static_cast<size_t>(M) * N);
    count_launch();
  }
  return cuda_ok(cudaGetLastError(), "debug_linear") ? 0 : -4;
}

// ---- CLIP image tower -------------------------------------------------------------------------------
namespace {
struct VPlan {
  bf16 *patches, *h, *attn, *ffn, *pool;
  float *pe, *x;
  void* qkv;
  size_t bytes;
};
VPlan make_vplan(const conzic_ctx* c, void* ws, int B) {
  const conzic_vision_config& v = c->vcfg;
  const int s = c->vis.split, H = v.hidden, F = v.ffn;
  const int g2 = (v.image_size / v.patch) * (v.image_size / v.patch), T = g2 + 1, Kp = 3 * v.patch * v.patch;
  const size_t M = static_cast<size_t>(B) * T;
  Bump b(ws, 0);
  VPlan p;
  p.patches = b.take<bf16>(static_cast<size_t>(B) * g2 * Kp * (1 + s));
  p.pe = b.take<float>(static_cast<size_t>(B) * g2 * H);
  p.x = b.take<float>(M * H);
  p.h = b.take<bf16>(M * H * (1 + s));
  p.attn = b.take<bf16>(M * H * (1 + s));
  p.ffn = b.take<bf16>(M * F * (1 + s));
  p.qkv = b.take<char>(M * 3 * H * (s ? 4 : 2));
  p.pool = b.take<bf16>(static_cast<size_t>(B) * H * (1 + s));
  p.bytes = align_up(b.off, 256);
  return p;
}
}  // namespace

int conzic_set_vision(conzic_ctx* c, const conzic_vision_config* vc, const void* const* w, int n, void* stream) {
  if (!c || !vc || !w) { set_error("set_vision: null argument"); return -1; }
  if (n != 8 + CONZIC_PER_LAYER * vc->layers) { set_error("set_vision: weight table length does not match the layer count"); return -1; }
  if (vc->hidden != vc->heads * 64 || (vc->patch % 4) || (vc->image_size % vc->patch)) { set_error("set_vision: unsupported shape"); return -1; }
  StateScope scope(&c->state);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  c->vcfg = *vc;
  const int H = vc->hidden, F = vc->ffn, Kp = 3 * vc->patch * vc->patch, s = c->vis.split;
  const int T = (vc->image_size / vc->patch) * (vc->image_size / vc->patch) + 1;
  bool ok = true;
  { const void* wp[1] = {w[0]}; int np[1] = {H}; ok = c->make_linear(&c->v_patch, wp, nullptr, np, 1, Kp, s, st); }
  c->v_cls = c->copy_f32(w[1], H, st);
  c->v_pos = c->copy_f32(w[2], static_cast<size_t>(T) * H, st);
  c->v_pre_g = c->copy_f32(w[3], H, st); c->v_pre_b = c->copy_f32(w[4], H, st);
  c->v_post_g = c->copy_f32(w[5], H, st); c->v_post_b = c->copy_f32(w[6], H, st);
  { const void* wp[1] = {w[7]}; int np[1] = {vc->proj}; ok = ok && c->make_linear(&c->vis.proj, wp, nullptr, np, 1, H, s, st); }
  ok = ok && c->v_cls && c->v_pos && c->v_pre_g && c->v_pre_b && c->v_post_g && c->v_post_b;
  c->vis.layers.clear();
  for (int l = 0; ok && l < vc->layers; ++l) {
    Layer ly;
    ok = c->make_clip_layer(&ly, w + 8 + CONZIC_PER_LAYER * l, H, F, s, st, nullptr);
    c->vis.layers.push_back(ly);
  }
  ok = ok && cuda_ok(cudaStreamSynchronize(st), "set_vision sync");
  c->has_vision = ok;
  return ok ? 0 : -3;
}

size_t conzic_vision_workspace_bytes(const conzic_ctx* c, int B) {
  if (!c || !c->has_vision) return 0;
  return make_vplan(c, nullptr, B).bytes;
}

int conzic_clip_image_encode(conzic_ctx* c, const float* pix, int B, float* out, void* ws, size_t ws_bytes, void* stream) {
  if (!c || !pix || !out || !ws) { set_error("clip_image_encode: null argument"); return -1; }
  if (!c->has_vision) { set_error("clip_image_encode: conzic_set_vision has not been called"); return -1; }
  if (B < 1) { set_error("clip_image_encode: B < 1"); return -1; }
  VPlan p = make_vplan(c, ws, B);
  if (ws_bytes < p.bytes) { set_error("clip_image_encode: workspace too small, need " + std::to_string(p.bytes)); return -1; }
  StateScope scope(&c->state);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const conzic_vision_config& v = c->vcfg;
  const GemmOpts& go = c->vis.gopt;
  const int s = c->vis.split, H = v.hidden, F = v.ffn, Kp = 3 * v.patch * v.patch;
  const int g2 = (v.image_size / v.patch) * (v.image_size / v.patch), T = g2 + 1;
  const int M = B * T, ldh = H * (1 + s), ldf = F * (1 + s);
  set_pdl_now(1);
  set_phase(PHASE_IMAGE);
  // patch embedding = GEMM over unfolded patches (conv with stride = kernel, no bias)
  launch_im2col(pix, B, v.image_size, v.patch, p.patches, Kp * (1 + s), s, st);
  Act pa{p.patches, Kp * (1 + s), Kp};
  Epi e0; e0.out_f32 = p.pe; e0.ldo_f32 = H;
  if (!launch_linear(pa, B * g2, c->v_patch, e0, go, st)) return -4;
  launch_vision_embed(p.pe, c->v_cls, c->v_pos, B, T, H, p.x, st);
  LNArgs pre{p.x, nullptr, M, H, c->v_pre_g, c->v_pre_b, v.ln_eps, p.x, nullptr, 0, 0};  // in place: a warp owns a row
  launch_layernorm(pre, st);
  for (size_t l = 0; l < c->vis.layers.size(); ++l) {
    const Layer& ly = c->vis.layers[l];
    LNArgs ln1{p.x, nullptr, M, H, ly.ln1_g, ly.ln1_b, v.ln_eps, nullptr, p.h, ldh, s};
    launch_layernorm(ln1, st);
    Act h{p.h, ldh, H};
    Epi e;
    if (s) e = epi_f32_out(ly.qkv, static_cast<float*>(p.qkv), 3 * H, nullptr, 0, ACT_NONE);
    else   e = epi_act_out(ly.qkv, static_cast<bf16*>(p.qkv), 3 * H, 0, ACT_NONE);
    if (!launch_linear(h, M, ly.qkv, e, go, st)) return -4;
    AttnArgs at;
    at.qkv = p.qkv; at.ld_qkv = 3 * H; at.qkv_f32 = s; at.p0 = nullptr;
    at.B = B; at.P = 0; at.K = 1; at.S = T; at.H = H; at.heads = v.heads; at.causal = 0;
    at.scale = 0.125f; at.out_act = p.attn; at.ld_act = ldh; at.split = s;
    if (!launch_attention(at, st)) return -4;
    Act a{p.attn, ldh, H};
    if (!launch_linear(a, M, ly.o, epi_f32_out(ly.o, p.x, H, p.x, H, ACT_NONE), go, st)) return -4;
    LNArgs ln2{p.x, nullptr, M, H, ly.ln2_g, ly.ln2_b, v.ln_eps, nullptr, p.h, ldh, s};
    launch_layernorm(ln2, st);
    if (!launch_linear(h, M, ly.f1, epi_act_out(ly.f1, p.ffn, ldf, F, ACT_QUICK_GELU), go, st)) return -4;
    Act f{p.ffn, ldf, F};
    if (!launch_linear(f, M, ly.f2, epi_f32_out(ly.f2, p.x, H, p.x, H, ACT_NONE), go, st)) return -4;
  }
  // pooled = post_layernorm(hidden[:, 0]) (row stride T*H) -> visual_projection
  LNArgs post{p.x, nullptr, B, H, c->v_post_g, c->v_post_b, v.ln_eps, nullptr, p.pool, ldh, s};
  post.x_row_stride = static_cast<size_t>(T) * H;
  launch_layernorm(post, st);
  Act pooled{p.pool, ldh, H};
  Epi e;
  e.out_f32 = out; e.ldo_f32 = v.proj;
  if (!launch_linear(pooled, B, c->vis.proj, e, go, st)) return -4;
  set_phase(PHASE_OTHER);
  return cuda_ok(cudaGetLastError(), "clip_image_encode") ? 0 : -4;
}

int conzic_image_preprocess(conzic_ctx* c, const uint8_t* images, int n, int H, int W, const conzic_resize_axis* hz,
                            const conzic_resize_axis* vt, int row_lo, int row_hi, const float* mean3, const float* std3,
                            float* pixel_values, void* ws, size_t ws_bytes, void* stream) {
  if (!c || !images || !hz || !vt || !mean3 || !std3 || !pixel_values || !ws) { set_error("image_preprocess: null argument"); return -1; }
  if (n < 1 || H < 1 || W < 1 || row_lo < 0 || row_hi > H || row_lo >= row_hi || hz->n_out < 1 || vt->n_out < 1 ||
      (!hz->identity && (hz->precision < 1 || hz->precision > 22)) || (!vt->identity && (vt->precision < 1 || vt->precision > 22))) {
    set_error("image_preprocess: bad geometry");
    return -1;
  }
  const size_t need = static_cast<size_t>(n) * (row_hi - row_lo) * hz->n_out * 3;
  if (ws_bytes < need) { set_error("image_preprocess: workspace too small, need " + std::to_string(need)); return -1; }
  StateScope scope(&c->state);
  set_phase(PHASE_IMAGE);
  set_pdl_now(1);
  ImagePreArgs a{};
  a.src = images; a.n = n; a.H = H; a.W = W;
  a.hz = ResizeAxis{hz->weights, hz->first, hz->count, hz->taps, hz->precision, hz->n_out, hz->identity};
  a.vt = ResizeAxis{vt->weights, vt->first, vt->count, vt->taps, vt->precision, vt->n_out, vt->identity};
  a.row_lo = row_lo; a.row_hi = row_hi;
  for (int i = 0; i < 3; ++i) { a.mean[i] = mean3[i]; a.std[i] = std3[i]; }
  a.tmp = static_cast<uint8_t*>(ws); a.out = pixel_values;
  launch_image_preprocess(a, static_cast<cudaStream_t>(stream));
  set_phase(PHASE_OTHER);
  return cuda_ok(cudaGetLastError(), "image_preprocess") ? 0 : -4;
}

}  // extern "C"
