// Non-GEMM pieces of the two towers: operand conversion, embeddings, LayerNorm, short-sequence attention.
// All are HBM/L2-bandwidth kernels: one warp per token row, float4 / bf16x2 vector accesses, fp32 math.
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

// Write 4 consecutive activations (col % 4 == 0) as bf16 (and the bf16 remainder plane when split).
__device__ __forceinline__ void store_act4(bf16* row, int K, int split, int col, float4 v) {
  __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&a);
  u.y = *reinterpret_cast<uint32_t*>(&b);
  *reinterpret_cast<uint2*>(row + col) = u;
  if (split) {
    __nv_bfloat162 c = __floats2bfloat162_rn(v.x - __bfloat162float(a.x), v.y - __bfloat162float(a.y));
    __nv_bfloat162 d = __floats2bfloat162_rn(v.z - __bfloat162float(b.x), v.w - __bfloat162float(b.y));
    u.x = *reinterpret_cast<uint32_t*>(&c);
    u.y = *reinterpret_cast<uint32_t*>(&d);
    *reinterpret_cast<uint2*>(row + K + col) = u;
  }
}

__global__ void f32_to_act_kernel(const float* __restrict__ src, int rows, int K, int lds, bf16* __restrict__ dst,
                                  int ldd, int split) {
  const int kq = K >> 2;
  const size_t total = static_cast<size_t>(rows) * kq;
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / kq), c = static_cast<int>(i % kq) * 4;
    float4 v = *reinterpret_cast<const float4*>(src + static_cast<size_t>(r) * lds + c);
    store_act4(dst + static_cast<size_t>(r) * ldd, K, split, c, v);
  }
}

// LayerNorm of one row held as NV float4 per lane (H = 128 * NV).  Two-pass variance like torch.
template <int NV>
__device__ __forceinline__ void ln_row(float4 (&v)[NV], int H, const float* __restrict__ gamma,
                                       const float* __restrict__ beta, float eps, int lane, float* out_f32,
                                       bf16* out_act, int split) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  const float mean = warp_sum(s) / H;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float rstd = rsqrtf(warp_sum(q) / H + eps);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int col = (i * 32 + lane) * 4;
    float4 g = __ldg(reinterpret_cast<const float4*>(gamma + col));
    float4 b = __ldg(reinterpret_cast<const float4*>(beta + col));
    float4 y;
    y.x = (v[i].x - mean) * rstd * g.x + b.x;
    y.y = (v[i].y - mean) * rstd * g.y + b.y;
    y.z = (v[i].z - mean) * rstd * g.z + b.z;
    y.w = (v[i].w - mean) * rstd * g.w + b.w;
    if (out_f32) *reinterpret_cast<float4*>(out_f32 + col) = y;
    if (out_act) store_act4(out_act, H, split, col, y);
  }
}

template <int NV>
__global__ void layernorm_kernel(LNArgs a) {
  const int warps = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int r = blockIdx.x * warps + (threadIdx.x >> 5); r < a.n_rows; r += gridDim.x * warps) {
    const int src = a.rows ? a.rows[r] : r;
    const float* x = a.x + static_cast<size_t>(src) * a.H;
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = *reinterpret_cast<const float4*>(x + (i * 32 + lane) * 4);
    ln_row<NV>(v, a.H, a.gamma, a.beta, a.eps, lane, a.out_f32 ? a.out_f32 + static_cast<size_t>(r) * a.H : nullptr,
               a.out_act ? a.out_act + static_cast<size_t>(r) * a.ld_act : nullptr, a.split);
  }
}

// HF:models/bert/modeling_bert.py:72-112: word + token_type(0) + position, LayerNorm.
template <int NV>
__global__ void bert_embed_ln_kernel(const int64_t* __restrict__ inp, int rows, int L, const float* __restrict__ word,
                                     const float* __restrict__ pos, const float* __restrict__ type,
                                     const float* __restrict__ gamma, const float* __restrict__ beta, float eps, int H,
                                     float* __restrict__ x_f32, bf16* __restrict__ act, int ld_act, int split) {
  const int warps = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int r = blockIdx.x * warps + (threadIdx.x >> 5); r < rows; r += gridDim.x * warps) {
    const int64_t id = inp[r];
    const int t = r % L;
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int col = (i * 32 + lane) * 4;
      float4 w = *reinterpret_cast<const float4*>(word + static_cast<size_t>(id) * H + col);
      float4 ty = __ldg(reinterpret_cast<const float4*>(type + col));
      float4 p = *reinterpret_cast<const float4*>(pos + static_cast<size_t>(t) * H + col);
      v[i] = make_float4((w.x + ty.x) + p.x, (w.y + ty.y) + p.y, (w.z + ty.z) + p.z, (w.w + ty.w) + p.w);
    }
    ln_row<NV>(v, H, gamma, beta, eps, lane, x_f32 + static_cast<size_t>(r) * H,
               act + static_cast<size_t>(r) * ld_act, split);
  }
}

// HF:models/clip/modeling_clip.py:253-256: token embedding + position embedding (no LayerNorm).
__global__ void clip_embed_kernel(const int32_t* __restrict__ ids_prefix, const int32_t* __restrict__ ids_suffix,
                                  const int32_t* __restrict__ p0, int B, int P, int K, int S, int maxpos,
                                  const float* __restrict__ tok, const float* __restrict__ pos, int H,
                                  float* __restrict__ x) {
  const int warps = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_pre = B * P;
  const int rows = n_pre + B * K * S;
  for (int r = blockIdx.x * warps + (threadIdx.x >> 5); r < rows; r += gridDim.x * warps) {
    int id, position;
    if (r < n_pre) {
      id = ids_prefix[r];
      position = r % P;
    } else {
      const int idx = r - n_pre;
      const int b = idx / (K * S), s = idx % S;
      id = ids_suffix[idx];
      position = (p0 ? p0[b] : 0) + s;
    }
    position = min(position, maxpos - 1);
    const float* te = tok + static_cast<size_t>(id) * H;
    const float* pe = pos + static_cast<size_t>(position) * H;
    float* o = x + static_cast<size_t>(r) * H;
    for (int c = lane * 4; c < H; c += 128) {
      float4 a = *reinterpret_cast<const float4*>(te + c);
      float4 b = *reinterpret_cast<const float4*>(pe + c);
      *reinterpret_cast<float4*>(o + c) = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Attention for many very short sequences (head_dim 64).  One warp per (sequence, head): K^T and V of the
// sequence (shared prefix + own rows) are staged in shared memory as fp32, then for each query row the
// lanes own keys for the score/softmax phase and own output dims for the P*V phase.
// HF:models/clip/modeling_clip.py:259-330 (causal), HF:models/bert/modeling_bert.py:129-137 (no mask).
// ---------------------------------------------------------------------------------------------------
constexpr int HD = 64;
constexpr int MAX_SLOTS = 3;  // up to 96 keys

template <bool F32>
__device__ __forceinline__ float2 load_pair(const void* base, size_t elem) {
  if (F32) {
    return *reinterpret_cast<const float2*>(reinterpret_cast<const float*>(base) + elem);
  } else {
    __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(reinterpret_cast<const bf16*>(base) + elem);
    return make_float2(__bfloat162float(h.x), __bfloat162float(h.y));
  }
}

template <bool F32>
__global__ void attention_kernel(AttnArgs a, int nk_cap) {
  extern __shared__ float sm[];
  const int warps = blockDim.x >> 5;
  const int wib = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nkp = nk_cap | 1;  // odd stride for the transposed K
  const int per_warp = (HD * nkp + HD * nk_cap + HD + nkp + 3) & ~3;  // keep every warp's slice 16B aligned
  float* Kt = sm + static_cast<size_t>(wib) * per_warp;  // [64][nkp]
  float* Vs = Kt + HD * nkp;                               // [nk][64]
  float* qs = Vs + HD * nk_cap;                            // [64]
  float* ps = qs + HD;                                     // [nkp]

  const int n_pre_seq = a.P > 0 ? a.B : 0;
  const int n_seq = n_pre_seq + a.B * a.K;
  const long long n_task = static_cast<long long>(n_seq) * a.heads;
  const int H = a.H;
  const int n_pre_rows = a.B * a.P;

  for (long long task = static_cast<long long>(blockIdx.x) * warps + wib; task < n_task;
       task += static_cast<long long>(gridDim.x) * warps) {
    const int head = static_cast<int>(task % a.heads);
    const int seq = static_cast<int>(task / a.heads);
    int pl, nq, own_base, b;
    if (seq < n_pre_seq) {
      b = seq; pl = 0; nq = a.P; own_base = b * a.P;
    } else {
      const int bk = seq - n_pre_seq;
      b = bk / a.K;
      pl = a.P > 0 ? (a.p0 ? min(a.p0[b], a.P) : a.P) : 0;
      nq = a.S;
      own_base = n_pre_rows + bk * a.S;
    }
    const int nk = pl + nq;
    const int pre_base = b * a.P;
    // ---- stage K^T and V
    for (int j = 0; j < nk; ++j) {
      const int row = j < pl ? pre_base + j : own_base + (j - pl);
      const size_t e = static_cast<size_t>(row) * a.ld_qkv + head * HD + 2 * lane;
      float2 k2 = load_pair<F32>(a.qkv, e + H);
      float2 v2 = load_pair<F32>(a.qkv, e + 2 * H);
      Kt[(2 * lane) * nkp + j] = k2.x;
      Kt[(2 * lane + 1) * nkp + j] = k2.y;
      *reinterpret_cast<float2*>(Vs + j * HD + 2 * lane) = v2;
    }
    __syncwarp();
    for (int t = 0; t < nq; ++t) {
      const int qrow = own_base + t;
      float2 q2 = load_pair<F32>(a.qkv, static_cast<size_t>(qrow) * a.ld_qkv + head * HD + 2 * lane);
      *reinterpret_cast<float2*>(qs + 2 * lane) = q2;
      __syncwarp();
      const int nvis = a.causal ? pl + t + 1 : nk;
      float sc[MAX_SLOTS];
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < MAX_SLOTS; ++c) {
        const int j = lane + 32 * c;
        float s = -INFINITY;
        if (j < nvis) {
          float acc = 0.f;
#pragma unroll 16
          for (int d = 0; d < HD; ++d) acc = fmaf(qs[d], Kt[d * nkp + j], acc);
          s = acc * a.scale;
        }
        sc[c] = s;
        mx = fmaxf(mx, s);
      }
      mx = warp_max(mx);
      float sum = 0.f;
#pragma unroll
      for (int c = 0; c < MAX_SLOTS; ++c) {
        const int j = lane + 32 * c;
        const float e = j < nvis ? expf(sc[c] - mx) : 0.f;
        sc[c] = e;
        sum += e;
      }
      sum = warp_sum(sum);
      const float inv = 1.0f / sum;
#pragma unroll
      for (int c = 0; c < MAX_SLOTS; ++c) {
        const int j = lane + 32 * c;
        if (j < nvis) ps[j] = sc[c] * inv;
      }
      __syncwarp();
      float o0 = 0.f, o1 = 0.f;
      for (int j = 0; j < nvis; ++j) {
        const float pj = ps[j];
        float2 v2 = *reinterpret_cast<const float2*>(Vs + j * HD + 2 * lane);
        o0 = fmaf(pj, v2.x, o0);
        o1 = fmaf(pj, v2.y, o1);
      }
      bf16* orow = a.out_act + static_cast<size_t>(qrow) * a.ld_act;
      const int col = head * HD + 2 * lane;
      __nv_bfloat162 h = __floats2bfloat162_rn(o0, o1);
      *reinterpret_cast<__nv_bfloat162*>(orow + col) = h;
      if (a.split) {
        *reinterpret_cast<__nv_bfloat162*>(orow + H + col) =
            __floats2bfloat162_rn(o0 - __bfloat162float(h.x), o1 - __bfloat162float(h.y));
      }
      __syncwarp();
    }
  }
}

int g_num_sms = 0;
int num_sms() {
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

int row_grid(int rows, int warps_per_block) {
  const int need = (rows + warps_per_block - 1) / warps_per_block;
  const int cap = num_sms() * 8;
  return need < cap ? (need > 0 ? need : 1) : cap;
}

}  // namespace

void launch_f32_to_act(const float* src, int rows, int K, int lds, bf16* dst, int ldd, int split, cudaStream_t st) {
  ++g_launches;
  ProfScope prof_(CAT_MISC, 0, st);
  const size_t total = static_cast<size_t>(rows) * (K / 4);
  int grid = static_cast<int>((total + 255) / 256);
  const int cap = num_sms() * 16;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  f32_to_act_kernel<<<grid, 256, 0, st>>>(src, rows, K, lds, dst, ldd, split);
}

void launch_layernorm(const LNArgs& a, cudaStream_t st) {
  ++g_launches;
  ProfScope prof_(CAT_LN, static_cast<double>(a.n_rows) * a.H, st);
  if (a.n_rows <= 0) return;
  const int grid = row_grid(a.n_rows, 8);
  switch (a.H / 128) {
    case 4: layernorm_kernel<4><<<grid, 256, 0, st>>>(a); break;
    case 6: layernorm_kernel<6><<<grid, 256, 0, st>>>(a); break;
    case 8: layernorm_kernel<8><<<grid, 256, 0, st>>>(a); break;
    default: set_error("layernorm: hidden size must be 512, 768 or 1024"); break;
  }
}

void launch_bert_embed_ln(const int64_t* inp, int rows, int L, const float* word, const float* pos, const float* type,
                          const float* gamma, const float* beta, float eps, int H, float* x_f32, bf16* act, int ld_act,
                          int split, cudaStream_t st) {
  ++g_launches;
  ProfScope prof_(CAT_EMBED, static_cast<double>(rows) * H, st);
  const int grid = row_grid(rows, 8);
  switch (H / 128) {
    case 4: bert_embed_ln_kernel<4><<<grid, 256, 0, st>>>(inp, rows, L, word, pos, type, gamma, beta, eps, H, x_f32, act, ld_act, split); break;
    case 6: bert_embed_ln_kernel<6><<<grid, 256, 0, st>>>(inp, rows, L, word, pos, type, gamma, beta, eps, H, x_f32, act, ld_act, split); break;
    case 8: bert_embed_ln_kernel<8><<<grid, 256, 0, st>>>(inp, rows, L, word, pos, type, gamma, beta, eps, H, x_f32, act, ld_act, split); break;
    default: set_error("bert_embed: hidden size must be 512, 768 or 1024"); break;
  }
}

void launch_clip_embed(const int32_t* ids_prefix, const int32_t* ids_suffix, const int32_t* p0, int B, int P, int K,
                       int S, int maxpos, const float* tok, const float* pos, int H, float* x_f32, cudaStream_t st) {
  ++g_launches;
  ProfScope prof_(CAT_EMBED, 0, st);
  const int rows = B * P + B * K * S;
  if (rows <= 0) return;
  clip_embed_kernel<<<row_grid(rows, 8), 256, 0, st>>>(ids_prefix, ids_suffix, p0, B, P, K, S, maxpos, tok, pos, H,
                                                       x_f32);
}

bool launch_attention(const AttnArgs& a, cudaStream_t st) {
  ++g_launches;
  ProfScope prof_(CAT_ATTN, 0, st);
  if (a.H != a.heads * HD) {
    set_error("attention: head_dim must be 64");
    return false;
  }
  const int nk_cap = a.P + a.S;
  if (nk_cap > 32 * MAX_SLOTS) {
    set_error("attention: more than 96 keys per sequence is not supported (sentence_len too large)");
    return false;
  }
  const int nkp = nk_cap | 1;
  const size_t per_warp = static_cast<size_t>((HD * nkp + HD * nk_cap + HD + nkp + 3) & ~3) * sizeof(float);
  int warps = static_cast<int>((200 * 1024) / per_warp);
  if (warps > 8) warps = 8;
  if (warps < 1) warps = 1;
  const size_t smem = per_warp * warps;
  const long long n_seq = (a.P > 0 ? a.B : 0) + static_cast<long long>(a.B) * a.K;
  const long long n_task = n_seq * a.heads;
  long long grid = (n_task + warps - 1) / warps;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  static size_t configured[2] = {0, 0};
  const int which = a.qkv_f32 ? 1 : 0;
  if (smem > 48 * 1024 && smem > configured[which]) {
    cudaError_t e = a.qkv_f32 ? cudaFuncSetAttribute(attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024)
                              : cudaFuncSetAttribute(attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    if (!cuda_ok(e, "cudaFuncSetAttribute(attention)")) return false;
    configured[which] = 220 * 1024;
  }
  if (a.qkv_f32)
    attention_kernel<true><<<static_cast<int>(grid), warps * 32, smem, st>>>(a, nk_cap);
  else
    attention_kernel<false><<<static_cast<int>(grid), warps * 32, smem, st>>>(a, nk_cap);
  return cuda_ok(cudaGetLastError(), "attention launch");
}

}  // namespace conzic
