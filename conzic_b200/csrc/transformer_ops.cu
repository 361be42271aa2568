// Non-GEMM pieces of the two towers: operand conversion, embeddings, LayerNorm, short-sequence attention.
// All are HBM/L2-bandwidth kernels: one warp per token row, float4 / bf16x2 vector accesses, fp32 math.
#include <cstdlib>

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
  PDL_ENTRY();
  const int warps = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int r = blockIdx.x * warps + (threadIdx.x >> 5); r < a.n_rows; r += gridDim.x * warps) {
    const int src = a.rows ? a.rows[r] : r;
    const float* x = a.x + static_cast<size_t>(src) * (a.x_row_stride ? a.x_row_stride : static_cast<size_t>(a.H));
    float4 v[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = *reinterpret_cast<const float4*>(x + (i * 32 + lane) * 4);
    if (a.partials) {  // residual + bias + split-K partial sums of the GEMM that precedes this LayerNorm
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int col = (i * 32 + lane) * 4;
        if (a.add_bias) {
          const float4 b = __ldg(reinterpret_cast<const float4*>(a.add_bias + col));
          v[i].x += b.x; v[i].y += b.y; v[i].z += b.z; v[i].w += b.w;
        }
        // up to 4 split-K partial sums (launch_linear's cap): all loads are issued before the first add, so a row
        // costs one memory latency instead of n_parts; the adds keep the order z = 0, 1, 2, 3
        float4 q[4];
#pragma unroll
        for (int z = 0; z < 4; ++z) {
          q[z] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (z < a.n_parts)
            q[z] = *reinterpret_cast<const float4*>(a.partials + z * a.part_stride + static_cast<size_t>(r) * a.H + col);
        }
#pragma unroll
        for (int z = 0; z < 4; ++z) {
          if (z < a.n_parts) { v[i].x += q[z].x; v[i].y += q[z].y; v[i].z += q[z].z; v[i].w += q[z].w; }
        }
      }
    }
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
  PDL_ENTRY();
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
                                  float* __restrict__ x, const float* __restrict__ ln_g, const float* __restrict__ ln_b,
                                  float ln_eps, bf16* __restrict__ ln_out, int ln_ld,
                                  const int32_t* __restrict__ cand_img, int n_cand) {
  PDL_ENTRY();
  const int warps = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_pre = B * P;
  const int rows = n_pre + (cand_img ? n_cand : B * K) * S;
  for (int r = blockIdx.x * warps + (threadIdx.x >> 5); r < rows; r += gridDim.x * warps) {
    int id, position;
    if (r < n_pre) {
      id = ids_prefix[r];
      position = r % P;
    } else {
      const int idx = r - n_pre;
      const int b = cand_img ? cand_img[idx / S] : idx / (K * S), s = idx % S;
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
    if (ln_out) {
      // LayerNorm 1 of the first block on the row this warp just produced (H == 512): same arithmetic as
      // layernorm_kernel<4> on the same fp32 values, without re-reading x from HBM
      float4 v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c = (i * 32 + lane) * 4;
        const float4 a = *reinterpret_cast<const float4*>(te + c);
        const float4 b = *reinterpret_cast<const float4*>(pe + c);
        v[i] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
      }
      ln_row<4>(v, H, ln_g, ln_b, ln_eps, lane, nullptr, ln_out + static_cast<size_t>(r) * ln_ld, 0);
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

constexpr int QS_LD = HD + 2;  // query rows 2 banks apart: the queries of one pass are read conflict free

// The queries of one (sequence, head) task, W lanes per query: W = 32 is one query per pass (up to 96 keys, 3 per
// lane); short sequences (<= 16 / <= 8 keys: BERT, the certified re-score) take 2 / 4 queries per pass, which halves /
// quarters the number of dependent passes of this latency-bound kernel.  Every query sees the same arithmetic in each
// form -- 4 interleaved partial dot products summed (a0 + a1) + (a2 + a3), probabilities x V accumulated key by key --
// and the form depends on the sequence's own key count only, so a caption gets the same bits wherever it is encoded.
template <int W>
__device__ __forceinline__ void attention_queries(const AttnArgs& a, const float* Kt, int nkp, const float* Vs, const float* Qs,
                                                  float* ps, int nq, int pl, int nk, int own_base, int head, int lane) {
  constexpr int G = 32 / W, DPL = HD / W, NSLOT = W == 32 ? MAX_SLOTS : 1;
  const int g = lane / W, jl = lane % W;
  const int H = a.H;
  for (int t0 = 0; t0 < nq; t0 += G) {
    const bool live = t0 + g < nq;
    const int t = live ? t0 + g : nq - 1;  // idle lane groups repeat the last query and store nothing
    const float* qs = Qs + t * QS_LD;
    const int nvis = a.causal ? pl + t + 1 : nk;
    float sc[NSLOT];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < NSLOT; ++c) {
      const int j = jl + W * c;
      float s = -INFINITY;
      if (j < nvis) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        const float* kt = Kt + j;
#pragma unroll 4
        for (int d = 0; d < HD; d += 4) {
          a0 = fmaf(qs[d], kt[d * nkp], a0);
          a1 = fmaf(qs[d + 1], kt[(d + 1) * nkp], a1);
          a2 = fmaf(qs[d + 2], kt[(d + 2) * nkp], a2);
          a3 = fmaf(qs[d + 3], kt[(d + 3) * nkp], a3);
        }
        s = ((a0 + a1) + (a2 + a3)) * a.scale;
      }
      sc[c] = s;
      mx = fmaxf(mx, s);
    }
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < NSLOT; ++c) {
      const int j = jl + W * c;
      const float e = j < nvis ? expf(sc[c] - mx) : 0.f;
      sc[c] = e;
      sum += e;
    }
#pragma unroll
    for (int o = W / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float inv = 1.0f / sum;
    float* pq = ps + g * W;  // W == 32: g == 0 and the row holds up to nkp probabilities
#pragma unroll
    for (int c = 0; c < NSLOT; ++c) {
      const int j = jl + W * c;
      if (W == 32 ? (j < nvis) : true) pq[j] = j < nvis ? sc[c] * inv : 0.f;  // short forms: zeros past the visible keys
    }
    __syncwarp();
    // probabilities x V: DPL output dims per lane; the short forms run to the longest query of the pass (the zeros
    // written above add exactly nothing)
    const int nloop = W == 32 ? nvis : (a.causal ? min(nk, pl + min(t0 + G, nq)) : nk);
    float o[DPL];
#pragma unroll
    for (int e = 0; e < DPL; ++e) o[e] = 0.f;
    for (int j = 0; j < nloop; ++j) {
      const float pj = pq[j];
      const float* v = Vs + j * HD + jl * DPL;
#pragma unroll
      for (int e = 0; e < DPL; ++e) o[e] = fmaf(pj, v[e], o[e]);
    }
    if (live) {
      bf16* orow = a.out_act + static_cast<size_t>(own_base + t) * a.ld_act + head * HD + jl * DPL;
#pragma unroll
      for (int e = 0; e < DPL; e += 2) {
        __nv_bfloat162 h = __floats2bfloat162_rn(o[e], o[e + 1]);
        *reinterpret_cast<__nv_bfloat162*>(orow + e) = h;
        if (a.split)
          *reinterpret_cast<__nv_bfloat162*>(orow + H + e) =
              __floats2bfloat162_rn(o[e] - __bfloat162float(h.x), o[e + 1] - __bfloat162float(h.y));
      }
    }
    __syncwarp();
  }
}

template <bool F32>
__global__ void attention_kernel(AttnArgs a, int nk_cap, int nq_cap) {
  PDL_ENTRY();
  extern __shared__ float sm[];
  const int warps = blockDim.x >> 5;
  const int wib = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nkp = nk_cap | 1;  // odd stride for the transposed K
  const int psz = nkp > 32 ? nkp : 32;
  const int per_warp = (HD * nkp + HD * nk_cap + QS_LD * nq_cap + psz + 3) & ~3;  // keep every warp's slice 16B aligned
  float* Kt = sm + static_cast<size_t>(wib) * per_warp;  // [64][nkp]
  float* Vs = Kt + HD * nkp;                               // [nk][64]
  float* Qs = Vs + HD * nk_cap;                            // [nq][QS_LD]
  float* ps = Qs + QS_LD * nq_cap;                         // [max(nkp, 32)]

  const int n_pre_seq = a.P > 0 ? a.B : 0;
  const int n_seq = n_pre_seq + (a.cand_img ? a.n_cand : a.B * a.K);
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
      b = a.cand_img ? a.cand_img[bk] : bk / a.K;
      pl = a.P > 0 ? (a.p0 ? min(a.p0[b], a.P) : a.P) : 0;
      nq = a.S;
      own_base = n_pre_rows + bk * a.S;
    }
    const int nk = pl + nq;
    const int pre_base = b * a.P;
    // ---- stage K^T, V and the queries: 8 rows per round, every global load of a round issued before its first
    // shared-memory store, so the warp pays one memory latency per 8 rows (this kernel is latency bound: a few
    // hundred short sequences; the row-by-row form waited on every row and again on every query)
    for (int j0 = 0; j0 < nk; j0 += 8) {
      float2 kk[8], vv[8], qq[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = min(j0 + u, nk - 1);  // clamp instead of predicate: keeps the arrays in registers
        const int row = j < pl ? pre_base + j : own_base + (j - pl);
        const size_t e = static_cast<size_t>(row) * a.ld_qkv + head * HD + 2 * lane;
        kk[u] = load_pair<F32>(a.qkv, e + H);
        vv[u] = load_pair<F32>(a.qkv, e + 2 * H);
        qq[u] = load_pair<F32>(a.qkv, e);
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int j = j0 + u;
        if (j < nk) {
          Kt[(2 * lane) * nkp + j] = kk[u].x;
          Kt[(2 * lane + 1) * nkp + j] = kk[u].y;
          *reinterpret_cast<float2*>(Vs + j * HD + 2 * lane) = vv[u];
          if (j >= pl) *reinterpret_cast<float2*>(Qs + (j - pl) * QS_LD + 2 * lane) = qq[u];  // own rows are the queries
        }
      }
    }
    __syncwarp();
    if (nk <= 8) attention_queries<8>(a, Kt, nkp, Vs, Qs, ps, nq, pl, nk, own_base, head, lane);
    else if (nk <= 16) attention_queries<16>(a, Kt, nkp, Vs, Qs, ps, nq, pl, nk, own_base, head, lane);
    else attention_queries<32>(a, Kt, nkp, Vs, Qs, ps, nq, pl, nk, own_base, head, lane);
    __syncwarp();
  }
}


// ---------------------------------------------------------------------------------------------------
// bf16 attention for many very short sequences on the warp-level tensor-core path (mma.sync m16n8k16; the
// sequences are far too short -- <= 16 queries x <= 32 keys typically -- to fill a 128-row tcgen05 tile).
// One warp owns (image b, head h, a group of up to 8 candidates).  It stages the K / V rows of the image's
// shared prefix once in its private shared-memory slice, then for every candidate appends the candidate's own
// rows, forms S = Q K^T with ldmatrix-fed MMAs, does the masked softmax on the accumulator fragments in fp32,
// and multiplies by V with the probabilities split into bf16 hi + lo parts (two MMAs) so P keeps fp32
// accuracy.  All global accesses are whole 128-byte head rows (8 lanes x 16 B); the smem tiles are XOR
// swizzled so both the 16-byte row writes and the ldmatrix reads are conflict free.
// ---------------------------------------------------------------------------------------------------

__device__ __forceinline__ uint32_t sw_off(int row, int piece) { return row * 128 + ((piece ^ (row & 7)) << 4); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                               uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
               "{%0, %1, %2, %3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
// hi = bf16(x), lo = bf16(x - hi) for a pair of probabilities
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  hi = *reinterpret_cast<uint32_t*>(&h);
  lo = pack_bf16(a - __bfloat162float(h.x), b - __bfloat162float(h.y));
}

// Copies `n_rows` head rows (64 bf16 = 8 pieces of 16 B) into a swizzled smem tile starting at tile row `dst0`.
// Source row i is global row (i < split ? base0 + i : base1 + i - split).
__device__ __forceinline__ void att_load_rows(const bf16* __restrict__ src, int ld, int col0, int base0, int split,
                                              int base1, int n_rows, uint8_t* tile, int dst0, int lane) {
  for (int idx = lane; idx < n_rows * 8; idx += 32) {
    const int i = idx >> 3, pc = idx & 7;
    const int grow = i < split ? base0 + i : base1 + (i - split);
    const uint4 v = *reinterpret_cast<const uint4*>(src + static_cast<size_t>(grow) * ld + col0 + pc * 8);
    *reinterpret_cast<uint4*>(tile + sw_off(dst0 + i, pc)) = v;
  }
}

// Batched variant for up to 16 rows: all loads are issued before the first shared-memory store, so the warp
// pays ONE global-memory latency per call (the plain loop above serialises load -> store per 4 rows).
// Rows are consecutive global rows starting at `base`.
__device__ __forceinline__ void att_fetch16(const bf16* __restrict__ src, int ld, int col0, int base, int n_rows,
                                            int lane, uint4 (&r)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = i * 32 + lane;
    const int row = min(idx >> 3, n_rows - 1);  // clamp instead of predicate: keeps r[] in registers, loads stay in bounds
    r[i] = *reinterpret_cast<const uint4*>(src + static_cast<size_t>(base + row) * ld + col0 + (idx & 7) * 8);
  }
}
__device__ __forceinline__ void att_put16(uint8_t* tile, int dst0, int n_rows, int lane, const uint4 (&r)[4]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = i * 32 + lane;
    if ((idx >> 3) < n_rows) *reinterpret_cast<uint4*>(tile + sw_off(dst0 + (idx >> 3), idx & 7)) = r[i];
  }
}

// NT key tiles of 8: up to NT*8 keys per tile; MINB: resident blocks per SM the register budget is sized for (the
// kernel is latency bound: 24 warps per SM instead of 16 when the tiles are small)
template <int NT, int MINB = 2>
__global__ void __launch_bounds__(256, MINB) attention_mma_kernel(AttnArgs a) {
  PDL_ENTRY();
  extern __shared__ __align__(1024) uint8_t att_smem[];
  constexpr int KV_BYTES = NT * 8 * 128;
  constexpr int VT = (NT + 1) / 2 * 2;  // P V consumes keys 16 at a time: the V tile is padded to an even tile count
  constexpr int V_BYTES = VT * 8 * 128;
  constexpr int WARP_BYTES = KV_BYTES + V_BYTES + 16 * 128;
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* sK = att_smem + wib * WARP_BYTES;
  uint8_t* sV = sK + KV_BYTES;
  uint8_t* sQ = sV + V_BYTES;  // 16 query rows; reused to stage the 16 output rows
  const uint32_t sK_a = static_cast<uint32_t>(__cvta_generic_to_shared(sK));
  const uint32_t sV_a = static_cast<uint32_t>(__cvta_generic_to_shared(sV));
  const uint32_t sQ_a = static_cast<uint32_t>(__cvta_generic_to_shared(sQ));

  const bf16* qkv = reinterpret_cast<const bf16*>(a.qkv);
  const int H = a.H, ld = a.ld_qkv;
  const int groups = (a.K + a.cand_per_task - 1) / a.cand_per_task;
  const int n_pre_tasks = a.P > 0 ? a.B * a.heads : 0;
  const long long n_tasks = n_pre_tasks + static_cast<long long>(a.B) * a.heads * groups;
  const long long task = static_cast<long long>(blockIdx.x) * (blockDim.x >> 5) + wib;
  if (task >= n_tasks) return;

  // A task: (image b, head, a range [k0, k1) of candidates).  `cpt` candidates share one 16-row query tile
  // when their sequences are short (cpt * nq <= 16): the tile then holds the image prefix keys followed by
  // the own keys of all cpt candidates, and the mask lets a query see the prefix and its own candidate only.
  int b, head, k0, k1, pl, nq, cpt;
  bool is_prefix;
  if (task < n_pre_tasks) {
    is_prefix = true;
    b = static_cast<int>(task / a.heads); head = static_cast<int>(task % a.heads);
    k0 = 0; k1 = 1; pl = 0; nq = a.P; cpt = 1;
  } else {
    is_prefix = false;
    // consecutive warps = consecutive heads of the same (image, candidate group): together they read whole
    // q / k / v rows (heads x 128 B contiguous), which keeps DRAM pages and L2 lines fully used
    const long long t = task - n_pre_tasks;
    head = static_cast<int>(t % a.heads);
    const long long bg = t / a.heads;
    const int grp = static_cast<int>(bg % groups);
    b = static_cast<int>(bg / groups);
    k0 = grp * a.cand_per_task; k1 = min(a.K, k0 + a.cand_per_task);
    pl = a.P > 0 ? (a.p0 ? min(a.p0[b], a.P) : a.P) : 0;
    nq = a.S; cpt = a.cpt;
  }
  const int col_q = head * 64, col_k = H + head * 64, col_v = 2 * H + head * 64;
  const int pre_base = b * a.P;
  const int n_pre_rows = a.B * a.P;

  // V rows that hold no key must be finite: P = 0 there, and 0 * garbage could be NaN
  for (int idx = lane; idx < (VT * 8 - pl) * 8; idx += 32)
    *reinterpret_cast<uint4*>(sV + sw_off(pl + (idx >> 3), idx & 7)) = make_uint4(0, 0, 0, 0);
  if (!is_prefix && pl > 0) {
    for (int r0 = 0; r0 < pl; r0 += 16) {
      uint4 rk[4], rv[4];
      const int nr = min(16, pl - r0);
      att_fetch16(qkv, ld, col_k, pre_base + r0, nr, lane, rk);
      att_fetch16(qkv, ld, col_v, pre_base + r0, nr, lane, rv);
      att_put16(sK, r0, nr, lane, rk);
      att_put16(sV, r0, nr, lane, rv);
    }
  }
  const int g = lane >> 2, qd = lane & 3;
  const int lm_r = lane & 7, lm_m = lane >> 3;  // ldmatrix: row within the 8x8 matrix, matrix index
  const float sl2 = a.scale * 1.4426950408889634f;

  for (int k = k0; k < k1; k += cpt) {
    const int nc = min(cpt, k1 - k);
    const int n_own = nc * nq;          // query rows = own key rows of this iteration (contiguous in memory)
    const int nk = pl + n_own;
    const int own_base = is_prefix ? pre_base : n_pre_rows + (b * a.K + k) * a.S;
    const bool single = n_own <= 16;
    __syncwarp();
    if (single) {
      uint4 rk[4], rv[4], rq[4];
      att_fetch16(qkv, ld, col_k, own_base, n_own, lane, rk);
      att_fetch16(qkv, ld, col_v, own_base, n_own, lane, rv);
      att_fetch16(qkv, ld, col_q, own_base, n_own, lane, rq);
      att_put16(sK, pl, n_own, lane, rk);
      att_put16(sV, pl, n_own, lane, rv);
      att_put16(sQ, 0, n_own, lane, rq);
    } else {
      for (int r0 = 0; r0 < n_own; r0 += 16) {
        uint4 rk[4], rv[4];
        const int nr = min(16, n_own - r0);
        att_fetch16(qkv, ld, col_k, own_base + r0, nr, lane, rk);
        att_fetch16(qkv, ld, col_v, own_base + r0, nr, lane, rv);
        att_put16(sK, pl + r0, nr, lane, rk);
        att_put16(sV, pl + r0, nr, lane, rv);
      }
    }
    for (int mt = 0; mt * 16 < n_own; ++mt) {
      const int q_rows = min(16, n_own - mt * 16);
      if (!single) {
        __syncwarp();
        uint4 rq[4];
        att_fetch16(qkv, ld, col_q, own_base + mt * 16, q_rows, lane, rq);
        att_put16(sQ, 0, q_rows, lane, rq);
      }
      __syncwarp();
      uint32_t qa[4][4];
#pragma unroll
      for (int ks = 0; ks < 4; ++ks)
        ldmatrix_x4(sQ_a + sw_off(lm_r + 8 * (lm_m & 1), 2 * ks + (lm_m >> 1)), qa[ks][0], qa[ks][1], qa[ks][2], qa[ks][3]);
      // ---- S = Q K^T
      float sc[NT][4];
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        sc[nt][0] = sc[nt][1] = sc[nt][2] = sc[nt][3] = 0.f;
        if (nt * 8 < nk) {
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            uint32_t b0, b1, b2, b3;
            ldmatrix_x4(sK_a + sw_off(nt * 8 + lm_r, 4 * hf + lm_m), b0, b1, b2, b3);
            mma_bf16_16816(sc[nt], qa[2 * hf][0], qa[2 * hf][1], qa[2 * hf][2], qa[2 * hf][3], b0, b1);
            mma_bf16_16816(sc[nt], qa[2 * hf + 1][0], qa[2 * hf + 1][1], qa[2 * hf + 1][2], qa[2 * hf + 1][3], b2, b3);
          }
        }
      }
      // ---- masked softmax over keys, rows g and g+8 of this query tile.  Query row r belongs to candidate
      // c = r / nq at position t = r % nq and sees keys [0, pl) and [pl + c*nq, pl + c*nq + t] (causal) or
      // every key (bidirectional, one sequence per tile).
      const int r0 = mt * 16 + g, r1 = r0 + 8;
      int lo0 = 0, hi0 = nk - 1, lo1 = 0, hi1 = nk - 1;
      if (a.causal) {
        const int c0 = r0 / nq, c1 = r1 / nq;
        lo0 = pl + c0 * nq; hi0 = min(nk - 1, lo0 + (r0 - c0 * nq));
        lo1 = pl + c1 * nq; hi1 = min(nk - 1, lo1 + (r1 - c1 * nq));
      }
      float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        const int j = nt * 8 + qd * 2;
        const bool v00 = a.causal ? (j < pl || (j >= lo0 && j <= hi0)) : (j < nk);
        const bool v01 = a.causal ? (j + 1 < pl || (j + 1 >= lo0 && j + 1 <= hi0)) : (j + 1 < nk);
        const bool v10 = a.causal ? (j < pl || (j >= lo1 && j <= hi1)) : (j < nk);
        const bool v11 = a.causal ? (j + 1 < pl || (j + 1 >= lo1 && j + 1 <= hi1)) : (j + 1 < nk);
        sc[nt][0] = v00 ? sc[nt][0] * sl2 : -INFINITY;   // scores in log2 units: exp(x) = exp2(x * log2 e)
        sc[nt][1] = v01 ? sc[nt][1] * sl2 : -INFINITY;
        sc[nt][2] = v10 ? sc[nt][2] * sl2 : -INFINITY;
        sc[nt][3] = v11 ? sc[nt][3] * sl2 : -INFINITY;
        m0 = fmaxf(m0, fmaxf(sc[nt][0], sc[nt][1]));
        m1 = fmaxf(m1, fmaxf(sc[nt][2], sc[nt][3]));
      }
      m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1)); m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
      m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1)); m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
      if (m0 == -INFINITY) m0 = 0.f;  // rows past the last query (never stored)
      if (m1 == -INFINITY) m1 = 0.f;
      float s0 = 0.f, s1 = 0.f;
#pragma unroll
      for (int nt = 0; nt < NT; ++nt) {
        sc[nt][0] = exp2f(sc[nt][0] - m0); sc[nt][1] = exp2f(sc[nt][1] - m0);
        sc[nt][2] = exp2f(sc[nt][2] - m1); sc[nt][3] = exp2f(sc[nt][3] - m1);
        s0 += sc[nt][0] + sc[nt][1];
        s1 += sc[nt][2] + sc[nt][3];
      }
      s0 += __shfl_xor_sync(0xffffffffu, s0, 1); s0 += __shfl_xor_sync(0xffffffffu, s0, 2);
      s1 += __shfl_xor_sync(0xffffffffu, s1, 1); s1 += __shfl_xor_sync(0xffffffffu, s1, 2);
      const float i0 = s0 > 0.f ? 1.0f / s0 : 0.f, i1 = s1 > 0.f ? 1.0f / s1 : 0.f;
      // ---- O = P V
      float o[8][4];
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) o[dt][0] = o[dt][1] = o[dt][2] = o[dt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < VT / 2; ++ks) {
        if (ks * 16 < nk) {
          uint32_t ph[4], plo[4];
          split_pair(sc[2 * ks][0] * i0, sc[2 * ks][1] * i0, ph[0], plo[0]);
          split_pair(sc[2 * ks][2] * i1, sc[2 * ks][3] * i1, ph[1], plo[1]);
          if (2 * ks + 1 < NT) {
            split_pair(sc[(2 * ks + 1) % NT][0] * i0, sc[(2 * ks + 1) % NT][1] * i0, ph[2], plo[2]);
            split_pair(sc[(2 * ks + 1) % NT][2] * i1, sc[(2 * ks + 1) % NT][3] * i1, ph[3], plo[3]);
          } else {  // odd NT: the last 8 keys of the padded V tile carry no probability
            ph[2] = ph[3] = plo[2] = plo[3] = 0u;
          }
#pragma unroll
          for (int d2 = 0; d2 < 4; ++d2) {
            uint32_t v0, v1, v2, v3;
            ldmatrix_x4_trans(sV_a + sw_off(ks * 16 + (lm_m & 1) * 8 + lm_r, d2 * 2 + (lm_m >> 1)), v0, v1, v2, v3);
            mma_bf16_16816(o[2 * d2], ph[0], ph[1], ph[2], ph[3], v0, v1);
            mma_bf16_16816(o[2 * d2], plo[0], plo[1], plo[2], plo[3], v0, v1);
            mma_bf16_16816(o[2 * d2 + 1], ph[0], ph[1], ph[2], ph[3], v2, v3);
            mma_bf16_16816(o[2 * d2 + 1], plo[0], plo[1], plo[2], plo[3], v2, v3);
          }
        }
      }
      // ---- stage the 16 x 64 output tile in the Q buffer, then store whole rows
      __syncwarp();
#pragma unroll
      for (int dt = 0; dt < 8; ++dt) {
        *reinterpret_cast<uint32_t*>(sQ + sw_off(g, dt) + qd * 4) = pack_bf16(o[dt][0], o[dt][1]);
        *reinterpret_cast<uint32_t*>(sQ + sw_off(g + 8, dt) + qd * 4) = pack_bf16(o[dt][2], o[dt][3]);
      }
      __syncwarp();
      for (int idx = lane; idx < q_rows * 8; idx += 32) {
        const int r = idx >> 3, pc = idx & 7;
        const uint4 v = *reinterpret_cast<const uint4*>(sQ + sw_off(r, pc));
        *reinterpret_cast<uint4*>(a.out_act + static_cast<size_t>(own_base + mt * 16 + r) * a.ld_act + head * 64 + pc * 8) = v;
      }
    }
  }
}

__global__ void im2col_kernel(const float* __restrict__ pix, int B, int S, int p, bf16* __restrict__ dst, int ldd,
                              int split) {
  const int g = S / p, K = 3 * p * p;
  const size_t total = static_cast<size_t>(B) * g * g * (K / 4);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % (K / 4)) * 4;
    const size_t row = i / (K / 4);
    const int b = static_cast<int>(row / (g * g)), pi = static_cast<int>(row % (g * g));
    const int c = k / (p * p), rem = k % (p * p), py = rem / p, px = rem % p;  // px % 4 == 0 (p % 4 == 0)
    const float4 v = *reinterpret_cast<const float4*>(
        pix + ((static_cast<size_t>(b) * 3 + c) * S + (pi / g) * p + py) * S + (pi % g) * p + px);
    store_act4(dst + row * ldd, K, split, k, v);
  }
}

__global__ void vision_embed_kernel(const float* __restrict__ patch, const float* __restrict__ cls,
                                    const float* __restrict__ pos, int B, int T, int H, float* __restrict__ x) {
  const size_t total = static_cast<size_t>(B) * T * (H / 4);
  for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int c = static_cast<int>(i % (H / 4)) * 4;
    const size_t row = i / (H / 4);
    const int b = static_cast<int>(row / T), t = static_cast<int>(row % T);
    const float4 a = t == 0 ? *reinterpret_cast<const float4*>(cls + c)
                            : *reinterpret_cast<const float4*>(patch + (static_cast<size_t>(b) * (T - 1) + (t - 1)) * H + c);
    const float4 q = *reinterpret_cast<const float4*>(pos + static_cast<size_t>(t) * H + c);
    *reinterpret_cast<float4*>(x + row * H + c) = make_float4(a.x + q.x, a.y + q.y, a.z + q.z, a.w + q.w);
  }
}

// dst row r = src row rows[r]; rows of `row_bytes` bytes (a multiple of 16).  One warp per row.
__global__ void gather_rows_kernel(const uint8_t* __restrict__ src, size_t row_bytes, const int32_t* __restrict__ rows,
                                   int n, uint8_t* __restrict__ dst) {
  PDL_ENTRY();
  const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
  for (int r = blockIdx.x * warps + (threadIdx.x >> 5); r < n; r += gridDim.x * warps) {
    const uint4* s = reinterpret_cast<const uint4*>(src + static_cast<size_t>(rows[r]) * row_bytes);
    uint4* d = reinterpret_cast<uint4*>(dst + static_cast<size_t>(r) * row_bytes);
    for (int i = lane; i < static_cast<int>(row_bytes >> 4); i += 32) d[i] = s[i];
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
  count_launch();
  ProfScope prof_(CAT_MISC, 0, st);
  const size_t total = static_cast<size_t>(rows) * (K / 4);
  int grid = static_cast<int>((total + 255) / 256);
  const int cap = num_sms() * 16;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  f32_to_act_kernel<<<grid, 256, 0, st>>>(src, rows, K, lds, dst, ldd, split);
}

void launch_im2col(const float* pix, int B, int S, int p, bf16* dst, int ldd, int split, cudaStream_t st) {
  count_launch();
  ProfScope prof_(CAT_EMBED, 0, st);
  const size_t total = static_cast<size_t>(B) * (S / p) * (S / p) * (3 * p * p / 4);
  int grid = static_cast<int>((total + 255) / 256);
  if (grid > num_sms() * 16) grid = num_sms() * 16;
  if (grid < 1) grid = 1;
  im2col_kernel<<<grid, 256, 0, st>>>(pix, B, S, p, dst, ldd, split);
}

void launch_vision_embed(const float* patch, const float* cls, const float* pos, int B, int T, int H, float* x,
                         cudaStream_t st) {
  count_launch();
  ProfScope prof_(CAT_EMBED, 0, st);
  const size_t total = static_cast<size_t>(B) * T * (H / 4);
  int grid = static_cast<int>((total + 255) / 256);
  if (grid > num_sms() * 16) grid = num_sms() * 16;
  if (grid < 1) grid = 1;
  vision_embed_kernel<<<grid, 256, 0, st>>>(patch, cls, pos, B, T, H, x);
}

void launch_gather_rows(const void* src, size_t row_bytes, const int32_t* rows, int n, void* dst, cudaStream_t st) {
  count_launch();
  ProfScope prof_(CAT_MISC, 0, st);
  if (n <= 0) return;
  launch_k(gather_rows_kernel, dim3(row_grid(n, 8)), dim3(256), 0, st, static_cast<const uint8_t*>(src), row_bytes, rows, n,
                                                     static_cast<uint8_t*>(dst));
}

void launch_layernorm(const LNArgs& a, cudaStream_t st) {
  count_launch();
  ProfScope prof_(CAT_LN, static_cast<double>(a.n_rows) * a.H, st);
  if (a.n_rows <= 0) return;
  if (a.partials && a.n_parts > 4) { set_error("layernorm: at most 4 split-K partial sums"); return; }
  // few rows (BERT: B * L ~ 1000): 4 warps per block so the rows spread over all SMs instead of 120 blocks
  const int wpb = a.n_rows < 4096 ? 4 : 8;
  const int grid = row_grid(a.n_rows, wpb);
  switch (a.H / 128) {
    case 4: launch_k(layernorm_kernel<4>, dim3(grid), dim3(wpb * 32), 0, st, a); break;
    case 6: launch_k(layernorm_kernel<6>, dim3(grid), dim3(wpb * 32), 0, st, a); break;
    case 8: launch_k(layernorm_kernel<8>, dim3(grid), dim3(wpb * 32), 0, st, a); break;
    default: set_error("layernorm: hidden size must be 512, 768 or 1024"); break;
  }
}

void launch_bert_embed_ln(const int64_t* inp, int rows, int L, const float* word, const float* pos, const float* type,
                          const float* gamma, const float* beta, float eps, int H, float* x_f32, bf16* act, int ld_act,
                          int split, cudaStream_t st) {
  count_launch();
  ProfScope prof_(CAT_EMBED, static_cast<double>(rows) * H, st);
  const int grid = row_grid(rows, 8);
  switch (H / 128) {
    case 4: launch_k(bert_embed_ln_kernel<4>, dim3(grid), dim3(256), 0, st, inp, rows, L, word, pos, type, gamma, beta, eps, H, x_f32, act, ld_act, split); break;
    case 6: launch_k(bert_embed_ln_kernel<6>, dim3(grid), dim3(256), 0, st, inp, rows, L, word, pos, type, gamma, beta, eps, H, x_f32, act, ld_act, split); break;
    case 8: launch_k(bert_embed_ln_kernel<8>, dim3(grid), dim3(256), 0, st, inp, rows, L, word, pos, type, gamma, beta, eps, H, x_f32, act, ld_act, split); break;
    default: set_error("bert_embed: hidden size must be 512, 768 or 1024"); break;
  }
}

void launch_clip_embed(const int32_t* ids_prefix, const int32_t* ids_suffix, const int32_t* p0, int B, int P, int K,
                       int S, int maxpos, const float* tok, const float* pos, int H, float* x_f32, cudaStream_t st,
                       const float* ln_g, const float* ln_b, float ln_eps, bf16* ln_out, int ln_ld,
                       const int32_t* cand_img, int n_cand) {
  count_launch();
  ProfScope prof_(CAT_EMBED, 0, st);
  const int rows = B * P + (cand_img ? n_cand : B * K) * S;
  if (rows <= 0) return;
  if (ln_out && H != 512) { set_error("clip_embed: the fused LayerNorm needs hidden size 512"); return; }
  launch_k(clip_embed_kernel, dim3(row_grid(rows, 8)), dim3(256), 0, st, ids_prefix, ids_suffix, p0, B, P, K, S, maxpos, tok, pos, H,
                                                       x_f32, ln_g, ln_b, ln_eps, ln_out, ln_ld, cand_img, n_cand);
}

constexpr int ATT_WARPS = 8;
constexpr size_t att_smem_bytes(int nt) { return static_cast<size_t>(ATT_WARPS) * ((nt + (nt + 1) / 2 * 2) * 1024 + 2048); }

bool attention_configure() {
  const int big = 220 * 1024;
  return cuda_ok(cudaFuncSetAttribute(attention_mma_kernel<2, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(att_smem_bytes(2))), "cudaFuncSetAttribute(attention_mma<2>)") &&
         cuda_ok(cudaFuncSetAttribute(attention_mma_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(att_smem_bytes(4))), "cudaFuncSetAttribute(attention_mma<4>)") &&
         cuda_ok(cudaFuncSetAttribute(attention_mma_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(att_smem_bytes(8))), "cudaFuncSetAttribute(attention_mma<8>)") &&
         cuda_ok(cudaFuncSetAttribute(attention_mma_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      static_cast<int>(att_smem_bytes(12))), "cudaFuncSetAttribute(attention_mma<12>)") &&
         cuda_ok(cudaFuncSetAttribute(attention_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, big),
                 "cudaFuncSetAttribute(attention<f32>)") &&
         cuda_ok(cudaFuncSetAttribute(attention_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, big),
                 "cudaFuncSetAttribute(attention<bf16>)");
}

bool launch_attention(const AttnArgs& a, cudaStream_t st) {
  count_launch();
  ProfScope prof_(CAT_ATTN, 0, st);
  if (a.H != a.heads * HD) {
    set_error("attention: head_dim must be 64");
    return false;
  }
  const int nk_cap = a.P + a.S;
  if (!a.qkv_f32 && !a.split && !a.cand_img && nk_cap <= 96 && (a.ld_qkv % 8) == 0 && (a.ld_act % 8) == 0) {
    AttnArgs aa = a;
    aa.cpt = (a.causal && a.S <= 8 && a.P + (16 / a.S) * a.S <= 96) ? 16 / a.S : 1;
    aa.cand_per_task = aa.cpt >= 4 ? 2 * aa.cpt : (aa.cpt > 1 ? ((8 + aa.cpt - 1) / aa.cpt) * aa.cpt : 8);
    const int keys_cap = a.P + (aa.cpt > 1 ? aa.cpt * a.S : a.S);
    const int groups = (a.K + aa.cand_per_task - 1) / aa.cand_per_task;
    const long long tasks = (a.P > 0 ? static_cast<long long>(a.B) * a.heads : 0) +
                            static_cast<long long>(a.B) * a.heads * groups;
    if (tasks <= 0) return true;
    // Tiles of <= 16 keys (the long-suffix steps, where this kernel's share is largest): 2 key tiles, 80 registers and
    // 48 KB per block -> 3 resident blocks (24 warps) per SM instead of 2; the kernel is latency bound, measured
    // -12 % attention time, bit-identical.  With 17..32 keys the same trade (more warps, spills in the two-candidate
    // tile loop) measured 2-3 % slower per step, so those keep 2 blocks of 128-register threads.
    const unsigned grid = static_cast<unsigned>((tasks + ATT_WARPS - 1) / ATT_WARPS);
    const int nt = keys_cap <= 16 ? 2 : (keys_cap <= 32 ? 4 : (keys_cap <= 64 ? 8 : 12));
    const size_t smem = att_smem_bytes(nt);
    if (nt == 2) launch_k(attention_mma_kernel<2, 3>, dim3(grid), dim3(ATT_WARPS * 32), smem, st, aa);
    else if (nt == 4) launch_k(attention_mma_kernel<4>, dim3(grid), dim3(ATT_WARPS * 32), smem, st, aa);
    else if (nt == 8) launch_k(attention_mma_kernel<8>, dim3(grid), dim3(ATT_WARPS * 32), smem, st, aa);
    else launch_k(attention_mma_kernel<12>, dim3(grid), dim3(ATT_WARPS * 32), smem, st, aa);
    return cuda_ok(cudaGetLastError(), "attention_mma launch");
  }
  if (nk_cap > 32 * MAX_SLOTS) {
    set_error("attention: more than 96 keys per sequence is not supported (sentence_len too large)");
    return false;
  }
  const int nkp = nk_cap | 1;
  const int nq_cap = a.P > a.S ? a.P : a.S;
  const int psz = nkp > 32 ? nkp : 32;
  const size_t per_warp = static_cast<size_t>((HD * nkp + HD * nk_cap + QS_LD * nq_cap + psz + 3) & ~3) * sizeof(float);
  int warps = static_cast<int>((200 * 1024) / per_warp);
  if (warps > 8) warps = 8;
  if (warps < 1) warps = 1;
  const long long n_seq = (a.P > 0 ? a.B : 0) + (a.cand_img ? static_cast<long long>(a.n_cand) : static_cast<long long>(a.B) * a.K);
  const long long n_task = n_seq * a.heads;
  // few tasks (BERT: images x heads; a certified re-score): smaller blocks so that every SM gets work -- the kernel is
  // latency bound, one warp walks the queries of its sequence one after the other
  while (warps > 2 && n_task / warps < 2 * static_cast<long long>(num_sms())) warps >>= 1;
  const size_t smem = per_warp * warps;
  long long grid = (n_task + warps - 1) / warps;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  if (grid > cap) grid = cap;
  if (grid < 1) grid = 1;
  if (a.qkv_f32)
    launch_k(attention_kernel<true>, dim3(static_cast<int>(grid)), dim3(warps * 32), smem, st, a, nk_cap, nq_cap);
  else
    launch_k(attention_kernel<false>, dim3(static_cast<int>(grid)), dim3(warps * 32), smem, st, a, nk_cap, nq_cap);
  return cuda_ok(cudaGetLastError(), "attention launch");
}

}  // namespace conzic
