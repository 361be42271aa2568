// out[M,N] = epilogue(A[M,K] * W[N,K]^T): the one dense contraction both towers are made of
// (HF:models/clip/modeling_clip.py:295-298,347-351 ; HF:models/bert/modeling_bert.py:179-181,295,340,353,481-501).
//
// Product kernel: warp-specialised sm_100a GEMM.
//   warp 0   : TMA producer  -- cp.async.bulk.tensor 2-D tiles (128B swizzle) of A and W into a smem ring
//   warp 1   : allocates TMEM, one elected thread issues tcgen05.mma (128 x BN x 16, bf16 -> fp32 in TMEM)
//   warps 2-5: epilogue      -- tcgen05.ld (thread = output row), bias / activation / residual, global store
// Several CTAs are resident per SM so one CTA's epilogue overlaps another's main loop.
// bf16x3 mode runs three passes over K (hi*hi, lo*hi, hi*lo) into the same accumulator.
#include <cstdlib>

#include "kernels.h"
#include "ptx.cuh"

namespace conzic {

namespace {

constexpr int BM = 128;
constexpr int BK = 64;  // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 192;

struct GemmParams {
  int M, N, K;  // K = logical reduction length (per plane)
  int split;
  Epi e;
  int ksplit = 1;            // split-K: blockIdx.z = z reduces k blocks [z*nkb/ksplit, (z+1)*nkb/ksplit) ...
  size_t part_stride = 0;    // ... into out_f32 + z * part_stride (summed by the LayerNorm kernel that follows)
};

template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int BAR_OFF = STAGES * (A_BYTES + B_BYTES);
  static constexpr int STG_OFF = BAR_OFF + (2 * STAGES + 1) * 8 + 16;  // 4 epilogue warps x 2 KB transpose buffers
  static constexpr int TOTAL = STG_OFF + 16 + 4 * 2048;
  static constexpr int DYN_BYTES = TOTAL + 1024;  // slack for manual 1024B alignment
};

__device__ __forceinline__ float apply_act(float v, int act, bool precise) {
  if (act == ACT_QUICK_GELU) {
    // x * sigmoid(1.702 x)   (HF:activations.py:122-123)
    if (precise) return v / (1.0f + expf(-1.702f * v));
    // sigmoid(z) = 0.5 * tanh(z / 2) + 0.5: one MUFU op per element instead of two (the fc1 epilogue is
    // MUFU-bound otherwise); tanh.approx is good to ~2^-11, the result is rounded to bf16 (2^-9) anyway
    float t;
    asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(0.851f * v));
    return v * fmaf(0.5f, t, 0.5f);
  }
  if (act == ACT_ERF_GELU) return 0.5f * v * (1.0f + erff(v * 0.70710678118654752440f));
  return v;
}

__device__ __forceinline__ void store_act_pair(bf16* base, int out_K, int split, int col, float v0, float v1) {
  // col is even; writes two adjacent elements
  __nv_bfloat162 h = __floats2bfloat162_rn(v0, v1);
  *reinterpret_cast<__nv_bfloat162*>(base + col) = h;
  if (split) {
    float r0 = v0 - __bfloat162float(h.x), r1 = v1 - __bfloat162float(h.y);
    *reinterpret_cast<__nv_bfloat162*>(base + out_K + col) = __floats2bfloat162_rn(r0, r1);
  }
}

// Epilogue for 32 consecutive columns [n0, n0+32) of one row held in registers.
__device__ __forceinline__ void epilogue_chunk(const GemmParams& p, int row, int n0, float (&v)[32]) {
  const Epi& e = p.e;
  const bool precise = p.split != 0;
  const bool full = (n0 + 32 <= p.N);
  if (full) {
    if (e.bias) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 b = __ldg(reinterpret_cast<const float4*>(e.bias + n0 + j));
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    }
    if (e.act != ACT_NONE) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], e.act, precise);
    }
    if (e.resid) {
      const float* r = e.resid + static_cast<size_t>(row) * e.ldr + n0;
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 b = *reinterpret_cast<const float4*>(r + j);
        v[j] += b.x; v[j + 1] += b.y; v[j + 2] += b.z; v[j + 3] += b.w;
      }
    }
    if (e.out_f32) {
      float* o = e.out_f32 + static_cast<size_t>(row) * e.ldo_f32 + n0;
#pragma unroll
      for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(o + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
    if (e.out_act) {
      bf16* o = e.out_act + static_cast<size_t>(row) * e.ldo_act;
      if (!p.split) {
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          __nv_bfloat162 h0 = __floats2bfloat162_rn(v[j], v[j + 1]);
          __nv_bfloat162 h1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
          __nv_bfloat162 h2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]);
          __nv_bfloat162 h3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
          uint4 u;
          u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
          u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
          *reinterpret_cast<uint4*>(o + n0 + j) = u;
        }
      } else if ((e.out_K & 7) == 0) {
        // hi and lo planes as 16-byte stores (the 4-byte form made this the slowest part of the bf16x3 fc1 GEMMs)
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint32_t uh[4], ul[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float a0 = v[j + 2 * q], a1 = v[j + 2 * q + 1];
            __nv_bfloat162 h = __floats2bfloat162_rn(a0, a1);
            __nv_bfloat162 l = __floats2bfloat162_rn(a0 - __bfloat162float(h.x), a1 - __bfloat162float(h.y));
            uh[q] = *reinterpret_cast<uint32_t*>(&h);
            ul[q] = *reinterpret_cast<uint32_t*>(&l);
          }
          *reinterpret_cast<uint4*>(o + n0 + j) = make_uint4(uh[0], uh[1], uh[2], uh[3]);
          *reinterpret_cast<uint4*>(o + e.out_K + n0 + j) = make_uint4(ul[0], ul[1], ul[2], ul[3]);
        }
      } else {
#pragma unroll
        for (int j = 0; j < 32; j += 2) store_act_pair(o, e.out_K, 1, n0 + j, v[j], v[j + 1]);
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) {  // unrolled with a predicate: a dynamic index would push v[] into local memory
      const int n = n0 + j;
      if (n >= p.N) continue;
      float x = v[j];
      if (e.bias) x += __ldg(e.bias + n);
      x = apply_act(x, e.act, precise);
      if (e.resid) x += e.resid[static_cast<size_t>(row) * e.ldr + n];
      if (e.out_f32) e.out_f32[static_cast<size_t>(row) * e.ldo_f32 + n] = x;
      if (e.out_act) {
        bf16* o = e.out_act + static_cast<size_t>(row) * e.ldo_act;
        bf16 h = __float2bfloat16_rn(x);
        o[n] = h;
        if (p.split) o[e.out_K + n] = __float2bfloat16_rn(x - __bfloat162float(h));
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// Persistent variant for the big CLIP linears (M = tens of thousands of token rows, N and K small).
//   * one CTA (CG == 1) or one CTA pair sharing every B tile through tcgen05 cta_group::2 (CG == 2) per SM;
//   * work unit = (128-row m tile per CTA, 256-column n tile); each CTA group walks a contiguous range of
//     the unit list, so the A tile of an m tile is fetched once and stays in shared memory for all its n
//     tiles (ARES, K <= 512) while only W streams through the TMA ring;
//   * the accumulator is double buffered in TMEM (2 x 256 columns): the 8 epilogue warps drain tile i while
//     the tensor core works on tile i+1.
// Per CTA and k block the L2 -> SM traffic is 32 KB / CG (B only) for 2*128*256*64 FLOP.
// ---------------------------------------------------------------------------------------------------
constexpr int PBN = 256;
constexpr int P_EPI_WARPS = 8;
constexpr int P_THREADS = 64 + 32 * P_EPI_WARPS;
constexpr int P_MAX_KB = 8;  // A-resident mode: K <= 512

struct PGemmParams {
  int M, N, K;
  int m_tiles;  // per-CTA-group tiles of CG*128 rows
  int n_tiles;
  Epi e;
  int split = 0;  // 1 = bf16x3 operands ([rows, 2K]: hi | lo planes): three passes over K (hi*hi, lo*hi, hi*lo) into the
                  // same accumulator, exact activations, hi + lo output planes (streamed-A pair kernel only)
  int tma_out = 0;  // persistent kernel, bf16-only output: 32 x 32 boxes leave through the TMA engine (tmO is valid)
};

template <int CG, bool ARES>
struct PSmem {
  static constexpr int A_SLOT = BM * BK * 2;               // 16 KB
  static constexpr int B_STAGE = (PBN / CG) * BK * 2;      // 32 KB or 16 KB (each CTA of a pair holds half)
  static constexpr int STAGE = ARES ? B_STAGE : (A_SLOT + B_STAGE);
  static constexpr int A_BYTES = ARES ? P_MAX_KB * A_SLOT : 0;
  static constexpr int STAGES = ARES ? 5 : (CG == 1 ? 4 : 6);
  static_assert(!ARES || CG == 2, "A-resident mode needs the CTA pair");
  static constexpr int STG_OFF = A_BYTES + STAGES * STAGE;  // per-epilogue-warp 32 x 64 B transpose buffers
  static constexpr int STG_BYTES = P_EPI_WARPS * 2048;
  static constexpr int BAR_OFF = STG_OFF + STG_BYTES;
  static constexpr int N_BARS = 2 * STAGES + P_MAX_KB + 4;
  static constexpr int TOTAL = BAR_OFF + N_BARS * 8 + 16;
  static constexpr int DYN_BYTES = TOTAL;  // the dynamic smem window is 1024-byte aligned (checked in the kernel)
  static_assert(DYN_BYTES <= 232448, "over the 227 KB shared-memory limit");
};

// Epilogue of the persistent kernel.  tcgen05.ld hands every thread one accumulator ROW; storing that way makes
// each warp store touch 32 different cache lines.  So every warp transposes its 32-row chunk through a private
// 2 KB shared-memory buffer (32 rows of 64 B, 16-byte pieces XOR-swizzled so both phases are conflict free) and
// then reads / writes global memory with 4 lanes per row: every access covers whole 32-byte sectors, 64
// contiguous bytes per row.  Bias lives in registers (lane l holds the 4 values of columns 4l..4l+3 of the
// warp's 128 columns) and reaches the lane that needs it by shuffle, so no shared memory is spent on it.
__device__ __forceinline__ uint32_t stg_off(int r, int piece) { return r * 64 + ((piece ^ ((r >> 1) & 3)) << 4); }

// The residual (which the engine aliases with the output: x += ...) is fetched into registers BEFORE the
// accumulator is read, so its latency hides behind tcgen05.ld and the loads are not serialised behind stores.
__device__ __forceinline__ void prefetch_resid(const PGemmParams& p, int lane, int row0, int n0, float4 (&rr)[4]) {
  const Epi& e = p.e;
  const int pc = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int grow = row0 + (lane >> 2) + 8 * i;
    rr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (e.resid && grow < p.M) {
      const float* src = e.resid + static_cast<size_t>(grow) * e.ldr + n0 + pc * 4;
      rr[i] = *reinterpret_cast<const float4*>(src);
    }
  }
}

// fp32 path: 16 columns [n0, n0+16) of rows [row0, row0+32); v = this lane's row (row0 + lane);
// bias4 = this lane's slice of the warp's bias, chunk = index of the 16-column chunk inside the warp's 128.
__device__ __forceinline__ void epilogue_f32_chunk(const PGemmParams& p, uint8_t* stg, int lane, int row0, int n0,
                                                   float (&v)[16], const float4 (&rr)[4], const float4& bias4,
                                                   int chunk, float (&st1)[4], float (&st2)[4]) {
  const Epi& e = p.e;
#pragma unroll
  for (int j = 0; j < 4; ++j)
    *reinterpret_cast<float4*>(stg + stg_off(lane, j)) = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  __syncwarp();
  const int pc = lane & 3;
  const int src = chunk * 4 + pc;
  float4 bb;
  bb.x = __shfl_sync(0xffffffffu, bias4.x, src); bb.y = __shfl_sync(0xffffffffu, bias4.y, src);
  bb.z = __shfl_sync(0xffffffffu, bias4.z, src); bb.w = __shfl_sync(0xffffffffu, bias4.w, src);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = (lane >> 2) + 8 * i;
    const int grow = row0 + r;
    float4 x = *reinterpret_cast<const float4*>(stg + stg_off(r, pc));
    x.x += bb.x; x.y += bb.y; x.z += bb.z; x.w += bb.w;
    if (e.act != ACT_NONE) {
      const bool precise = p.split != 0;
      x.x = apply_act(x.x, e.act, precise); x.y = apply_act(x.y, e.act, precise);
      x.z = apply_act(x.z, e.act, precise); x.w = apply_act(x.w, e.act, precise);
    }
    if (grow < p.M) {
      const int col = n0 + pc * 4;
      x.x += rr[i].x; x.y += rr[i].y; x.z += rr[i].z; x.w += rr[i].w;
      if (e.lnf_out) {  // LayerNorm statistics of the values being written
        st1[i] += (x.x + x.y) + (x.z + x.w);
        st2[i] += (x.x * x.x + x.y * x.y) + (x.z * x.z + x.w * x.w);
      }
      if (e.out_f32) {
        float* dst = e.out_f32 + static_cast<size_t>(grow) * e.ldo_f32 + col;
        *reinterpret_cast<float4*>(dst) = x;
      }
      if (e.out_act) {
        __nv_bfloat162 h0 = __floats2bfloat162_rn(x.x, x.y), h1 = __floats2bfloat162_rn(x.z, x.w);
        uint2 u;
        u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
        *reinterpret_cast<uint2*>(e.out_act + static_cast<size_t>(grow) * e.ldo_act + col) = u;
      }
    }
  }
  __syncwarp();
}

// bf16-only path (no residual, no fp32 output): 32 columns [n0, n0+32) per chunk = 64-byte rows of bf16.
__device__ __forceinline__ void epilogue_bf16_chunk(const PGemmParams& p, uint8_t* stg, int lane, int row0, int n0,
                                                    float (&v)[32], const float4& bias4, int chunk) {
  const Epi& e = p.e;
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {  // columns 4jj..4jj+3 of the chunk: bias held by lane chunk*8 + jj
    const int src = chunk * 8 + jj;
    v[4 * jj] += __shfl_sync(0xffffffffu, bias4.x, src);
    v[4 * jj + 1] += __shfl_sync(0xffffffffu, bias4.y, src);
    v[4 * jj + 2] += __shfl_sync(0xffffffffu, bias4.z, src);
    v[4 * jj + 3] += __shfl_sync(0xffffffffu, bias4.w, src);
  }
  if (p.split) {
    // bf16x3 output: exact activation, then the hi plane and the remainder plane, each through the transposing buffer
    if (e.act != ACT_NONE) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], e.act, true);
    }
    const int pc = lane & 3;
#pragma unroll
    for (int plane = 0; plane < 2; ++plane) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        uint32_t u[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float a0 = v[8 * j + 2 * q], a1 = v[8 * j + 2 * q + 1];
          __nv_bfloat162 h = __floats2bfloat162_rn(a0, a1);
          if (plane == 0) {
            u[q] = *reinterpret_cast<uint32_t*>(&h);
          } else {
            __nv_bfloat162 l = __floats2bfloat162_rn(a0 - __bfloat162float(h.x), a1 - __bfloat162float(h.y));
            u[q] = *reinterpret_cast<uint32_t*>(&l);
          }
        }
        *reinterpret_cast<uint4*>(stg + stg_off(lane, j)) = make_uint4(u[0], u[1], u[2], u[3]);
      }
      __syncwarp();
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int r = (lane >> 2) + 8 * i;
        const int grow = row0 + r;
        const uint4 x = *reinterpret_cast<const uint4*>(stg + stg_off(r, pc));
        if (grow < p.M)
          *reinterpret_cast<uint4*>(e.out_act + static_cast<size_t>(grow) * e.ldo_act + plane * e.out_K + n0 + pc * 8) = x;
      }
      __syncwarp();
    }
    return;
  }
  if (e.act == ACT_QUICK_GELU) {
    // x * sigmoid(1.702 x) = x * (0.5 * tanh(0.851 x) + 0.5), written as three passes over the 32 values so the
    // 32 MUFU ops are independent (one pass per element serialises on the ~30-cycle FMUL -> MUFU -> FFMA chain
    // and makes this epilogue slower than the tile's MMAs)
    float t[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) t[j] = 0.851f * v[j];
#pragma unroll
    for (int j = 0; j < 32; ++j) asm("tanh.approx.f32 %0, %1;" : "=f"(t[j]) : "f"(t[j]));
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= fmaf(0.5f, t[j], 0.5f);
  } else if (e.act != ACT_NONE) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], e.act, false);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * j], v[8 * j + 1]);
    __nv_bfloat162 h1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]);
    __nv_bfloat162 h3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
    u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
    *reinterpret_cast<uint4*>(stg + stg_off(lane, j)) = u;
  }
  __syncwarp();
  const int pc = lane & 3;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = (lane >> 2) + 8 * i;
    const int grow = row0 + r;
    const uint4 x = *reinterpret_cast<const uint4*>(stg + stg_off(r, pc));
    if (grow < p.M) {
      bf16* dst = e.out_act + static_cast<size_t>(grow) * e.ldo_act + n0 + pc * 8;
      *reinterpret_cast<uint4*>(dst) = x;
    }
  }
  __syncwarp();
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmParams p) {
  PDL_ENTRY();
  using SL = SmemLayout<BN, STAGES>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * SL::A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SL::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const int n_blk = blockIdx.x;
  const int m_blk = blockIdx.y;
  const int nkb_all = p.K / BK;
  const int kb0 = static_cast<int>(blockIdx.z) * nkb_all / p.ksplit;
  const int nkb = (static_cast<int>(blockIdx.z) + 1) * nkb_all / p.ksplit - kb0;  // k blocks of this split
  const int iters = p.split ? 3 * nkb : nkb;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1) {
    if (lane == 0) {
      for (int s = 0; s < STAGES; ++s) {
        mbar_init(&full_bar[s], 1);
        mbar_init(&empty_bar[s], 1);
      }
      mbar_init(accum_bar, 1);
      fence_mbar_init();
    }
    __syncwarp();
    tmem_alloc(tmem_slot, BN);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int i = 0; i < iters; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&empty_bar[s], ph ^ 1);
        const int pass = i / nkb, kb = kb0 + (i - pass * nkb);
        const int a_col = (pass == 1 ? p.K : 0) + kb * BK;
        const int b_col = (pass == 2 ? p.K : 0) + kb * BK;
        mbar_arrive_expect_tx(&full_bar[s], SL::A_BYTES + SL::B_BYTES);
        tma_load_2d(sA + s * SL::A_BYTES, &tmA, &full_bar[s], a_col, m_blk * BM);
        tma_load_2d(sB + s * SL::B_BYTES, &tmB, &full_bar[s], b_col, n_blk * BN);
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN);
      for (int i = 0; i < iters; ++i) {
        const int s = i % STAGES;
        const uint32_t ph = (i / STAGES) & 1;
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a0 = smem_u32(sA + s * SL::A_BYTES);
        const uint32_t b0 = smem_u32(sB + s * SL::B_BYTES);
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          const uint64_t da = make_smem_desc_sw128(a0 + k * (UMMA_K * 2));
          const uint64_t db = make_smem_desc_sw128(b0 + k * (UMMA_K * 2));
          umma_bf16(tmem_base, da, db, idesc, (i > 0 || k > 0) ? 1u : 0u);
        }
        umma_commit(&empty_bar[s]);  // frees this smem stage once the MMAs above have read it
      }
      umma_commit(accum_bar);  // accumulator complete
    }
    __syncwarp();
  } else {
    const int q = warp & 3;  // TMEM lane quarter this warp may read
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const int row = m_blk * BM + q * 32 + lane;
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(q * 32) << 16);
    if (BN == 128 && !p.split && (n_blk + 1) * BN <= p.N) {
      // bf16 mode, full tile: the smem-transposed, coalesced epilogue of the persistent kernel (the row-per-thread
      // form below made these small, latency-bound launches spend most of their time in scattered accesses)
      PGemmParams pp;
      pp.M = p.M; pp.N = p.N; pp.K = p.K; pp.m_tiles = 0; pp.n_tiles = 0; pp.e = p.e;
      if (p.ksplit > 1) pp.e.out_f32 += static_cast<size_t>(blockIdx.z) * p.part_stride;
      uint8_t* stg = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem + SL::STG_OFF) + 15) & ~uintptr_t(15)) + q * 2048;
      const int row0 = m_blk * BM + q * 32;
      const int nbase = n_blk * BN;
      float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.e.bias) bias4 = __ldg(reinterpret_cast<const float4*>(p.e.bias + nbase + lane * 4));
      if (p.e.out_act && !p.e.out_f32 && !p.e.resid) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld32(lane_addr + c * 32, r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          epilogue_bf16_chunk(pp, stg, lane, row0, nbase + c * 32, v, bias4, c);
        }
      } else {
        float st1[4] = {0.f, 0.f, 0.f, 0.f}, st2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
          const int n0 = nbase + c * 16;
          float4 rr[4];
          prefetch_resid(pp, lane, row0, n0, rr);
          uint32_t r[16];
          tmem_ld16(lane_addr + c * 16, r);
          tmem_ld_wait();
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
          epilogue_f32_chunk(pp, stg, lane, row0, n0, v, rr, bias4, c, st1, st2);
        }
      }
    } else
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(lane_addr + c0, r);
      tmem_ld_wait();
      const int n0 = n_blk * BN + c0;
      if (row < p.M && n0 < p.N) {
        float v[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
        if (p.ksplit > 1) {
          GemmParams pz = p;
          pz.e.out_f32 += static_cast<size_t>(blockIdx.z) * p.part_stride;
          epilogue_chunk(pz, row, n0, v);
        } else {
          epilogue_chunk(p, row, n0, v);
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

// bf16-only output through the TMA engine: the same 32-row x 64-byte staging tile as epilogue_bf16_chunk (its XOR
// swizzle is the TMA engine's 64-byte swizzle), but instead of reading it back and storing 16 bytes per lane -- 64
// contiguous bytes per row and instruction, which keeps the load/store unit busy for most of a tile's MMA time (ncu:
// the epilogue warps of fc1 never wait for an accumulator) -- one cp.async.bulk.tensor store moves the box.
__device__ __forceinline__ void epilogue_bf16_box(const PGemmParams& p, const CUtensorMap* tmO, uint8_t* stg, int lane,
                                                  int row0, int n0, float (&v)[32], const float4& bias4, int chunk) {
  const Epi& e = p.e;
#pragma unroll
  for (int jj = 0; jj < 8; ++jj) {  // columns 4jj..4jj+3 of the chunk: bias held by lane chunk*8 + jj
    const int src = chunk * 8 + jj;
    v[4 * jj] += __shfl_sync(0xffffffffu, bias4.x, src);
    v[4 * jj + 1] += __shfl_sync(0xffffffffu, bias4.y, src);
    v[4 * jj + 2] += __shfl_sync(0xffffffffu, bias4.z, src);
    v[4 * jj + 3] += __shfl_sync(0xffffffffu, bias4.w, src);
  }
  if (e.act == ACT_QUICK_GELU) {  // see epilogue_bf16_chunk
    float t[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) t[j] = 0.851f * v[j];
#pragma unroll
    for (int j = 0; j < 32; ++j) asm("tanh.approx.f32 %0, %1;" : "=f"(t[j]) : "f"(t[j]));
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= fmaf(0.5f, t[j], 0.5f);
  } else if (e.act != ACT_NONE) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = apply_act(v[j], e.act, false);
  }
  // the previous box of this warp must have been read by its store before the staging tile is overwritten
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    __nv_bfloat162 h0 = __floats2bfloat162_rn(v[8 * j], v[8 * j + 1]);
    __nv_bfloat162 h1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
    __nv_bfloat162 h2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]);
    __nv_bfloat162 h3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
    uint4 u;
    u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
    u.z = *reinterpret_cast<uint32_t*>(&h2); u.w = *reinterpret_cast<uint32_t*>(&h3);
    *reinterpret_cast<uint4*>(stg + stg_off(lane, j)) = u;
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane == 0) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tmO)), "r"(smem_u32(stg)), "r"(n0), "r"(row0) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
}

template <int CG, bool ARES>
__global__ void __launch_bounds__(P_THREADS, 1)
gemm_persist_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmO, const PGemmParams p) {
  PDL_ENTRY();
  using SL = PSmem<CG, ARES>;
  constexpr int EW = P_EPI_WARPS;
  constexpr int WCOLS = PBN / 2;  // columns per epilogue warp
  constexpr int STAGES = SL::STAGES;
  extern __shared__ __align__(1024) uint8_t psmem[];
  uint8_t* smem = psmem;
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // SWIZZLE_128B tiles need 1024-byte alignment
  uint8_t* sA = smem;
  uint8_t* sStage = smem + SL::A_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SL::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* empty_a = empty_bar + STAGES;
  uint64_t* tfull_bar = empty_a + P_MAX_KB;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = (CG == 2) ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const long long n_groups = gridDim.x / CG;
  const long long group = blockIdx.x / CG;
  const long long U = static_cast<long long>(p.m_tiles) * p.n_tiles;
  const long long u0 = group * U / n_groups, u1 = (group + 1) * U / n_groups;
  const int nkb1 = p.K / BK;                              // k blocks of one pass
  const int nkb = (!ARES && p.split) ? 3 * nkb1 : nkb1;   // bf16x3: hi*hi, lo*hi, hi*lo into the same accumulator

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int k = 0; k < P_MAX_KB; ++k) mbar_init(&empty_a[k], 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&tfull_bar[b], 1);
      mbar_init(&tempty_bar[b], CG * EW);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_cg<CG>(tmem_slot, 512);
    tmem_relinquish_cg<CG>();
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer (one lane; in a pair both CTAs load their own rows of A and their half of B) =====
    if (lane == 0) {
      int s = 0; uint32_t ph = 0; int cur_m = -1; uint32_t m_started = 0;
      for (long long u = u0; u < u1; ++u) {
        const int m = static_cast<int>(u / p.n_tiles), n = static_cast<int>(u % p.n_tiles);
        const bool newm = (m != cur_m);
        cur_m = m;
        const bool load_a = !ARES || newm;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (ARES && newm) mbar_wait(&empty_a[kb], (m_started & 1) ^ 1);
          const uint32_t bytes = SL::B_STAGE + (load_a ? SL::A_SLOT : 0);
          if (p.e.resid) {
            // pull this unit's residual tile (128 rows x 1 KB) into L2 well before the epilogue reads it
            const int rows_per_kb = (BM + nkb - 1) / nkb;
            const int r_lo = (m * CG + static_cast<int>(cta_rank)) * BM + kb * rows_per_kb;
            const int r_hi = min(min(r_lo + rows_per_kb, (m * CG + static_cast<int>(cta_rank) + 1) * BM), p.M);
            const int c0 = n * PBN;
            const int nbytes = min(PBN, p.N - c0) * 4;
            if ((nbytes & 15) == 0 && nbytes > 0)
              for (int r = r_lo; r < r_hi; ++r)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.e.resid + static_cast<size_t>(r) * p.e.ldr + c0),
                             "r"(nbytes) : "memory");
          }
          if (leader) mbar_arrive_expect_tx(&full_bar[s], CG * bytes);
          const uint32_t bar = (CG == 2) ? mapa_shared(smem_u32(&full_bar[s]), 0) : smem_u32(&full_bar[s]);
          uint8_t* st = sStage + s * SL::STAGE;
          const int pass = kb / nkb1, kcol = (kb - pass * nkb1) * BK;
          if (load_a)
            tma_load_2d_cg<CG>(ARES ? sA + kb * SL::A_SLOT : st, &tmA, bar, kcol + (pass == 1 ? p.K : 0),
                               (m * CG + static_cast<int>(cta_rank)) * BM);
          tma_load_2d_cg<CG>(ARES ? st : st + SL::A_SLOT, &tmB, bar, kcol + (pass == 2 ? p.K : 0),
                             n * PBN + static_cast<int>(cta_rank) * (PBN / CG));
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        if (newm) ++m_started;
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===== MMA issuer (one lane of the leader CTA) =====
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = make_idesc_bf16(BM * CG, PBN);
      int s = 0; uint32_t ph = 0; uint32_t acc = 0, acc_ph = 0;
      for (long long u = u0; u < u1; ++u) {
        const int n = static_cast<int>(u % p.n_tiles);
        const bool last_of_m = (n == p.n_tiles - 1) || (u == u1 - 1);
        mbar_wait(&tempty_bar[acc], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * PBN;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          uint8_t* st = sStage + s * SL::STAGE;
          const uint32_t a0 = smem_u32(ARES ? sA + kb * SL::A_SLOT : st);
          const uint32_t b0 = smem_u32(ARES ? st : st + SL::A_SLOT);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            const uint64_t da = make_smem_desc_sw128(a0 + k * (UMMA_K * 2));
            const uint64_t db = make_smem_desc_sw128(b0 + k * (UMMA_K * 2));
            umma_bf16_cg<CG>(d, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit_cg<CG>(&empty_bar[s]);
          if (ARES && last_of_m) umma_commit_cg<CG>(&empty_a[kb]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit_cg<CG>(&tfull_bar[acc]);
        acc ^= 1;
        if (acc == 0) acc_ph ^= 1;
      }
    }
    __syncwarp();
  } else {
    // ===== epilogue: 8 warps, warp%4 = TMEM lane quarter, (warp-2)/4 = column half of the 256-wide tile =====
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;  // which WCOLS-wide slice of the 256-column tile
    uint32_t acc = 0, acc_ph = 0;
    const uint32_t tempty_addr0 = (CG == 2) ? mapa_shared(smem_u32(&tempty_bar[0]), 0) : smem_u32(&tempty_bar[0]);
    uint8_t* stg = smem + SL::STG_OFF + (warp - 2) * 2048;
    const bool bf16_only = p.e.out_act && !p.e.out_f32 && !p.e.resid;
    for (long long u = u0; u < u1; ++u) {
      const int m = static_cast<int>(u / p.n_tiles), n = static_cast<int>(u % p.n_tiles);
      const int row0 = (m * CG + static_cast<int>(cta_rank)) * BM + q * 32;
      const int nbase = n * PBN + half * WCOLS;
      // this lane's 4 bias values of the warp's 128 columns (loaded while the tensor core is still busy)
      float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (p.e.bias && nbase + lane * 4 + 4 <= p.N) bias4 = __ldg(reinterpret_cast<const float4*>(p.e.bias + nbase + lane * 4));
      mbar_wait(&tfull_bar[acc], acc_ph);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + acc * PBN + half * WCOLS;
      auto release = [&]() {
        // everything this warp needs from the accumulator is in registers: hand the buffer back to the MMA
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) mbar_arrive_cluster_relaxed(tempty_addr0 + acc * 8);
          else mbar_arrive_relaxed(&tempty_bar[acc]);
        }
      };
      if (bf16_only && p.tma_out && nbase + PBN / 2 <= p.N) {
        // two register sets: the accumulator columns of chunk c + 1 are on their way while chunk c is activated, packed
        // and handed to the TMA engine; the accumulator goes back to the MMA warp before the last chunk is processed
        uint32_t ra[32], rb[32];
        float v[32];
        tmem_ld32(taddr, ra);
#pragma unroll
        for (int c = 0; c < 4; c += 2) {
          tmem_ld_wait();
          tmem_ld32(taddr + (c + 1) * 32, rb);
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(ra[j]);
          epilogue_bf16_box(p, &tmO, stg, lane, row0, nbase + c * 32, v, bias4, c);
          tmem_ld_wait();
          if (c + 2 < 4) tmem_ld32(taddr + (c + 2) * 32, ra);
          else release();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(rb[j]);
          epilogue_bf16_box(p, &tmO, stg, lane, row0, nbase + (c + 1) * 32, v, bias4, c + 1);
        }
      } else if (bf16_only && nbase + PBN / 2 <= p.N) {
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          tmem_ld_wait();
          if (c == 3) release();
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
          epilogue_bf16_chunk(p, stg, lane, row0, nbase + c * 32, v, bias4, c);
        }
      } else if (nbase + PBN / 2 <= p.N) {
        float st1[4] = {0.f, 0.f, 0.f, 0.f}, st2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 1
        for (int c = 0; c < 8; ++c) {
          const int n0 = nbase + c * 16;
          float4 rr[4];
          prefetch_resid(p, lane, row0, n0, rr);
          uint32_t r[16];
          tmem_ld16(taddr + c * 16, r);
          tmem_ld_wait();
          if (c == 7) release();
          float v[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
          epilogue_f32_chunk(p, stg, lane, row0, n0, v, rr, bias4, c, st1, st2);
        }
      } else {
        // ragged last n tile (N not a multiple of 128): plain row-per-thread epilogue
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld32(taddr + c * 32, r);
          tmem_ld_wait();
          if (c == 3) release();
          const int n0 = nbase + c * 32;
          if (row0 + lane < p.M && n0 < p.N) {
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
            GemmParams gp;
            gp.M = p.M; gp.N = p.N; gp.K = p.K; gp.split = p.split; gp.e = p.e;
            epilogue_chunk(gp, row0 + lane, n0, v);
          }
        }
      }
      acc ^= 1;
      if (acc == 0) acc_ph ^= 1;
    }
    if (p.tma_out) {  // the staging tile must outlive the last store's read of it
      if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      __syncwarp();
    }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_cg<CG>(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------------------------------
// Streamed-A variant for N = 512 (fc2 of the CLIP blocks, K = 2048): one unit = a 256-row pair tile x ALL 512
// columns, i.e. both 256-column TMEM accumulators belong to the same unit.  The A tile (the [M, 2048] MLP
// intermediate, 4 KB per token row) is then streamed from HBM ONCE instead of once per 256-column n tile --
// measured with ncu, the second pass of gemm_persist_kernel<2,false> missed L2 and doubled the kernel's DRAM
// reads.  Cost: the accumulators are no longer double buffered across units; the epilogue hands the two halves
// back separately so the next unit's MMAs restart on half 0 while half 1 is still being drained.
// ---------------------------------------------------------------------------------------------------
// Two epilogue forms.  EPI 0: EW = 16 warps move the tile with per-lane loads and stores through a small transposing
// staging buffer (64 contiguous bytes per row and instruction); the fallback for epilogues the other form does not do,
// and the A/B reference.  EPI 2: EW = 8 warps (two per TMEM lane quarter, one per accumulator half); acc + bias leaves
// as 32 x 32 fp32 boxes that the L2 ADDS to the residual stream in place (cp.reduce.async.bulk.tensor .add.f32: x += ...,
// no residual load at all, the load/store units only see shared memory); the LayerNorm statistics come from the
// row-coalesced re-read that the LayerNorm pass makes anyway.  EPI 3: the same with 4 K stages and one output box per
// warp instead of 3 and 2 (long K).  (Measured and dropped: the same boxes with the residual
// loaded by TMA and added in registers, 4 warps -- 2 % slower per step; 16-column boxes with 4 K stages and the
// LayerNorm one unit late -- the late re-read misses L2; 8 warps with per-lane accesses.)
template <int EW, int EPI>
struct WideSmem {
  static constexpr int A_SLOT = BM * BK * 2;           // 16 KB
  static constexpr int B_SLOT = (PBN / 2) * BK * 2;    // 16 KB: this CTA's 128 rows of one 256-row B tile
  static constexpr int STAGE = A_SLOT + 2 * B_SLOT;    // 48 KB
  static constexpr int STAGES = (EPI == 0 || EPI == 3) ? 4 : 3;
  static constexpr int NOB = EPI == 3 ? 1 : 2;          // EPI 2 / 3: output boxes per warp
  static constexpr int STG_OFF = STAGES * STAGE;       // EPI 0: one 2 KB staging buffer per epilogue warp
  static constexpr int BOX = 32 * 32 * 4;               // EPI 2: one 32-row x 32-column fp32 box
  static constexpr int EPI_WARP = NOB * BOX;
  static constexpr int BIAS_OFF = STG_OFF + (EPI ? EW * EPI_WARP : EW * 2048);
  static constexpr int BAR_OFF = BIAS_OFF + (EPI ? 2 * PBN * 4 : 0);
  static constexpr int N_BARS = 2 * STAGES + 4;
  static constexpr int DYN_BYTES = BAR_OFF + N_BARS * 8 + 16;
  static constexpr int THREADS = 64 + 32 * EW;
  static_assert(DYN_BYTES <= 232448, "over the 227 KB shared-memory limit");
  static_assert(EPI == 0 || EPI == 2 || EPI == 3, "epilogue form");
  static_assert(EPI == 0 || EW == 8, "the TMA reduce epilogue uses two warps per TMEM lane quarter");
};

template <int EW, int EPI>
__global__ void __launch_bounds__(64 + 32 * EW, 1)
gemm_wide_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                 const __grid_constant__ CUtensorMap tmO, const PGemmParams p) {
  PDL_ENTRY();
  using SL = WideSmem<EW, EPI>;
  constexpr int CG = 2;
  constexpr int STAGES = SL::STAGES;
  extern __shared__ __align__(1024) uint8_t psmem[];
  uint8_t* smem = psmem;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + SL::BAR_OFF);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;   // [0]: both halves of the unit are complete
  uint64_t* tempty_bar = tfull_bar + 2;       // [h]: half h has been read by every epilogue warp of the pair
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  const int lane = threadIdx.x & 31;
  const uint32_t cta_rank = cluster_ctarank();
  const bool leader = cta_rank == 0;
  const int n_groups = gridDim.x / CG;
  const int group = blockIdx.x / CG;
  const int nkb = p.K / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA); tma_prefetch_desc(&tmB);
    if (EPI) tma_prefetch_desc(&tmO);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&tfull_bar[0], 1); mbar_init(&tfull_bar[1], 1);
    // arrivals per half: EW == 16 -> 8 of the 16 warps read each half; reduce epilogue -> 4 of its 8 warps
    mbar_init(&tempty_bar[0], CG * (EPI ? 4 : 8)); mbar_init(&tempty_bar[1], CG * (EPI ? 4 : 8));
    fence_mbar_init();
  }
  if (EPI && warp >= 2) {  // the bias of all 512 columns, read as shared-memory broadcasts by the epilogue
    float* bias_s = reinterpret_cast<float*>(smem + SL::BIAS_OFF);
    for (int i = threadIdx.x - 64; i < 2 * PBN; i += 32 * EW) bias_s[i] = p.e.bias ? __ldg(p.e.bias + i) : 0.f;
  }
  if (warp == 2) {
    tmem_alloc_cg<CG>(tmem_slot, 512);
    tmem_relinquish_cg<CG>();
  }
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int m = group; m < p.m_tiles; m += n_groups) {
        const int arow = (m * CG + static_cast<int>(cta_rank)) * BM;
        // Short K (O-proj: 8 k blocks): a unit's MMAs restart on STAGES loaded k blocks and then wait out the HBM
        // latency of the rest; requesting the NEXT unit's A tile into L2 now turns that into L2 latency.
        if (nkb <= 8 && m + n_groups < p.m_tiles) {
          const int nrow = ((m + n_groups) * CG + static_cast<int>(cta_rank)) * BM;
          if (nrow < p.M)
            for (int kb = 0; kb < nkb; ++kb) tma_prefetch_2d(&tmA, kb * BK, nrow);
        }
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty_bar[s], ph ^ 1);
          if (p.e.resid) {
            const int rows_per_kb = (BM + nkb - 1) / nkb;
            const int r_lo = arow + kb * rows_per_kb;
            const int r_hi = min(min(r_lo + rows_per_kb, arow + BM), p.M);
            for (int r = r_lo; r < r_hi; ++r)
              asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p.e.resid + static_cast<size_t>(r) * p.e.ldr),
                           "r"(2 * PBN * 4) : "memory");
          }
          if (leader) mbar_arrive_expect_tx(&full_bar[s], CG * SL::STAGE);
          const uint32_t bar = mapa_shared(smem_u32(&full_bar[s]), 0);
          uint8_t* st = smem + s * SL::STAGE;
          tma_load_2d_cg<CG>(st, &tmA, bar, kb * BK, arow);
          tma_load_2d_cg<CG>(st + SL::A_SLOT, &tmB, bar, kb * BK, static_cast<int>(cta_rank) * BM);
          tma_load_2d_cg<CG>(st + SL::A_SLOT + SL::B_SLOT, &tmB, bar, kb * BK, PBN + static_cast<int>(cta_rank) * BM);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = make_idesc_bf16(BM * CG, PBN);
      int s = 0; uint32_t ph = 0; uint32_t uph = 0;
      for (int m = group; m < p.m_tiles; m += n_groups, uph ^= 1) {
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a0 = smem_u32(smem + s * SL::STAGE);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (kb == 0) {  // half h of the previous unit must have been drained
              mbar_wait(&tempty_bar[h], uph ^ 1);
              tc_fence_after();
            }
            const uint32_t b0 = a0 + SL::A_SLOT + h * SL::B_SLOT;
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              const uint64_t da = make_smem_desc_sw128(a0 + k * (UMMA_K * 2));
              const uint64_t db = make_smem_desc_sw128(b0 + k * (UMMA_K * 2));
              umma_bf16_cg<CG>(tmem_base + h * PBN, da, db, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            }
          }
          umma_commit_cg<CG>(&empty_bar[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit_cg<CG>(&tfull_bar[0]);
      }
    }
    __syncwarp();
  } else if constexpr (EPI >= 2) {
    // ---- TMA reduce epilogue: warp (q, half) drains the 8 boxes of accumulator half `half` for rows [q*32, q*32+32):
    // acc + bias -> swizzled box -> cp.reduce.async.bulk.tensor (.add.f32: the residual is added by the L2, in place)
    // or a plain TMA store when there is no residual.
    const int q = warp & 3;
    const int we = warp - 2;
    const int half = we >> 2;
    uint8_t* Ob = smem + SL::STG_OFF + we * SL::EPI_WARP;
    const float* bias_s = reinterpret_cast<const float*>(smem + SL::BIAS_OFF);
    const Epi& e = p.e;
    const bool has_res = e.resid != nullptr;
    const bool lnf = e.lnf_out != nullptr;
    const uint32_t tempty_addr = mapa_shared(smem_u32(&tempty_bar[half]), 0);
    uint32_t uph = 0;
    float4 g4[4], b4[4];
    if (lnf) {
#pragma unroll
      for (int sg = 0; sg < 4; ++sg) {
        g4[sg] = __ldg(reinterpret_cast<const float4*>(e.lnf_g + sg * 128 + lane * 4));
        b4[sg] = __ldg(reinterpret_cast<const float4*>(e.lnf_b + sg * 128 + lane * 4));
      }
    }
    const uint32_t swz = static_cast<uint32_t>(lane & 7);
    for (int m = group; m < p.m_tiles; m += n_groups, uph ^= 1) {
      const int row0 = (m * CG + static_cast<int>(cta_rank)) * BM + q * 32;
      mbar_wait(&tfull_bar[0], uph);
      tc_fence_after();
      // one box: registers of accumulator columns -> + bias -> swizzled shared-memory box -> TMA reduce / store
      auto emit_box = [&](const uint32_t (&r)[32], int tt) {
        const int t = half * 8 + tt;
        float4 bb[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) bb[j] = *reinterpret_cast<const float4*>(bias_s + t * 32 + 4 * j);
        // the output box written two boxes ago must have been read by its TMA operation before it is overwritten
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(SL::NOB - 1) : "memory");
        __syncwarp();
        uint8_t* O = Ob + (tt % SL::NOB) * SL::BOX + lane * 128;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float4 x;
          x.x = __uint_as_float(r[4 * j]) + bb[j].x;
          x.y = __uint_as_float(r[4 * j + 1]) + bb[j].y;
          x.z = __uint_as_float(r[4 * j + 2]) + bb[j].z;
          x.w = __uint_as_float(r[4 * j + 3]) + bb[j].w;
          *reinterpret_cast<float4*>(O + ((static_cast<uint32_t>(j) ^ swz) << 4)) = x;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the box is read by the async proxy next
        __syncwarp();
        if (lane == 0) {
          if (has_res)
            asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                         ::"l"(reinterpret_cast<uint64_t>(&tmO)), "r"(smem_u32(Ob + (tt % SL::NOB) * SL::BOX)), "r"(t * 32), "r"(row0)
                         : "memory");
          else
            asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                         ::"l"(reinterpret_cast<uint64_t>(&tmO)), "r"(smem_u32(Ob + (tt % SL::NOB) * SL::BOX)), "r"(t * 32), "r"(row0)
                         : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      };
      // two register sets: the accumulator columns of box tt + 1 are on their way while box tt is emitted; the half goes
      // back to the MMA warp as soon as its last columns are in registers (before the last box is emitted)
      const uint32_t tcol = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + half * PBN;
      uint32_t ra[32], rb[32];
      tmem_ld32(tcol, ra);
#pragma unroll
      for (int tt = 0; tt < 8; tt += 2) {
        tmem_ld_wait();
        tmem_ld32(tcol + (tt + 1) * 32, rb);
        emit_box(ra, tt);
        tmem_ld_wait();
        if (tt + 2 < 8) {
          tmem_ld32(tcol + (tt + 2) * 32, ra);
        } else {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_relaxed(tempty_addr);
        }
        emit_box(rb, tt + 1);
      }
      if (lnf) {
        // Fused LayerNorm: once both warps of the lane quarter have seen their boxes complete, each takes 16 of the
        // 32 rows: whole-row coalesced re-read (L2), statistics by warp reduction, bf16(LN(x) * g + b).
        if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __syncwarp();
        asm volatile("bar.sync %0, 64;" ::"r"(1 + q) : "memory");
        const float inv = 1.0f / static_cast<float>(2 * PBN);
        constexpr int LNR = 4;  // rows per round: 16 loads of 16 bytes in flight per lane (8 rows: slower, measured)
#pragma unroll 1
        for (int r0 = half * 16; r0 < half * 16 + 16; r0 += LNR) {
          float4 x[LNR][4];
#pragma unroll
          for (int u = 0; u < LNR; ++u) {
            const int grow = row0 + r0 + u;
#pragma unroll
            for (int sg = 0; sg < 4; ++sg) {
              x[u][sg] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (grow < p.M)
                x[u][sg] = __ldcg(reinterpret_cast<const float4*>(e.out_f32 + static_cast<size_t>(grow) * e.ldo_f32 + sg * 128 + lane * 4));
            }
          }
          float s1[LNR], s2[LNR];
#pragma unroll
          for (int u = 0; u < LNR; ++u) {
            s1[u] = 0.f; s2[u] = 0.f;
#pragma unroll
            for (int sg = 0; sg < 4; ++sg) {
              const float4 v = x[u][sg];
              s1[u] += (v.x + v.y) + (v.z + v.w);
              s2[u] += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
            for (int u = 0; u < LNR; ++u) {
              s1[u] += __shfl_xor_sync(0xffffffffu, s1[u], o);
              s2[u] += __shfl_xor_sync(0xffffffffu, s2[u], o);
            }
          }
#pragma unroll
          for (int u = 0; u < LNR; ++u) {
            const int grow = row0 + r0 + u;
            const float mu = s1[u] * inv;
            const float var = fmaxf(s2[u] * inv - mu * mu, 0.f);
            const float rs = rsqrtf(var + e.lnf_eps);
            const float nm = -mu * rs;
            if (grow < p.M) {
#pragma unroll
              for (int sg = 0; sg < 4; ++sg) {
                const float4 v = x[u][sg];
                const float y0 = fmaf(v.x, rs, nm) * g4[sg].x + b4[sg].x, y1 = fmaf(v.y, rs, nm) * g4[sg].y + b4[sg].y;
                const float y2 = fmaf(v.z, rs, nm) * g4[sg].z + b4[sg].z, y3 = fmaf(v.w, rs, nm) * g4[sg].w + b4[sg].w;
                __nv_bfloat162 h0 = __floats2bfloat162_rn(y0, y1), h1 = __floats2bfloat162_rn(y2, y3);
                uint2 u2;
                u2.x = *reinterpret_cast<uint32_t*>(&h0); u2.y = *reinterpret_cast<uint32_t*>(&h1);
                *reinterpret_cast<uint2*>(e.lnf_out + static_cast<size_t>(grow) * e.lnf_ld + sg * 128 + lane * 4) = u2;
              }
            }
          }
        }
      }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // shared memory stays valid until read
    __syncwarp();
  } else {
    // ---- per-lane epilogue (EPI 0; instantiated with EW == 16 only, the EW == 8 arithmetic is kept for reference).
    // warp % 4 = TMEM lane quarter.  EW == 16: (warp - 2) / 4 = which 128-column slice of the 512 columns (one half per
    // warp); EW == 8: (warp - 2) / 4 = which 128-column slice of EACH 256-column half
    const int q = warp & 3;
    const int slice = (warp - 2) >> 2;
    const int sub = EW == 8 ? slice : (slice & 1);  // columns [sub*128, sub*128+128) of a 256-column half
    const int h_lo = EW == 8 ? 0 : (slice >> 1), h_hi = EW == 8 ? 2 : (slice >> 1) + 1;
    uint32_t uph = 0;
    const bool lnf = p.e.lnf_out != nullptr;
    const uint32_t tempty_addr0 = mapa_shared(smem_u32(&tempty_bar[0]), 0);
    uint8_t* stg = smem + SL::STG_OFF + (warp - 2) * 2048;
    // this warp's bias values (columns do not depend on the tile): fetched once, before the first accumulator is
    // waited for -- inside the tile loop the load's latency sat on the un-overlapped epilogue's critical path
    float4 bias_h[2] = {make_float4(0.f, 0.f, 0.f, 0.f), make_float4(0.f, 0.f, 0.f, 0.f)};
    if (p.e.bias) {
#pragma unroll
      for (int h = 0; h < 2; ++h)
        if (h >= h_lo && h < h_hi) bias_h[h] = __ldg(reinterpret_cast<const float4*>(p.e.bias + h * PBN + sub * (PBN / 2) + lane * 4));
    }
    for (int m = group; m < p.m_tiles; m += n_groups, uph ^= 1) {
      const int row0 = (m * CG + static_cast<int>(cta_rank)) * BM + q * 32;
      // the first chunk's residual does not depend on the accumulator either: request it before waiting for the MMAs
      float4 rr[4];
      prefetch_resid(p, lane, row0, h_lo * PBN + sub * (PBN / 2), rr);
      mbar_wait(&tfull_bar[0], uph);
      tc_fence_after();
      float ls1[4] = {0.f, 0.f, 0.f, 0.f}, ls2[4] = {0.f, 0.f, 0.f, 0.f};  // row sums over this warp's columns (fused LN)
      // 16-column chunks; the residual of chunk t + 1 is requested before chunk t is drained, so that its latency
      // (the epilogue's largest stall in the ncu source view) is covered by one whole chunk of work
      constexpr int NT = EW == 8 ? 16 : 8;
#pragma unroll 1
      for (int t = 0; t < NT; ++t) {
        const int h = h_lo + (t >> 3), c = t & 7;
        const int n0 = h * PBN + sub * (PBN / 2) + c * 16;
        float4 rn[4];
        if (t + 1 < NT) prefetch_resid(p, lane, row0, (h_lo + ((t + 1) >> 3)) * PBN + sub * (PBN / 2) + ((t + 1) & 7) * 16, rn);
        const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + h * PBN + sub * (PBN / 2);
        uint32_t r[16];
        tmem_ld16(taddr + c * 16, r);
        tmem_ld_wait();
        if (c == 7) {  // this warp's part of accumulator half h is in registers: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster_relaxed(tempty_addr0 + h * 8);
        }
        float v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
        epilogue_f32_chunk(p, stg, lane, row0, n0, v, rr, h == 0 ? bias_h[0] : bias_h[1], c, ls1, ls2);
        if (t + 1 < NT) {
#pragma unroll
          for (int i = 0; i < 4; ++i) rr[i] = rn[i];
        }
      }
      if (lnf) {
        // Fused LayerNorm of the rows just written.  The EW / 4 warps of this TMEM lane quarter each covered
        // 512 / (EW / 4) of a row's columns: exchange (sum, sum of squares) through the warps' own staging buffers
        // (idle between units), re-read the values this lane stored, write bf16(LN(x) * g + b).
        const Epi& e = p.e;
        constexpr int NSL = EW / 4;
        float2* mine = reinterpret_cast<float2*>(stg);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float a = ls1[i], b = ls2[i];
          a += __shfl_xor_sync(0xffffffffu, a, 1); a += __shfl_xor_sync(0xffffffffu, a, 2);
          b += __shfl_xor_sync(0xffffffffu, b, 1); b += __shfl_xor_sync(0xffffffffu, b, 2);
          if ((lane & 3) == 0) mine[(lane >> 2) + 8 * i] = make_float2(a, b);
        }
        asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "r"(32 * NSL) : "memory");  // the warps of lane quarter q
        float mean[4], rstd[4];
        const int wq = (warp - 2) & 3;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          float a = 0.f, b = 0.f;
#pragma unroll
          for (int sl = 0; sl < NSL; ++sl) {  // same order in every warp: identical statistics for all columns of a row
            const float2 o = reinterpret_cast<const float2*>(smem + SL::STG_OFF + (4 * sl + wq) * 2048)[(lane >> 2) + 8 * i];
            a += o.x; b += o.y;
          }
          const float inv = 1.0f / static_cast<float>(2 * PBN);
          const float mu = a * inv;
          const float var = fmaxf(b * inv - mu * mu, 0.f);
          rstd[i] = rsqrtf(var + e.lnf_eps);
          mean[i] = -mu * rstd[i];  // y = x * rstd + (-mean * rstd)
        }
        asm volatile("bar.sync %0, %1;" ::"r"(1 + q), "r"(32 * NSL) : "memory");  // all read: staging may be reused
        const int pc = lane & 3;
        auto write_ln = [&](int col, const float4& g4, const float4& b4, int i, float x0, float x1, float x2, float x3) {
          const int grow = row0 + (lane >> 2) + 8 * i;
          if (grow >= p.M) return;
          const float y0 = fmaf(x0, rstd[i], mean[i]) * g4.x + b4.x, y1 = fmaf(x1, rstd[i], mean[i]) * g4.y + b4.y;
          const float y2 = fmaf(x2, rstd[i], mean[i]) * g4.z + b4.z, y3 = fmaf(x3, rstd[i], mean[i]) * g4.w + b4.w;
          __nv_bfloat162 h0 = __floats2bfloat162_rn(y0, y1), h1 = __floats2bfloat162_rn(y2, y3);
          uint2 u;
          u.x = *reinterpret_cast<uint32_t*>(&h0); u.y = *reinterpret_cast<uint32_t*>(&h1);
          *reinterpret_cast<uint2*>(e.lnf_out + static_cast<size_t>(grow) * e.lnf_ld + col) = u;
        };
        {
          // re-read what this lane stored (same addresses, program order; L2 hits mostly), GC chunks = 4 GC loads
          // in flight per lane so the pass is bandwidth- and not latency-bound; it overlaps the next unit's MMAs
          constexpr int NCH = EW == 8 ? 16 : 8, GC = EW == 8 ? 4 : 2;
#pragma unroll 1
          for (int g0 = 0; g0 < NCH; g0 += GC) {
            float4 x[GC][4];
#pragma unroll
            for (int j = 0; j < GC; ++j) {
              const int hc = g0 + j;
              const int col = (EW == 8 ? (hc >> 3) : h_lo) * PBN + sub * (PBN / 2) + (hc & 7) * 16 + pc * 4;
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const int grow = row0 + (lane >> 2) + 8 * i;
                x[j][i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (grow < p.M)
                  x[j][i] = __ldcg(reinterpret_cast<const float4*>(e.out_f32 + static_cast<size_t>(grow) * e.ldo_f32 + col));
              }
            }
#pragma unroll
            for (int j = 0; j < GC; ++j) {
              const int hc = g0 + j;
              const int col = (EW == 8 ? (hc >> 3) : h_lo) * PBN + sub * (PBN / 2) + (hc & 7) * 16 + pc * 4;
              const float4 g4 = __ldg(reinterpret_cast<const float4*>(e.lnf_g + col));
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(e.lnf_b + col));
#pragma unroll
              for (int i = 0; i < 4; ++i) write_ln(col, g4, b4, i, x[j][i].x, x[j][i].y, x[j][i].z, x[j][i].w);
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  cluster_sync_all();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_cg<CG>(tmem_base, 512);
  }
}

// Slow reference kernel on CUDA cores with the same operand format and epilogue (tests / bring-up only).
__global__ void gemm_simt_debug_kernel(const bf16* __restrict__ A, int lda, const bf16* __restrict__ W, int ldw,
                                       const GemmParams p) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (n >= p.N || m >= p.M) return;
  const bf16* a = A + static_cast<size_t>(m) * lda;
  const bf16* w = W + static_cast<size_t>(n) * ldw;
  float acc = 0.f;
  for (int k = 0; k < p.K; ++k) acc = fmaf(__bfloat162float(a[k]), __bfloat162float(w[k]), acc);
  if (p.split) {
    for (int k = 0; k < p.K; ++k) acc = fmaf(__bfloat162float(a[p.K + k]), __bfloat162float(w[k]), acc);
    for (int k = 0; k < p.K; ++k) acc = fmaf(__bfloat162float(a[k]), __bfloat162float(w[p.K + k]), acc);
  }
  const Epi& e = p.e;
  float x = acc;
  if (e.bias) x += e.bias[n];
  x = apply_act(x, e.act, p.split != 0);
  if (e.resid) x += e.resid[static_cast<size_t>(m) * e.ldr + n];
  if (e.out_f32) e.out_f32[static_cast<size_t>(m) * e.ldo_f32 + n] = x;
  if (e.out_act) {
    bf16* o = e.out_act + static_cast<size_t>(m) * e.ldo_act;
    bf16 h = __float2bfloat16_rn(x);
    o[n] = h;
    if (p.split) o[e.out_K + n] = __float2bfloat16_rn(x - __bfloat162float(h));
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;

template <int BN, int STAGES>
bool configure_one() {
  return cuda_ok(cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      SmemLayout<BN, STAGES>::DYN_BYTES),
                 "cudaFuncSetAttribute(gemm)");
}

template <int BN, int STAGES>
void launch_one(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, cudaStream_t st) {
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM, p.ksplit);
  launch_k(gemm_tcgen05_kernel<BN, STAGES>, dim3(grid), dim3(GEMM_THREADS), SmemLayout<BN, STAGES>::DYN_BYTES, st, ta, tb, p);
}

}  // namespace

bool tma_init() {
  if (g_encode) return true;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    set_error("cuTensorMapEncodeTiled not available from the driver (need an sm_100a-capable driver)");
    return false;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  return true;
}

bool make_tmap_bf16_2d(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                       uint32_t box_rows) {
  if (!tma_init()) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed, CUresult=" + std::to_string(static_cast<int>(r)) +
              " rows=" + std::to_string(rows) + " cols=" + std::to_string(cols) + " ld=" + std::to_string(ld_elems));
    return false;
  }
  return true;
}

// fp32 [rows, cols] tensor as 32-row x 32-column boxes (128-byte rows, 128-byte swizzle): the wide kernel's TMA epilogue
static bool make_tmap_f32_box32(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld_elems) {
  if (!tma_init()) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_elems * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (fp32 boxes) failed, CUresult=" + std::to_string(static_cast<int>(r)));
    return false;
  }
  return true;
}

template <int CG, bool ARES>
bool configure_persist() {
  return cuda_ok(cudaFuncSetAttribute(gemm_persist_kernel<CG, ARES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      PSmem<CG, ARES>::DYN_BYTES),
                 "cudaFuncSetAttribute(gemm_persist)");
}

// bf16 [rows, cols] output as 32-row x 32-column boxes (64-byte rows, 64-byte swizzle): the persistent kernel's TMA stores
static bool make_tmap_bf16_box32(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld_elems) {
  if (!tma_init()) return false;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (bf16 boxes) failed, CUresult=" + std::to_string(static_cast<int>(r)));
    return false;
  }
  return true;
}

template <int CG, bool ARES>
bool launch_persist(const CUtensorMap& ta, const CUtensorMap& tb, const PGemmParams& p_in, int sms, bool tma_out, cudaStream_t st) {
  PGemmParams p = p_in;
  const Epi& e = p.e;
  CUtensorMap to = ta;
  // bf16-only output of whole 128-column slices: boxes through the TMA engine
  if (tma_out && !p.split && e.out_act && !e.out_f32 && !e.resid && (p.N % (PBN / 2)) == 0 && (e.ldo_act % 8) == 0 &&
      (reinterpret_cast<uintptr_t>(e.out_act) & 15) == 0) {
    if (!make_tmap_bf16_box32(&to, e.out_act, static_cast<uint64_t>(p.M), static_cast<uint64_t>(p.N), static_cast<uint64_t>(e.ldo_act)))
      return false;
    p.tma_out = 1;
  }
  const long long U = static_cast<long long>(p.m_tiles) * p.n_tiles;
  long long groups = sms / CG;
  if (groups > U) groups = U;
  if (groups < 1) groups = 1;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(groups * CG));
  cfg.blockDim = dim3(P_THREADS);
  cfg.dynamicSmemBytes = PSmem<CG, ARES>::DYN_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  return cuda_ok(cudaLaunchKernelEx(&cfg, gemm_persist_kernel<CG, ARES>, ta, tb, to, p), "gemm_persist launch");
}

int g_sm_count = 0;
int sm_count() {
  if (g_sm_count == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
    if (g_sm_count <= 0) g_sm_count = 148;
  }
  return g_sm_count;
}

template <int EW, int EPI>
static bool configure_wide_one() {
  return cuda_ok(cudaFuncSetAttribute(gemm_wide_kernel<EW, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                      WideSmem<EW, EPI>::DYN_BYTES), "cudaFuncSetAttribute(gemm_wide)");
}
static bool configure_wide() {
  return configure_wide_one<16, 0>() && configure_wide_one<8, 2>() && configure_wide_one<8, 3>();
}
// lsu == false: TMA reduce epilogue (8 warps; the L2 adds acc + bias to the residual stream in place -- needs resid == out
// or no residual, fp32 output only); otherwise / lsu == true: 16 epilogue warps with per-lane loads and stores
template <int EW, int EPI>
static bool launch_wide_one(cudaLaunchConfig_t& cfg, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap& to,
                            const PGemmParams& p) {
  cfg.blockDim = dim3(WideSmem<EW, EPI>::THREADS);
  cfg.dynamicSmemBytes = WideSmem<EW, EPI>::DYN_BYTES;
  return cuda_ok(cudaLaunchKernelEx(&cfg, gemm_wide_kernel<EW, EPI>, ta, tb, to, p), "gemm_wide launch");
}
static bool launch_wide(const CUtensorMap& ta, const CUtensorMap& tb, const PGemmParams& p, bool lsu, cudaStream_t st) {
  int groups = sm_count() / 2;
  if (groups > p.m_tiles) groups = p.m_tiles;
  if (groups < 1) groups = 1;
  const Epi& e = p.e;
  const bool tma_ok = e.out_f32 && !e.out_act && e.act == ACT_NONE && e.ldo_f32 == 2 * PBN && (!e.resid || e.resid == e.out_f32) &&
                      (reinterpret_cast<uintptr_t>(e.out_f32) & 15) == 0;
  if (!tma_ok) lsu = true;
  CUtensorMap to = ta;
  if (!lsu && !make_tmap_f32_box32(&to, e.out_f32, static_cast<uint64_t>(p.M), 2 * PBN, 2 * PBN)) return false;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(static_cast<unsigned>(groups * 2));
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 2 : 1;
  if (lsu) return launch_wide_one<16, 0>(cfg, ta, tb, to, p);
  // long K (fc2): the MMA phase dominates a unit -- a 4th K stage is worth more than the second output box (tower -1.4 %,
  // same-box A/B); short K (O-proj): the drain dominates, two boxes per warp (the 4-stage form measured neutral there)
  return p.K >= 1024 ? launch_wide_one<8, 3>(cfg, ta, tb, to, p) : launch_wide_one<8, 2>(cfg, ta, tb, to, p);
}

bool gemm_configure() {
  if (!configure_wide()) return false;
  if (!(configure_persist<1, false>() && configure_persist<2, true>() && configure_persist<2, false>())) return false;
  return configure_one<64, 8>() && configure_one<128, 2>() && configure_one<128, 3>() && configure_one<128, 4>() && configure_one<128, 6>() &&
         configure_one<256, 2>() && configure_one<256, 4>();
}

bool launch_linear(const Act& A, int M, const LinearW& W, const Epi& epi, const GemmOpts& o, cudaStream_t st) {
  if (M <= 0) return true;
  if (A.K != W.K || (W.K % BK) != 0) {
    set_error("linear: K mismatch or K not a multiple of 64");
    return false;
  }
  GemmParams p;
  p.M = M; p.N = W.N; p.K = W.K; p.split = o.split; p.e = epi;
  if (o.ksplit > 1 && !o.persist && o.impl == 0) {
    if (epi.bias || epi.resid || epi.out_act || epi.act != ACT_NONE || !epi.out_f32 || o.ksplit > W.K / BK) {
      set_error("linear: split-K writes raw fp32 partial sums only");
      return false;
    }
    p.ksplit = o.ksplit;
    p.part_stride = static_cast<size_t>(M) * epi.ldo_f32;
  }
  count_launch();
  // the persistent pair kernel (CLIP tower) and the gridded kernel (BERT, bf16x3 passes) are timed separately
  ProfScope prof_((o.persist && !o.split && o.impl == 0) ? CAT_GEMM : CAT_GEMM_SMALL,  // bf16 tower vs BERT / bf16x3 work
                  2.0 * M * static_cast<double>(W.N) * W.K * (o.split ? 3 : 1), st);
  if ((epi.out_f32 && (epi.ldo_f32 & 3)) || (epi.resid && (epi.ldr & 3)) || (epi.out_act && (epi.ldo_act & 7)) ||
      (A.ld & 7)) {
    set_error("linear: leading dimensions must keep rows 16-byte aligned");
    return false;
  }
  if (o.impl == 1) {
    dim3 grid((W.N + 127) / 128, M);
    gemm_simt_debug_kernel<<<grid, 128, 0, st>>>(A.p, A.ld, W.w, o.split ? 2 * W.K : W.K, p);
    return cuda_ok(cudaGetLastError(), "gemm_simt_debug launch");
  }
  CUtensorMap ta;
  if (!make_tmap_bf16_2d(&ta, A.p, static_cast<uint64_t>(M), static_cast<uint64_t>(W.K) * (o.split ? 2 : 1),
                         static_cast<uint64_t>(A.ld), BM))
    return false;
  if (o.persist && (!o.split || o.cg == 2)) {
    const int cg = o.cg == 2 ? 2 : 1;
    PGemmParams pp;
    pp.M = M; pp.N = W.N; pp.K = W.K; pp.e = epi; pp.split = o.split;
    pp.m_tiles = (M + BM * cg - 1) / (BM * cg);
    pp.n_tiles = (W.N + PBN - 1) / PBN;
    if (o.split) {
      // bf16x3: the streamed-A pair kernel runs the three operand passes into one accumulator (same k order as the
      // gridded kernel, bit-identical results: tests/test_gpu_parity.py); no resident A (the hi | lo tile is twice as
      // wide), no wide form, no fused LayerNorm.  With fewer 256 x 256 units than half the CTA pairs (narrow N on a few
      // thousand rows: the certified re-score's O-proj / fc2) the 128 x 128 gridded kernel spreads the work better.
      if (epi.lnf_out) { set_error("linear: no fused LayerNorm in bf16x3 mode"); return false; }
      const long long units = static_cast<long long>(pp.m_tiles) * pp.n_tiles;
      if (units * 2 >= sm_count() / 2 || (W.N % BM) != 0 || o.force_pair)
        return launch_persist<2, false>(ta, W.tmap128, pp, sm_count(), false, st);
      p.split = 1;
      // N = 512 on ~2 k rows (the re-score's O-proj / fc2): 128 x 64 tiles double the CTAs that stream the 24 - 96 k blocks
      // (each CTA is bound by what it can pull from L2); same k order per element, bit-identical
      const long long ctas64 = static_cast<long long>((M + BM - 1) / BM) * ((W.N + 63) / 64);
      if ((W.N % 64) == 0 && ctas64 <= sm_count()) launch_one<64, 8>(ta, W.tmap64, p, st);
      else launch_one<128, 6>(ta, W.tmap128, p, st);
      return cuda_ok(cudaGetLastError(), "gemm_tcgen05 launch");
    }
    const bool ares = cg == 2 && W.K <= P_MAX_KB * BK;  // a single CTA has no room for a resident A tile + 32 KB B stages
    const CUtensorMap& tb = cg == 2 ? W.tmap128 : W.tmap256;
    // fp32-output GEMM with N == 512: one 512-column unit per tile, so a streamed A (fc2, K = 2048) leaves HBM once
    // and a CTA pair owns whole rows (which is what a LayerNorm written by the same epilogue needs)
    const bool wide_ok = cg == 2 && W.N == 2 * PBN && epi.out_f32 != nullptr;
    if (wide_ok && ((!ares && W.K >= 1024) || epi.lnf_out)) {
      pp.n_tiles = 1;
      return launch_wide(ta, tb, pp, o.wide_lsu != 0, st);
    }
    if (epi.lnf_out) {
      set_error("linear: a fused LayerNorm output needs the wide pair kernel (N == 512, fp32 output, CTA pairs)");
      return false;
    }
    if (cg == 2) return ares ? launch_persist<2, true>(ta, tb, pp, sm_count(), o.lsu_out == 0, st) : launch_persist<2, false>(ta, tb, pp, sm_count(), o.lsu_out == 0, st);
    return launch_persist<1, false>(ta, tb, pp, sm_count(), o.lsu_out == 0, st);
  }
  if (epi.lnf_out) {
    set_error("linear: a fused LayerNorm output needs the persistent wide pair kernel");
    return false;
  }
  const int key = o.bn * 10 + o.stages;
  switch (key) {
    case 1282: launch_one<128, 2>(ta, W.tmap128, p, st); break;
    case 1283: launch_one<128, 3>(ta, W.tmap128, p, st); break;
    case 1284: launch_one<128, 4>(ta, W.tmap128, p, st); break;
    case 1286: launch_one<128, 6>(ta, W.tmap128, p, st); break;
    case 2562: launch_one<256, 2>(ta, W.tmap256, p, st); break;
    case 2564: launch_one<256, 4>(ta, W.tmap256, p, st); break;
    default:
      set_error("linear: unsupported tile variant bn=" + std::to_string(o.bn) + " stages=" + std::to_string(o.stages));
      return false;
  }
  return cuda_ok(cudaGetLastError(), "gemm_tcgen05 launch");
}

}  // namespace conzic
