// out[M,N] = epilogue(A[M,K] * W[N,K]^T): the one dense contraction both towers are made of
// (HF:models/clip/modeling_clip.py:295-298,347-351 ; HF:models/bert/modeling_bert.py:179-181,295,340,353,481-501).
//
// Product kernel: warp-specialised sm_100a GEMM.
//   warp 0   : TMA producer  -- cp.async.bulk.tensor 2-D tiles (128B swizzle) of A and W into a smem ring
//   warp 1   : allocates TMEM, one elected thread issues tcgen05.mma (128 x BN x 16, bf16 -> fp32 in TMEM)
//   warps 2-5: epilogue      -- tcgen05.ld (thread = output row), bias / activation / residual, global store
// Several CTAs are resident per SM so one CTA's epilogue overlaps another's main loop.
// bf16x3 mode runs three passes over K (hi*hi, lo*hi, hi*lo) into the same accumulator.
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
};

template <int BN, int STAGES>
struct SmemLayout {
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int BAR_OFF = STAGES * (A_BYTES + B_BYTES);
  static constexpr int TOTAL = BAR_OFF + (2 * STAGES + 1) * 8 + 16;
  static constexpr int DYN_BYTES = TOTAL + 1024;  // slack for manual 1024B alignment
};

__device__ __forceinline__ float apply_act(float v, int act, bool precise) {
  if (act == ACT_QUICK_GELU) {
    // x * sigmoid(1.702 x)   (HF:activations.py:122-123)
    if (precise) return v / (1.0f + expf(-1.702f * v));
    return __fdividef(v, 1.0f + __expf(-1.702f * v));
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
      } else {
#pragma unroll
        for (int j = 0; j < 32; j += 2) store_act_pair(o, e.out_K, 1, n0 + j, v[j], v[j + 1]);
      }
    }
  } else {
    for (int j = 0; j < 32; ++j) {
      const int n = n0 + j;
      if (n >= p.N) break;
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

template <int BN, int STAGES>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const GemmParams p) {
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
  const int nkb = p.K / BK;
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
        const int pass = i / nkb, kb = i - pass * nkb;
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
        epilogue_chunk(p, row, n0, v);
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
  dim3 grid((p.N + BN - 1) / BN, (p.M + BM - 1) / BM);
  gemm_tcgen05_kernel<BN, STAGES><<<grid, GEMM_THREADS, SmemLayout<BN, STAGES>::DYN_BYTES, st>>>(ta, tb, p);
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

bool gemm_configure() {
  return configure_one<128, 2>() && configure_one<128, 3>() && configure_one<128, 4>() && configure_one<128, 6>() &&
         configure_one<256, 2>() && configure_one<256, 4>();
}

bool launch_linear(const Act& A, int M, const LinearW& W, const Epi& epi, const GemmOpts& o, cudaStream_t st,
                   uint64_t* launches) {
  if (M <= 0) return true;
  if (A.K != W.K || (W.K % BK) != 0) {
    set_error("linear: K mismatch or K not a multiple of 64");
    return false;
  }
  GemmParams p;
  p.M = M; p.N = W.N; p.K = W.K; p.split = o.split; p.e = epi;
  (void)launches;
  ++g_launches;
  ProfScope prof_(CAT_GEMM, 2.0 * M * static_cast<double>(W.N) * W.K * (o.split ? 3 : 1), st);
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
