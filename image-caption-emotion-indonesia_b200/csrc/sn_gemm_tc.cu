// K2 on the 5th-generation tensor cores: bf16 operands, fp32 accumulation in TMEM.
//
//   C[M,N] = op(A) op(B) (+ bias) (+ beta*C)        op in {NT, NN, TN}, optionally 1..4 groups
//
// One 128x128 output tile per CTA (grid sized from the tile count; the per-step GEMMs of this path are
// 60..1300 tiles).  Warp roles (192 threads):
//   warp 0      TMA producer: cp.async.bulk.tensor tiles of A and B into a 4-stage shared-memory ring
//               (128-byte swizzle), completion on the stage's "full" mbarrier
//   warp 1      MMA issuer: one elected lane issues tcgen05.mma.cta_group::1.kind::f16 (M=128, N=128, K=16)
//               four times per 64-deep stage, then tcgen05.commit -> the stage's "empty" mbarrier;
//               allocates / frees the 128 TMEM columns of the accumulator
//   warps 2..5  epilogue: tcgen05.ld the accumulator (lane quadrant = warp_id % 4), add bias / beta*C,
//               store fp32 and/or bf16 straight to global memory
// Operand layouts: a K-major operand tile is 128 rows x 128 B (one TMA box); an MN-major operand tile
// (the "T" side of TN, the second operand of NN) is two boxes of 64 k-rows x 128 B, consumed through an
// MN-major UMMA descriptor -- no materialised transposes anywhere.
// K / M / N tails are zero-filled by TMA and masked in the epilogue.
#include <cuda.h>
#include <cuda_bf16.h>

#include "sn_common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 64;
constexpr int STAGES = 3;   // 3 x 32 KB: two CTAs per SM -> one CTA's epilogue overlaps the other's MMA main loop
constexpr int UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2, B_BYTES = BN * BK * 2;       // 16 KB each
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int TMEM_COLS = 128;
constexpr int NTHREADS = 192;

struct TmaSet { CUtensorMap m[4]; };

struct TcArgs {
  int M, N, K;
  int splits;                      // split-K factor (>1: fp32 atomic accumulation into a zeroed C)
  int a_mn_major, b_mn_major;     // 0: K-major tile, 1: MN-major tile
  float* C; __nv_bfloat16* Cb; int64_t ldc, ldcb;
  const float* bias; float beta;
  int64_t strideC, strideCb, strideBias;
};

// ---- PTX wrappers ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
        "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
        "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14), LBO>>4 [16,30),
// SBO>>4 [32,46), version=1 [46,48), layout_type [61,64) (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

__global__ void __launch_bounds__(NTHREADS, 2)
gemm_tc_kernel(const __grid_constant__ TmaSet tma_a, const __grid_constant__ TmaSet tma_b, TcArgs g) {
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B swizzle atoms
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
  const uint32_t smem_base = smem_u32(smem);
  const uint32_t full0 = smem_u32(bars), empty0 = smem_u32(bars + STAGES), tfull = smem_u32(bars + 2 * STAGES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int grp = blockIdx.z / g.splits, split = blockIdx.z - grp * g.splits;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int total_kb = (g.K + BK - 1) / BK;
  const int kb_begin = (int)(((int64_t)total_kb * split) / g.splits);
  const int kb_end = (int)(((int64_t)total_kb * (split + 1)) / g.splits);
  const int num_kb = kb_end - kb_begin;
  const CUtensorMap* map_a = &tma_a.m[grp];
  const CUtensorMap* map_b = &tma_b.m[grp];

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(full0 + 8 * s, 1); mbar_init(empty0 + 8 * s, 1); }
    mbar_init(tfull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map_a) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(map_b) : "memory");
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % STAGES;
        const uint32_t ph = (kb / STAGES) & 1;
        mbar_wait(empty0 + 8 * s, ph ^ 1);
        const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_BYTES;
        mbar_expect_tx(full0 + 8 * s, STAGE_BYTES);
        const int k0 = (kb_begin + kb) * BK;
        if (!g.a_mn_major) {
          tma_load_2d(sa, map_a, full0 + 8 * s, k0, m0);                 // box {64 k, 128 rows}
        } else {
          tma_load_2d(sa, map_a, full0 + 8 * s, m0, k0);                 // box {64 m, 64 k-rows}
          tma_load_2d(sa + A_BYTES / 2, map_a, full0 + 8 * s, m0 + 64, k0);
        }
        if (!g.b_mn_major) {
          tma_load_2d(sb, map_b, full0 + 8 * s, k0, n0);
        } else {
          tma_load_2d(sb, map_b, full0 + 8 * s, n0, k0);
          tma_load_2d(sb + B_BYTES / 2, map_b, full0 + 8 * s, n0 + 64, k0);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1,
    // a_major bit15, b_major bit16, N>>3 [17,23), M>>4 [24,29)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)g.a_mn_major << 15) |
                           ((uint32_t)g.b_mn_major << 16) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int s = kb % STAGES;
      const uint32_t ph = (kb / STAGES) & 1;
      mbar_wait(full0 + 8 * s, ph);
      tc_fence_after();
      if (lane == 0) {
        const uint32_t sa = smem_base + s * STAGE_BYTES, sb = sa + A_BYTES;
#pragma unroll
        for (int k = 0; k < BK / UMMA_K; ++k) {
          // K-major: advance 32 B inside the 128 B swizzle row; SBO = 1024 B between 8-row groups.
          // MN-major: advance two 8-k-row groups (2 KB); LBO = 8 KB between the two 64-wide MN chunks.
          const uint64_t ad = g.a_mn_major ? make_desc(sa + k * 2048, A_BYTES / 2, 1024) : make_desc(sa + k * 32, 16, 1024);
          const uint64_t bd = g.b_mn_major ? make_desc(sb + k * 2048, B_BYTES / 2, 1024) : make_desc(sb + k * 32, 16, 1024);
          umma_bf16(tmem_base, ad, bd, idesc, (kb | k) ? 1u : 0u);
        }
        umma_commit(empty0 + 8 * s);                 // frees the smem stage when these MMAs retire
        if (kb == num_kb - 1) umma_commit(tfull);    // accumulator complete
      }
      __syncwarp();
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;                          // TMEM lane quadrant this warp may access
    mbar_wait(tfull, 0);
    tc_fence_after();
    const int row = m0 + q * 32 + lane;
    float* Cg = g.C ? g.C + grp * g.strideC : nullptr;
    __nv_bfloat16* Cbg = g.Cb ? g.Cb + grp * g.strideCb : nullptr;
    const float* bias = (g.bias && split == 0) ? g.bias + grp * g.strideBias : nullptr;
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + c0, r);
      if (g.splits > 1) {
        // split-K: partial tile -> fp32 reduction in L2 (C was zeroed by the host wrapper)
        if (row < g.M && num_kb > 0) {
          float* dst = Cg + (int64_t)row * g.ldc + n0 + c0;
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (n0 + c0 + j < g.N) atomicAdd(dst + j, __uint_as_float(r[j]) + (bias ? bias[n0 + c0 + j] : 0.f));
        }
        continue;
      }
      if (row < g.M) {
        const int nb = n0 + c0;
        if (nb + 32 <= g.N) {
          float v[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]) + (bias ? __ldg(bias + nb + j) : 0.f);
          if (Cg) {
            float* dst = Cg + (int64_t)row * g.ldc + nb;
            if (g.beta != 0.f) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] += g.beta * dst[j];
            }
            if ((reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) dst[j] = v[j];
            }
          }
          if (Cbg) {
            __nv_bfloat16* dstb = Cbg + (int64_t)row * g.ldcb + nb;
            if ((reinterpret_cast<uintptr_t>(dstb) & 15) == 0) {
#pragma unroll
              for (int j = 0; j < 32; j += 8) {
                __nv_bfloat162 p0 = __floats2bfloat162_rn(v[j], v[j + 1]), p1 = __floats2bfloat162_rn(v[j + 2], v[j + 3]);
                __nv_bfloat162 p2 = __floats2bfloat162_rn(v[j + 4], v[j + 5]), p3 = __floats2bfloat162_rn(v[j + 6], v[j + 7]);
                uint4 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
                pk.z = *reinterpret_cast<uint32_t*>(&p2); pk.w = *reinterpret_cast<uint32_t*>(&p3);
                *reinterpret_cast<uint4*>(dstb + j) = pk;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) dstb[j] = __float2bfloat16(v[j]);
            }
          }
        } else {
          for (int j = 0; j < 32; ++j) {
            const int n = nb + j;
            if (n < g.N) {
              float v = __uint_as_float(r[j]) + (bias ? bias[n] : 0.f);
              if (Cg) {
                float* dst = Cg + (int64_t)row * g.ldc + n;
                if (g.beta != 0.f) v += g.beta * *dst;
                *dst = v;
              }
              if (Cbg) Cbg[(int64_t)row * g.ldcb + n] = __float2bfloat16(v);
            }
          }
        }
      }
    }
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- host side: tensor maps -----------------------------------------------------------------------------
typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn get_encode() {
  static EncodeFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeFn>(p);
  }
  return fn;
}

// 2-D bf16 tensor map: inner (contiguous) extent `inner`, `rows` rows of pitch `pitch_elems`
int32_t encode_2d(CUtensorMap* map, const void* base, int64_t inner, int64_t rows, int64_t pitch_elems, int box_inner,
                  int box_rows) {
  EncodeFn enc = get_encode();
  if (!enc) return sn::fail(-4, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)inner, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)pitch_elems * 2};
  cuuint32_t box[2] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return sn::fail(-5, "cuTensorMapEncodeTiled failed (%d): base=%p inner=%lld rows=%lld pitch=%lld", (int)r,
                                         base, (long long)inner, (long long)rows, (long long)pitch_elems);
  return 0;
}

}  // namespace

extern "C" int32_t sn_gemm_bf16_splitk(int32_t op, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda,
                                       const void* B, int64_t ldb, float* C, int64_t ldc, void* Cb, int64_t ldcb,
                                       const float* bias, float beta, int32_t batch, int64_t strideA, int64_t strideB,
                                       int64_t strideC, int64_t strideCb, int64_t strideBias, int32_t splits,
                                       void* stream);

extern "C" int32_t sn_gemm_bf16(int32_t op, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda, const void* B,
                                int64_t ldb, float* C, int64_t ldc, void* Cb, int64_t ldcb, const float* bias, float beta,
                                int32_t batch, int64_t strideA, int64_t strideB, int64_t strideC, int64_t strideCb,
                                int64_t strideBias, void* stream) {
  return sn_gemm_bf16_splitk(op, M, N, K, A, lda, B, ldb, C, ldc, Cb, ldcb, bias, beta, batch, strideA, strideB, strideC,
                             strideCb, strideBias, 1, stream);
}

extern "C" int32_t sn_gemm_bf16_splitk(int32_t op, int64_t M, int64_t N, int64_t K, const void* A, int64_t lda,
                                       const void* B, int64_t ldb, float* C, int64_t ldc, void* Cb, int64_t ldcb,
                                       const float* bias, float beta, int32_t batch, int64_t strideA, int64_t strideB,
                                       int64_t strideC, int64_t strideCb, int64_t strideBias, int32_t splits,
                                       void* stream) {
  SN_REQUIRE(op >= 0 && op <= 2, "sn_gemm_bf16: bad op %d", op);
  SN_REQUIRE(batch >= 1 && batch <= 4, "sn_gemm_bf16: 1..4 groups supported, got %d", batch);
  SN_REQUIRE(M >= 0 && N >= 0 && K > 0, "sn_gemm_bf16: bad dims");
  if (M == 0 || N == 0) return 0;
  SN_REQUIRE(A && B && (C || Cb), "sn_gemm_bf16: null operand");
  SN_REQUIRE((lda % 8) == 0 && (ldb % 8) == 0 && (strideA % 8) == 0 && (strideB % 8) == 0,
             "sn_gemm_bf16: bf16 leading dimensions / group strides must be multiples of 8 elements (TMA 16-byte rule): "
             "lda=%lld ldb=%lld", (long long)lda, (long long)ldb);
  SN_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0, "sn_gemm_bf16: operands must be 16-byte aligned");
  const int64_t total_kb = (K + BK - 1) / BK;
  if (splits <= 0) {
    // auto: fill ~2 waves of the 148 SMs when the tile count alone cannot (fp32 output, no beta only)
    const int64_t tiles = ((M + BM - 1) / BM) * ((N + BN - 1) / BN) * batch;
    splits = 1;
    if (!Cb && beta == 0.f && tiles < 100 && total_kb >= 8) {
      int64_t want = (2 * (int64_t)sn::dev_info().sm_count + tiles / 2) / tiles;
      int64_t cap = total_kb / 4;
      if (want > cap) want = cap;
      if (want > 8) want = 8;
      if (want > 1) splits = (int)want;
    }
  }
  if (splits > total_kb) splits = (int)total_kb;
  if (splits < 1) splits = 1;
  if (splits > 1) {
    SN_REQUIRE(C && !Cb && beta == 0.f, "sn_gemm_bf16: split-K needs an fp32-only output and beta == 0");
    // zero the output tile range (row pitch ldc): one memset per group when rows are dense, else per row
    for (int gi = 0; gi < batch; ++gi) {
      float* base = C + gi * strideC;
      if (ldc == N) {
        SN_CUDA(cudaMemsetAsync(base, 0, sizeof(float) * (size_t)M * (size_t)N, (cudaStream_t)stream));
      } else {
        SN_CUDA(cudaMemset2DAsync(base, sizeof(float) * (size_t)ldc, 0, sizeof(float) * (size_t)N, (size_t)M,
                                  (cudaStream_t)stream));
      }
    }
  }
  TmaSet ta, tb;
  TcArgs g;
  g.M = (int)M; g.N = (int)N; g.K = (int)K; g.splits = splits;
  g.a_mn_major = (op == SN_OP_TN) ? 1 : 0;
  g.b_mn_major = (op == SN_OP_NT) ? 0 : 1;
  g.C = C; g.Cb = (__nv_bfloat16*)Cb; g.ldc = ldc; g.ldcb = ldcb; g.bias = bias; g.beta = beta;
  g.strideC = strideC; g.strideCb = strideCb; g.strideBias = strideBias;
  const __nv_bfloat16* Ab = (const __nv_bfloat16*)A;
  const __nv_bfloat16* Bb = (const __nv_bfloat16*)B;
  for (int i = 0; i < 4; ++i) {
    int gi = i < batch ? i : 0;
    int32_t rc;
    if (!g.a_mn_major) rc = encode_2d(&ta.m[i], Ab + gi * strideA, K, M, lda, BK, BM);      // A[M,K]
    else rc = encode_2d(&ta.m[i], Ab + gi * strideA, M, K, lda, 64, BK);                      // A stored [K,M]
    if (rc) return rc;
    if (!g.b_mn_major) rc = encode_2d(&tb.m[i], Bb + gi * strideB, K, N, ldb, BK, BN);      // B[N,K]
    else rc = encode_2d(&tb.m[i], Bb + gi * strideB, N, K, ldb, 64, BK);                      // B stored [K,N]
    if (rc) return rc;
  }
  static thread_local bool configured = false;
  if (!configured) {
    SN_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    configured = true;
  }
  dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)((M + BM - 1) / BM), (unsigned)(batch * splits));
  gemm_tc_kernel<<<grid, NTHREADS, SMEM_BYTES, (cudaStream_t)stream>>>(ta, tb, g);
  return sn::check_launch("sn_gemm_bf16");
}

// ---- fp32 -> bf16 cast with row padding (weight shadows, activations) -------------------------------------
namespace {
__global__ void cast_pad_kernel(const float* __restrict__ src, int64_t R, int64_t Cc, int64_t lds,
                                __nv_bfloat16* __restrict__ dst, int64_t Cp, int64_t ldd) {
  const int64_t total = R * Cp;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    int64_t r = i / Cp, c = i - r * Cp;
    dst[r * ldd + c] = __float2bfloat16(c < Cc ? src[r * lds + c] : 0.f);
  }
}
// dense, unpadded, 8 elements per thread: 2 x 128-bit loads -> one 128-bit store
__global__ void cast_vec8_kernel(const float* __restrict__ src, int64_t n8, __nv_bfloat16* __restrict__ dst) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src) + 2 * i);
    const float4 b = __ldg(reinterpret_cast<const float4*>(src) + 2 * i + 1);
    __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
    uint4 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&p0); pk.y = *reinterpret_cast<uint32_t*>(&p1);
    pk.z = *reinterpret_cast<uint32_t*>(&p2); pk.w = *reinterpret_cast<uint32_t*>(&p3);
    reinterpret_cast<uint4*>(dst)[i] = pk;
  }
}
}  // namespace

extern "C" int32_t sn_cast_bf16_ex(const float* src, int64_t R, int64_t Cc, int64_t lds, void* dst, int64_t Cp, int64_t ldd,
                                   int32_t max_blocks, void* stream) {
  SN_REQUIRE(R >= 0 && Cc >= 0 && Cp >= Cc && ldd >= Cp, "sn_cast_bf16: bad dims");
  if (R == 0 || Cp == 0) return 0;
  int64_t total = R * Cp;
  int64_t cap = (int64_t)sn::dev_info().sm_count * 16;
  if (max_blocks > 0 && max_blocks < cap) cap = max_blocks;
  const bool dense = Cc == Cp && lds == Cc && ldd == Cp && (total % 8) == 0 && ((uintptr_t)src & 15) == 0 &&
                     ((uintptr_t)dst & 15) == 0;
  if (dense) {
    int64_t blocks = (total / 8 + 255) / 256;
    if (blocks > cap) blocks = cap;
    cast_vec8_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, total / 8, (__nv_bfloat16*)dst);
    return sn::check_launch("sn_cast_bf16");
  }
  int64_t blocks = (total + 255) / 256;
  if (blocks > cap) blocks = cap;
  cast_pad_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(src, R, Cc, lds, (__nv_bfloat16*)dst, Cp, ldd);
  return sn::check_launch("sn_cast_bf16");
}

extern "C" int32_t sn_cast_bf16(const float* src, int64_t R, int64_t Cc, int64_t lds, void* dst, int64_t Cp, int64_t ldd,
                                void* stream) {
  return sn_cast_bf16_ex(src, R, Cc, lds, dst, Cp, ldd, 0, stream);
}
