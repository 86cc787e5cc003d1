// fp32-accurate GEMMs on the bf16 tensor cores (SN_PREC_BF16X6): every fp32 operand element is split into three bf16
// limbs x = x0 + x1 + x2 (x0 = bf16(x), x1 = bf16(x - x0), x2 = bf16(x - x0 - x1); 3 x 8 significand bits = the 24 of
// fp32, so the split is exact up to limb underflow) and the product is evaluated as the six limb products of order
// <= 2^-16:  a0b0 + a0b1 + a1b0 + a1b1 + a0b2 + a2b0   (dropped: a1b2, a2b1, a2b2 <= 2^-24 relative),
// accumulated in fp32 in TMEM.  The six products become ONE ordinary bf16 GEMM with a 6x longer contraction by laying
// the limbs out along K:   A' = [a1 | a0 | a2 | a0 | a1 | a0],  B' = [b1 | b2 | b0 | b1 | b0 | b0]
// so the tcgen05 kernels (sn_gemm2.cu / sn_gemm_tc.cu) are used unchanged.  This kernel writes A' / B'.
// replaces, in fp32 mode, the fp32 FFMA GEMM (sn_gemm) behind every nn.Linear of the path (stylenet/model.py:119-150,193).
#include <cuda_bf16.h>

#include "sn_common.cuh"

namespace {

__device__ __forceinline__ void limbs(float x, __nv_bfloat16 (&l)[3]) {
  l[0] = __float2bfloat16_rn(x);
  const float r1 = x - __bfloat162float(l[0]);
  l[1] = __float2bfloat16_rn(r1);
  const float r2 = r1 - __bfloat162float(l[1]);
  l[2] = __float2bfloat16_rn(r2);
}

// which limb goes into slot s of the left (pattern 0) / right (pattern 1) operand
__device__ __forceinline__ int slot_limb(int pattern, int s) {
  // products in the order a1b1, a0b2, a2b0, a0b1, a1b0, a0b0: the tensor core adds into its fp32 accumulator with a
  // truncating alignment (error ~2^-23 of the ACCUMULATOR per MMA), so the small limb products go first, while the
  // accumulator is still small, and the dominant a0b0 block last
  const int left[6] = {1, 0, 2, 0, 1, 0}, right[6] = {1, 2, 0, 1, 0, 0};
  return pattern ? right[s] : left[s];
}

// K along the columns: src [R, G*K] (row pitch ld) -> dst [R, G*6*Kp]; group g, slot s, column k at g*6*Kp + s*Kp + k
__global__ void split_cols_kernel(const float* __restrict__ src, int64_t R, int G, int K, int64_t ld,
                                  __nv_bfloat16* __restrict__ dst, int Kp, int pattern) {
  const int64_t row = blockIdx.y;
  const int64_t ldd = (int64_t)G * 6 * Kp;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < G * Kp; i += gridDim.x * blockDim.x) {
    const int g = i / Kp, k = i - g * Kp;
    __nv_bfloat16 l[3];
    limbs(k < K ? src[row * ld + (int64_t)g * K + k] : 0.f, l);
    __nv_bfloat16* d = dst + row * ldd + (int64_t)g * 6 * Kp + k;
#pragma unroll
    for (int s = 0; s < 6; ++s) d[(int64_t)s * Kp] = l[slot_limb(pattern, s)];
  }
}

// K along the rows: src [G*K, C] (row pitch ld; groups of K rows) -> dst [G*6*Kp, Cp]; group g, slot s, row k at
// g*6*Kp + s*Kp + k (rows k >= K and columns >= C are zero)
__global__ void split_rows_kernel(const float* __restrict__ src, int G, int K, int64_t C, int64_t ld,
                                  __nv_bfloat16* __restrict__ dst, int Kp, int64_t Cp, int pattern) {
  const int g = blockIdx.z, k = blockIdx.y;
  for (int64_t c = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; c < Cp; c += (int64_t)gridDim.x * blockDim.x) {
    __nv_bfloat16 l[3];
    limbs((k < K && c < C) ? src[((int64_t)g * K + k) * ld + c] : 0.f, l);
#pragma unroll
    for (int s = 0; s < 6; ++s) dst[((int64_t)g * 6 * Kp + (int64_t)s * Kp + k) * Cp + c] = l[slot_limb(pattern, s)];
  }
}

}  // namespace

extern "C" {

int32_t sn_split_limbs_cols(const float* src, int64_t R, int64_t G, int64_t K, int64_t ld, void* dst, int64_t Kp,
                            int32_t pattern, void* stream) {
  SN_REQUIRE(src && dst && R >= 0 && G >= 1 && K >= 1 && Kp >= K && Kp % 8 == 0, "sn_split_limbs_cols: bad argument");
  SN_REQUIRE(pattern == 0 || pattern == 1, "sn_split_limbs_cols: pattern 0 (left) or 1 (right)");
  SN_REQUIRE(R <= 65535 * 1024LL, "sn_split_limbs_cols: too many rows");
  if (R == 0) return 0;
  // grid.y is limited to 65535 rows: loop over row blocks
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t ldd = G * 6 * Kp;
  for (int64_t r0 = 0; r0 < R; r0 += 65535) {
    const int64_t nr = R - r0 < 65535 ? R - r0 : 65535;
    const unsigned gx = (unsigned)((G * Kp + 255) / 256);
    split_cols_kernel<<<dim3(gx < 64 ? gx : 64, (unsigned)nr), 256, 0, st>>>(
        src + r0 * ld, nr, (int)G, (int)K, ld, (__nv_bfloat16*)dst + r0 * ldd, (int)Kp, pattern);
  }
  return sn::check_launch("sn_split_limbs_cols");
}

int32_t sn_split_limbs_rows(const float* src, int64_t G, int64_t K, int64_t C, int64_t ld, void* dst, int64_t Kp,
                            int64_t Cp, int32_t pattern, void* stream) {
  SN_REQUIRE(src && dst && G >= 1 && G <= 65535 && K >= 1 && Kp >= K && C >= 1 && Cp >= C && Cp % 8 == 0,
             "sn_split_limbs_rows: bad argument");
  SN_REQUIRE(pattern == 0 || pattern == 1, "sn_split_limbs_rows: pattern 0 (left) or 1 (right)");
  SN_REQUIRE(Kp <= 65535, "sn_split_limbs_rows: contraction length %lld too long for one launch", (long long)Kp);
  const unsigned gx = (unsigned)((Cp + 255) / 256);
  split_rows_kernel<<<dim3(gx < 64 ? gx : 64, (unsigned)Kp, (unsigned)G), 256, 0, (cudaStream_t)stream>>>(
      src, (int)G, (int)K, C, ld, (__nv_bfloat16*)dst, (int)Kp, Cp, pattern);
  return sn::check_launch("sn_split_limbs_rows");
}

}  // extern "C"
