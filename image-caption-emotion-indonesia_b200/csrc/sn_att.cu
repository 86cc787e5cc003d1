// K4: soft attention step -- scores, softmax over pixels, context, f_beta gate -- fused per sample.
// encoder_att(features) (att1) is time-invariant and is hoisted out of the time loop by the host
// (the reference recomputes it every step, stylenet/model_att.py:59); this kernel consumes it.
// One CTA per sample; feature/att1 rows are read with coalesced 128-bit loads; HBM-bound
// (algorithmic bytes per sample-step: P*(A+D)*4 read, (D+P)*4 written).
#include "sn_common.cuh"

namespace {

constexpr int NT = 256;

__device__ __forceinline__ float block_sum(float v, float* red) {
  v = sn::warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = 0.f;
#pragma unroll
  for (int w = 0; w < NT / 32; ++w) t += red[w];
  return t;
}

// e[p] = wfull . relu(att1[p,:] + att2) + bfull  for all pixels (warp per pixel)
__device__ __forceinline__ void scores(const float* __restrict__ att1, const float* __restrict__ att2s,
                                       const float* __restrict__ ws, float bfull, int P, int A, float* e) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int p = warp; p < P; p += NT / 32) {
    const float* row = att1 + (int64_t)p * A;
    float s = 0.f;
    for (int a = lane; a < A; a += 32) s = fmaf(ws[a], fmaxf(row[a] + att2s[a], 0.f), s);
    s = sn::warp_sum(s);
    if (lane == 0) e[p] = s + bfull;
  }
}

__global__ void __launch_bounds__(NT) att_step_fwd_kernel(const float* __restrict__ att1, const float* __restrict__ att2,
                                                          const float* __restrict__ feat, const float* __restrict__ wfull,
                                                          float bfull, const float* __restrict__ gate_pre, int P, int A,
                                                          int D, float* __restrict__ alpha, int64_t ld_alpha,
                                                          float* __restrict__ ctx, int64_t ldc) {
  extern __shared__ float sm[];
  float* att2s = sm;          // [A]
  float* ws = sm + A;         // [A]
  float* e = ws + A;          // [P]
  __shared__ float red[NT / 32];
  const int b = blockIdx.x, tid = threadIdx.x;
  for (int a = tid; a < A; a += NT) { att2s[a] = att2[(int64_t)b * A + a]; ws[a] = wfull[a]; }
  __syncthreads();
  scores(att1 + (int64_t)b * P * A, att2s, ws, bfull, P, A, e);
  __syncthreads();
  // softmax over pixels
  float mx = -INFINITY;
  for (int p = tid; p < P; p += NT) mx = fmaxf(mx, e[p]);
  mx = sn::warp_max(mx);
  __syncthreads();
  if ((tid & 31) == 0) red[tid >> 5] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < NT / 32; ++w) mx = fmaxf(mx, red[w]);
  float se = 0.f;
  for (int p = tid; p < P; p += NT) { float v = expf(e[p] - mx); e[p] = v; se += v; }
  se = block_sum(se, red);
  const float inv = 1.f / se;
  for (int p = tid; p < P; p += NT) {
    float al = e[p] * inv;
    e[p] = al;
    alpha[(int64_t)b * ld_alpha + p] = al;
  }
  __syncthreads();
  // gated context
  const float* fb = feat + (int64_t)b * P * D;
  for (int d = tid; d < D; d += NT) {
    float s = 0.f;
    for (int p = 0; p < P; ++p) s = fmaf(e[p], fb[(int64_t)p * D + d], s);
    ctx[(int64_t)b * ldc + d] = sn::sigmoidf_(gate_pre[(int64_t)b * D + d]) * s;
  }
}

__global__ void __launch_bounds__(NT) att_step_bwd_kernel(const float* __restrict__ att1, const float* __restrict__ att2,
                                                          const float* __restrict__ feat, const float* __restrict__ wfull,
                                                          float bfull, const float* __restrict__ gate_pre,
                                                          const float* __restrict__ alpha, int64_t ld_alpha,
                                                          const float* __restrict__ dctx, int64_t ldc,
                                                          const float* __restrict__ dalpha_extra, int64_t ld_da, int P,
                                                          int A, int D, float* __restrict__ datt2,
                                                          float* __restrict__ dgate_pre, float* __restrict__ datt1,
                                                          float* __restrict__ dwfull, float* __restrict__ dfeat) {
  extern __shared__ float sm[];
  float* att2s = sm;              // [A]
  float* ws = att2s + A;          // [A]
  float* al = ws + A;             // [P]
  float* dal = al + P;            // [P]  d alpha, then d e
  float* dawe = dal + P;          // [D]
  float* part = dawe + D;         // [8][A] cross-warp partials
  __shared__ float red[NT / 32];
  const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int a = tid; a < A; a += NT) { att2s[a] = att2[(int64_t)b * A + a]; ws[a] = wfull[a]; }
  for (int p = tid; p < P; p += NT) al[p] = alpha[(int64_t)b * ld_alpha + p];
  __syncthreads();
  const float* fb = feat + (int64_t)b * P * D;
  // pass A: awe_raw (recomputed), gate grads, d awe
  for (int d = tid; d < D; d += NT) {
    float s = 0.f;
    for (int p = 0; p < P; ++p) s = fmaf(al[p], fb[(int64_t)p * D + d], s);
    float g = sn::sigmoidf_(gate_pre[(int64_t)b * D + d]);
    float dc = dctx[(int64_t)b * ldc + d];
    dgate_pre[(int64_t)b * D + d] = dc * s * g * (1.f - g);
    dawe[d] = dc * g;
  }
  __syncthreads();
  // pass B: d alpha_p = dawe . feat_p (+ regulariser); optional d feat
  for (int p = warp; p < P; p += NT / 32) {
    const float* row = fb + (int64_t)p * D;
    float s = 0.f;
    for (int d = lane; d < D; d += 32) s = fmaf(dawe[d], row[d], s);
    s = sn::warp_sum(s);
    if (lane == 0) dal[p] = s + (dalpha_extra ? dalpha_extra[(int64_t)b * ld_da + p] : 0.f);
    if (dfeat) {
      float* drow = dfeat + ((int64_t)b * P + p) * D;
      const float ap = al[p];
      for (int d = lane; d < D; d += 32) drow[d] += ap * dawe[d];
    }
  }
  __syncthreads();
  // softmax backward: de_p = alpha_p (dalpha_p - sum_q alpha_q dalpha_q)
  float dot = 0.f;
  for (int p = tid; p < P; p += NT) dot += al[p] * dal[p];
  dot = block_sum(dot, red);
  for (int p = tid; p < P; p += NT) dal[p] = al[p] * (dal[p] - dot);
  __syncthreads();
  // pass C: through relu / full_att: datt1 += , datt2 = sum_p, dwfull += sum_p
  for (int a = tid; a < 2 * (NT / 32) * A; a += NT) part[a] = 0.f;
  __syncthreads();
  float* p_att2 = part + warp * A;
  float* p_w = part + (NT / 32) * A + warp * A;
  const float* a1b = att1 + (int64_t)b * P * A;
  float* d1b = datt1 + (int64_t)b * P * A;
  for (int p = warp; p < P; p += NT / 32) {
    const float de = dal[p];
    for (int a = lane; a < A; a += 32) {
      float v = a1b[(int64_t)p * A + a] + att2s[a];
      float r = fmaxf(v, 0.f);
      float dpre = v > 0.f ? de * ws[a] : 0.f;
      d1b[(int64_t)p * A + a] += dpre;
      p_att2[a] += dpre;     // lane-private slots (a == lane mod 32) inside this warp's row
      p_w[a] += de * r;
    }
  }
  __syncthreads();
  for (int a = tid; a < A; a += NT) {
    float s2 = 0.f, sw = 0.f;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) { s2 += part[w * A + a]; sw += part[(NT / 32) * A + w * A + a]; }
    datt2[(int64_t)b * A + a] = s2;
    atomicAdd(dwfull + a, sw);
  }
}

}  // namespace

extern "C" {

int32_t sn_att_step_fwd(const float* att1, const float* att2, const float* feat, const float* wfull, float bfull,
                        const float* gate_pre, int64_t nb, int64_t P, int64_t A, int64_t D, float* alpha,
                        int64_t ld_alpha, float* ctx, int64_t ldc, void* stream) {
  SN_REQUIRE(nb >= 0 && P > 0 && A > 0 && D > 0, "sn_att_step_fwd: bad dims");
  if (nb == 0) return 0;
  size_t smem = (size_t)(2 * A + P) * sizeof(float);
  SN_REQUIRE(smem <= 48 * 1024, "sn_att_step_fwd: A=%lld P=%lld exceed shared memory", (long long)A, (long long)P);
  att_step_fwd_kernel<<<(unsigned)nb, NT, smem, (cudaStream_t)stream>>>(att1, att2, feat, wfull, bfull, gate_pre, (int)P,
                                                                       (int)A, (int)D, alpha, ld_alpha, ctx, ldc);
  return sn::check_launch("sn_att_step_fwd");
}

int32_t sn_att_step_bwd(const float* att1, const float* att2, const float* feat, const float* wfull, float bfull,
                        const float* gate_pre, const float* alpha, int64_t ld_alpha, const float* dctx, int64_t ldc,
                        const float* dalpha_extra, int64_t ld_da, int64_t nb, int64_t P, int64_t A, int64_t D,
                        float* datt2, float* dgate_pre, float* datt1, float* dwfull, float* dfeat, void* stream) {
  SN_REQUIRE(nb >= 0 && P > 0 && A > 0 && D > 0, "sn_att_step_bwd: bad dims");
  if (nb == 0) return 0;
  size_t smem = (size_t)(2 * A + 2 * P + D + 2 * (NT / 32) * A) * sizeof(float);
  SN_REQUIRE(smem <= 200 * 1024, "sn_att_step_bwd: dims exceed shared memory");
  if (smem > 48 * 1024) {
    SN_CUDA(cudaFuncSetAttribute(att_step_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  att_step_bwd_kernel<<<(unsigned)nb, NT, smem, (cudaStream_t)stream>>>(
      att1, att2, feat, wfull, bfull, gate_pre, alpha, ld_alpha, dctx, ldc, dalpha_extra, ld_da, (int)P, (int)A, (int)D,
      datt2, dgate_pre, datt1, dwfull, dfeat);
  return sn::check_launch("sn_att_step_bwd");
}

}  // extern "C"
