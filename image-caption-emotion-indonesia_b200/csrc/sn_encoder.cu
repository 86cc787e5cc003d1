// Encoder tail -> decoder hand-off: the two small stages between the ResNet trunk and the
// caption decoder, each as one pass over its input.
//   * attention models (stylenet/model_att.py:19-28): AdaptiveAvgPool2d((S,S)) + permute(0,2,3,1).  The reference
//     hands the decoder a NON-contiguous NHWC view of an NCHW tensor, which the decoder then copies (.view needs
//     .contiguous()) and reduces again for init_h / init_c (mean over pixels, model_att.py:185-194).
//     sn_pool_nhwc_fwd reads the NCHW trunk output once and writes the pooled map directly as contiguous NHWC fp32
//     (+ optional bf16 GEMM-operand copy) together with its mean over the P = S*S pixels.
//   * non-attention models (stylenet/model.py:19-26): Linear(2048, E) + BatchNorm1d(E, momentum = 0.01).  The Linear is
//     a plain GEMM (sn_gemm); sn_bn1d_fwd / sn_bn1d_bwd are the batch-norm halves (batch statistics + running-stat
//     update in training, running statistics in eval), one thread per feature column.
#include <cuda_bf16.h>

#include "sn_common.cuh"

namespace {

constexpr int PC = 32;      // channels per CTA (one 128-byte NHWC segment per output pixel)

// grid (D / PC, B); block 256.  Phase 1: the PC x (h*w) input tile is read with lanes along the contiguous h*w axis.
// Phase 2: every (output pixel, channel) averages its adaptive bin from shared memory; lanes along the channel axis so
// the NHWC stores are full 128-byte segments.
__global__ void __launch_bounds__(256) pool_nhwc_kernel(const float* __restrict__ x, int D, int h, int w, int S,
                                                        float* __restrict__ out, __nv_bfloat16* __restrict__ outb,
                                                        float* __restrict__ mean) {
  extern __shared__ float tile[];              // [PC][hw + 1]
  const int hw = h * w, ld = hw + 1;
  const int d0 = blockIdx.x * PC, b = blockIdx.y;
  const int nd = min(PC, D - d0);
  const float* src = x + ((int64_t)b * D + d0) * hw;
  for (int i = threadIdx.x; i < nd * hw; i += blockDim.x) {
    const int c = i / hw, p = i - c * hw;
    tile[c * ld + p] = __ldg(src + i);
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  const int P = S * S;
  float msum = 0.f;
  for (int p = warp; p < P; p += nwarp) {
    const int oy = p / S, ox = p - oy * S;
    // torch adaptive pooling bins: [floor(o*in/out), ceil((o+1)*in/out))
    const int y0 = (oy * h) / S, y1 = ((oy + 1) * h + S - 1) / S;
    const int x0 = (ox * w) / S, x1 = ((ox + 1) * w + S - 1) / S;
    if (lane < nd) {
      float s = 0.f;
      for (int yy = y0; yy < y1; ++yy)
        for (int xx = x0; xx < x1; ++xx) s += tile[lane * ld + yy * w + xx];
      const float v = s / (float)((y1 - y0) * (x1 - x0));
      const int64_t o = ((int64_t)b * P + p) * D + d0 + lane;
      out[o] = v;
      if (outb) outb[o] = __float2bfloat16(v);
      msum += v;
    }
  }
  if (mean) {
    // per-warp partial sums of the channel `lane` -> one per CTA through shared memory (tile is free now)
    __syncthreads();
    tile[warp * PC + lane] = msum;
    __syncthreads();
    if (warp == 0 && lane < nd) {
      float s = 0.f;
      for (int q = 0; q < nwarp; ++q) s += tile[q * PC + lane];
      mean[(int64_t)b * D + d0 + lane] = s / (float)P;
    }
  }
}

// backward of pool + permute: dx[b,d,y,x] = sum over the output bins covering (y,x) of dout[b,p,d] / |bin|
__global__ void __launch_bounds__(256) pool_nhwc_bwd_kernel(const float* __restrict__ dout, int D, int h, int w, int S,
                                                            float* __restrict__ dx) {
  extern __shared__ float tile[];              // [PC][P + 1] of dout / |bin|
  const int hw = h * w, P = S * S, ld = P + 1;
  const int d0 = blockIdx.x * PC, b = blockIdx.y;
  const int nd = min(PC, D - d0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  for (int p = warp; p < P; p += nwarp) {
    const int oy = p / S, ox = p - oy * S;
    const int y0 = (oy * h) / S, y1 = ((oy + 1) * h + S - 1) / S;
    const int x0 = (ox * w) / S, x1 = ((ox + 1) * w + S - 1) / S;
    if (lane < nd) tile[lane * ld + p] = __ldg(dout + ((int64_t)b * P + p) * D + d0 + lane) / (float)((y1 - y0) * (x1 - x0));
  }
  __syncthreads();
  float* dst = dx + ((int64_t)b * D + d0) * hw;
  for (int i = threadIdx.x; i < nd * hw; i += blockDim.x) {
    const int c = i / hw, q = i - c * hw;
    const int yy = q / w, xx = q - yy * w;
    float s = 0.f;
    for (int oy = 0; oy < S; ++oy) {
      const int y0 = (oy * h) / S, y1 = ((oy + 1) * h + S - 1) / S;
      if (yy < y0 || yy >= y1) continue;
      for (int ox = 0; ox < S; ++ox) {
        const int x0 = (ox * w) / S, x1 = ((ox + 1) * w + S - 1) / S;
        if (xx >= x0 && xx < x1) s += tile[c * ld + oy * S + ox];
      }
    }
    dst[i] = s;
  }
}

// BatchNorm1d over [B, E]: one thread per column (B is 64-96 on this path; lanes along E -> coalesced rows)
__global__ void bn1d_fwd_kernel(const float* __restrict__ x, int64_t B, int64_t E, const float* __restrict__ gamma,
                                const float* __restrict__ beta, float* __restrict__ run_mean, float* __restrict__ run_var,
                                float momentum, float eps, int training, float* __restrict__ y,
                                float* __restrict__ save_mean, float* __restrict__ save_invstd) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  float mu, invstd;
  if (training) {
    float s = 0.f;
    for (int64_t b = 0; b < B; ++b) s += x[b * E + e];
    mu = s / (float)B;
    float q = 0.f;
    for (int64_t b = 0; b < B; ++b) { const float d = x[b * E + e] - mu; q += d * d; }
    const float var = q / (float)B;                       // biased: used to normalise
    invstd = rsqrtf(var + eps);
    if (run_mean) run_mean[e] = (1.f - momentum) * run_mean[e] + momentum * mu;
    if (run_var) run_var[e] = (1.f - momentum) * run_var[e] + momentum * (B > 1 ? q / (float)(B - 1) : var);   // unbiased
  } else {
    mu = run_mean[e];
    invstd = rsqrtf(run_var[e] + eps);
  }
  const float g = gamma ? gamma[e] : 1.f, bt = beta ? beta[e] : 0.f;
  for (int64_t b = 0; b < B; ++b) y[b * E + e] = (x[b * E + e] - mu) * invstd * g + bt;
  if (save_mean) save_mean[e] = mu;
  if (save_invstd) save_invstd[e] = invstd;
}

__global__ void bn1d_bwd_kernel(const float* __restrict__ x, const float* __restrict__ dy, int64_t B, int64_t E,
                                const float* __restrict__ gamma, const float* __restrict__ save_mean,
                                const float* __restrict__ save_invstd, int training, float* __restrict__ dx,
                                float* __restrict__ dgamma, float* __restrict__ dbeta) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  const float mu = save_mean[e], invstd = save_invstd[e], g = gamma ? gamma[e] : 1.f;
  float sdy = 0.f, sdyx = 0.f;
  for (int64_t b = 0; b < B; ++b) {
    const float d = dy[b * E + e];
    sdy += d;
    sdyx += d * (x[b * E + e] - mu) * invstd;
  }
  if (dgamma) dgamma[e] = sdyx;
  if (dbeta) dbeta[e] = sdy;
  if (dx) {
    if (training) {
      const float inv_b = 1.f / (float)B;
      for (int64_t b = 0; b < B; ++b) {
        const float xh = (x[b * E + e] - mu) * invstd;
        dx[b * E + e] = g * invstd * (dy[b * E + e] - sdy * inv_b - xh * sdyx * inv_b);
      }
    } else {
      for (int64_t b = 0; b < B; ++b) dx[b * E + e] = g * invstd * dy[b * E + e];
    }
  }
}

}  // namespace

extern "C" {

int32_t sn_pool_nhwc_fwd(const float* x, int64_t B, int64_t D, int64_t h, int64_t w, int64_t S, float* out,
                         void* out_bf16, float* mean, void* stream) {
  SN_REQUIRE(x && out, "sn_pool_nhwc_fwd: null argument");
  SN_REQUIRE(B >= 0 && D > 0 && h > 0 && w > 0 && S > 0, "sn_pool_nhwc_fwd: bad dims");
  SN_REQUIRE(h * w <= 1024, "sn_pool_nhwc_fwd: feature map %lldx%lld too large (h*w <= 1024)", (long long)h, (long long)w);
  if (B == 0) return 0;
  size_t smem = sizeof(float) * (size_t)PC * (size_t)(h * w + 1);
  if (smem < sizeof(float) * 8 * PC) smem = sizeof(float) * 8 * PC;
  SN_CUDA(cudaFuncSetAttribute(pool_nhwc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((D + PC - 1) / PC), (unsigned)B);
  pool_nhwc_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(x, (int)D, (int)h, (int)w, (int)S, out,
                                                             (__nv_bfloat16*)out_bf16, mean);
  return sn::check_launch("sn_pool_nhwc_fwd");
}

int32_t sn_pool_nhwc_bwd(const float* dout, int64_t B, int64_t D, int64_t h, int64_t w, int64_t S, float* dx,
                         void* stream) {
  SN_REQUIRE(dout && dx, "sn_pool_nhwc_bwd: null argument");
  SN_REQUIRE(B >= 0 && D > 0 && h > 0 && w > 0 && S > 0 && S * S <= 1024, "sn_pool_nhwc_bwd: bad dims");
  if (B == 0) return 0;
  size_t smem = sizeof(float) * (size_t)PC * (size_t)(S * S + 1);
  SN_CUDA(cudaFuncSetAttribute(pool_nhwc_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((D + PC - 1) / PC), (unsigned)B);
  pool_nhwc_bwd_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>(dout, (int)D, (int)h, (int)w, (int)S, dx);
  return sn::check_launch("sn_pool_nhwc_bwd");
}

int32_t sn_bn1d_fwd(const float* x, int64_t B, int64_t E, const float* gamma, const float* beta, float* running_mean,
                    float* running_var, float momentum, float eps, int32_t training, float* y, float* save_mean,
                    float* save_invstd, void* stream) {
  SN_REQUIRE(x && y && B > 0 && E > 0, "sn_bn1d_fwd: bad argument");
  SN_REQUIRE(training || (running_mean && running_var), "sn_bn1d_fwd: eval mode needs running statistics");
  bn1d_fwd_kernel<<<(unsigned)((E + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      x, B, E, gamma, beta, running_mean, running_var, momentum, eps, training, y, save_mean, save_invstd);
  return sn::check_launch("sn_bn1d_fwd");
}

int32_t sn_bn1d_bwd(const float* x, const float* dy, int64_t B, int64_t E, const float* gamma, const float* save_mean,
                    const float* save_invstd, int32_t training, float* dx, float* dgamma, float* dbeta, void* stream) {
  SN_REQUIRE(x && dy && save_mean && save_invstd && B > 0 && E > 0, "sn_bn1d_bwd: bad argument");
  bn1d_bwd_kernel<<<(unsigned)((E + 127) / 128), 128, 0, (cudaStream_t)stream>>>(
      x, dy, B, E, gamma, save_mean, save_invstd, training, dx, dgamma, dbeta);
  return sn::check_launch("sn_bn1d_bwd");
}

}  // extern "C"
