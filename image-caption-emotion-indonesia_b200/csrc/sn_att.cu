// K4: soft attention step -- scores, softmax over pixels, context, f_beta gate -- fused per sample.
// encoder_att(features) (att1) is time-invariant and is hoisted out of the time loop by the host
// (the reference recomputes it every step, stylenet/model_att.py:59); this kernel consumes it.
// One CTA per sample; feature/att1 rows are read with coalesced 128-bit loads; HBM-bound
// (algorithmic bytes per sample-step: P*(A+D)*4 read, (D+P)*4 written).
#include <cooperative_groups.h>
#include <cuda_bf16.h>

#include "sn_common.cuh"

namespace cg = cooperative_groups;

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

// ------------------------------------------------------------------------------------------------------
// Cluster versions: CL = 4 CTAs per sample (a thread-block cluster).  The pixel axis (scores, attention-net
// backward) and the feature axis D (context, gate, d alpha partials) are each split 4 ways; the P scores /
// partial d alpha vectors are exchanged through distributed shared memory.  4x the CTAs of the per-sample
// kernels -> enough memory-level parallelism to stream the feature map near HBM speed.
// ------------------------------------------------------------------------------------------------------
constexpr int CL = 4;

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(NT)
att_step_fwd_cl_kernel(const float* __restrict__ att1, const float* __restrict__ att2, const float* __restrict__ feat,
                       const float* __restrict__ wfull, float bfull, const float* __restrict__ gate_pre, int P, int A,
                       int D, float* __restrict__ alpha, int64_t ld_alpha, float* __restrict__ ctx, int64_t ldc) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ float sm[];
  float* att2s = sm;          // [A]
  float* ws = sm + A;         // [A]
  float* e = ws + A;          // [P]  all scores of the sample (filled by the 4 CTAs through DSMEM)
  __shared__ float red[NT / 32];
  const int r = (int)cluster.block_rank();
  const int b = blockIdx.x / CL, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int a = tid; a < A; a += NT) { att2s[a] = att2[(int64_t)b * A + a]; ws[a] = wfull[a]; }
  __syncthreads();
  // scores of this CTA's pixels, broadcast into every CTA's e[]
  const float* a1b = att1 + (int64_t)b * P * A;
  for (int p = r + CL * warp; p < P; p += CL * (NT / 32)) {
    const float* row = a1b + (int64_t)p * A;
    float s = 0.f;
    for (int a = lane; a < A; a += 32) s = fmaf(ws[a], fmaxf(__ldg(row + a) + att2s[a], 0.f), s);
    s = sn::warp_sum(s) + bfull;
    if (lane < CL) cluster.map_shared_rank(e, lane)[p] = s;
  }
  cluster.sync();
  // softmax over pixels (each CTA redundantly; P is tiny)
  float mx = -INFINITY;
  for (int p = tid; p < P; p += NT) mx = fmaxf(mx, e[p]);
  mx = sn::warp_max(mx);
  if (lane == 0) red[warp] = mx;
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
    if (r == 0) alpha[(int64_t)b * ld_alpha + p] = al;
  }
  __syncthreads();
  // gated context for this CTA's D chunk, two features per thread (64-bit loads, coalesced across the warp)
  const int chunk = D / CL, d0 = r * chunk;
  const float* fb = feat + (int64_t)b * P * D + d0;
  for (int dd = 2 * tid; dd < chunk; dd += 2 * NT) {
    float sx = 0.f, sy = 0.f;
#pragma unroll 7
    for (int p = 0; p < P; ++p) {
      const float2 f = __ldg(reinterpret_cast<const float2*>(fb + (int64_t)p * D + dd));
      sx = fmaf(e[p], f.x, sx);
      sy = fmaf(e[p], f.y, sy);
    }
    const float2 gp = *reinterpret_cast<const float2*>(gate_pre + (int64_t)b * D + d0 + dd);
    ctx[(int64_t)b * ldc + d0 + dd] = sn::sigmoidf_(gp.x) * sx;
    ctx[(int64_t)b * ldc + d0 + dd + 1] = sn::sigmoidf_(gp.y) * sy;
  }
}

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(NT)
att_step_bwd_cl_kernel(const float* __restrict__ att1, const float* __restrict__ att2, const float* __restrict__ feat,
                       const float* __restrict__ wfull, const float* __restrict__ gate_pre,
                       const float* __restrict__ alpha, int64_t ld_alpha, const float* __restrict__ dctx, int64_t ldc,
                       const float* __restrict__ dalpha_extra, int64_t ld_da, int P, int A, int D,
                       float* __restrict__ datt2, float* __restrict__ dgate_pre, float* __restrict__ datt1,
                       float* __restrict__ dwfull, float* __restrict__ dfeat) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ float sm[];
  float* att2s = sm;              // [A]
  float* ws = att2s + A;          // [A]
  float* al = ws + A;             // [P]
  float* dal = al + P;            // [P]        d alpha (sum over the cluster), then d e
  float* slots = dal + P;         // [CL][P]    partial d alpha of every CTA of the cluster (DSMEM targets)
  float* part = slots + CL * P;   // [2][8][A]  cross-warp partials of pass C
  __shared__ float red[NT / 32];
  const int r = (int)cluster.block_rank();
  const int b = blockIdx.x / CL, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int a = tid; a < A; a += NT) { att2s[a] = att2[(int64_t)b * A + a]; ws[a] = wfull[a]; }
  for (int p = tid; p < P; p += NT) { al[p] = alpha[(int64_t)b * ld_alpha + p]; dal[p] = 0.f; }
  __syncthreads();
  // pass A (own D chunk, single pass over the feature map): awe_raw, gate gradient, partial d alpha
  const int chunk = D / CL, d0 = r * chunk;
  const float* fb = feat + (int64_t)b * P * D + d0;
  for (int dd0 = 0; dd0 < chunk; dd0 += 2 * NT) {
    const int dd = dd0 + 2 * tid;
    const bool act = dd < chunk;
    float gx = 0.f, gy = 0.f, dax = 0.f, day = 0.f, sx = 0.f, sy = 0.f;
    if (act) {
      const float2 gp = *reinterpret_cast<const float2*>(gate_pre + (int64_t)b * D + d0 + dd);
      gx = sn::sigmoidf_(gp.x); gy = sn::sigmoidf_(gp.y);
      dax = dctx[(int64_t)b * ldc + d0 + dd] * gx;          // d awe = dctx * gate
      day = dctx[(int64_t)b * ldc + d0 + dd + 1] * gy;
    }
    for (int p = 0; p < P; ++p) {
      float2 f = make_float2(0.f, 0.f);
      if (act) f = __ldg(reinterpret_cast<const float2*>(fb + (int64_t)p * D + dd));
      sx = fmaf(al[p], f.x, sx);
      sy = fmaf(al[p], f.y, sy);
      float t = sn::warp_sum(fmaf(dax, f.x, day * f.y));
      if (lane == 0) atomicAdd(dal + p, t);
      if (dfeat && act) {
        float* df = dfeat + ((int64_t)b * P + p) * D + d0 + dd;
        df[0] += al[p] * dax; df[1] += al[p] * day;
      }
    }
    if (act) {
      const float dcx = dctx[(int64_t)b * ldc + d0 + dd], dcy = dctx[(int64_t)b * ldc + d0 + dd + 1];
      dgate_pre[(int64_t)b * D + d0 + dd] = dcx * sx * gx * (1.f - gx);
      dgate_pre[(int64_t)b * D + d0 + dd + 1] = dcy * sy * gy * (1.f - gy);
    }
  }
  __syncthreads();
  // exchange the partial d alpha vectors: slot r of every CTA <- this CTA's partial
  for (int i = tid; i < CL * P; i += NT) {
    const int q = i / P, p = i - q * P;
    cluster.map_shared_rank(slots, q)[r * P + p] = dal[p];
  }
  cluster.sync();
  for (int p = tid; p < P; p += NT) {
    float s = dalpha_extra ? dalpha_extra[(int64_t)b * ld_da + p] : 0.f;
#pragma unroll
    for (int q = 0; q < CL; ++q) s += slots[q * P + p];
    dal[p] = s;
  }
  __syncthreads();
  // softmax backward: de_p = alpha_p (dalpha_p - sum_q alpha_q dalpha_q)
  float dot = 0.f;
  for (int p = tid; p < P; p += NT) dot += al[p] * dal[p];
  dot = block_sum(dot, red);
  for (int p = tid; p < P; p += NT) dal[p] = al[p] * (dal[p] - dot);
  for (int a = tid; a < 2 * (NT / 32) * A; a += NT) part[a] = 0.f;
  __syncthreads();
  // pass C (own pixels): through relu / full_att
  float* p_att2 = part + warp * A;
  float* p_w = part + (NT / 32) * A + warp * A;
  const float* a1b = att1 + (int64_t)b * P * A;
  float* d1b = datt1 + (int64_t)b * P * A;
  for (int p = r + CL * warp; p < P; p += CL * (NT / 32)) {
    const float de = dal[p];
    for (int a = lane; a < A; a += 32) {
      const float v = a1b[(int64_t)p * A + a] + att2s[a];
      const float dpre = v > 0.f ? de * ws[a] : 0.f;
      d1b[(int64_t)p * A + a] += dpre;
      p_att2[a] += dpre;
      p_w[a] += de * fmaxf(v, 0.f);
    }
  }
  __syncthreads();
  for (int a = tid; a < A; a += NT) {
    float s2 = 0.f, sw = 0.f;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) { s2 += part[w * A + a]; sw += part[(NT / 32) * A + w * A + a]; }
    atomicAdd(datt2 + (int64_t)b * A + a, s2);      // datt2 rows are zeroed by the host wrapper
    atomicAdd(dwfull + a, sw);
  }
  cluster.sync();     // keep every CTA's shared memory alive until all remote accesses are done
}

// ------------------------------------------------------------------------------------------------------
// bf16-feature-map versions of the cluster kernels (bf16 mode): the feature map is the largest tensor of the step
// (B x P x D) and is re-read by every time step, forward and backward.  As bf16 it is half the bytes, and it is read
// with 16-byte loads: 64 threads x 8 features cover the CTA's 512-wide D chunk of one pixel row, the four 64-thread
// groups of the CTA take every fourth pixel.  att1 / att2 / the relu mask stay fp32 (see decoders_att.py).
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&v);
#pragma unroll
  for (int i = 0; i < 4; ++i) { const float2 t = __bfloat1622float2(h[i]); f[2 * i] = t.x; f[2 * i + 1] = t.y; }
}

constexpr int DCH = 512;         // D chunk per CTA handled by the bf16 kernels (D == CL * DCH)

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(NT)
att_step_fwd_b16_kernel(const float* __restrict__ att1, const float* __restrict__ att2,
                        const __nv_bfloat16* __restrict__ feat, const float* __restrict__ wfull, float bfull,
                        const float* __restrict__ gate_pre, int P, int A, int D, float* __restrict__ alpha,
                        int64_t ld_alpha, float* __restrict__ ctx, int64_t ldc) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ float sm[];
  float* att2s = sm;          // [A]
  float* ws = sm + A;         // [A]
  float* e = ws + A;          // [P]
  float* part = e + ((P + 3) & ~3);     // [4][DCH] partial contexts of the four pixel groups
  __shared__ float red[NT / 32];
  const int r = (int)cluster.block_rank();
  const int b = blockIdx.x / CL, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int a = tid; a < A; a += NT) { att2s[a] = att2[(int64_t)b * A + a]; ws[a] = wfull[a]; }
  __syncthreads();
  const float* a1b = att1 + (int64_t)b * P * A;
  for (int p = r + CL * warp; p < P; p += CL * (NT / 32)) {
    const float4* row = reinterpret_cast<const float4*>(a1b + (int64_t)p * A);
    float s = 0.f;
    for (int a4 = lane; a4 < A / 4; a4 += 32) {
      const float4 v = __ldg(row + a4);
      const float4 w = *reinterpret_cast<const float4*>(ws + 4 * a4);
      const float4 h = *reinterpret_cast<const float4*>(att2s + 4 * a4);
      s = fmaf(w.x, fmaxf(v.x + h.x, 0.f), s); s = fmaf(w.y, fmaxf(v.y + h.y, 0.f), s);
      s = fmaf(w.z, fmaxf(v.z + h.z, 0.f), s); s = fmaf(w.w, fmaxf(v.w + h.w, 0.f), s);
    }
    s = sn::warp_sum(s) + bfull;
    if (lane < CL) cluster.map_shared_rank(e, lane)[p] = s;
  }
  cluster.sync();
  float mx = -INFINITY;
  for (int p = tid; p < P; p += NT) mx = fmaxf(mx, e[p]);
  mx = sn::warp_max(mx);
  if (lane == 0) red[warp] = mx;
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
    if (r == 0) alpha[(int64_t)b * ld_alpha + p] = al;
  }
  __syncthreads();
  // context of this CTA's D chunk: thread (pg, dl) = 8 features of every 4th pixel
  const int pg = tid >> 6, dl = tid & 63, d0 = r * DCH;
  const __nv_bfloat16* fb = feat + (int64_t)b * P * D + d0 + dl * 8;
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int p = pg; p < P; p += 4) {
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(fb + (int64_t)p * D)), f);
    const float a = e[p];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(a, f[i], acc[i]);
  }
  *reinterpret_cast<float4*>(part + pg * DCH + dl * 8) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  *reinterpret_cast<float4*>(part + pg * DCH + dl * 8 + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  __syncthreads();
  {
    const int dd = 2 * tid;
    const float sx = (part[dd] + part[DCH + dd]) + (part[2 * DCH + dd] + part[3 * DCH + dd]);
    const float sy = (part[dd + 1] + part[DCH + dd + 1]) + (part[2 * DCH + dd + 1] + part[3 * DCH + dd + 1]);
    const float2 gp = *reinterpret_cast<const float2*>(gate_pre + (int64_t)b * D + d0 + dd);
    *reinterpret_cast<float2*>(ctx + (int64_t)b * ldc + d0 + dd) = make_float2(sn::sigmoidf_(gp.x) * sx, sn::sigmoidf_(gp.y) * sy);
  }
}

__global__ void __cluster_dims__(CL, 1, 1) __launch_bounds__(NT)
att_step_bwd_b16_kernel(const float* __restrict__ att1, const float* __restrict__ att2,
                        const __nv_bfloat16* __restrict__ feat, const float* __restrict__ wfull,
                        const float* __restrict__ gate_pre, const float* __restrict__ alpha, int64_t ld_alpha,
                        const float* __restrict__ dctx, int64_t ldc, const float* __restrict__ dalpha_extra, int64_t ld_da,
                        int P, int A, int D, float* __restrict__ datt2, float* __restrict__ dgate_pre,
                        float* __restrict__ datt1, float* __restrict__ dwfull) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ float sm[];
  const int P4 = (P + 3) & ~3;
  float* att2s = sm;              // [A]
  float* ws = att2s + A;          // [A]
  float* al = ws + A;             // [P4]
  float* dal = al + P4;           // [P4]       d alpha (sum over the cluster), then d e
  float* slots = dal + P4;        // [CL][P4]   partial d alpha of every CTA of the cluster (DSMEM targets)
  float* dawe = slots + CL * P4;  // [DCH]      d awe of this CTA's chunk
  float* part = dawe + DCH;       // max([4][DCH], [2][8][A]) partial contexts, later the pass C partials
  __shared__ float red[NT / 32];
  const int r = (int)cluster.block_rank();
  const int b = blockIdx.x / CL, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int a = tid; a < A; a += NT) { att2s[a] = att2[(int64_t)b * A + a]; ws[a] = wfull[a]; }
  for (int p = tid; p < P; p += NT) { al[p] = alpha[(int64_t)b * ld_alpha + p]; dal[p] = 0.f; }
  const int d0 = r * DCH;
  float gx, gy;
  {
    const int dd = 2 * tid;
    const float2 gp = *reinterpret_cast<const float2*>(gate_pre + (int64_t)b * D + d0 + dd);
    gx = sn::sigmoidf_(gp.x); gy = sn::sigmoidf_(gp.y);
    dawe[dd] = dctx[(int64_t)b * ldc + d0 + dd] * gx;              // d awe = dctx * gate
    dawe[dd + 1] = dctx[(int64_t)b * ldc + d0 + dd + 1] * gy;
  }
  __syncthreads();
  // pass A: ONE pass over the bf16 feature chunk: partial awe (for the gate gradient) and partial d alpha
  const int pg = tid >> 6, dl = tid & 63;
  const __nv_bfloat16* fb = feat + (int64_t)b * P * D + d0 + dl * 8;
  float da[8], acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 8; ++i) da[i] = dawe[dl * 8 + i];
  for (int p = pg; p < P; p += 4) {
    float f[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(fb + (int64_t)p * D)), f);
    const float a = al[p];
    float t = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) { acc[i] = fmaf(a, f[i], acc[i]); t = fmaf(da[i], f[i], t); }
    t = sn::warp_sum(t);
    if (lane == 0) atomicAdd(dal + p, t);          // two warps per pixel group add into the same pixel
  }
  *reinterpret_cast<float4*>(part + pg * DCH + dl * 8) = make_float4(acc[0], acc[1], acc[2], acc[3]);
  *reinterpret_cast<float4*>(part + pg * DCH + dl * 8 + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
  __syncthreads();
  {
    const int dd = 2 * tid;
    const float sx = (part[dd] + part[DCH + dd]) + (part[2 * DCH + dd] + part[3 * DCH + dd]);
    const float sy = (part[dd + 1] + part[DCH + dd + 1]) + (part[2 * DCH + dd + 1] + part[3 * DCH + dd + 1]);
    const float dcx = dctx[(int64_t)b * ldc + d0 + dd], dcy = dctx[(int64_t)b * ldc + d0 + dd + 1];
    *reinterpret_cast<float2*>(dgate_pre + (int64_t)b * D + d0 + dd) =
        make_float2(dcx * sx * gx * (1.f - gx), dcy * sy * gy * (1.f - gy));
  }
  // exchange the partial d alpha vectors: slot r of every CTA <- this CTA's partial
  for (int i = tid; i < CL * P; i += NT) {
    const int q = i / P, p = i - q * P;
    cluster.map_shared_rank(slots, q)[r * P4 + p] = dal[p];
  }
  cluster.sync();
  for (int p = tid; p < P; p += NT) {
    float s = dalpha_extra ? dalpha_extra[(int64_t)b * ld_da + p] : 0.f;
#pragma unroll
    for (int q = 0; q < CL; ++q) s += slots[q * P4 + p];
    dal[p] = s;
  }
  __syncthreads();
  float dot = 0.f;
  for (int p = tid; p < P; p += NT) dot += al[p] * dal[p];
  dot = block_sum(dot, red);
  for (int p = tid; p < P; p += NT) dal[p] = al[p] * (dal[p] - dot);
  for (int a = tid; a < 2 * (NT / 32) * A; a += NT) part[a] = 0.f;
  __syncthreads();
  // pass C (own pixels): through relu / full_att, 128-bit accesses of att1 / d att1
  float* p_att2 = part + warp * A;
  float* p_w = part + (NT / 32) * A + warp * A;
  const float* a1b = att1 + (int64_t)b * P * A;
  float* d1b = datt1 + (int64_t)b * P * A;
  for (int p = r + CL * warp; p < P; p += CL * (NT / 32)) {
    const float de = dal[p];
    for (int a4 = lane; a4 < A / 4; a4 += 32) {
      const float4 v1 = *reinterpret_cast<const float4*>(a1b + (int64_t)p * A + 4 * a4);
      float4 d1 = *reinterpret_cast<const float4*>(d1b + (int64_t)p * A + 4 * a4);
      const float v[4] = {v1.x + att2s[4 * a4], v1.y + att2s[4 * a4 + 1], v1.z + att2s[4 * a4 + 2], v1.w + att2s[4 * a4 + 3]};
      float dp[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        dp[i] = v[i] > 0.f ? de * ws[4 * a4 + i] : 0.f;
        p_att2[4 * a4 + i] += dp[i];                 // lane-private slots inside this warp's row
        p_w[4 * a4 + i] += de * fmaxf(v[i], 0.f);
      }
      d1.x += dp[0]; d1.y += dp[1]; d1.z += dp[2]; d1.w += dp[3];
      *reinterpret_cast<float4*>(d1b + (int64_t)p * A + 4 * a4) = d1;
    }
  }
  __syncthreads();
  for (int a = tid; a < A; a += NT) {
    float s2 = 0.f, sw = 0.f;
#pragma unroll
    for (int w = 0; w < NT / 32; ++w) { s2 += part[w * A + a]; sw += part[(NT / 32) * A + w * A + a]; }
    atomicAdd(datt2 + (int64_t)b * A + a, s2);      // datt2 rows are zeroed by the host wrapper
    atomicAdd(dwfull + a, sw);
  }
  cluster.sync();     // keep every CTA's shared memory alive until all remote accesses are done
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
  if (D % (2 * CL) == 0 && (ldc % 2) == 0 && (((uintptr_t)feat | (uintptr_t)gate_pre) & 7) == 0) {
    att_step_fwd_cl_kernel<<<(unsigned)(nb * CL), NT, smem, (cudaStream_t)stream>>>(
        att1, att2, feat, wfull, bfull, gate_pre, (int)P, (int)A, (int)D, alpha, ld_alpha, ctx, ldc);
    return sn::check_launch("sn_att_step_fwd(cluster)");
  }
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
  if (D % (2 * CL) == 0 && (((uintptr_t)feat | (uintptr_t)gate_pre) & 7) == 0) {
    size_t smem_cl = (size_t)(2 * A + 2 * P + CL * P + 2 * (NT / 32) * A) * sizeof(float);
    if (smem_cl <= 200 * 1024) {
      if (smem_cl > 48 * 1024)
        SN_CUDA(cudaFuncSetAttribute(att_step_bwd_cl_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      SN_CUDA(cudaMemsetAsync(datt2, 0, sizeof(float) * (size_t)nb * (size_t)A, (cudaStream_t)stream));
      att_step_bwd_cl_kernel<<<(unsigned)(nb * CL), NT, smem_cl, (cudaStream_t)stream>>>(
          att1, att2, feat, wfull, gate_pre, alpha, ld_alpha, dctx, ldc, dalpha_extra, ld_da, (int)P, (int)A, (int)D,
          datt2, dgate_pre, datt1, dwfull, dfeat);
      return sn::check_launch("sn_att_step_bwd(cluster)");
    }
  }
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

// bf16 feature map (bf16 mode).  Requirements: D == 4 * 512 (the cluster kernels' chunking), A % 4 == 0, 16-byte aligned
// att1 / datt1 / feat rows; returns -2 ("not applicable") otherwise so that the caller can use the fp32 kernels.
int32_t sn_att_step_fwd_b16(const float* att1, const float* att2, const void* feat_bf16, const float* wfull, float bfull,
                            const float* gate_pre, int64_t nb, int64_t P, int64_t A, int64_t D, float* alpha,
                            int64_t ld_alpha, float* ctx, int64_t ldc, void* stream) {
  SN_REQUIRE(nb >= 0 && P > 0 && A > 0 && D > 0, "sn_att_step_fwd_b16: bad dims");
  if (D != CL * DCH || A % 4 != 0 || (ldc % 2) != 0 || (((uintptr_t)feat_bf16 | (uintptr_t)att1) & 15) != 0)
    return sn::fail(-2, "sn_att_step_fwd_b16: not applicable (needs D = %d, A %% 4 == 0, aligned rows)", CL * DCH);
  if (nb == 0) return 0;
  const size_t smem = (size_t)(2 * A + ((P + 3) & ~3) + 4 * DCH) * sizeof(float);
  SN_REQUIRE(smem <= 200 * 1024, "sn_att_step_fwd_b16: dims exceed shared memory");
  if (smem > 48 * 1024)
    SN_CUDA(cudaFuncSetAttribute(att_step_fwd_b16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  att_step_fwd_b16_kernel<<<(unsigned)(nb * CL), NT, smem, (cudaStream_t)stream>>>(
      att1, att2, (const __nv_bfloat16*)feat_bf16, wfull, bfull, gate_pre, (int)P, (int)A, (int)D, alpha, ld_alpha, ctx, ldc);
  return sn::check_launch("sn_att_step_fwd_b16");
}

int32_t sn_att_step_bwd_b16(const float* att1, const float* att2, const void* feat_bf16, const float* wfull,
                            const float* gate_pre, const float* alpha, int64_t ld_alpha, const float* dctx, int64_t ldc,
                            const float* dalpha_extra, int64_t ld_da, int64_t nb, int64_t P, int64_t A, int64_t D,
                            float* datt2, float* dgate_pre, float* datt1, float* dwfull, void* stream) {
  SN_REQUIRE(nb >= 0 && P > 0 && A > 0 && D > 0, "sn_att_step_bwd_b16: bad dims");
  if (D != CL * DCH || A % 4 != 0 || (((uintptr_t)feat_bf16 | (uintptr_t)att1 | (uintptr_t)datt1) & 15) != 0)
    return sn::fail(-2, "sn_att_step_bwd_b16: not applicable (needs D = %d, A %% 4 == 0, aligned rows)", CL * DCH);
  if (nb == 0) return 0;
  const size_t P4 = (size_t)((P + 3) & ~3);
  const size_t tail = (size_t)(4 * DCH) > (size_t)(2 * (NT / 32) * A) ? (size_t)(4 * DCH) : (size_t)(2 * (NT / 32) * A);
  const size_t smem = ((size_t)2 * A + 2 * P4 + CL * P4 + DCH + tail) * sizeof(float);
  SN_REQUIRE(smem <= 200 * 1024, "sn_att_step_bwd_b16: dims exceed shared memory");
  if (smem > 48 * 1024)
    SN_CUDA(cudaFuncSetAttribute(att_step_bwd_b16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  SN_CUDA(cudaMemsetAsync(datt2, 0, sizeof(float) * (size_t)nb * (size_t)A, (cudaStream_t)stream));
  att_step_bwd_b16_kernel<<<(unsigned)(nb * CL), NT, smem, (cudaStream_t)stream>>>(
      att1, att2, (const __nv_bfloat16*)feat_bf16, wfull, gate_pre, alpha, ld_alpha, dctx, ldc, dalpha_extra, ld_da,
      (int)P, (int)A, (int)D, datt2, dgate_pre, datt1, dwfull);
  return sn::check_launch("sn_att_step_bwd_b16");
}

}  // extern "C"
