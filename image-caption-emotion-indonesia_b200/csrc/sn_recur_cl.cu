// K3, bf16 mode, CLUSTER form: the recurrence of a slice of 16 samples runs inside ONE thread-block cluster for all T
// steps; clusters never talk to each other (samples are independent), so there is no grid-wide barrier, no flag in
// global memory and no co-residency requirement -- a cluster that does not fit yet simply starts later.
//
// replaces the hot loop stylenet/model.py:180-187 + the W_g(h) / gate math of forward_step (model.py:147-153), resp.
// nn.LSTMCell (nic/model.py:77), like sn_recur_bf16.cu, for hidden sizes H in {128, 256, 512}.
//
// Decomposition (H = 512: clusters of CS = H/32 = 16 CTAs, 512 threads each):
//   * CTA `rank` owns 32 hidden units = 128 rows of W_hh (4 gates x 32 units).  Its bf16 W_hh slice (128 KB) lives in
//     REGISTERS for the whole launch (64 per thread, as mma.sync B fragments): per step the tensor cores read only the
//     16 KB h_{t-1} tile from shared memory, never the weights.
//   * forward, step t:  Z[16 samples, 128 rows] = h_{t-1}[16, H] x Wslice^T as m16n8k16 tiles; warp (octet o, k-quarter q)
//     holds the four gates of 8 units for a quarter of K, so a lane's accumulators are i,f,o,c~ of 2 units x 2 samples.
//     The four k-quarters meet in shared memory; then thread (sample = warp, unit = lane) adds XP_t + b_hh, applies the
//     gate nonlinearities (c_t stays in a register for all T steps) and writes h_t / c_t / gates with full-row
//     coalescing.  h_t (bf16, 64 B per sample per CTA) is broadcast to the 16 CTAs with st.async through distributed
//     shared memory; every store carries its own completion (mbarrier complete_tx), so the consumer waits on ONE local
//     mbarrier per step -- no barrier.cluster round trip in the loop.
//   * backward, step t: the CTA multiplies ITS OWN dZ_{t+1} rows (128 gate rows it produced itself, kept in shared
//     memory) with the same W_hh rows: partial dh_rec[16 samples, all H units] over K = 128.  Warp w's accumulators are
//     exactly the 32 units CTA w owns, so the reduce-scatter is one st.async of fp32 partials per lane and n-tile to
//     CTA w; the owner sums the CS partials in fp32, applies the cell backward (dc carried in a register) and writes
//     dZ_t.  No input exchange at all.
// Buffers are double-buffered by step parity; reuse is ordered by the data dependence itself (a CTA can only be two
// steps ahead after every peer has consumed the older buffer), see the comments at the waits.
#include <cuda_bf16.h>

#include "sn_common.cuh"

namespace {

constexpr int NTH = 512;         // threads per CTA (16 warps)
constexpr int NS = 16;           // samples per cluster (= one m16 tile = one finisher warp per sample)
constexpr int RP = 132;          // fp32 pitch of one sample row of k-quarter partials: 32 units x 4 gates + 4 (conflict-free 128-bit stores)
constexpr int DP = 136;          // bf16 row pitch of the CTA's own dZ rows (128 + 8)

struct CArgs {
  int cell, B, t0, t1;
  const int* bs; const int* off;
  const float* XP; const __nv_bfloat16* Wb; const float* bhh; const float* h_init;
  float* Hall; __nv_bfloat16* Hb; __nv_bfloat16* Hprevb; float* Call; float* gates; float* c_state;
  const float* c_init; const float* dHall; float* dZ; __nv_bfloat16* dZb; float* dh_carry; float* dc_carry;
  int* start_flag;      // optional [3]: {arrivals, generations started, generations consumed by sn_gate_wait}
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t cluster_id_x() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r));
  return r;
}
__device__ __forceinline__ uint32_t mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// acquire at cluster scope: the data were written by st.async of OTHER CTAs (complete_tx releases at cluster scope)
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "CL_WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra CL_WAIT_DONE;\n\t"
      "bra CL_WAIT_LOOP;\n\t"
      "CL_WAIT_DONE:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
// 16-byte store into the shared memory of a CTA of the cluster; completion is counted on that CTA's mbarrier
__device__ __forceinline__ void st_async16(uint32_t dst_cluster, uint32_t v0, uint32_t v1, uint32_t v2, uint32_t v3,
                                           uint32_t bar_cluster) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(dst_cluster), "r"(v0), "r"(v1), "r"(v2), "r"(v3), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void st_async8(uint32_t dst_cluster, uint32_t v0, uint32_t v1, uint32_t bar_cluster) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.b32 [%0], {%1, %2}, [%3];"
               ::"r"(dst_cluster), "r"(v0), "r"(v1), "r"(bar_cluster) : "memory");
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t saddr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(saddr));
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }
// gate nonlinearities on the SFU (ex2 + rcp): absolute error ~1e-7, far inside the bf16 rounding of the exchanged h
__device__ __forceinline__ float sigmoid_fast(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_fast(float x) { return 1.f - __fdividef(2.f, 1.f + __expf(2.f * x)); }

// Stage 64 rows x H bf16 of W_hh (row r of the stage = global row `grow(r)`) into shared memory with 16-byte cp.async
// (fully coalesced), row pitch H+8.  The register fragments are then read with ldmatrix: the prologue is bound by the
// L2 -> SM copy of the 128 KB slice instead of by thousands of 4-byte loads.
template <int H, typename RowFn>
__device__ __forceinline__ void stage_rows(uint32_t stage, const __nv_bfloat16* Wb, RowFn grow) {
  constexpr int CH = H / 8;                                  // 16-byte chunks per row
  for (int i = threadIdx.x; i < 64 * CH; i += NTH) {
    const int r = i / CH, c = i - r * CH;
    cp_async16(stage + (uint32_t)(r * (H + 8) + c * 8) * 2, Wb + (int64_t)grow(r) * H + c * 8);
  }
  cp_async_wait_all();
  __syncthreads();
}

template <int H>
struct FwdSmem {
  static constexpr int HP = H + 8;                                   // bf16 row pitch of the h tile (conflict-free ldmatrix)
  static constexpr size_t hbuf = 0;                                  // [2][NS][HP] bf16
  static constexpr size_t red = hbuf + (size_t)2 * NS * HP * 2;      // [2][4 kq][NS][RP]: (unit, gate) gate-fastest, fp32
  static constexpr size_t hst = red + (size_t)2 * 4 * NS * RP * 4;   // [NS][32] bf16
  static constexpr size_t bars = hst + (size_t)NS * 32 * 2;          // [2] mbarrier
  static constexpr size_t tab = bars + 16;                           // per step: {live samples, first packed row}
  static constexpr size_t total = tab;                               // + 8 bytes per step (added at launch)
};

// ================================================================================================
// forward
// ================================================================================================
template <int H>
__global__ void __launch_bounds__(NTH, 1) recur_fwd_cl_kernel(CArgs a) {
  constexpr int CS = H / 32, KT = H / 64, HP = FwdSmem<H>::HP, KQ = H / 4;
  extern __shared__ __align__(128) unsigned char smem[];
  __nv_bfloat16* hbuf = reinterpret_cast<__nv_bfloat16*>(smem + FwdSmem<H>::hbuf);
  float* red = reinterpret_cast<float*>(smem + FwdSmem<H>::red);
  __nv_bfloat16* hst = reinterpret_cast<__nv_bfloat16*>(smem + FwdSmem<H>::hst);
  const uint32_t bar0 = smem_u32(smem + FwdSmem<H>::bars);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = (int)cluster_ctarank();
  const int s0 = (int)cluster_id_x() * NS, u0 = rank * 32;
  const int oct = warp & 3, kq = warp >> 2;                    // MMA role: units u0 + oct*8 .. +7, k in [kq*KQ, +KQ)
  const int pos_o = a.cell == SN_CELL_LSTM ? 3 : 2, pos_c = a.cell == SN_CELL_LSTM ? 2 : 3;

  // W_hh slice -> registers, as B fragments: b0 = W[row n][k0 + 2q .. +1], b1 = W[row n][k0 + 8 + 2q .. +1].
  // Two rounds of 64 rows (gates 2r, 2r+1 x 32 units) staged through the (not yet used) h / partial buffers.
  uint32_t Wr[4][KT][2];
  {
    const uint32_t stage = smem_u32(smem);
#pragma unroll
    for (int rd = 0; rd < 2; ++rd) {
      stage_rows<H>(stage, a.Wb, [&](int r) { return (2 * rd + (r >> 5)) * H + u0 + (r & 31); });
#pragma unroll
      for (int gg = 0; gg < 2; ++gg) {
        const uint32_t base = stage + (uint32_t)((gg * 32 + oct * 8 + (lane & 7)) * HP + kq * KQ + (lane >> 3) * 8) * 2;
#pragma unroll
        for (int kt = 0; kt < KT; kt += 2) {
          uint32_t f[4];
          ldmatrix_x4(f, base + kt * 32);
          Wr[2 * rd + gg][kt][0] = f[0]; Wr[2 * rd + gg][kt][1] = f[1];
          Wr[2 * rd + gg][kt + 1][0] = f[2]; Wr[2 * rd + gg][kt + 1][1] = f[3];
        }
      }
      __syncthreads();
    }
  }

  // finisher role: sample fs of the cluster, unit fu of the CTA
  const int fs = warp, fu = lane, gs = s0 + fs, ug = u0 + fu;
  const bool live = gs < a.B;
  float c_reg = live ? a.c_state[(int64_t)gs * H + ug] : 0.f;
  // h_init holds rows only for the samples alive at step t0 (the caller passes the Hall rows of step t0-1)
  const int nv_first = clampi(a.bs[a.t0] - s0, 0, NS);
  __nv_bfloat16 hprev = __float2bfloat16((fs < nv_first && a.h_init) ? a.h_init[(int64_t)gs * H + ug] : 0.f);
  float bh[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) bh[g] = a.bhh ? __ldg(a.bhh + g * H + ug) : 0.f;

  // per-step tables in shared memory: a global load of batch_sizes[t] / offsets[t] inside the loop would sit on the
  // critical path of every step (the acquire at cluster scope invalidates L1, so each one is an L2 round trip)
  int2* tab = reinterpret_cast<int2*>(smem + FwdSmem<H>::tab);
  for (int i = tid; i < a.t1 - a.t0; i += NTH)
    tab[i] = make_int2(clampi(a.bs[a.t0 + i] - s0, 0, NS), a.off[a.t0 + i] + s0);
  // h before step t0 into the buffer step t0 reads (every CTA fills its own copy); the other buffer starts as zeros
  {
    const int p0 = a.t0 & 1;
    for (int i = tid; i < NS * (H / 2); i += NTH) {
      const int r = i / (H / 2), c = (i - r * (H / 2)) * 2;
      float2 v = make_float2(0.f, 0.f);
      if (a.h_init && r < nv_first) v = *reinterpret_cast<const float2*>(a.h_init + (int64_t)(s0 + r) * H + c);
      *reinterpret_cast<__nv_bfloat162*>(hbuf + ((size_t)p0 * NS + r) * HP + c) = __floats2bfloat162_rn(v.x, v.y);
      *reinterpret_cast<__nv_bfloat162*>(hbuf + ((size_t)(p0 ^ 1) * NS + r) * HP + c) = __floats2bfloat162_rn(0.f, 0.f);
    }
  }
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync();          // barriers initialised and buffers filled in every CTA before anyone sends

  uint32_t ph0 = 0, ph1 = 0;       // phase parity of bar[0], bar[1]
  for (int t = a.t0; t < a.t1; ++t) {
    const int2 tb = tab[t - a.t0];
    const int nv = tb.x;
    if (nv == 0) break;                                        // batch sizes never grow: this cluster is done
    const int nvn = (t + 1 < a.t1) ? tab[t + 1 - a.t0].x : 0;
    const int p = t & 1, pn = p ^ 1;                           // buffer p holds h_{t-1}; h_t goes to buffer pn
    // arm the barrier of h_t.  Its previous phase (h_{t-2}) was consumed by every thread before the __syncthreads of
    // step t-1, and it cannot complete before this arrival however early peers send.
    if (tid == 0 && nvn > 0) mbar_expect_tx(bar0 + 8 * pn, (uint32_t)(CS * nvn * 64));
    const bool valid = fs < nv;
    const int64_t row = (int64_t)tb.y + fs;
    float xp[4] = {0.f, 0.f, 0.f, 0.f};
    if (valid) {
      const float* x = a.XP + row * 4 * H + ug;
#pragma unroll
      for (int g = 0; g < 4; ++g) xp[g] = __ldg(x + g * H);
    }
    if (t > a.t0) {
      // all CS x nv rows of h_{t-1} have landed in buffer p (sent at the end of step t-1 by every CTA)
      if (p) { mbar_wait_cluster(bar0 + 8, ph1); ph1 ^= 1; } else { mbar_wait_cluster(bar0, ph0); ph0 ^= 1; }
    }
    float acc[4][4];
#pragma unroll
    for (int g = 0; g < 4; ++g) acc[g][0] = acc[g][1] = acc[g][2] = acc[g][3] = 0.f;
    {
      const uint32_t abase = smem_u32(hbuf + ((size_t)p * NS + (lane & 15)) * HP + kq * KQ + (lane >> 4) * 8);
#pragma unroll
      for (int kt = 0; kt < KT; ++kt) {
        uint32_t af[4];
        ldmatrix_x4(af, abase + kt * 32);
#pragma unroll
        for (int g = 0; g < 4; ++g) mma_bf16(acc[g], af, Wr[g][kt][0], Wr[g][kt][1]);
      }
    }
    {
      // partial of k-quarter kq: rows = samples lane/4 (+8), columns = units oct*8 + 2*(lane%4) (+1); the four gates of a
      // (sample, unit) are one float4
      float* rp = red + (((size_t)p * 4 + kq) * NS + (lane >> 2)) * RP + (oct * 8 + (lane & 3) * 2) * 4;
      *reinterpret_cast<float4*>(rp) = make_float4(acc[0][0], acc[1][0], acc[2][0], acc[3][0]);
      *reinterpret_cast<float4*>(rp + 4) = make_float4(acc[0][1], acc[1][1], acc[2][1], acc[3][1]);
      *reinterpret_cast<float4*>(rp + 8 * RP) = make_float4(acc[0][2], acc[1][2], acc[2][2], acc[3][2]);
      *reinterpret_cast<float4*>(rp + 8 * RP + 4) = make_float4(acc[0][3], acc[1][3], acc[2][3], acc[3][3]);
    }
    // `red` and the h buffers are double-buffered by step parity: a warp can
    // write red[p] / see peers overwrite hbuf[p] again only at step t+2, i.e. after the barrier of step t+1, which
    // every warp reaches after it finished reading them here.
    __syncthreads();
    __nv_bfloat16 hb = __float2bfloat16(0.f);
    if (valid) {
      float z[4];
      {
        const float* r = red + ((size_t)p * 4 * NS + fs) * RP + fu * 4;
        const float4 q0 = *reinterpret_cast<const float4*>(r), q1 = *reinterpret_cast<const float4*>(r + (size_t)NS * RP);
        const float4 q2 = *reinterpret_cast<const float4*>(r + (size_t)2 * NS * RP);
        const float4 q3 = *reinterpret_cast<const float4*>(r + (size_t)3 * NS * RP);
        z[0] = ((q0.x + q1.x) + (q2.x + q3.x)) + xp[0] + bh[0];
        z[1] = ((q0.y + q1.y) + (q2.y + q3.y)) + xp[1] + bh[1];
        z[2] = ((q0.z + q1.z) + (q2.z + q3.z)) + xp[2] + bh[2];
        z[3] = ((q0.w + q1.w) + (q2.w + q3.w)) + xp[3] + bh[3];
      }
      const float zo = a.cell == SN_CELL_LSTM ? z[3] : z[2], zc = a.cell == SN_CELL_LSTM ? z[2] : z[3];
      const float gi = sigmoid_fast(z[0]), gf = sigmoid_fast(z[1]), go = sigmoid_fast(zo), gc = tanh_fast(zc);
      const float c = gf * c_reg + gi * gc;
      const float h = a.cell == SN_CELL_LSTM ? go * tanh_fast(c) : go * c;
      c_reg = c;
      hb = __float2bfloat16(h);
      a.Hb[row * H + ug] = hb;
      if (a.Hprevb) a.Hprevb[row * H + ug] = hprev;
      if (a.Hall) a.Hall[row * H + ug] = h;
      if (a.Call) a.Call[row * H + ug] = c;
      if (a.gates) {
        float* gp = a.gates + row * 4 * H + ug;
        gp[0] = gi; gp[H] = gf; gp[pos_o * H] = go; gp[pos_c * H] = gc;
      }
      hprev = hb;
    }
    if (nvn > 0) {
      // broadcast this CTA's h_t block (NS samples x 32 units, bf16 = 1 KB) to all CS CTAs (this one included).
      // Every st.async INSTRUCTION targets a single CTA (warp w serves CTA w when CS = 16): a warp-wide store whose
      // lanes address 16 different CTAs is split into 16 transactions and made the forward send-bound (profiles/).
      hst[fs * 32 + fu] = hb;
      __syncthreads();
      for (int j = lane; j < 4 * CS; j += 32) {
        const int i = warp * (4 * CS) + j;                     // (destination, 16-byte chunk) pair
        const int dest = i >> 6, ch = i & 63, srow = ch >> 2, part = ch & 3;
        if (srow < nvn) {
          const uint4 v = *reinterpret_cast<const uint4*>(hst + srow * 32 + part * 8);
          const uint32_t dst_local = smem_u32(hbuf + ((size_t)pn * NS + srow) * HP + u0 + part * 8);
          st_async16(mapa(dst_local, dest), v.x, v.y, v.z, v.w, mapa(bar0 + 8 * pn, dest));
        }
      }
    }
  }
  if (live) a.c_state[(int64_t)gs * H + ug] = c_reg;
  cluster_sync();          // nobody leaves while a peer could still address its shared memory
}

template <int H>
struct BwdSmem {
  static constexpr int CS = H / 32;
  static constexpr size_t recv = 0;                                        // [2][CS][NS][32] bf16 partials
  static constexpr size_t dzs = recv + (size_t)2 * CS * NS * 32 * 2;       // [2][NS][DP] bf16
  static constexpr size_t bars = dzs + (size_t)2 * NS * DP * 2;
  static constexpr size_t tab = bars + 16;                                 // per step: {live samples, first packed row}
  static constexpr size_t stage = (size_t)64 * (H + 8) * 2;                // prologue staging of 64 W_hh rows (aliases recv/dzs)
  static constexpr size_t total = tab > stage ? tab : stage;               // + 8 bytes per step (added at launch)
};

// ================================================================================================
// backward
// ================================================================================================
template <int H>
__global__ void __launch_bounds__(NTH, 1) recur_bwd_cl_kernel(CArgs a) {
  constexpr int CS = H / 32, NTB = H / 128;                   // n-tiles (8 units) per warp: 16 warps cover all H units
  extern __shared__ __align__(128) unsigned char smem[];
  __nv_bfloat16* recv = reinterpret_cast<__nv_bfloat16*>(smem + BwdSmem<H>::recv);
  __nv_bfloat16* dzs = reinterpret_cast<__nv_bfloat16*>(smem + BwdSmem<H>::dzs);
  const uint32_t bar0 = smem_u32(smem + BwdSmem<H>::bars);

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = (int)cluster_ctarank();
  const int s0 = (int)cluster_id_x() * NS, u0 = rank * 32;
  const int K4 = 4 * H;
  // "every CTA of this launch is resident": lets work queued behind sn_gate_wait start on the SMs this kernel leaves free
  // only AFTER its clusters have been placed (a cluster needs 16 completely free SMs of one GPC)
  if (tid == 0 && a.start_flag) {
    if (atomicAdd(a.start_flag, 1) == (int)gridDim.x - 1) {
      a.start_flag[0] = 0;
      __threadfence();
      atomicAdd(a.start_flag + 1, 1);
    }
  }

  // W_hh rows of this CTA (k = gate*32 + unit -> row gate*H + u0 + unit), all H columns, as B fragments:
  // b0 = {W[k0 + 2q][n], W[k0 + 2q + 1][n]}, b1 = same at k0 + 8;  n = unit (warp*NTB + nt)*8 + lane/4.
  // Staged 64 rows at a time ([k][n] row-major) and read with ldmatrix.trans.
  uint32_t Wr[8][NTB][2];
  {
    const uint32_t stage = smem_u32(smem);
    constexpr int HP = H + 8;
#pragma unroll
    for (int rd = 0; rd < 2; ++rd) {
      stage_rows<H>(stage, a.Wb, [&](int r) { return (2 * rd + (r >> 5)) * H + u0 + (r & 31); });
#pragma unroll
      for (int k2 = 0; k2 < 4; k2 += 2) {                       // k-tiles 4*rd + k2, +1 : stage rows k2*16 .. k2*16 + 31
#pragma unroll
        for (int nt = 0; nt < NTB; ++nt) {
          // matrices: (k-tile k2, rows +0..7), (k2, +8..15), (k2+1, +0..7), (k2+1, +8..15); 8 columns n
          uint32_t f[4];
          ldmatrix_x4_trans(f, stage + (uint32_t)((k2 * 16 + (lane >> 3) * 8 + (lane & 7)) * HP + (warp * NTB + nt) * 8) * 2);
          Wr[4 * rd + k2][nt][0] = f[0]; Wr[4 * rd + k2][nt][1] = f[1];
          Wr[4 * rd + k2 + 1][nt][0] = f[2]; Wr[4 * rd + k2 + 1][nt][1] = f[3];
        }
      }
      __syncthreads();
    }
  }

  const int fs = warp, fu = lane, gs = s0 + fs, ug = u0 + fu;
  const bool live = gs < a.B;
  const int64_t sidx = (int64_t)gs * H + ug;
  float dc_reg = live ? a.dc_carry[sidx] : 0.f;
  // where unit fu sits inside a 64-byte (32 x bf16) row of partials, see the sender below
  const int fpos = (((fu >> 4) * 2 + ((fu >> 2) & 1)) * 8) + ((fu >> 3) & 1) * 4 + (fu & 3);

  int2* tab = reinterpret_cast<int2*>(smem + BwdSmem<H>::total);
  for (int i = tid; i < a.t1 - a.t0; i += NTH)
    tab[i] = make_int2(clampi(a.bs[a.t0 + i] - s0, 0, NS), a.off[a.t0 + i] + s0);
  const int row_before = a.t0 > 0 ? a.off[a.t0 - 1] + s0 : 0;       // packed row of sample s0 at step t0-1 (c_{t0-1})
  for (int i = tid; i < 2 * NS * DP / 2; i += NTH) reinterpret_cast<uint32_t*>(dzs)[i] = 0u;
  if (tid == 0) {
    mbar_init(bar0, 1);
    mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster_sync();

  uint32_t ph0 = 0, ph1 = 0;
  // iteration t finishes step t: dh_rec comes from the partials of dZ_{t+1} (sent at the end of iteration t+1); its own
  // dZ_t partials are sent at the end for iteration t-1 (or for the tail, which returns dL/dh_{t0-1})
  for (int t = a.t1 - 1; t >= a.t0; --t) {
    const int2 tb = tab[t - a.t0];
    const int nv = tb.x;
    if (nv == 0) continue;                                     // this cluster's samples all ended before step t
    const int nrec = (t + 1 < a.t1) ? tab[t + 1 - a.t0].x : 0;
    const int p = t & 1, pr = p ^ 1;                           // partials of dZ_t travel in buffer p, those of dZ_{t+1} in pr
    if (tid == 0) mbar_expect_tx(bar0 + 8 * p, (uint32_t)(CS * nv * 64));
    const bool valid = fs < nv;
    const int64_t row = (int64_t)tb.y + fs;
    float g4[4] = {0.f, 0.f, 0.f, 0.f}, cc = 0.f, cprev = 0.f, dhl = 0.f;
    if (valid) {
      const float* gp = a.gates + row * K4 + ug;
#pragma unroll
      for (int g = 0; g < 4; ++g) g4[g] = __ldg(gp + g * H);
      cc = __ldg(a.Call + row * H + ug);
      if (t > 0) cprev = __ldg(a.Call + ((int64_t)(t > a.t0 ? tab[t - 1 - a.t0].y : row_before) + fs) * H + ug);
      else cprev = a.c_init ? __ldg(a.c_init + sidx) : 0.f;
      dhl = __ldg(a.dHall + row * H + ug);
    }
    float dh_rec = 0.f;
    if (nrec > 0) {
      if (pr) { mbar_wait_cluster(bar0 + 8, ph1); ph1 ^= 1; } else { mbar_wait_cluster(bar0, ph0); ph0 ^= 1; }
      if (fs < nrec) {
        const __nv_bfloat16* r = recv + ((size_t)pr * CS * NS + fs) * 32 + fpos;
        float s = 0.f;
#pragma unroll
        for (int src = 0; src < CS; ++src) s += __bfloat162float(r[(size_t)src * NS * 32]);
        dh_rec = s;
      }
    }
    if (valid) {
      if (t == a.t1 - 1) dh_rec = a.dh_carry[sidx];
      const float gi = g4[0], gf = g4[1];
      const float go = a.cell == SN_CELL_LSTM ? g4[3] : g4[2], gc = a.cell == SN_CELL_LSTM ? g4[2] : g4[3];
      const float dh = dhl + dh_rec;
      float d_o, dc;
      if (a.cell == SN_CELL_LSTM) {
        const float tc = tanhf(cc);
        d_o = dh * tc;
        dc = dc_reg + dh * go * (1.f - tc * tc);
      } else {
        d_o = dh * cc;
        dc = dc_reg + dh * go;
      }
      const float di = dc * gc, df = dc * cprev, dg = dc * gi;
      dc_reg = dc * gf;
      float z[4];
      const float z_o = d_o * go * (1.f - go), z_c = dg * (1.f - gc * gc);
      z[0] = di * gi * (1.f - gi); z[1] = df * gf * (1.f - gf);
      z[2] = a.cell == SN_CELL_LSTM ? z_c : z_o; z[3] = a.cell == SN_CELL_LSTM ? z_o : z_c;
      __nv_bfloat16* dzb = a.dZb + row * K4 + ug;
      __nv_bfloat16* own = dzs + ((size_t)p * NS + fs) * DP + fu;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const __nv_bfloat16 zb = __float2bfloat16(z[g]);
        dzb[g * H] = zb;
        own[g * 32] = zb;
        if (a.dZ) a.dZ[row * K4 + g * H + ug] = z[g];
      }
    }
    // own dZ_t rows complete.  dzs / recv are double-buffered by step parity: they are written again at iteration t-2,
    // i.e. after the barrier of iteration t-1, which every warp passes only after its MMAs / sums of this iteration.
    __syncthreads();
    {
      float acc[NTB][4];
#pragma unroll
      for (int nt = 0; nt < NTB; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
      const uint32_t abase = smem_u32(dzs + ((size_t)p * NS + (lane & 15)) * DP + (lane >> 4) * 8);
#pragma unroll
      for (int kt = 0; kt < 8; ++kt) {
        uint32_t af[4];
        ldmatrix_x4(af, abase + kt * 32);
#pragma unroll
        for (int nt = 0; nt < NTB; ++nt) mma_bf16(acc[nt], af, Wr[kt][nt][0], Wr[kt][nt][1]);
      }
      // lane pairs swap halves so that each lane owns 4 consecutive units of ONE sample (even lanes keep sample lane/4,
      // odd lanes sample lane/4 + 8); the partials travel as bf16 (the owner sums the CS of them in fp32 -- the same
      // rounding the exchanged dZ itself already has), two n-tiles per 16-byte st.async.  Position of local unit
      // u = lnt*8 + r inside the 64-byte row: chunk (lnt/2)*2 + (r/4), then (lnt%2)*4 + r%4  (see fpos above).
      const bool odd = lane & 1;
      const int srow = (lane >> 2) + (odd ? 8 : 0);
      const int ub = ((lane & 3) >> 1) * 4;
      uint32_t pk[NTB][2];
#pragma unroll
      for (int nt = 0; nt < NTB; ++nt) {
        const float x0 = odd ? acc[nt][0] : acc[nt][2], x1 = odd ? acc[nt][1] : acc[nt][3];
        const float y0 = __shfl_xor_sync(0xffffffffu, x0, 1), y1 = __shfl_xor_sync(0xffffffffu, x1, 1);
        const float v0 = odd ? y0 : acc[nt][0], v1 = odd ? y1 : acc[nt][1];
        const float v2 = odd ? acc[nt][2] : y0, v3 = odd ? acc[nt][3] : y1;
        __nv_bfloat162 lo = __floats2bfloat162_rn(v0, v1), hi = __floats2bfloat162_rn(v2, v3);
        pk[nt][0] = *reinterpret_cast<uint32_t*>(&lo);
        pk[nt][1] = *reinterpret_cast<uint32_t*>(&hi);
      }
      if (srow < nv) {
        const __nv_bfloat16* rowp = recv + (((size_t)p * CS + rank) * NS + srow) * 32;
        if (NTB >= 2) {
#pragma unroll
          for (int nt = 0; nt + 1 < NTB; nt += 2) {
            const int gnt = warp * NTB + nt;                    // global n-tile (8 units); even
            const int dest = gnt >> 2, lnt = gnt & 3;
            const uint32_t dst_local = smem_u32(rowp + ((lnt >> 1) * 2 + (ub >> 2)) * 8);
            st_async16(mapa(dst_local, dest), pk[nt][0], pk[nt][1], pk[nt + 1][0], pk[nt + 1][1], mapa(bar0 + 8 * p, dest));
          }
        } else {
          const int gnt = warp, dest = gnt >> 2, lnt = gnt & 3;
          const uint32_t dst_local = smem_u32(rowp + ((lnt >> 1) * 2 + (ub >> 2)) * 8 + (lnt & 1) * 4);
          st_async8(mapa(dst_local, dest), pk[0][0], pk[0][1], mapa(bar0 + 8 * p, dest));
        }
      }
    }
  }
  // tail: dL/dh_{t0-1} = dZ_{t0} W_hh for the samples alive at t0, zero for the rest; dL/dc_{t0-1}
  {
    const int nv0 = tab[0].x;
    float dh_rec = 0.f;
    if (nv0 > 0) {
      const int p = a.t0 & 1;
      if (p) mbar_wait_cluster(bar0 + 8, ph1); else mbar_wait_cluster(bar0, ph0);
      if (fs < nv0) {
        const __nv_bfloat16* r = recv + ((size_t)p * CS * NS + fs) * 32 + fpos;
        float s = 0.f;
#pragma unroll
        for (int src = 0; src < CS; ++src) s += __bfloat162float(r[(size_t)src * NS * 32]);
        dh_rec = s;
      }
    }
    if (live) {
      a.dh_carry[sidx] = dh_rec;
      a.dc_carry[sidx] = dc_reg;
    }
  }
  cluster_sync();
}

// flag[1] = launches of the reverse recurrence whose CTAs all became resident (bumped by the kernel itself),
// flag[2] = launches this gate has accounted for.  The gate is queued long before the recurrence can start (it only
// depends on the vocabulary backward), so a bump that is already there when the gate starts belongs to an EARLIER step
// whose gate timed out: drop it, then wait for this step's bump.  A timeout undoes the accounting, so one missed
// hand-shake (e.g. the first step, while kernels are still being loaded) cannot shift every later step by one.
__global__ void gate_wait_kernel(int* flag, long long timeout_cycles) {
  const long long t0 = clock64();
  const int stale = sn::ld_acquire(flag + 1);
  if (stale > flag[2]) flag[2] = stale;
  const int target = flag[2] + 1;
  flag[2] = target;
  while (sn::ld_acquire(flag + 1) < target) {
    if (clock64() - t0 > timeout_cycles) {       // never hang: the ordering is an optimisation, not a dependency
      flag[2] = target - 1;
      return;
    }
    __nanosleep(200);
  }
}

template <typename Kern>
int32_t launch_cl(Kern kernel, int CS, size_t smem, const CArgs& a, cudaStream_t stream, const char* what, int* max_clusters) {
  cudaLaunchConfig_t cfg = {};
  const int nclusters = (a.B + NS - 1) / NS;
  cfg.gridDim = dim3((unsigned)(nclusters * CS));
  cfg.blockDim = dim3(NTH);
  smem += (size_t)(a.t1 - a.t0 + 1) * 8;                 // per-step tables behind the fixed layout
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  SN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (CS > 8) SN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  if (max_clusters) {
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kernel, &cfg);
    if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
    *max_clusters = n;
    return 0;
  }
  CArgs args = a;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kernel, args);
  if (e != cudaSuccess) return sn::fail((int32_t)e, "%s: cluster launch failed: %s", what, cudaGetErrorString(e));
  return 0;
}

int32_t dispatch(bool bwd, int64_t H, const CArgs& a, cudaStream_t st, int* max_clusters) {
  switch (H) {
    case 512:
      return bwd ? launch_cl(recur_bwd_cl_kernel<512>, 16, BwdSmem<512>::total, a, st, "sn_recur_bwd_cl", max_clusters)
                 : launch_cl(recur_fwd_cl_kernel<512>, 16, FwdSmem<512>::total, a, st, "sn_recur_fwd_cl", max_clusters);
    case 256:
      return bwd ? launch_cl(recur_bwd_cl_kernel<256>, 8, BwdSmem<256>::total, a, st, "sn_recur_bwd_cl", max_clusters)
                 : launch_cl(recur_fwd_cl_kernel<256>, 8, FwdSmem<256>::total, a, st, "sn_recur_fwd_cl", max_clusters);
    case 128:
      return bwd ? launch_cl(recur_bwd_cl_kernel<128>, 4, BwdSmem<128>::total, a, st, "sn_recur_bwd_cl", max_clusters)
                 : launch_cl(recur_fwd_cl_kernel<128>, 4, FwdSmem<128>::total, a, st, "sn_recur_fwd_cl", max_clusters);
    default:
      return sn::fail(-1, "sn_recur_*_cl: hidden size %lld not supported by the cluster form (128, 256, 512)", (long long)H);
  }
}

}  // namespace

extern "C" {

int32_t sn_recur_cl_max_clusters(int64_t H) {
  if (H != 128 && H != 256 && H != 512) return 0;
  static int cache[3] = {-1, -1, -1};
  const int slot = H == 512 ? 0 : (H == 256 ? 1 : 2);
  if (cache[slot] >= 0) return cache[slot];
  CArgs a = {};
  a.B = NS;
  int nf = 0, nb = 0;
  if (dispatch(false, H, a, nullptr, &nf) != 0 || dispatch(true, H, a, nullptr, &nb) != 0) return 0;
  cache[slot] = nf < nb ? nf : nb;
  return cache[slot];
}

int32_t sn_recur_fwd_cl(int32_t cell, int64_t H, int64_t B, const int32_t* batch_sizes, const int32_t* offsets,
                        int32_t t0, int32_t t1, const float* XP, const void* Whh_bf16, const float* bhh,
                        const float* h_init, float* Hall, void* Hb, void* Hprevb, float* Call, float* gates,
                        float* c_state, void* stream) {
  SN_REQUIRE(cell == SN_CELL_FACTORED || cell == SN_CELL_LSTM, "sn_recur_fwd_cl: bad cell %d", cell);
  SN_REQUIRE(t0 >= 0 && t1 >= t0 && B > 0, "sn_recur_fwd_cl: bad step range");
  SN_REQUIRE(XP && Whh_bf16 && Hb && c_state && batch_sizes && offsets, "sn_recur_fwd_cl: null argument");
  if (t1 == t0) return 0;
  CArgs a = {};
  a.cell = cell; a.B = (int)B; a.t0 = t0; a.t1 = t1; a.bs = batch_sizes; a.off = offsets;
  a.XP = XP; a.Wb = (const __nv_bfloat16*)Whh_bf16; a.bhh = bhh; a.h_init = h_init; a.Hall = Hall;
  a.Hb = (__nv_bfloat16*)Hb; a.Hprevb = (__nv_bfloat16*)Hprevb; a.Call = Call; a.gates = gates; a.c_state = c_state;
  return dispatch(false, H, a, (cudaStream_t)stream, nullptr);
}

int32_t sn_recur_bwd_cl(int32_t cell, int64_t H, int64_t B, const int32_t* batch_sizes, const int32_t* offsets,
                        int32_t t0, int32_t t1, const void* Whh_bf16, const float* c_init, const float* Call,
                        const float* gates, const float* dHall, float* dZ, void* dZb, float* dh_carry,
                        float* dc_carry, int32_t* start_flag, void* stream) {
  SN_REQUIRE(cell == SN_CELL_FACTORED || cell == SN_CELL_LSTM, "sn_recur_bwd_cl: bad cell %d", cell);
  SN_REQUIRE(t0 >= 0 && t1 >= t0 && B > 0, "sn_recur_bwd_cl: bad step range");
  SN_REQUIRE(Whh_bf16 && Call && gates && dHall && dZb && dh_carry && dc_carry && batch_sizes && offsets,
             "sn_recur_bwd_cl: null argument");
  if (t1 == t0) return 0;
  CArgs a = {};
  a.cell = cell; a.B = (int)B; a.t0 = t0; a.t1 = t1; a.bs = batch_sizes; a.off = offsets;
  a.Wb = (const __nv_bfloat16*)Whh_bf16; a.c_init = c_init; a.Call = const_cast<float*>(Call);
  a.gates = const_cast<float*>(gates); a.dHall = dHall; a.dZ = dZ; a.dZb = (__nv_bfloat16*)dZb;
  a.dh_carry = dh_carry; a.dc_carry = dc_carry; a.start_flag = start_flag;
  return dispatch(true, H, a, (cudaStream_t)stream, nullptr);
}

int32_t sn_gate_wait(int32_t* flag3, int64_t timeout_us, void* stream) {
  SN_REQUIRE(flag3 && timeout_us > 0, "sn_gate_wait: bad argument");
  gate_wait_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(flag3, (long long)timeout_us * 2000);
  return sn::check_launch("sn_gate_wait");
}

}  // extern "C"
