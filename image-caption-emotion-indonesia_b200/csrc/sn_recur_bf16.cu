// K3, bf16 mode: the persistent recurrence with the per-step h_{t-1} W_hh^T (forward) / dZ_{t+1} W_hh
// (backward) contraction on the tensor cores.  Operands bf16 (W_hh slice resident in shared memory for the
// whole launch, h / dZ exchanged between SMs as bf16 through L2), accumulation, gate pre-activations, c_t,
// gate math and every carried gradient in fp32 (SURVEY.md Appendix A "bf16 mode definition").
//
// Same decomposition as sn_recur.cu: CTA (ub, bb) owns a block of hidden units for the samples of batch
// block bb for all T steps; one release/acquire counter per (batch block, step).  Per step the CTA computes
//   fwd:  D[32 = 4 gates x 8 units, n samples]  = Ws[32, H]   * h_{t-1}[n, H]^T
//   bwd:  D[16 units,               n samples]  = WsT[16, 4H] * dZ_{t+1}[n, 4H]^T
// as m16n8k16 tiles (one 8-sample tile per warp), fragments via ldmatrix from padded shared memory; the
// accumulator fragment of a lane holds all four gates of ONE unit for two samples, so the gate
// nonlinearities / cell update run in registers with no exchange (lane-local gate fusion).
// At B<=96 a step moves ~100 KB per CTA and does <1 us of math: the kernel is bound by the per-step
// exchange latency, not by the tensor pipe -- which is why the simple warp-level MMA is used here and
// tcgen05/TMEM (whose single-thread issue + TMEM round trip adds latency) is kept for the big GEMMs.
#include <cuda_bf16.h>

#include "sn_common.cuh"

namespace {

constexpr int NW = 8, NT = NW * 32;
constexpr int G = 64;            // samples per group: 8 tiles of 8, one per warp
constexpr int PADB = 8;          // bf16 row padding (16 B): conflict-free ldmatrix

struct RArgs {
  int cell, H, B, t0, t1, T;
  const int* bs; const int* off;
  const float* XP; const __nv_bfloat16* Wb; const float* bhh;
  const float* h_init; const float* c_init;
  float* Hall; __nv_bfloat16* Hb; __nv_bfloat16* Hprevb; float* Call; float* gates; float* c_state;
  const float* dHall; float* dZ; __nv_bfloat16* dZb; float* dh_carry; float* dc_carry;
  int* flags; int n_ub, nbb, BB;
  int nbuf;        // forward: 2 = double-buffered staging across sample groups, 1 = single buffer (large H)
};

__device__ __forceinline__ void wait_flag(const int* flag, int target) {
  if (threadIdx.x == 0) { while (sn::ld_acquire(flag) < target) { } }
  __syncthreads();
}
__device__ __forceinline__ void signal_flag(int* flag) {
  __syncthreads();
  if (threadIdx.x == 0) sn::red_release_add(flag, 1);   // release.gpu: cumulative over the CTA's writes (bar.sync above)
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p) {
  unsigned a = (unsigned)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// stage rows [0,nrows) x bf16 cols [k0, k0+KC) (KC multiple of 8) into smem rows of pitch ldi (elements)
__device__ __forceinline__ void stage_async(__nv_bfloat16* INs, int ldi, const __nv_bfloat16* src, int64_t ld,
                                            int nrows, int k0, int KC) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < nrows; r += NW) {
    const __nv_bfloat16* s = src + (int64_t)r * ld + k0;
    __nv_bfloat16* d = INs + r * ldi;
    for (int c = lane << 3; c < KC; c += 256) cp_async16(d + c, s + c);
  }
}

// ================================================================================================
// forward
// ================================================================================================
__global__ void __launch_bounds__(NT, 1) recur_fwd_bf16_kernel(RArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int H = a.H, ldw = H + PADB;
  __nv_bfloat16* Ws = reinterpret_cast<__nv_bfloat16*>(smem_raw);     // [32][ldw]  row j*8+u = Whh[j*H+u0+u, :]
  __nv_bfloat16* INs = Ws + 32 * ldw;                                  // [2][G][ldw] (double buffered)
  const int ub = blockIdx.x % a.n_ub, bb = blockIdx.x / a.n_ub;
  const int u0 = ub * 8, sb0 = bb * a.BB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pos_o = a.cell == SN_CELL_LSTM ? 3 : 2, pos_c = a.cell == SN_CELL_LSTM ? 2 : 3;
  const int u = u0 + (lane >> 2);          // the unit whose four gates this lane's accumulators hold
  const int sj = (lane & 3) * 2;           // its two samples inside the warp's 8-sample tile

  for (int i = tid; i < 32 * (H >> 3); i += NT) {
    int r = i / (H >> 3), c = (i - r * (H >> 3)) << 3;
    int j = r >> 3, uu = r & 7;
    *reinterpret_cast<uint4*>(Ws + r * ldw + c) =
        __ldg(reinterpret_cast<const uint4*>(a.Wb + (int64_t)(j * H + u0 + uu) * H + c));
  }
  __syncthreads();

  __nv_bfloat16* INbuf[2] = {INs, a.nbuf == 2 ? INs + G * ldw : INs};
  const float bh0 = a.bhh ? __ldg(a.bhh + u) : 0.f, bh1 = a.bhh ? __ldg(a.bhh + H + u) : 0.f;
  const float bh2 = a.bhh ? __ldg(a.bhh + 2 * H + u) : 0.f, bh3 = a.bhh ? __ldg(a.bhh + 3 * H + u) : 0.f;

  for (int t = a.t0; t < a.t1; ++t) {
    const int bt = a.bs[t];
    const int nv = min(max(bt - sb0, 0), a.BB);
    if (nv > 0) {
      const int64_t row0 = (int64_t)a.off[t] + sb0;
      const int n0 = warp * 8;
      // XP (+ recurrent bias) and c_{t-1} of this lane's two samples of a group: loaded one group AHEAD so the
      // global-memory latency hides under the wait / the previous group's MMAs
      float pz[2][4], pc[2];
      auto prefetch = [&](int g0) {
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int sl = g0 + n0 + sj + j;
          pc[j] = 0.f;
          pz[j][0] = pz[j][1] = pz[j][2] = pz[j][3] = 0.f;
          if (sl < nv) {
            const float* xp = a.XP + (row0 + sl) * 4 * H + u;
            // raw loads only: any arithmetic here would stall on the load and defeat the prefetch
            pz[j][0] = __ldg(xp); pz[j][1] = __ldg(xp + H);
            pz[j][2] = __ldg(xp + 2 * H); pz[j][3] = __ldg(xp + 3 * H);
            pc[j] = a.c_state[(int64_t)(sb0 + sl) * H + u];
          }
        }
      };
      prefetch(0);
      const bool first = (t == a.t0);
      if (!first) wait_flag(a.flags + bb * a.T + (t - 1), a.n_ub);
      const __nv_bfloat16* hsrc = first ? nullptr : a.Hb + ((int64_t)a.off[t - 1] + sb0) * H;
      // stage(g): asynchronous copy of group g's h_{t-1} rows into buffer g&1
      auto stage = [&](int g0, __nv_bfloat16* dst) {
        const int ng = min(G, nv - g0);
        if (!first) {
          stage_async(dst, ldw, hsrc + (int64_t)g0 * H, H, ng, 0, H);
        } else {
          // state before step t0 comes in fp32 (zeros when NULL): convert while staging
          for (int i = tid; i < ng * (H >> 1); i += NT) {
            int r = i / (H >> 1), c = (i - r * (H >> 1)) << 1;
            float2 v = make_float2(0.f, 0.f);
            if (a.h_init) v = *reinterpret_cast<const float2*>(a.h_init + (int64_t)(sb0 + g0 + r) * H + c);
            *reinterpret_cast<__nv_bfloat162*>(dst + r * ldw + c) = __floats2bfloat162_rn(v.x, v.y);
          }
        }
        cp_async_commit();
      };
      stage(0, INbuf[0]);
      int buf = 0;
      for (int g0 = 0; g0 < nv; g0 += G, buf ^= 1) {
        const int ng = min(G, nv - g0);
        const bool more = g0 + G < nv;
        if (more && a.nbuf == 2) { stage(g0 + G, INbuf[buf ^ 1]); cp_async_wait<1>(); } else { cp_async_wait<0>(); }
        __syncthreads();
        const __nv_bfloat16* IN = INbuf[buf];
        if (a.Hprevb) {      // h_{t-1} rows (bf16): operand of dW_hh = dZ^T Hprev
          for (int i = tid; i < ng; i += NT)
            *reinterpret_cast<uint4*>(a.Hprevb + (row0 + g0 + i) * H + u0) = *reinterpret_cast<const uint4*>(IN + i * ldw + u0);
        }
        // this group's operands were prefetched one iteration ago; move them aside and prefetch the next group's
        float z[2][4], cp_[2];
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          cp_[j] = pc[j];
#pragma unroll
          for (int q = 0; q < 4; ++q) z[j][q] = pz[j][q];
        }
        if (more) prefetch(g0 + G);
        float acc[2][4], acc2[2][4];       // two k-phases per m-tile: 4 independent MMA dependency chains
#pragma unroll
        for (int m = 0; m < 2; ++m)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[m][q] = acc2[m][q] = 0.f;
        if (n0 < ng) {
          // A (Ws): lanes 0-15 -> rows 0-15 @k0, lanes 16-31 -> rows 0-15 @k0+8
          const __nv_bfloat16* a_ptr = Ws + (lane & 15) * ldw + (lane >> 4) * 8;
          // B (IN): matrices {k0, k0+8, k0+16, k0+24} x samples n0..n0+7
          const __nv_bfloat16* b_ptr = IN + (n0 + (lane & 7)) * ldw + (lane >> 3) * 8;
#pragma unroll 2
          for (int k0 = 0; k0 < H; k0 += 32) {
            uint32_t bfr[4], a0[4], a1[4], a2[4], a3[4];
            ldmatrix_x4(bfr, b_ptr + k0);
            ldmatrix_x4(a0, a_ptr + k0);
            ldmatrix_x4(a1, a_ptr + 16 * ldw + k0);
            ldmatrix_x4(a2, a_ptr + k0 + 16);
            ldmatrix_x4(a3, a_ptr + 16 * ldw + k0 + 16);
            mma_bf16(acc[0], a0, bfr[0], bfr[1]);
            mma_bf16(acc[1], a1, bfr[0], bfr[1]);
            mma_bf16(acc2[0], a2, bfr[2], bfr[3]);
            mma_bf16(acc2[1], a3, bfr[2], bfr[3]);
          }
#pragma unroll
          for (int m = 0; m < 2; ++m)
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[m][q] += acc2[m][q];
        }
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int sl = g0 + n0 + sj + j;
          if (n0 + sj + j < ng) {
            const int64_t row = row0 + sl;
            // accumulator rows: m-tile 0 = gate blocks 0,1 ; m-tile 1 = gate blocks 2,3
            const float zi = acc[0][j] + z[j][0] + bh0, zf = acc[0][2 + j] + z[j][1] + bh1;
            const float za = acc[1][j] + z[j][2] + bh2, zb = acc[1][2 + j] + z[j][3] + bh3;
            const float zo = a.cell == SN_CELL_LSTM ? zb : za;
            const float zc = a.cell == SN_CELL_LSTM ? za : zb;
            const float gi = sn::sigmoidf_(zi), gf = sn::sigmoidf_(zf), go = sn::sigmoidf_(zo), gc = tanhf(zc);
            const float c = gf * cp_[j] + gi * gc;
            const float h = a.cell == SN_CELL_LSTM ? go * tanhf(c) : go * c;
            a.c_state[(int64_t)(sb0 + sl) * H + u] = c;
            a.Hb[row * H + u] = __float2bfloat16(h);
            if (a.Hall) a.Hall[row * H + u] = h;
            if (a.Call) a.Call[row * H + u] = c;
            if (a.gates) {
              float* gp = a.gates + row * 4 * H + u;
              gp[0] = gi; gp[H] = gf; gp[pos_o * H] = go; gp[pos_c * H] = gc;
            }
          }
        }
        __syncthreads();     // buffer `buf` may be overwritten by the staging issued in the next iteration
        if (more && a.nbuf == 1) stage(g0 + G, INbuf[0]);
      }
    }
    signal_flag(a.flags + bb * a.T + t);
  }
}

// ================================================================================================
// backward
// ================================================================================================
constexpr int UBB = 16;          // units per CTA in backward (a full m16 tile)
constexpr int GB = 32;           // samples per group: 4 tiles of 8; warps = 4 n-tiles x 2 k-halves
constexpr int KCB = 512;         // bf16 K chunk of dZ_{t+1} per pipeline stage

__global__ void __launch_bounds__(NT, 1) recur_bwd_bf16_kernel(RArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int H = a.H, K = 4 * H, ldw = K + PADB;
  const int KC = K < KCB ? K : KCB;
  const int ldi = KC + PADB;
  __nv_bfloat16* Ws = reinterpret_cast<__nv_bfloat16*>(smem_raw);      // [16][ldw]  Ws[u][k] = Whh[k, u0+u]
  __nv_bfloat16* IN0 = Ws + UBB * ldw;                                  // [2][GB][ldi]
  __nv_bfloat16* IN1 = IN0 + GB * ldi;
  float* part = reinterpret_cast<float*>(IN1 + GB * ldi);               // [2 k-halves][4 tiles][32 lanes][4]
  const int ub = blockIdx.x % a.n_ub, bb = blockIdx.x / a.n_ub;
  const int u0 = ub * UBB, sb0 = bb * a.BB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pos_o = a.cell == SN_CELL_LSTM ? 3 : 2, pos_c = a.cell == SN_CELL_LSTM ? 2 : 3;
  const int ntile = warp & 3, khalf = warp >> 2;
  // after the cross-warp reduction thread tid finishes pairs p = 2*tid, 2*tid+1:  s = p / 16, u = p % 16
  const int ps = (2 * tid) >> 4, pu = (2 * tid) & 15;

  // Ws[u][k] = Whh[k, u0+u]: each thread reads 8 consecutive units of one row k (16 B) and scatters them
  for (int i = tid; i < K * 2; i += NT) {
    const int k = i >> 1, half = i & 1;
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(a.Wb + (int64_t)k * H + u0 + half * 8));
    const __nv_bfloat16* e = reinterpret_cast<const __nv_bfloat16*>(&v);
#pragma unroll
    for (int q = 0; q < 8; ++q) Ws[(half * 8 + q) * ldw + k] = e[q];
  }
  __syncthreads();

  for (int t = a.t1 - 1; t >= a.t0 - 1; --t) {
    const bool tail = (t < a.t0);
    const int bt = tail ? a.bs[a.t0] : a.bs[t];
    const int nv = min(max(bt - sb0, 0), a.BB);
    const int bnext = (t + 1 < a.t1) ? a.bs[t + 1] : 0;
    const int nrec = min(max(bnext - sb0, 0), a.BB);
    if (tail) {
      for (int i = tid; i < a.BB * UBB; i += NT) {
        int s = i >> 4, uu = i & 15;
        if (s >= nrec && sb0 + s < a.B) a.dh_carry[(int64_t)(sb0 + s) * H + u0 + uu] = 0.f;
      }
    }
    if (nv > 0) {
      const int64_t row0 = tail ? 0 : (int64_t)a.off[t] + sb0;
      const int64_t rown0 = (t + 1 < a.t1) ? (int64_t)a.off[t + 1] + sb0 : 0;
      // prefetch step-local operands of this thread's two (sample, unit) pairs of the first group
      float p_g[2][4], p_c[2], p_cp[2], p_dh[2], p_dc[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        p_c[j] = p_cp[j] = p_dh[j] = p_dc[j] = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) p_g[j][q] = 0.f;
        const int sl = ps, uu = u0 + pu + j;
        if (!tail && sl < min(GB, nv)) {
          const int64_t row = row0 + sl, sidx = (int64_t)(sb0 + sl) * H + uu;
          const float* gp = a.gates + row * K + uu;
#pragma unroll
          for (int q = 0; q < 4; ++q) p_g[j][q] = __ldg(gp + q * H);
          p_c[j] = __ldg(a.Call + row * H + uu);
          if (t > 0) p_cp[j] = __ldg(a.Call + ((int64_t)a.off[t - 1] + sb0 + sl) * H + uu);
          else p_cp[j] = a.c_init ? __ldg(a.c_init + sidx) : 0.f;
          p_dh[j] = __ldg(a.dHall + row * H + uu);
          p_dc[j] = a.dc_carry[sidx];
        }
      }
      if (nrec > 0) wait_flag(a.flags + bb * a.T + (t + 1), a.n_ub);
      for (int g0 = 0; g0 < nv; g0 += GB) {
        const int ng = min(GB, nv - g0);
        const int ngrec = min(max(nrec - g0, 0), GB);
        if (ngrec > 0) {
          float acc[4] = {0.f, 0.f, 0.f, 0.f}, accb[4] = {0.f, 0.f, 0.f, 0.f};   // two MMA dependency chains
          const __nv_bfloat16* src = a.dZb + (rown0 + g0) * K;
          const int nchunk = K / KC;
          stage_async(IN0, ldi, src, K, ngrec, 0, KC);
          cp_async_commit();
          for (int ci = 0; ci < nchunk; ++ci) {
            __nv_bfloat16* cur = (ci & 1) ? IN1 : IN0;
            __nv_bfloat16* nxt = (ci & 1) ? IN0 : IN1;
            if (ci + 1 < nchunk) {
              stage_async(nxt, ldi, src, K, ngrec, (ci + 1) * KC, KC);
              cp_async_commit();
              cp_async_wait<1>();
            } else {
              cp_async_wait<0>();
            }
            __syncthreads();
            if (ntile * 8 < ngrec) {
              const int kh = KC >> 1;                      // this warp's half of the chunk
              const __nv_bfloat16* a_ptr = Ws + (lane & 15) * ldw + ci * KC + khalf * kh + (lane >> 4) * 8;
              const __nv_bfloat16* b_ptr = cur + (ntile * 8 + (lane & 7)) * ldi + khalf * kh + (lane >> 3) * 8;
#pragma unroll 2
              for (int k0 = 0; k0 < kh; k0 += 32) {
                uint32_t bfr[4], a0[4], a1[4];
                ldmatrix_x4(bfr, b_ptr + k0);
                ldmatrix_x4(a0, a_ptr + k0);
                ldmatrix_x4(a1, a_ptr + k0 + 16);
                mma_bf16(acc, a0, bfr[0], bfr[1]);
                mma_bf16(accb, a1, bfr[2], bfr[3]);
              }
            }
            __syncthreads();
          }
          *reinterpret_cast<float4*>(part + ((khalf * 4 + ntile) * 32 + lane) * 4) =
              make_float4(acc[0] + accb[0], acc[1] + accb[1], acc[2] + accb[2], acc[3] + accb[3]);
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int sloc = ps, uu = pu + j;                // sample inside the group, unit inside the CTA
          if (sloc < ng) {
            const int sl = g0 + sloc;
            const int64_t sidx = (int64_t)(sb0 + sl) * H + u0 + uu;
            float dh_rec = 0.f;
            if (sl < nrec) {
              // accumulator element of (unit uu, sample sloc): tile sloc/8, lane (uu%8)*4 + (sloc%8)/2,
              // register (uu/8)*2 + (sloc%2); the two k-halves are summed here
              const int tl = sloc >> 3, ln = ((uu & 7) << 2) + ((sloc & 7) >> 1), rg = ((uu >> 3) << 1) + (sloc & 1);
              dh_rec = part[((0 * 4 + tl) * 32 + ln) * 4 + rg] + part[((1 * 4 + tl) * 32 + ln) * 4 + rg];
            }
            if (tail) {
              if (sl < nrec) a.dh_carry[sidx] = dh_rec;
              continue;
            }
            if (t == a.t1 - 1) dh_rec = a.dh_carry[sidx];
            const int64_t row = row0 + sl;
            const int ug = u0 + uu;
            float gi, gf, go, gc, c, cprev, dhl, dcar;
            if (g0 == 0) {
              gi = p_g[j][0]; gf = p_g[j][1];
              go = a.cell == SN_CELL_LSTM ? p_g[j][3] : p_g[j][2];
              gc = a.cell == SN_CELL_LSTM ? p_g[j][2] : p_g[j][3];
              c = p_c[j]; cprev = p_cp[j]; dhl = p_dh[j]; dcar = p_dc[j];
            } else {
              const float* gp = a.gates + row * K + ug;
              gi = gp[0]; gf = gp[H]; go = gp[pos_o * H]; gc = gp[pos_c * H];
              c = a.Call[row * H + ug];
              if (t > 0) cprev = a.Call[((int64_t)a.off[t - 1] + sb0 + sl) * H + ug];
              else cprev = a.c_init ? a.c_init[sidx] : 0.f;
              dhl = a.dHall[row * H + ug];
              dcar = a.dc_carry[sidx];
            }
            const float dh = dhl + dh_rec;
            float d_o, dc;
            if (a.cell == SN_CELL_LSTM) {
              float tc = tanhf(c);
              d_o = dh * tc;
              dc = dcar + dh * go * (1.f - tc * tc);
            } else {
              d_o = dh * c;
              dc = dcar + dh * go;
            }
            const float di = dc * gc, df = dc * cprev, dg = dc * gi;
            a.dc_carry[sidx] = dc * gf;
            const float z0 = di * gi * (1.f - gi), z1 = df * gf * (1.f - gf);
            const float zo = d_o * go * (1.f - go), zc = dg * (1.f - gc * gc);
            __nv_bfloat16* dzb = a.dZb + row * K + ug;
            dzb[0] = __float2bfloat16(z0); dzb[H] = __float2bfloat16(z1);
            dzb[pos_o * H] = __float2bfloat16(zo); dzb[pos_c * H] = __float2bfloat16(zc);
            if (a.dZ) {
              float* dz = a.dZ + row * K + ug;
              dz[0] = z0; dz[H] = z1; dz[pos_o * H] = zo; dz[pos_c * H] = zc;
            }
          }
        }
        __syncthreads();
      }
    }
    if (!tail) signal_flag(a.flags + bb * a.T + t);
  }
}

int32_t plan(bool bwd, int64_t H, int64_t B, int* n_ub, int* nbb, int* BB, size_t* smem) {
  const sn::DevInfo& d = sn::dev_info();
  const int ub = bwd ? UBB : 8;
  if (H % 32 != 0 || H < 32) return sn::fail(-1, "sn_recur_*_bf16: hidden size %lld must be a multiple of 32", (long long)H);
  *n_ub = (int)(H / ub);
  if (*n_ub > d.sm_count) return sn::fail(-1, "sn_recur_*_bf16: hidden size %lld needs %d CTAs > %d SMs", (long long)H, *n_ub, d.sm_count);
  int nb = d.sm_count / *n_ub;
  int maxnb = (int)((B + 7) / 8);
  if (nb > maxnb) nb = maxnb;
  if (nb < 1) nb = 1;
  *nbb = nb;
  *BB = (int)((B + nb - 1) / nb);
  if (!bwd) {
    *smem = ((size_t)32 * (H + PADB) + (size_t)2 * G * (H + PADB)) * 2;   // W slice + double-buffered h staging
    if (*smem > (size_t)d.smem_optin) *smem = ((size_t)32 * (H + PADB) + (size_t)G * (H + PADB)) * 2;   // single
  } else {
    const int64_t K = 4 * H, KC = K < KCB ? K : KCB;
    *smem = ((size_t)UBB * (K + PADB) + (size_t)2 * GB * (KC + PADB)) * 2 + 2 * 4 * 32 * 4 * sizeof(float);
  }
  if (*smem > (size_t)d.smem_optin) return sn::fail(-1, "sn_recur_*_bf16: hidden size %lld does not fit shared memory", (long long)H);
  return 0;
}

template <typename Kern>
int32_t launch(Kern kernel, RArgs& a, size_t smem, cudaStream_t stream, const char* what) {
  SN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  void* params[] = {&a};
  dim3 grid((unsigned)(a.n_ub * a.nbb)), block(NT);
  cudaError_t e;
  if (sn::recur_cooperative()) {
    e = cudaLaunchCooperativeKernel((const void*)kernel, grid, block, params, smem, stream);
  } else {
    if ((int)grid.x > sn::dev_info().sm_count) return sn::fail(-1, "%s: grid of %u CTAs exceeds the %d SMs", what, grid.x, sn::dev_info().sm_count);
    e = cudaLaunchKernel((const void*)kernel, grid, block, params, smem, stream);   // see sn::recur_cooperative()
  }
  if (e != cudaSuccess) return sn::fail((int32_t)e, "%s: cooperative launch failed: %s", what, cudaGetErrorString(e));
  return 0;
}

}  // namespace

extern "C" {

int32_t sn_recur_fwd_bf16(int32_t cell, int64_t H, int64_t B, const int32_t* batch_sizes, const int32_t* offsets,
                          int32_t t0, int32_t t1, const float* XP, const void* Whh_bf16, const float* bhh,
                          const float* h_init, float* Hall, void* Hb, void* Hprevb, float* Call, float* gates,
                          float* c_state, void* ws, void* stream) {
  SN_REQUIRE(cell == SN_CELL_FACTORED || cell == SN_CELL_LSTM, "sn_recur_fwd_bf16: bad cell %d", cell);
  SN_REQUIRE(t0 >= 0 && t1 >= t0 && B > 0, "sn_recur_fwd_bf16: bad step range");
  SN_REQUIRE(XP && Whh_bf16 && Hb && c_state && ws && batch_sizes && offsets, "sn_recur_fwd_bf16: null argument");
  if (t1 == t0) return 0;
  RArgs a = {};
  size_t smem;
  int32_t rc = plan(false, H, B, &a.n_ub, &a.nbb, &a.BB, &smem);
  if (rc) return rc;
  a.cell = cell; a.H = (int)H; a.B = (int)B; a.t0 = t0; a.t1 = t1; a.T = t1;
  a.bs = batch_sizes; a.off = offsets; a.XP = XP; a.Wb = (const __nv_bfloat16*)Whh_bf16; a.bhh = bhh;
  a.h_init = h_init; a.Hall = Hall; a.Hb = (__nv_bfloat16*)Hb; a.Hprevb = (__nv_bfloat16*)Hprevb;
  a.Call = Call; a.gates = gates; a.c_state = c_state; a.flags = (int*)ws;
  a.nbuf = smem >= ((size_t)32 * (H + PADB) + (size_t)2 * G * (H + PADB)) * 2 ? 2 : 1;
  cudaStream_t st = (cudaStream_t)stream;
  SN_CUDA(cudaMemsetAsync(ws, 0, sizeof(int) * (size_t)a.nbb * (size_t)t1, st));
  return launch(recur_fwd_bf16_kernel, a, smem, st, "sn_recur_fwd_bf16");
}

int32_t sn_recur_bwd_bf16(int32_t cell, int64_t H, int64_t B, const int32_t* batch_sizes, const int32_t* offsets,
                          int32_t t0, int32_t t1, const void* Whh_bf16, const float* c_init, const float* Call,
                          const float* gates, const float* dHall, float* dZ, void* dZb, float* dh_carry,
                          float* dc_carry, void* ws, void* stream) {
  SN_REQUIRE(cell == SN_CELL_FACTORED || cell == SN_CELL_LSTM, "sn_recur_bwd_bf16: bad cell %d", cell);
  SN_REQUIRE(t0 >= 0 && t1 >= t0 && B > 0, "sn_recur_bwd_bf16: bad step range");
  SN_REQUIRE(Whh_bf16 && Call && gates && dHall && dZb && dh_carry && dc_carry && ws, "sn_recur_bwd_bf16: null argument");
  if (t1 == t0) return 0;
  RArgs a = {};
  size_t smem;
  int32_t rc = plan(true, H, B, &a.n_ub, &a.nbb, &a.BB, &smem);
  if (rc) return rc;
  a.cell = cell; a.H = (int)H; a.B = (int)B; a.t0 = t0; a.t1 = t1; a.T = t1 + 1;
  a.bs = batch_sizes; a.off = offsets; a.Wb = (const __nv_bfloat16*)Whh_bf16; a.c_init = c_init;
  a.Call = const_cast<float*>(Call); a.gates = const_cast<float*>(gates); a.dHall = dHall; a.dZ = dZ;
  a.dZb = (__nv_bfloat16*)dZb; a.dh_carry = dh_carry; a.dc_carry = dc_carry; a.flags = (int*)ws;
  cudaStream_t st = (cudaStream_t)stream;
  SN_CUDA(cudaMemsetAsync(ws, 0, sizeof(int) * (size_t)a.nbb * (size_t)(t1 + 1), st));
  return launch(recur_bwd_bf16_kernel, a, smem, st, "sn_recur_bwd_bf16");
}

}  // extern "C"
