// K3: persistent recurrence kernels (fp32 exact path).
//
// One cooperative launch runs all T steps of   z_t = XP_t + h_{t-1} Whh^T ; gates ; c_t ; h_t .
// The grid is (H/8 unit blocks) x (nbb batch blocks) <= #SMs, one CTA per SM:
//   * CTA (ub, bb) owns hidden units [8ub, 8ub+8) (all four gate rows of them -> gate fusion is
//     lane-local) for the samples of batch block bb, for ALL time steps;
//   * its 32 rows of Whh stay resident in shared memory for the whole launch (read from HBM once);
//   * h_t is exchanged between the CTAs of one batch block through Hall (L2) with one release/acquire
//     counter per (batch block, step): no grid-wide barrier, CTAs of different batch blocks never wait
//     on each other;
//   * inside a warp: 8 unit lanes x 4 k lanes, SW samples per warp -> 4*SW accumulators per lane,
//     k-partials combined with two warp shuffles, then the gate nonlinearities run in registers.
// The backward kernel is the mirror image in reverse time with the column slice Whh[:, units] resident
// (K = 4H), producing dZ (= dXP) and carrying dc in place.
#include "sn_common.cuh"

namespace {

constexpr int UB = 8;          // hidden units per CTA
constexpr int NW = 8;          // warps per CTA
constexpr int NT = NW * 32;

struct RecurArgs {
  int cell, H, B, t0, t1, T;
  const int* bs;
  const int* off;
  const float* XP;
  const float* W;
  const float* h_init;
  const float* c_init;
  const float* bhh;
  float* Hall;
  float* Call;
  float* Hprev;
  float* gates;
  float* c_state;
  const float* dHall;
  float* dZ;
  float* dh_carry;
  float* dc_carry;
  int* flags;
  int n_ub, nbb, BB;
};

__device__ __forceinline__ void wait_flag(const int* flag, int target) {
  if (threadIdx.x == 0) {
    while (sn::ld_acquire(flag) < target) { }
  }
  __syncthreads();
}
__device__ __forceinline__ void signal_flag(int* flag) {
  __syncthreads();
  if (threadIdx.x == 0) sn::red_release_add(flag, 1);   // release.gpu: cumulative over the CTA's writes (bar.sync above)
}

// acc[j][s] += sum_k Ws[(j*8+rl)][k] * INs[(sw0+s)][k] over this lane's k subset of [0,KC)
template <int RPL, int SW>
__device__ __forceinline__ void matvec_accum(const float* __restrict__ Ws, int ldw, int wk0,
                                             const float* __restrict__ INs, int ldi, int KC, int rl, int kl,
                                             int sw0, float (&acc)[RPL][SW]) {
  for (int k = kl * 4; k < KC; k += 16) {
    float4 w[RPL];
#pragma unroll
    for (int j = 0; j < RPL; ++j) w[j] = *reinterpret_cast<const float4*>(Ws + (j * UB + rl) * ldw + wk0 + k);
#pragma unroll
    for (int s = 0; s < SW; ++s) {
      float4 x = *reinterpret_cast<const float4*>(INs + (sw0 + s) * ldi + k);
#pragma unroll
      for (int j = 0; j < RPL; ++j) {
        float a = acc[j][s];
        a = fmaf(w[j].x, x.x, a);
        a = fmaf(w[j].y, x.y, a);
        a = fmaf(w[j].z, x.z, a);
        a = fmaf(w[j].w, x.w, a);
        acc[j][s] = a;
      }
    }
  }
}

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gsrc) {
  unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// asynchronous staging of rows [0,nrows) x cols [k0,k0+KC) into smem (L2-coherent .cg path, all copies of
// a thread in flight at once); rows past nrows are left untouched (their results are never stored)
__device__ __forceinline__ void stage_rows_async(float* __restrict__ INs, int ldi, const float* __restrict__ src,
                                                 int64_t ld, int nrows, int k0, int KC) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int r = warp; r < nrows; r += NW) {            // division-free: a warp per row, 16-byte chunks per lane
    const float* s = src + (int64_t)r * ld + k0;
    float* d = INs + r * ldi;
    for (int c = lane << 2; c < KC; c += 128) cp_async16(d + c, s + c);
  }
}

template <int SW>
__global__ void __launch_bounds__(NT, 1) recur_fwd_kernel(RecurArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int H = a.H, ldw = H + 4;
  constexpr int G = NW * SW;
  float* Ws = smem;                  // [32][ldw]
  float* INs = smem + 32 * ldw;      // [G][ldw]
  const int ub = blockIdx.x % a.n_ub, bb = blockIdx.x / a.n_ub;
  const int u0 = ub * UB, sb0 = bb * a.BB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rl = lane & 7, kl = lane >> 3;
  const int pos_o = a.cell == SN_CELL_LSTM ? 3 : 2, pos_c = a.cell == SN_CELL_LSTM ? 2 : 3;

  // resident weight slice: row (j*8+u) <- Whh[j*H + u0 + u, :]
  for (int i = tid; i < 32 * (H >> 2); i += NT) {
    int r = i / (H >> 2), c = (i - r * (H >> 2)) << 2;
    int j = r >> 3, u = r & 7;
    *reinterpret_cast<float4*>(Ws + r * ldw + c) =
        __ldg(reinterpret_cast<const float4*>(a.W + (int64_t)(j * H + u0 + u) * H + c));
  }
  __syncthreads();

  constexpr int NSLOT = (SW + 3) / 4;     // samples of a group this lane finishes: s = kl + 4*j
  for (int t = a.t0; t < a.t1; ++t) {
    const int bt = a.bs[t];
    const int nv = min(max(bt - sb0, 0), a.BB);
    int* flag_t = a.flags + bb * a.T + t;
    if (nv > 0) {
      const int64_t row0 = (int64_t)a.off[t] + sb0;
      const int sw0 = warp * SW;
      // ---- prefetch everything that does not depend on h_{t-1}: XP (+bias) and c_{t-1} of this lane's
      //      (sample, unit) slots of the first group -- in flight while we wait for the other CTAs
      float pz[NSLOT][4], pc[NSLOT];
#pragma unroll
      for (int j = 0; j < NSLOT; ++j) {
        const int sl = sw0 + kl + 4 * j;
        pc[j] = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) pz[j][q] = 0.f;
        if (kl + 4 * j < SW && sl < min(G, nv)) {
          const float* xp = a.XP + (row0 + sl) * 4 * H + u0 + rl;
#pragma unroll
          for (int q = 0; q < 4; ++q) pz[j][q] = __ldg(xp + q * H);     // raw loads: arithmetic here would stall the prefetch
          pc[j] = a.c_state[(int64_t)(sb0 + sl) * H + u0 + rl];
        }
      }
      const float* hsrc;
      if (t == a.t0) {
        hsrc = a.h_init ? a.h_init + (int64_t)sb0 * H : nullptr;
      } else {
        hsrc = a.Hall + ((int64_t)a.off[t - 1] + sb0) * H;
        wait_flag(a.flags + bb * a.T + (t - 1), a.n_ub);
      }
      for (int g0 = 0; g0 < nv; g0 += G) {
        const int ng = min(G, nv - g0);
        if (hsrc) {
          stage_rows_async(INs, ldw, hsrc + (int64_t)g0 * H, H, ng, 0, H);
          cp_async_commit();
          cp_async_wait<0>();
        } else {
          for (int i = tid; i < ng * (H >> 2); i += NT) {
            int r = i / (H >> 2), c = (i - r * (H >> 2)) << 2;
            *reinterpret_cast<float4*>(INs + r * ldw + c) = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
        __syncthreads();
        if (a.Hprev) {
          for (int i = tid; i < ng * UB; i += NT) {
            int s = i >> 3, u = i & 7;
            a.Hprev[(row0 + g0 + s) * H + u0 + u] = INs[s * ldw + u0 + u];
          }
        }
        float acc[4][SW];
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int s = 0; s < SW; ++s) acc[j][s] = 0.f;
        if (sw0 < ng) matvec_accum<4, SW>(Ws, ldw, 0, INs, ldw, H, rl, kl, sw0, acc);
#pragma unroll
        for (int j = 0; j < 4; ++j)
#pragma unroll
          for (int s = 0; s < SW; ++s) {
            float v = acc[j][s];
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            acc[j][s] = v;
          }
#pragma unroll
        for (int s = 0; s < SW; ++s) {
          if ((s & 3) == kl && sw0 + s < ng) {
            const int sl = g0 + sw0 + s;             // sample index inside the batch block
            const int64_t row = row0 + sl;
            const int u = u0 + rl;
            float z0, z1, z2, z3, cprev;
            float* cst = a.c_state + (int64_t)(sb0 + sl) * H + u;
            if (g0 == 0) {
              z0 = pz[s >> 2][0]; z1 = pz[s >> 2][1]; z2 = pz[s >> 2][2]; z3 = pz[s >> 2][3];
              cprev = pc[s >> 2];
            } else {
              const float* xp = a.XP + row * 4 * H + u;
              z0 = xp[0]; z1 = xp[H]; z2 = xp[2 * H]; z3 = xp[3 * H];
              cprev = *cst;
            }
            if (a.bhh) { z0 += a.bhh[u]; z1 += a.bhh[H + u]; z2 += a.bhh[2 * H + u]; z3 += a.bhh[3 * H + u]; }
            const float zi = acc[0][s] + z0, zf = acc[1][s] + z1;
            const float za = acc[2][s] + z2, zb = acc[3][s] + z3;
            const float zo = a.cell == SN_CELL_LSTM ? zb : za;
            const float zc = a.cell == SN_CELL_LSTM ? za : zb;
            const float gi = sn::sigmoidf_(zi), gf = sn::sigmoidf_(zf), go = sn::sigmoidf_(zo), gc = tanhf(zc);
            const float c = gf * cprev + gi * gc;
            const float h = a.cell == SN_CELL_LSTM ? go * tanhf(c) : go * c;
            *cst = c;
            a.Hall[row * H + u] = h;
            if (a.Call) a.Call[row * H + u] = c;
            if (a.gates) {
              float* gp = a.gates + row * 4 * H + u;
              gp[0] = gi; gp[H] = gf; gp[pos_o * H] = go; gp[pos_c * H] = gc;
            }
          }
        }
        __syncthreads();   // INs reused by the next group / step
      }
    }
    signal_flag(flag_t);
  }
}

// K chunk of dZ_{t+1} staged per pipeline stage (double buffered): largest power of two <= 256 dividing 4H
__host__ __device__ inline int bwd_kc(int H) {
  int kc = 256;
  while ((4 * H) % kc != 0) kc >>= 1;
  return kc;
}

template <int SW>
__global__ void __launch_bounds__(NT, 1) recur_bwd_kernel(RecurArgs a) {
  extern __shared__ __align__(16) float smem[];
  const int H = a.H, K = 4 * H, ldw = K + 4;
  constexpr int G = NW * SW;
  const int KC = bwd_kc(H);
  const int ldi = KC + 4;
  float* Ws = smem;                        // [8][ldw]   Ws[u][k] = Whh[k, u0+u]
  float* INs0 = smem + UB * ldw;           // [2][G][ldi]
  float* INs1 = INs0 + G * ldi;
  const int ub = blockIdx.x % a.n_ub, bb = blockIdx.x / a.n_ub;
  const int u0 = ub * UB, sb0 = bb * a.BB;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rl = lane & 7, kl = lane >> 3;
  const int pos_o = a.cell == SN_CELL_LSTM ? 3 : 2, pos_c = a.cell == SN_CELL_LSTM ? 2 : 3;
  constexpr int NSLOT = (SW + 3) / 4;

  for (int i = tid; i < K * UB; i += NT) {
    int k = i >> 3, u = i & 7;
    Ws[u * ldw + k] = __ldg(a.W + (int64_t)k * H + u0 + u);
  }
  __syncthreads();

  // t runs t1-1 .. t0, plus one extra pass (t == t0-1) that only produces dh_carry = dZ_{t0} Whh
  for (int t = a.t1 - 1; t >= a.t0 - 1; --t) {
    const bool tail = (t < a.t0);
    const int bt = tail ? a.bs[a.t0] : a.bs[t];
    const int nv = min(max(bt - sb0, 0), a.BB);
    const int bnext = (t + 1 < a.t1) ? a.bs[t + 1] : 0;   // samples with a recurrent gradient from t+1
    const int nrec = min(max(bnext - sb0, 0), a.BB);
    if (tail) {
      for (int i = tid; i < a.BB * UB; i += NT) {
        int s = i >> 3, u = i & 7;
        if (s >= nrec && sb0 + s < a.B) a.dh_carry[(int64_t)(sb0 + s) * H + u0 + u] = 0.f;
      }
    }
    if (nv > 0) {
      const int64_t row0 = tail ? 0 : (int64_t)a.off[t] + sb0;
      const int64_t rown0 = (t + 1 < a.t1) ? (int64_t)a.off[t + 1] + sb0 : 0;
      const int sw0 = warp * SW;
      // ---- prefetch the step-local operands of this lane's slots (independent of dZ_{t+1})
      float p_g[NSLOT][4], p_c[NSLOT], p_cp[NSLOT], p_dh[NSLOT], p_dc[NSLOT];
#pragma unroll
      for (int j = 0; j < NSLOT; ++j) {
        const int sl = sw0 + kl + 4 * j;
        p_c[j] = p_cp[j] = p_dh[j] = p_dc[j] = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) p_g[j][q] = 0.f;
        if (!tail && kl + 4 * j < SW && sl < min(G, nv)) {
          const int u = u0 + rl;
          const int64_t row = row0 + sl;
          const int64_t sidx = (int64_t)(sb0 + sl) * H + u;
          const float* gp = a.gates + row * K + u;
#pragma unroll
          for (int q = 0; q < 4; ++q) p_g[j][q] = __ldg(gp + q * H);
          p_c[j] = __ldg(a.Call + row * H + u);
          if (t > 0) p_cp[j] = __ldg(a.Call + ((int64_t)a.off[t - 1] + sb0 + sl) * H + u);
          else p_cp[j] = a.c_init ? __ldg(a.c_init + sidx) : 0.f;
          p_dh[j] = __ldg(a.dHall + row * H + u);
          p_dc[j] = a.dc_carry[sidx];
        }
      }
      if (nrec > 0) wait_flag(a.flags + bb * a.T + (t + 1), a.n_ub);
      for (int g0 = 0; g0 < nv; g0 += G) {
        const int ng = min(G, nv - g0);
        const int ngrec = min(max(nrec - g0, 0), G);
        float acc[1][SW];
#pragma unroll
        for (int s = 0; s < SW; ++s) acc[0][s] = 0.f;
        if (ngrec > 0) {
          const float* src = a.dZ + (rown0 + g0) * K;
          const int nchunk = K / KC;
          stage_rows_async(INs0, ldi, src, K, ngrec, 0, KC);
          cp_async_commit();
          for (int ci = 0; ci < nchunk; ++ci) {
            float* cur = (ci & 1) ? INs1 : INs0;
            float* nxt = (ci & 1) ? INs0 : INs1;
            if (ci + 1 < nchunk) {
              stage_rows_async(nxt, ldi, src, K, ngrec, (ci + 1) * KC, KC);
              cp_async_commit();
              cp_async_wait<1>();
            } else {
              cp_async_wait<0>();
            }
            __syncthreads();
            if (sw0 < ngrec) matvec_accum<1, SW>(Ws, ldw, ci * KC, cur, ldi, KC, rl, kl, sw0, acc);
            __syncthreads();
          }
#pragma unroll
          for (int s = 0; s < SW; ++s) {
            float v = acc[0][s];
            v += __shfl_xor_sync(0xffffffffu, v, 8);
            v += __shfl_xor_sync(0xffffffffu, v, 16);
            acc[0][s] = v;
          }
        }
#pragma unroll
        for (int s = 0; s < SW; ++s) {
          if ((s & 3) == kl && sw0 + s < ng) {
            const int sl = g0 + sw0 + s;
            const int u = u0 + rl;
            const int64_t sidx = (int64_t)(sb0 + sl) * H + u;
            float dh_rec = (sl < nrec) ? acc[0][s] : 0.f;
            if (tail) {
              if (sl < nrec) a.dh_carry[sidx] = dh_rec;
              continue;
            }
            if (t == a.t1 - 1) dh_rec = a.dh_carry[sidx];
            const int64_t row = row0 + sl;
            float gi, gf, go, gc, c, cprev, dhl, dcar;
            if (g0 == 0) {
              const int j = s >> 2;
              gi = p_g[j][0]; gf = p_g[j][1];
              go = a.cell == SN_CELL_LSTM ? p_g[j][3] : p_g[j][2];
              gc = a.cell == SN_CELL_LSTM ? p_g[j][2] : p_g[j][3];
              c = p_c[j]; cprev = p_cp[j]; dhl = p_dh[j]; dcar = p_dc[j];
            } else {
              const float* gp = a.gates + row * K + u;
              gi = gp[0]; gf = gp[H]; go = gp[pos_o * H]; gc = gp[pos_c * H];
              c = a.Call[row * H + u];
              if (t > 0) cprev = a.Call[((int64_t)a.off[t - 1] + sb0 + sl) * H + u];
              else cprev = a.c_init ? a.c_init[sidx] : 0.f;
              dhl = a.dHall[row * H + u];
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
            float* dz = a.dZ + row * K + u;
            dz[0] = di * gi * (1.f - gi);
            dz[H] = df * gf * (1.f - gf);
            dz[pos_o * H] = d_o * go * (1.f - go);
            dz[pos_c * H] = dg * (1.f - gc * gc);
          }
        }
      }
    }
    if (!tail) signal_flag(a.flags + bb * a.T + t);
  }
}

struct Plan {
  int n_ub, nbb, BB, SW;
  size_t smem;
};

int32_t make_plan(bool bwd, int64_t H, int64_t B, Plan* p) {
  const sn::DevInfo& d = sn::dev_info();
  if (H % 16 != 0 || H < 16) return sn::fail(-1, "sn_recur: hidden size %lld must be a multiple of 16", (long long)H);
  p->n_ub = (int)(H / UB);
  if (p->n_ub > d.sm_count) return sn::fail(-1, "sn_recur: hidden size %lld needs %d CTAs > %d SMs", (long long)H, p->n_ub, d.sm_count);
  int nbb = d.sm_count / p->n_ub;
  int max_nbb = (int)((B + 7) / 8);
  if (nbb > max_nbb) nbb = max_nbb;
  if (nbb < 1) nbb = 1;
  p->nbb = nbb;
  p->BB = (int)((B + nbb - 1) / nbb);
  const int cands[3] = {6, 4, 2};
  int best = 0;
  for (int i = 0; i < 3; ++i) {
    int sw = cands[i];
    const int64_t kc = bwd_kc((int)H);
    size_t smem = bwd ? ((size_t)UB * (4 * H + 4) + (size_t)2 * NW * sw * (kc + 4)) * 4
                      : ((size_t)32 * (H + 4) + (size_t)NW * sw * (H + 4)) * 4;
    if (smem > (size_t)d.smem_optin) continue;
    // smallest SW whose group covers the batch block, else the largest that fits
    if (best == 0) { best = sw; p->smem = smem; }
    if (NW * sw >= p->BB) { best = sw; p->smem = smem; }
  }
  if (best == 0) return sn::fail(-1, "sn_recur: hidden size %lld does not fit shared memory", (long long)H);
  p->SW = best;
  return 0;
}

template <typename K>
int32_t launch_coop(K kernel, const Plan& p, RecurArgs& a, cudaStream_t stream, const char* what) {
  SN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem));
  void* params[] = {&a};
  dim3 grid((unsigned)(p.n_ub * p.nbb)), block(NT);
  cudaError_t e;
  if (sn::recur_cooperative()) {
    e = cudaLaunchCooperativeKernel((const void*)kernel, grid, block, params, p.smem, stream);
  } else {
    if ((int)grid.x > sn::dev_info().sm_count) return sn::fail(-1, "%s: grid of %u CTAs exceeds the %d SMs", what, grid.x, sn::dev_info().sm_count);
    e = cudaLaunchKernel((const void*)kernel, grid, block, params, p.smem, stream);   // see sn::recur_cooperative()
  }
  if (e != cudaSuccess) return sn::fail((int32_t)e, "%s: cooperative launch failed: %s", what, cudaGetErrorString(e));
  return 0;
}

}  // namespace

extern "C" {

int64_t sn_recur_ws_bytes(int64_t B, int64_t T) {
  (void)B;
  return (int64_t)sizeof(int) * 160 * (T + 1);
}

int32_t sn_recur_fwd(int32_t cell, int64_t H, int64_t B, const int32_t* batch_sizes, const int32_t* offsets,
                     int32_t t0, int32_t t1, const float* XP, const float* Whh, const float* h_init,
                     const float* bhh, float* Hall, float* Call, float* Hprev, float* gates, float* c_state,
                     void* ws, void* stream) {
  SN_REQUIRE(cell == SN_CELL_FACTORED || cell == SN_CELL_LSTM, "sn_recur_fwd: bad cell %d", cell);
  SN_REQUIRE(t0 >= 0 && t1 >= t0 && B > 0, "sn_recur_fwd: bad step range [%d,%d) or B", t0, t1);
  SN_REQUIRE(XP && Whh && Hall && c_state && ws && batch_sizes && offsets, "sn_recur_fwd: null argument");
  if (t1 == t0) return 0;
  Plan p;
  int32_t rc = make_plan(false, H, B, &p);
  if (rc) return rc;
  RecurArgs a = {};
  a.cell = cell; a.H = (int)H; a.B = (int)B; a.t0 = t0; a.t1 = t1; a.T = t1;
  a.bs = batch_sizes; a.off = offsets; a.XP = XP; a.W = Whh; a.h_init = h_init; a.bhh = bhh;
  a.Hall = Hall; a.Call = Call; a.Hprev = Hprev; a.gates = gates; a.c_state = c_state;
  a.flags = (int*)ws; a.n_ub = p.n_ub; a.nbb = p.nbb; a.BB = p.BB;
  cudaStream_t st = (cudaStream_t)stream;
  SN_CUDA(cudaMemsetAsync(ws, 0, sizeof(int) * (size_t)p.nbb * (size_t)t1, st));
  switch (p.SW) {
    case 6: return launch_coop(recur_fwd_kernel<6>, p, a, st, "sn_recur_fwd");
    case 4: return launch_coop(recur_fwd_kernel<4>, p, a, st, "sn_recur_fwd");
    default: return launch_coop(recur_fwd_kernel<2>, p, a, st, "sn_recur_fwd");
  }
}

int32_t sn_recur_bwd(int32_t cell, int64_t H, int64_t B, const int32_t* batch_sizes, const int32_t* offsets,
                     int32_t t0, int32_t t1, const float* Whh, const float* c_init, const float* Call,
                     const float* gates, const float* dHall, float* dZ, float* dh_carry, float* dc_carry,
                     void* ws, void* stream) {
  SN_REQUIRE(cell == SN_CELL_FACTORED || cell == SN_CELL_LSTM, "sn_recur_bwd: bad cell %d", cell);
  SN_REQUIRE(t0 >= 0 && t1 >= t0 && B > 0, "sn_recur_bwd: bad step range [%d,%d) or B", t0, t1);
  SN_REQUIRE(Whh && Call && gates && dHall && dZ && dh_carry && dc_carry && ws, "sn_recur_bwd: null argument");
  if (t1 == t0) return 0;
  Plan p;
  int32_t rc = make_plan(true, H, B, &p);
  if (rc) return rc;
  RecurArgs a = {};
  a.cell = cell; a.H = (int)H; a.B = (int)B; a.t0 = t0; a.t1 = t1; a.T = t1 + 1;
  a.bs = batch_sizes; a.off = offsets; a.W = Whh; a.c_init = c_init;
  a.Call = const_cast<float*>(Call); a.gates = const_cast<float*>(gates); a.dHall = dHall; a.dZ = dZ;
  a.dh_carry = dh_carry; a.dc_carry = dc_carry;
  a.flags = (int*)ws; a.n_ub = p.n_ub; a.nbb = p.nbb; a.BB = p.BB;
  cudaStream_t st = (cudaStream_t)stream;
  SN_CUDA(cudaMemsetAsync(ws, 0, sizeof(int) * (size_t)p.nbb * (size_t)(t1 + 1), st));
  switch (p.SW) {
    case 6: return launch_coop(recur_bwd_kernel<6>, p, a, st, "sn_recur_bwd");
    case 4: return launch_coop(recur_bwd_kernel<4>, p, a, st, "sn_recur_bwd");
    default: return launch_coop(recur_bwd_kernel<2>, p, a, st, "sn_recur_bwd");
  }
}

}  // extern "C"
