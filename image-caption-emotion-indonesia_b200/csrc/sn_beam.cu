// K8: one beam-search step for a batch of images, device-side bookkeeping (no host sync per step).
// Restates the per-step body of sample() (stylenet/model.py:232-285): log_softmax, add running scores,
// top-k over the flattened (live beams x V) candidates (row 0 only at step 1), split finished /
// unfinished beams, shrink k, and pick the best finished sequence at termination.  One CTA per image.
#include "sn_common.cuh"

namespace {

constexpr int NT = 512;
constexpr int KMAX = 8;

struct BeamArgs {
  const float* logits; int64_t ld; int V;
  int kmax, step, max_len, end_token, L;
  int* k_live; float* run_score; int* prev_word; int* src_row; int* cur_buf;
  int* seqs;        // [2][n_img*kmax][L]
  int* done_seq;    // [n_img*kmax][L]
  int* done_len; float* done_score; int* n_done;
  int* out_seq; int* out_len; int* n_unfinished;
  int n_img;
  const int* step_dev;   // optional: step number in device memory (CUDA-graph replay), overrides `step`
};

// bookkeeping of one image after its k winners are known (model.py:249-285); ONE thread
__device__ void beam_bookkeep(BeamArgs& a, int img, int k, const float* w_val, const int* w_idx) {
  const int row0 = img * a.kmax;
  const int V = a.V;
  // bookkeeping (model.py:249-285)
  const int L = a.L, cur = a.cur_buf[img], nxt = cur ^ 1;
  const int64_t plane = (int64_t)a.n_img * a.kmax * L;
  const int* sq_old = a.seqs + cur * plane + (int64_t)row0 * L;
  int* sq_new = a.seqs + nxt * plane + (int64_t)row0 * L;
  int nd = a.n_done[img];
  int new_k = 0;
  float ns[KMAX]; int nw[KMAX], nsrc[KMAX];
  const int len_old = a.step;   // tokens so far incl. <start>
  for (int j = 0; j < k; ++j) {
    const int flat = w_idx[j];
    const int src = flat / V, word = flat - src * V;
    const int* from = sq_old + (int64_t)src * L;
    if (word == a.end_token) {
      int* to = a.done_seq + (int64_t)(row0 + nd) * L;
      for (int q = 0; q < len_old; ++q) to[q] = from[q];
      to[len_old] = word;
      a.done_len[row0 + nd] = len_old + 1;
      a.done_score[row0 + nd] = w_val[j];
      ++nd;
    } else {
      int* to = sq_new + (int64_t)new_k * L;
      for (int q = 0; q < len_old; ++q) to[q] = from[q];
      to[len_old] = word;
      ns[new_k] = w_val[j]; nw[new_k] = word; nsrc[new_k] = row0 + src;
      ++new_k;
    }
  }
  for (int j = 0; j < a.kmax; ++j) {
    if (j < new_k) { a.run_score[row0 + j] = ns[j]; a.prev_word[row0 + j] = nw[j]; a.src_row[row0 + j] = nsrc[j]; }
    else { a.prev_word[row0 + j] = a.end_token; a.src_row[row0 + j] = row0 + j; }
  }
  a.n_done[img] = nd;
  a.cur_buf[img] = nxt;
  const bool finished = (new_k == 0) || (a.step > a.max_len);   // model.py:273,283
  if (finished) {
    int* out = a.out_seq + (int64_t)img * L;
    if (nd == 0) { out[0] = a.end_token; a.out_len[img] = 1; }   // model.py:288-289
    else {
      int best = 0;
      for (int j = 1; j < nd; ++j) if (a.done_score[row0 + j] > a.done_score[row0 + best]) best = j;   // first max
      const int* from = a.done_seq + (int64_t)(row0 + best) * L;
      const int n = a.done_len[row0 + best];
      for (int q = 0; q < n; ++q) out[q] = from[q];
      a.out_len[img] = n;
    }
    a.k_live[img] = 0;
    atomicSub(a.n_unfinished, 1);
  } else {
    a.k_live[img] = new_k;
  }
}

__global__ void __launch_bounds__(NT) beam_step_kernel(BeamArgs a) {
  __shared__ float s_mx[KMAX], s_ls[KMAX];
  __shared__ float red_v[NT / 32];
  __shared__ int red_i[NT / 32];
  __shared__ float c_val[NT * KMAX];
  __shared__ int c_idx[NT * KMAX];
  __shared__ float w_val[KMAX];
  __shared__ int w_idx[KMAX];
  const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (a.step_dev) a.step = *a.step_dev;
  const int k = a.k_live[img];
  const int row0 = img * a.kmax;
  if (k == 0) {
    if (tid < a.kmax) a.src_row[row0 + tid] = row0 + tid;
    return;
  }
  const int nrows = a.step == 1 ? 1 : k;   // model.py:239-241
  const int V = a.V;
  // per-row log-sum-exp
  for (int r = 0; r < nrows; ++r) {
    const float* x = a.logits + (int64_t)(row0 + r) * a.ld;
    float mx = -INFINITY;
    for (int v = tid; v < V; v += NT) mx = fmaxf(mx, x[v]);
    mx = sn::warp_max(mx);
    if (lane == 0) red_v[warp] = mx;
    __syncthreads();
    mx = red_v[0];
#pragma unroll
    for (int w = 1; w < NT / 32; ++w) mx = fmaxf(mx, red_v[w]);
    __syncthreads();
    float se = 0.f;
    for (int v = tid; v < V; v += NT) se += expf(x[v] - mx);
    se = sn::warp_sum(se);
    if (lane == 0) red_v[warp] = se;
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
      for (int w = 0; w < NT / 32; ++w) t += red_v[w];
      s_mx[r] = mx; s_ls[r] = logf(t);
    }
    __syncthreads();
  }
  // thread-local top-k (descending; ties keep the lower flat index)
  float tv[KMAX]; int ti[KMAX];
#pragma unroll
  for (int j = 0; j < KMAX; ++j) { tv[j] = -INFINITY; ti[j] = 0x7fffffff; }
  for (int r = 0; r < nrows; ++r) {
    const float* x = a.logits + (int64_t)(row0 + r) * a.ld;
    const float rs = a.run_score[row0 + r], mx = s_mx[r], ls = s_ls[r];
    for (int v = tid; v < V; v += NT) {
      float val = rs + ((x[v] - mx) - ls);
      int idx = r * V + v;
      if (val > tv[KMAX - 1] || (val == tv[KMAX - 1] && idx < ti[KMAX - 1])) {
        tv[KMAX - 1] = val; ti[KMAX - 1] = idx;
#pragma unroll
        for (int j = KMAX - 1; j > 0; --j) {
          bool sw = tv[j] > tv[j - 1] || (tv[j] == tv[j - 1] && ti[j] < ti[j - 1]);
          if (sw) { float fv = tv[j]; tv[j] = tv[j - 1]; tv[j - 1] = fv; int iv = ti[j]; ti[j] = ti[j - 1]; ti[j - 1] = iv; }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < KMAX; ++j) { c_val[tid * KMAX + j] = tv[j]; c_idx[tid * KMAX + j] = ti[j]; }
  __syncthreads();
  // k rounds of block arg-max over the NT*KMAX candidates
  for (int round = 0; round < k; ++round) {
    float bv = -INFINITY; int bi = 0x7fffffff; int bp = -1;
    for (int q = tid; q < NT * KMAX; q += NT) {
      float v = c_val[q]; int i = c_idx[q];
      if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; bp = q; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      int op = __shfl_xor_sync(0xffffffffu, bp, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; bp = op; }
    }
    if (lane == 0) { red_v[warp] = bv; red_i[warp] = bp; c_idx[0] = c_idx[0]; }
    __syncthreads();
    if (tid == 0) {
      float fv = red_v[0]; int fp = red_i[0];
      int fi = fp >= 0 ? c_idx[fp] : 0x7fffffff;
      for (int w = 1; w < NT / 32; ++w) {
        int p2 = red_i[w];
        if (p2 < 0) continue;
        float v2 = red_v[w]; int i2 = c_idx[p2];
        if (v2 > fv || (v2 == fv && i2 < fi)) { fv = v2; fi = i2; fp = p2; }
      }
      w_val[round] = fv; w_idx[round] = fi;
      if (fp >= 0) { c_val[fp] = -INFINITY; c_idx[fp] = 0x7fffffff; }
    }
    __syncthreads();
  }
  if (tid != 0) return;
  beam_bookkeep(a, img, k, w_val, w_idx);
}


// ---- the same step for FEW images, spread over NCH x more CTAs (single-image beam search is bound by this kernel when
// one CTA scans k x V logits alone): (A) per-(row, chunk) max / sum-exp, (B) per-(image, chunk) top-k of the final
// scores, (C) per-image merge of the NCH x k candidates + the bookkeeping above.
constexpr int NTS = 256;

struct SplitArgs {
  BeamArgs b;
  float* part;      // [rows][nch][2]   (max, sum exp(x - max)) per chunk
  float* cand_v;    // [n_img][nch][KMAX]
  int* cand_i;
  int nch;
  int* counter;     // optional: blocks-done counter; the last block of the finish kernel bumps *step_dev (saves a launch)
};

__device__ __forceinline__ void chunk_range(int V, int nch, int ch, int& c0, int& c1) {
  const int per = (V + nch - 1) / nch;
  c0 = ch * per;
  c1 = min(V, c0 + per);
}

__global__ void __launch_bounds__(NTS) beam_lse_partial_kernel(SplitArgs s) {
  BeamArgs& a = s.b;
  __shared__ float red[NTS / 32];
  const int ch = blockIdx.x, row = blockIdx.y, img = row / a.kmax, r = row - img * a.kmax;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int step = a.step_dev ? *a.step_dev : a.step;
  const int k = a.k_live[img];
  const int nrows = k == 0 ? 0 : (step == 1 ? 1 : k);
  if (r >= nrows) return;
  int c0, c1;
  chunk_range(a.V, s.nch, ch, c0, c1);
  const float* x = a.logits + (int64_t)row * a.ld;
  float mx = -INFINITY;
  for (int v = c0 + tid; v < c1; v += NTS) mx = fmaxf(mx, x[v]);
  mx = sn::warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int w = 1; w < NTS / 32; ++w) mx = fmaxf(mx, red[w]);
  __syncthreads();
  float se = 0.f;
  for (int v = c0 + tid; v < c1; v += NTS) se += expf(x[v] - mx);
  se = sn::warp_sum(se);
  if (lane == 0) red[warp] = se;
  __syncthreads();
  if (tid == 0) {
    float t = 0.f;
    for (int w = 0; w < NTS / 32; ++w) t += red[w];
    s.part[((int64_t)row * s.nch + ch) * 2] = mx;
    s.part[((int64_t)row * s.nch + ch) * 2 + 1] = t;
  }
}

__global__ void __launch_bounds__(NTS) beam_topk_partial_kernel(SplitArgs s) {
  BeamArgs& a = s.b;
  __shared__ float s_mx[KMAX], s_ls[KMAX];
  __shared__ float red_v[NTS / 32];
  __shared__ int red_i[NTS / 32];
  __shared__ float c_val[NTS * KMAX];
  __shared__ int c_idx[NTS * KMAX];
  const int ch = blockIdx.x, img = blockIdx.y;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int step = a.step_dev ? *a.step_dev : a.step;
  const int k = a.k_live[img];
  if (k == 0) return;
  const int row0 = img * a.kmax;
  const int nrows = step == 1 ? 1 : k;
  const int V = a.V;
  if (tid < nrows) {
    // every CTA of the image combines the chunk partials in the same order -> identical (mx, ls) everywhere
    const float* p = s.part + (int64_t)(row0 + tid) * s.nch * 2;
    float mx = -INFINITY;
    for (int c = 0; c < s.nch; ++c) mx = fmaxf(mx, p[2 * c]);
    float t = 0.f;
    for (int c = 0; c < s.nch; ++c) t += p[2 * c + 1] * expf(p[2 * c] - mx);
    s_mx[tid] = mx; s_ls[tid] = logf(t);
  }
  __syncthreads();
  int c0, c1;
  chunk_range(V, s.nch, ch, c0, c1);
  float tv[KMAX]; int ti[KMAX];
#pragma unroll
  for (int j = 0; j < KMAX; ++j) { tv[j] = -INFINITY; ti[j] = 0x7fffffff; }
  for (int r = 0; r < nrows; ++r) {
    const float* x = a.logits + (int64_t)(row0 + r) * a.ld;
    const float rs = a.run_score[row0 + r], mx = s_mx[r], ls = s_ls[r];
    for (int v = c0 + tid; v < c1; v += NTS) {
      float val = rs + ((x[v] - mx) - ls);
      int idx = r * V + v;
      if (val > tv[KMAX - 1] || (val == tv[KMAX - 1] && idx < ti[KMAX - 1])) {
        tv[KMAX - 1] = val; ti[KMAX - 1] = idx;
#pragma unroll
        for (int j = KMAX - 1; j > 0; --j) {
          bool sw = tv[j] > tv[j - 1] || (tv[j] == tv[j - 1] && ti[j] < ti[j - 1]);
          if (sw) { float fv = tv[j]; tv[j] = tv[j - 1]; tv[j - 1] = fv; int iv = ti[j]; ti[j] = ti[j - 1]; ti[j - 1] = iv; }
        }
      }
    }
  }
#pragma unroll
  for (int j = 0; j < KMAX; ++j) { c_val[tid * KMAX + j] = tv[j]; c_idx[tid * KMAX + j] = ti[j]; }
  __syncthreads();
  float* out_v = s.cand_v + ((int64_t)img * s.nch + ch) * KMAX;
  int* out_i = s.cand_i + ((int64_t)img * s.nch + ch) * KMAX;
  for (int round = 0; round < KMAX; ++round) {
    if (round >= k) {
      if (tid == 0) { out_v[round] = -INFINITY; out_i[round] = 0x7fffffff; }
      continue;
    }
    float bv = -INFINITY; int bi = 0x7fffffff; int bp = -1;
    for (int q = tid; q < NTS * KMAX; q += NTS) {
      float v = c_val[q]; int i = c_idx[q];
      if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; bp = q; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      int op = __shfl_xor_sync(0xffffffffu, bp, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; bp = op; }
    }
    if (lane == 0) { red_v[warp] = bv; red_i[warp] = bp; }
    __syncthreads();
    if (tid == 0) {
      float fv = red_v[0]; int fp = red_i[0];
      int fi = fp >= 0 ? c_idx[fp] : 0x7fffffff;
      for (int w = 1; w < NTS / 32; ++w) {
        int p2 = red_i[w];
        if (p2 < 0) continue;
        float v2 = red_v[w]; int i2 = c_idx[p2];
        if (v2 > fv || (v2 == fv && i2 < fi)) { fv = v2; fi = i2; fp = p2; }
      }
      out_v[round] = fv; out_i[round] = fi;
      if (fp >= 0) { c_val[fp] = -INFINITY; c_idx[fp] = 0x7fffffff; }
    }
    __syncthreads();
  }
}

__device__ __forceinline__ void finish_count(SplitArgs& s) {
  if (s.counter == nullptr) return;
  __threadfence();
  if (atomicAdd(s.counter, 1) == (int)gridDim.x - 1) {       // every block has read the step number: advance it
    *s.counter = 0;
    *const_cast<int*>(s.b.step_dev) += 1;
  }
}

__global__ void __launch_bounds__(32) beam_finish_kernel(SplitArgs s) {
  BeamArgs& a = s.b;
  __shared__ float w_val[KMAX];
  __shared__ int w_idx[KMAX];
  __shared__ float cv[32 * KMAX];
  __shared__ int ci[32 * KMAX];
  const int img = blockIdx.x, lane = threadIdx.x;
  if (a.step_dev) a.step = *a.step_dev;
  const int k = a.k_live[img];
  const int row0 = img * a.kmax;
  if (k == 0) {
    if (lane < a.kmax) a.src_row[row0 + lane] = row0 + lane;
    if (lane == 0) finish_count(s);
    return;
  }
  const int ncand = s.nch * KMAX;                 // <= 32 * KMAX
  for (int q = lane; q < 32 * KMAX; q += 32) {
    cv[q] = q < ncand ? s.cand_v[(int64_t)img * ncand + q] : -INFINITY;
    ci[q] = q < ncand ? s.cand_i[(int64_t)img * ncand + q] : 0x7fffffff;
  }
  __syncwarp();
  for (int round = 0; round < k; ++round) {
    float bv = -INFINITY; int bi = 0x7fffffff; int bp = -1;
    for (int q = lane; q < 32 * KMAX; q += 32) {
      float v = cv[q]; int i = ci[q];
      if (v > bv || (v == bv && i < bi)) { bv = v; bi = i; bp = q; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      int op = __shfl_xor_sync(0xffffffffu, bp, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; bp = op; }
    }
    if (lane == 0) {
      w_val[round] = bv; w_idx[round] = bi;
      if (bp >= 0) { cv[bp] = -INFINITY; ci[bp] = 0x7fffffff; }
    }
    __syncwarp();
  }
  if (lane != 0) return;
  beam_bookkeep(a, img, k, w_val, w_idx);
  finish_count(s);
}

}  // namespace

extern "C" int32_t sn_beam_step(const float* logits, int64_t ld, int64_t V, int32_t n_img, int32_t kmax, int32_t step,
                                int32_t max_len, int32_t end_token, int32_t* k_live, float* run_score,
                                int32_t* prev_word, int32_t* src_row, int32_t* cur_buf, int32_t* seqs,
                                int32_t* done_seq, int32_t* done_len, float* done_score, int32_t* n_done,
                                int32_t* out_seq, int32_t* out_len, int32_t* n_unfinished, const int32_t* step_dev,
                                void* stream) {
  SN_REQUIRE(kmax >= 1 && kmax <= KMAX, "sn_beam_step: beam width %d not in [1,%d]", kmax, KMAX);
  SN_REQUIRE(n_img >= 0 && V > 0 && step >= 1, "sn_beam_step: bad dims");
  SN_REQUIRE((int64_t)kmax * V < 0x7fffffff, "sn_beam_step: k*V overflows int32");
  if (n_img == 0) return 0;
  BeamArgs a;
  a.logits = logits; a.ld = ld; a.V = (int)V; a.kmax = kmax; a.step = step; a.max_len = max_len;
  a.end_token = end_token; a.L = max_len + 2;
  a.k_live = k_live; a.run_score = run_score; a.prev_word = prev_word; a.src_row = src_row; a.cur_buf = cur_buf;
  a.seqs = seqs; a.done_seq = done_seq; a.done_len = done_len; a.done_score = done_score; a.n_done = n_done;
  a.out_seq = out_seq; a.out_len = out_len; a.n_unfinished = n_unfinished; a.n_img = n_img;
  a.step_dev = step_dev;
  beam_step_kernel<<<(unsigned)n_img, NT, 0, (cudaStream_t)stream>>>(a);
  return sn::check_launch("sn_beam_step");
}

extern "C" int64_t sn_beam_split_ws_floats(int32_t n_img, int32_t kmax, int32_t nch) {
  // part [rows][nch][2] + cand_v [n_img][nch][KMAX] + cand_i (int32, same count)
  return (int64_t)n_img * kmax * nch * 2 + 2 * (int64_t)n_img * nch * KMAX + 4;      // + the blocks-done counter
}

extern "C" int32_t sn_beam_step_split(const float* logits, int64_t ld, int64_t V, int32_t n_img, int32_t kmax, int32_t step,
                                      int32_t max_len, int32_t end_token, int32_t* k_live, float* run_score,
                                      int32_t* prev_word, int32_t* src_row, int32_t* cur_buf, int32_t* seqs,
                                      int32_t* done_seq, int32_t* done_len, float* done_score, int32_t* n_done,
                                      int32_t* out_seq, int32_t* out_len, int32_t* n_unfinished, const int32_t* step_dev,
                                      int32_t nch, float* ws, int32_t advance_step, void* stream) {
  SN_REQUIRE(kmax >= 1 && kmax <= KMAX, "sn_beam_step_split: beam width %d not in [1,%d]", kmax, KMAX);
  SN_REQUIRE(n_img >= 0 && V > 0 && step >= 1 && ws, "sn_beam_step_split: bad argument");
  SN_REQUIRE(nch >= 1 && nch <= 32, "sn_beam_step_split: 1..32 chunks");
  SN_REQUIRE((int64_t)kmax * V < 0x7fffffff, "sn_beam_step_split: k*V overflows int32");
  if (n_img == 0) return 0;
  SplitArgs s;
  BeamArgs& a = s.b;
  a.logits = logits; a.ld = ld; a.V = (int)V; a.kmax = kmax; a.step = step; a.max_len = max_len;
  a.end_token = end_token; a.L = max_len + 2;
  a.k_live = k_live; a.run_score = run_score; a.prev_word = prev_word; a.src_row = src_row; a.cur_buf = cur_buf;
  a.seqs = seqs; a.done_seq = done_seq; a.done_len = done_len; a.done_score = done_score; a.n_done = n_done;
  a.out_seq = out_seq; a.out_len = out_len; a.n_unfinished = n_unfinished; a.n_img = n_img;
  a.step_dev = step_dev;
  s.nch = nch;
  s.part = ws;
  s.cand_v = ws + (int64_t)n_img * kmax * nch * 2;
  s.cand_i = reinterpret_cast<int*>(s.cand_v + (int64_t)n_img * nch * KMAX);
  // advance_step: the finish kernel adds 1 to *step_dev once every image is done with this step (the counter word must
  // be zero on the first call: the caller zero-fills the work space once)
  s.counter = (advance_step && step_dev) ? s.cand_i + (int64_t)n_img * nch * KMAX : nullptr;
  cudaStream_t st = (cudaStream_t)stream;
  beam_lse_partial_kernel<<<dim3((unsigned)nch, (unsigned)(n_img * kmax)), NTS, 0, st>>>(s);
  beam_topk_partial_kernel<<<dim3((unsigned)nch, (unsigned)n_img), NTS, 0, st>>>(s);
  beam_finish_kernel<<<(unsigned)n_img, 32, 0, st>>>(s);
  return sn::check_launch("sn_beam_step_split");
}
