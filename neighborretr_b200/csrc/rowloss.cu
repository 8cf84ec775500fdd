// Row-block loss kernels of the retrieval head: centrality-weighted InfoNCE, neighbour-adjusting
// loss (per-row top-k selection), KL consistency and the uniform-regularisation cross-entropy,
// forward and backward.  One CTA owns one row of the [rows, B] similarity block, stages it in
// shared memory once and makes every pass (max, log-sum-exp, top-k, min/max, losses) out of it.
//
// Reference semantics (paths relative to the reference repo):
//   CentralityWeightingLoss   NeighborRetr/models/until_module.py:303-328
//   NeighborAdjustingLoss     NeighborRetr/models/until_module.py:56-211
//   UniformRegularizationLoss NeighborRetr/models/until_module.py:263-291 (CE part; Sinkhorn in sinkhorn.cu)
//   KLDivergenceLoss          NeighborRetr/models/until_module.py:339-359
// Roofline: HBM/L2-bound, 4*B bytes per row and matrix read once.
#include <stdlib.h>
#include "common.cuh"
#include "nrhead_internal.h"

namespace nr {

// threads of a CTA-per-row kernel.  NR_ROW_THREADS=1024 selects the 1024-thread instantiation for rows beyond 2048
// columns (measured on B200 at B = 8192: 2052 vs 1637 us forward, 674 vs 595 us backward — slower, so 256 stays)
constexpr int ROW_THREADS_SMALL = 256;
constexpr int ROW_THREADS_BIG = 1024;
static bool row_big(int64_t B) {
  const char* e = getenv("NR_ROW_THREADS");
  return B > 2048 && e && atoi(e) == 1024;
}

struct RowArgs {
  const float* X; int64_t ldx;     // [rows, B] local similarity rows
  const float* G; int64_t ldg;     // [rows, B] global similarity rows
  const float* cbank;              // [B] bank centrality (by column)
  const float* w;                  // [rows] centrality weights
  const float* sk_u;               // [rows] Sinkhorn row duals of these rows
  const float* sk_v;               // [B] Sinkhorn column duals
  int rows, B, row0;
  const float* logit_scale;        // device scalar (exp'd); nullptr => 1
  int k; float tau_nbr, tau_uni, beta; int flags;
};

__device__ __forceinline__ float load_ls(const float* p) { return p ? __ldg(p) : 1.0f; }

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
template <int ROW_THREADS>
__global__ void __launch_bounds__(ROW_THREADS)
row_losses_fwd_kernel(RowArgs a, float* __restrict__ row_out, int32_t* __restrict__ nbr_idx,
                      float* __restrict__ saved) {
  extern __shared__ __align__(16) float sm[];
  __shared__ float red[32];
  __shared__ uint2 redk[32];
  __shared__ float nb_x[NR_MAX_K], nb_a[NR_MAX_K];
  __shared__ int nb_j[NR_MAX_K];
  const int B = a.B, tid = threadIdx.x, i = blockIdx.x, gi = a.row0 + i;
  float* x = sm;
  float* xs = sm + B;
  float* g = sm + 2 * B;
  const bool need_g = (a.flags & (NR_LOSS_KL | NR_LOSS_UNIFORM)) != 0;
  const float* Xr = a.X + (int64_t)i * a.ldx;
  const float* Gr = need_g ? a.G + (int64_t)i * a.ldg : nullptr;
  stage_row<ROW_THREADS>(Xr, x, B, tid);
  if (need_g) stage_row<ROW_THREADS>(Gr, g, B, tid);
  __syncthreads();
  if (a.flags & NR_LOSS_NEIGHBOR) {          // selection copy: the row without its diagonal
    for (int j = tid; j < B; j += ROW_THREADS) xs[j] = (j == gi) ? NR_NEG_INF : x[j];
    __syncthreads();
  }
  float* sv = saved + (int64_t)i * NR_NSAVE;
  float* ro = row_out;
  const int rows = a.rows;
  const float ls = load_ls(a.logit_scale);

  // ---- maxima
  float mx = NR_NEG_INF, mg = NR_NEG_INF;
  for (int j = tid; j < B; j += ROW_THREADS) {
    mx = fmaxf(mx, x[j]);
    if (need_g) mg = fmaxf(mg, g[j]);
  }
  mx = block_max(mx, red);
  if (need_g) mg = block_max(mg, red);

  // ---- log-sum-exps
  float s_c = 0.f, s_x = 0.f, s_g = 0.f, s_tg = 0.f;
  for (int j = tid; j < B; j += ROW_THREADS) {
    float dx = x[j] - mx;
    s_c += expf(ls * dx);
    s_x += expf(dx);
    if (need_g) {
      float dg = g[j] - mg;
      s_g += expf(dg);
      s_tg += expf(a.tau_uni * dg);
    }
  }
  s_c = block_sum(s_c, red);
  s_x = block_sum(s_x, red);
  float lse_c = ls * mx + logf(s_c);      // LSE_b(ls * x_b)   (ls > 0)
  float lse_x = mx + logf(s_x);
  float lse_g = 0.f, lse_tg = 0.f;
  if (need_g) {
    s_g = block_sum(s_g, red);
    s_tg = block_sum(s_tg, red);
    lse_g = mg + logf(s_g);
    lse_tg = a.tau_uni * mg + logf(s_tg);
  }

  // ---- centrality-weighted InfoNCE row  (until_module.py:315-324): -w * log_softmax(ls*x)[diag]
  if (a.flags & NR_LOSS_CENTRALITY) {
    if (tid == 0) {
      float wv = a.w ? a.w[i] : 1.f;
      ro[0 * rows + i] = -wv * (ls * x[gi] - lse_c);
      sv[0] = lse_c;
    }
  }

  // ---- KL row (until_module.py:351-357): sum_b p_b (log p_b - log q_b), p = softmax(x), q = softmax(g)
  float klrow = 0.f, lurow = 0.f, sumT = 0.f;
  if (need_g) {
    const float nu = -logf(2.0f * (float)B);          // Sinkhorn norm = -log(m+n), until_module.py:238
    const float ui = (a.flags & NR_LOSS_UNIFORM) ? a.sk_u[i] : 0.f;
    for (int j = tid; j < B; j += ROW_THREADS) {
      float lp = x[j] - lse_x, lq = g[j] - lse_g;
      klrow += expf(lp) * (lp - lq);
      if (a.flags & NR_LOSS_UNIFORM) {
        // target T = beta*Q + (1-beta)*I, Q = exp(G + u + v - norm)   (until_module.py:253-266)
        float t = a.beta * expf(g[j] + ui + a.sk_v[j] - nu) + ((j == gi) ? (1.f - a.beta) : 0.f);
        sumT += t;
        lurow -= t * (a.tau_uni * g[j] - lse_tg);
      }
    }
    klrow = block_sum(klrow, red);
    if (a.flags & NR_LOSS_UNIFORM) {
      lurow = block_sum(lurow, red);
      sumT = block_sum(sumT, red);
    }
    if (tid == 0) {
      if (a.flags & NR_LOSS_KL) ro[2 * rows + i] = klrow;
      if (a.flags & NR_LOSS_UNIFORM) ro[3 * rows + i] = lurow;
      sv[1] = lse_x; sv[2] = lse_g; sv[3] = klrow; sv[4] = lse_tg; sv[5] = sumT;
    }
  }

  // ---- neighbour-adjusting row (until_module.py:161-211)
  if (a.flags & NR_LOSS_NEIGHBOR) {
    const int k = a.k;
    // top-k of the off-diagonal entries: k rounds of block arg-max, ties -> lower column.  Every thread keeps the best
    // (ord, column) of ITS columns in registers; a round is one block arg-max of those (redux.sync, two barriers),
    // and only the warp of the thread that owned the winner rescans that thread's B/256 columns, one per lane.
    // (Before: every thread rescanned the whole staged row with 64-bit keys in every round; ncu at B = 8192: 71 k
    // instructions per row, 41 % of the stall samples at the barrier behind a single rescanning thread.)
    const int lane = tid & 31, warp = tid >> 5;
    auto scan_cols = [&](int j0, int step, uint32_t& bo, uint32_t& bj) {
      bo = 0u; bj = NR_NO_INDEX;
      for (int j = j0; j < B; j += step) {                  // ascending columns + strict compare: ties -> lower column
        const float v = xs[j];
        if (v != NR_NEG_INF) {
          const uint32_t o = float_ord(v);
          if (o > bo) { bo = o; bj = (uint32_t)j; }
        }
      }
    };
    uint32_t bo, bj;
    scan_cols(tid, ROW_THREADS, bo, bj);
    for (int r = 0; r < k; ++r) {
      uint32_t mo = bo, mj = bj;
      block_argmax_ord(mo, mj, redk);
      const int jsel = (int)mj;
      if (tid == 0) {
        nb_j[r] = jsel;
        nb_x[r] = x[jsel];
        nbr_idx[(int64_t)i * k + r] = jsel;
      }
      const int t_own = jsel % ROW_THREADS;
      if (warp == (t_own >> 5)) {                           // xs[jsel] is only re-read by this warp until the barrier below
        if (tid == t_own) xs[jsel] = NR_NEG_INF;
        __syncwarp();
        uint32_t so, sj;
        scan_cols(t_own + lane * ROW_THREADS, 32 * ROW_THREADS, so, sj);
        warp_argmax_ord(so, sj);
        if (tid == t_own) { bo = so; bj = sj; }
      }
    }
    __syncthreads();
    // min / max over the NON-extended entries of x and of the bank centrality c (until_module.py:77-85)
    uint32_t xlo_o = 0u, xlo_j = NR_NO_INDEX, xhi_o = 0u, xhi_j = NR_NO_INDEX;
    uint32_t clo_o = 0u, clo_j = NR_NO_INDEX, chi_o = 0u, chi_j = NR_NO_INDEX;
    for (int j = tid; j < B; j += ROW_THREADS) {
      if (xs[j] != NR_NEG_INF) {
        const uint32_t ov = float_ord(x[j]), oc = float_ord(a.cbank[j]);
        if (~ov > xlo_o) { xlo_o = ~ov; xlo_j = (uint32_t)j; }
        if (ov > xhi_o) { xhi_o = ov; xhi_j = (uint32_t)j; }
        if (~oc > clo_o) { clo_o = ~oc; clo_j = (uint32_t)j; }
        if (oc > chi_o) { chi_o = oc; chi_j = (uint32_t)j; }
      }
    }
    block_argmax_ord(xlo_o, xlo_j, redk);
    block_argmax_ord(xhi_o, xhi_j, redk);
    block_argmax_ord(clo_o, clo_j, redk);
    block_argmax_ord(chi_o, chi_j, redk);
    const float lo_x = ord_float(~xlo_o), hi_x = ord_float(xhi_o);
    const float lo_c = ord_float(~clo_o), hi_c = ord_float(chi_o);
    if (tid < k) {
      int j = nb_j[tid];
      float nx = (nb_x[tid] - lo_x) / (hi_x - lo_x);
      float nc = (a.cbank[j] - lo_c) / (hi_c - lo_c);
      nb_a[tid] = nx - nc;                                   // Eq. 5: de-centrality similarity
    }
    __syncthreads();
    if (tid == 0) {
      float xd = x[gi];
      float ma = NR_NEG_INF, mext = xd;
      for (int r = 0; r < k; ++r) { ma = fmaxf(ma, a.tau_nbr * nb_a[r]); mext = fmaxf(mext, nb_x[r]); }
      float sa = 0.f, sext = expf(xd - mext);
      for (int r = 0; r < k; ++r) { sa += expf(a.tau_nbr * nb_a[r] - ma); sext += expf(nb_x[r] - mext); }
      float lse_ext = mext + logf(sext);
      float num = xd - lse_ext, den = 1.f;                   // diagonal weight 1 (until_module.py:157)
      for (int r = 0; r < k; ++r) {
        float p = expf(a.tau_nbr * nb_a[r] - ma) / sa;
        num += p * (nb_x[r] - lse_ext);
        den += p;
      }
      ro[1 * rows + i] = -num / den;
      sv[6] = lo_x; sv[7] = hi_x; sv[8] = lo_c; sv[9] = hi_c; sv[10] = lse_ext; sv[11] = den;
      sv[12] = __int_as_float((int)xlo_j);
      sv[13] = __int_as_float((int)xhi_j);
      sv[14] = __int_as_float((int)clo_j);
      sv[15] = __int_as_float((int)chi_j);
    }
  }
}

// ------------------------------------------------------------------------------------------
// forward, warp-per-row variant (B <= 32*VPL <= 1024): the row of X (and of G) lives in the registers of ONE warp,
// VPL values per lane (column j = lane + 32*q), so every reduction is a shuffle tree and the k top-k rounds need
// no block barrier at all; 4 rows (warps) per CTA.  Same arithmetic, same tie rules, same saved scalars as
// row_losses_fwd_kernel (sums are taken in a different order: results agree to fp32 rounding).
// ------------------------------------------------------------------------------------------
constexpr int ROWW_WARPS = 4;

template <int VPL>
__global__ void __launch_bounds__(ROWW_WARPS * 32)
row_losses_fwd_warp_kernel(RowArgs a, float* __restrict__ row_out, int32_t* __restrict__ nbr_idx,
                           float* __restrict__ saved) {
  __shared__ float nb_x_s[ROWW_WARPS][NR_MAX_K], nb_a_s[ROWW_WARPS][NR_MAX_K];
  __shared__ int nb_j_s[ROWW_WARPS][NR_MAX_K];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * ROWW_WARPS + warp;
  if (i >= a.rows) return;                             // whole warp leaves together; no block-level barriers below
  const int B = a.B, gi = a.row0 + i, rows = a.rows;
  float* nb_x = nb_x_s[warp]; float* nb_a = nb_a_s[warp]; int* nb_j = nb_j_s[warp];
  const bool need_g = (a.flags & (NR_LOSS_KL | NR_LOSS_UNIFORM)) != 0;
  const float* Xr = a.X + (int64_t)i * a.ldx;
  const float* Gr = need_g ? a.G + (int64_t)i * a.ldg : nullptr;
  float x[VPL], g[VPL];
#pragma unroll
  for (int q = 0; q < VPL; ++q) {
    const int j = lane + 32 * q;
    x[q] = j < B ? Xr[j] : NR_NEG_INF;
    g[q] = (need_g && j < B) ? Gr[j] : NR_NEG_INF;
  }
  float* sv = saved + (int64_t)i * NR_NSAVE;
  float* ro = row_out;
  const float ls = load_ls(a.logit_scale);
  float xdiag = 0.f;                                   // x[gi]: the lane holding it broadcasts (q unrolled: no dynamic indexing)
#pragma unroll
  for (int q = 0; q < VPL; ++q) {
    const float t = __shfl_sync(0xffffffffu, x[q], gi & 31);
    if (q == (gi >> 5)) xdiag = t;
  }
  // ---- maxima and log-sum-exps
  float mx = NR_NEG_INF, mg = NR_NEG_INF;
#pragma unroll
  for (int q = 0; q < VPL; ++q) { mx = fmaxf(mx, x[q]); mg = fmaxf(mg, g[q]); }
  mx = warp_max(mx);
  mg = warp_max(mg);
  float s_c = 0.f, s_x = 0.f, s_g = 0.f, s_tg = 0.f;
#pragma unroll
  for (int q = 0; q < VPL; ++q) {
    if (lane + 32 * q < B) {
      const float dx = x[q] - mx;
      s_c += expf(ls * dx);
      s_x += expf(dx);
      if (need_g) {
        const float dg = g[q] - mg;
        s_g += expf(dg);
        s_tg += expf(a.tau_uni * dg);
      }
    }
  }
  s_c = warp_sum(s_c); s_x = warp_sum(s_x);
  const float lse_c = ls * mx + logf(s_c);
  const float lse_x = mx + logf(s_x);
  float lse_g = 0.f, lse_tg = 0.f;
  if (need_g) {
    s_g = warp_sum(s_g); s_tg = warp_sum(s_tg);
    lse_g = mg + logf(s_g);
    lse_tg = a.tau_uni * mg + logf(s_tg);
  }
  // ---- centrality-weighted InfoNCE row (until_module.py:315-324)
  if ((a.flags & NR_LOSS_CENTRALITY) && lane == 0) {
    const float wv = a.w ? a.w[i] : 1.f;
    ro[0 * rows + i] = -wv * (ls * xdiag - lse_c);
    sv[0] = lse_c;
  }
  // ---- KL row (until_module.py:351-357) and uniform cross-entropy with the Sinkhorn target (:253-291)
  if (need_g) {
    float klrow = 0.f, lurow = 0.f, sumT = 0.f;
    const float nu = -logf(2.0f * (float)B);
    const float ui = (a.flags & NR_LOSS_UNIFORM) ? a.sk_u[i] : 0.f;
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
      const int j = lane + 32 * q;
      if (j < B) {
        const float lp = x[q] - lse_x, lq = g[q] - lse_g;
        klrow += expf(lp) * (lp - lq);
        if (a.flags & NR_LOSS_UNIFORM) {
          const float t = a.beta * expf(g[q] + ui + a.sk_v[j] - nu) + ((j == gi) ? (1.f - a.beta) : 0.f);
          sumT += t;
          lurow -= t * (a.tau_uni * g[q] - lse_tg);
        }
      }
    }
    klrow = warp_sum(klrow);
    if (a.flags & NR_LOSS_UNIFORM) { lurow = warp_sum(lurow); sumT = warp_sum(sumT); }
    if (lane == 0) {
      if (a.flags & NR_LOSS_KL) ro[2 * rows + i] = klrow;
      if (a.flags & NR_LOSS_UNIFORM) ro[3 * rows + i] = lurow;
      sv[1] = lse_x; sv[2] = lse_g; sv[3] = klrow; sv[4] = lse_tg; sv[5] = sumT;
    }
  }
  // ---- neighbour-adjusting row (until_module.py:161-211)
  if (a.flags & NR_LOSS_NEIGHBOR) {
    const int k = a.k;
    uint32_t alive = 0;                                  // bit q: column lane+32q is still a candidate
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
      const int j = lane + 32 * q;
      if (j < B && j != gi) alive |= 1u << q;
    }
    // top-k of the off-diagonal entries: k rounds of warp arg-max (redux.sync on the order-preserving integer image
    // of the value, then on the column among the lanes that hold the maximum), ties -> lower column
    uint32_t xo[VPL];
#pragma unroll
    for (int q = 0; q < VPL; ++q) xo[q] = float_ord(x[q]);
    for (int r = 0; r < k; ++r) {
      uint32_t bo = 0u, bj = NR_NO_INDEX;
#pragma unroll
      for (int q = 0; q < VPL; ++q) {
        if (((alive >> q) & 1u) && xo[q] > bo) { bo = xo[q]; bj = (uint32_t)(lane + 32 * q); }
      }
      warp_argmax_ord(bo, bj);
      const int jsel = (int)bj;
      if ((jsel & 31) == lane) alive &= ~(1u << (jsel >> 5));
      if (lane == 0) {
        nb_j[r] = jsel;
        nb_x[r] = ord_float(bo);
        nbr_idx[(int64_t)i * k + r] = jsel;
      }
    }
    // min / max over the NON-extended entries of x and of the bank centrality c (until_module.py:77-85)
    uint32_t xlo_o = 0u, xlo_j = NR_NO_INDEX, xhi_o = 0u, xhi_j = NR_NO_INDEX;
    uint32_t clo_o = 0u, clo_j = NR_NO_INDEX, chi_o = 0u, chi_j = NR_NO_INDEX;
#pragma unroll
    for (int q = 0; q < VPL; ++q) {
      if ((alive >> q) & 1u) {
        const uint32_t j = (uint32_t)(lane + 32 * q);
        const uint32_t ov = xo[q], oc = float_ord(a.cbank[j]);
        if (~ov > xlo_o) { xlo_o = ~ov; xlo_j = j; }
        if (ov > xhi_o) { xhi_o = ov; xhi_j = j; }
        if (~oc > clo_o) { clo_o = ~oc; clo_j = j; }
        if (oc > chi_o) { chi_o = oc; chi_j = j; }
      }
    }
    warp_argmax_ord(xlo_o, xlo_j); warp_argmax_ord(xhi_o, xhi_j);
    warp_argmax_ord(clo_o, clo_j); warp_argmax_ord(chi_o, chi_j);
    const float lo_x = ord_float(~xlo_o), hi_x = ord_float(xhi_o);
    const float lo_c = ord_float(~clo_o), hi_c = ord_float(chi_o);
    __syncwarp();
    for (int r = lane; r < k; r += 32) {
      const int j = nb_j[r];
      const float nx = (nb_x[r] - lo_x) / (hi_x - lo_x);
      const float nc = (a.cbank[j] - lo_c) / (hi_c - lo_c);
      nb_a[r] = nx - nc;                                     // Eq. 5: de-centrality similarity
    }
    __syncwarp();
    // softmax over the neighbours and the extended-set log-softmax, one neighbour per lane (strided for k > 32)
    float ma = NR_NEG_INF, mext = xdiag;
    for (int r = lane; r < k; r += 32) { ma = fmaxf(ma, a.tau_nbr * nb_a[r]); mext = fmaxf(mext, nb_x[r]); }
    ma = warp_max(ma); mext = warp_max(mext);
    float sa = 0.f, sext = 0.f;
    for (int r = lane; r < k; r += 32) { sa += expf(a.tau_nbr * nb_a[r] - ma); sext += expf(nb_x[r] - mext); }
    sa = warp_sum(sa);
    sext = warp_sum(sext) + expf(xdiag - mext);
    const float lse_ext = mext + logf(sext);
    float num = 0.f, den = 0.f;
    for (int r = lane; r < k; r += 32) {
      const float p = expf(a.tau_nbr * nb_a[r] - ma) / sa;
      num += p * (nb_x[r] - lse_ext);
      den += p;
    }
    num = warp_sum(num) + (xdiag - lse_ext);               // diagonal weight 1 (until_module.py:157)
    den = warp_sum(den) + 1.f;
    if (lane == 0) {
      ro[1 * rows + i] = -num / den;
      sv[6] = lo_x; sv[7] = hi_x; sv[8] = lo_c; sv[9] = hi_c; sv[10] = lse_ext; sv[11] = den;
      sv[12] = __int_as_float((int)xlo_j);
      sv[13] = __int_as_float((int)xhi_j);
      sv[14] = __int_as_float((int)clo_j);
      sv[15] = __int_as_float((int)chi_j);
    }
  }
}

// ------------------------------------------------------------------------------------------
// backward: dX (dense row), dG (dense row), dc (atomic over rows), dw, d logit_scale
// gscale[4] = upstream multipliers of the per-row terms {centrality, neighbour, kl, uniform}
// ------------------------------------------------------------------------------------------
template <int ROW_THREADS>
__global__ void __launch_bounds__(ROW_THREADS)
row_losses_bwd_kernel(RowArgs a, const int32_t* __restrict__ nbr_idx, const float* __restrict__ saved,
                      const float* __restrict__ gscale, float* __restrict__ dX, int64_t lddx,
                      float* __restrict__ dG, int64_t lddg, float* __restrict__ dc,
                      float* __restrict__ dw, float* __restrict__ dls) {
  extern __shared__ __align__(16) float sm[];
  __shared__ float red[32];
  const int B = a.B, tid = threadIdx.x, i = blockIdx.x, gi = a.row0 + i;
  float* x = sm;
  float* dx = sm + B;
  float* g = sm + 2 * B;
  const bool need_g = (a.flags & (NR_LOSS_KL | NR_LOSS_UNIFORM)) != 0;
  const float* Xr = a.X + (int64_t)i * a.ldx;
  const float* Gr = need_g ? a.G + (int64_t)i * a.ldg : nullptr;
  const float* sv = saved + (int64_t)i * NR_NSAVE;
  const float ls = load_ls(a.logit_scale);
  const float gs_c = (a.flags & NR_LOSS_CENTRALITY) ? gscale[0] : 0.f;
  const float gs_n = (a.flags & NR_LOSS_NEIGHBOR) ? gscale[1] : 0.f;
  const float gs_k = (a.flags & NR_LOSS_KL) ? gscale[2] : 0.f;
  const float gs_u = (a.flags & NR_LOSS_UNIFORM) ? gscale[3] : 0.f;
  stage_row<ROW_THREADS>(Xr, x, B, tid);
  if (need_g) stage_row<ROW_THREADS>(Gr, g, B, tid);
  __syncthreads();
  const float lse_c = sv[0], lse_x = sv[1], lse_g = sv[2], klrow = sv[3], lse_tg = sv[4], sumT = sv[5];
  const float wv = (a.flags & NR_LOSS_CENTRALITY) ? (a.w ? a.w[i] : 1.f) : 0.f;
  const float nu = -logf(2.0f * (float)B);
  const float ui = (a.flags & NR_LOSS_UNIFORM) ? a.sk_u[i] : 0.f;
  float* dGr = (need_g && dG) ? dG + (int64_t)i * lddg : nullptr;

  float sxq = 0.f;   // sum_b softmax(ls x)_b * x_b   (for d logit_scale)
  for (int j = tid; j < B; j += ROW_THREADS) {
    float acc = 0.f;
    if (a.flags & NR_LOSS_CENTRALITY) {
      float q = expf(ls * x[j] - lse_c);
      acc += gs_c * wv * ls * (q - ((j == gi) ? 1.f : 0.f));
      sxq += q * x[j];
    }
    if (need_g) {
      float lp = x[j] - lse_x, lq = g[j] - lse_g;
      float p = expf(lp), q = expf(lq);
      float dg = 0.f;
      if (a.flags & NR_LOSS_KL) {
        acc += gs_k * p * ((lp - lq) - klrow);
        dg += gs_k * (q - p);
      }
      if (a.flags & NR_LOSS_UNIFORM) {
        float t = a.beta * expf(g[j] + ui + a.sk_v[j] - nu) + ((j == gi) ? (1.f - a.beta) : 0.f);
        float qt = expf(a.tau_uni * g[j] - lse_tg);
        dg += gs_u * a.tau_uni * (qt * sumT - t);
      }
      if (dGr) dGr[j] = dg;
    }
    dx[j] = acc;
  }
  if (a.flags & NR_LOSS_CENTRALITY) {
    sxq = block_sum(sxq, red);
    if (tid == 0) {
      if (dw) dw[i] = -gs_c * (ls * x[gi] - lse_c);
      if (dls) atomicAdd(dls, -gs_c * wv * (x[gi] - sxq));
    }
  }
  __syncthreads();

  if ((a.flags & NR_LOSS_NEIGHBOR) && tid < 32) {
    // sparse part on one warp (lane strides over the k neighbours): k+1 extended entries plus the arg-min /
    // arg-max of the non-neighbours
    const int k = a.k, lane = tid;
    const int32_t* nb = nbr_idx + (int64_t)i * k;
    const float lo_x = sv[6], hi_x = sv[7], lo_c = sv[8], hi_c = sv[9], lse_ext = sv[10], den = sv[11];
    const int i_lo = __float_as_int(sv[12]), i_hi = __float_as_int(sv[13]);
    const int j_lo = __float_as_int(sv[14]), j_hi = __float_as_int(sv[15]);
    const float rx = hi_x - lo_x, rc = hi_c - lo_c;
    float ma = NR_NEG_INF;
    for (int r = lane; r < k; r += 32) {
      const int j = nb[r];
      ma = fmaxf(ma, a.tau_nbr * ((x[j] - lo_x) / rx - (a.cbank[j] - lo_c) / rc));
    }
    ma = warp_max(ma);
    float sa = 0.f;
    for (int r = lane; r < k; r += 32) {
      const int j = nb[r];
      sa += expf(a.tau_nbr * ((x[j] - lo_x) / rx - (a.cbank[j] - lo_c) / rc) - ma);
    }
    sa = warp_sum(sa);
    float mean_dp = 0.f;            // sum_r p_r * dL/dp_r,  dL/dp_r = -lp_r/den
    for (int r = lane; r < k; r += 32) {
      const int j = nb[r];
      const float p = expf(a.tau_nbr * ((x[j] - lo_x) / rx - (a.cbank[j] - lo_c) / rc) - ma) / sa;
      mean_dp += p * (-(x[j] - lse_ext) / den);
    }
    mean_dp = warp_sum(mean_dp);
    float sum_dlp = 0.f, g_lo_x = 0.f, g_hi_x = 0.f, g_lo_c = 0.f, g_hi_c = 0.f;
    for (int r = lane; r < k; r += 32) {
      const int j = nb[r];
      const float nxv = (x[j] - lo_x) / rx, ncv = (a.cbank[j] - lo_c) / rc;
      const float p = expf(a.tau_nbr * (nxv - ncv) - ma) / sa;
      const float dlp = -p / den;
      sum_dlp += dlp;
      const float ga = a.tau_nbr * p * (-(x[j] - lse_ext) / den - mean_dp);
      dx[j] += gs_n * (dlp + ga / rx);          // neighbour columns are distinct: no race
      g_lo_x += ga * (nxv - 1.f) / rx;
      g_hi_x += -ga * nxv / rx;
      if (dc) atomicAdd(dc + j, gs_n * (-ga / rc));
      g_lo_c += -ga * (ncv - 1.f) / rc;
      g_hi_c += ga * ncv / rc;
    }
    sum_dlp = warp_sum(sum_dlp) - 1.f / den;    // + the diagonal (weight 1)
    g_lo_x = warp_sum(g_lo_x); g_hi_x = warp_sum(g_hi_x);
    g_lo_c = warp_sum(g_lo_c); g_hi_c = warp_sum(g_hi_c);
    __syncwarp();
    // -q_i * sum_dlp over the extended set
    for (int r = lane; r < k; r += 32) {
      const int j = nb[r];
      dx[j] -= gs_n * expf(x[j] - lse_ext) * sum_dlp;
    }
    __syncwarp();
    if (lane == 0) {
      dx[gi] += gs_n * (-1.f / den) - gs_n * expf(x[gi] - lse_ext) * sum_dlp;
      dx[i_lo] += gs_n * g_lo_x;
      dx[i_hi] += gs_n * g_hi_x;
      if (dc) {
        atomicAdd(dc + j_lo, gs_n * g_lo_c);
        atomicAdd(dc + j_hi, gs_n * g_hi_c);
      }
    }
  }
  __syncthreads();
  float* dXr = dX + (int64_t)i * lddx;
  for (int j = tid; j < B; j += ROW_THREADS) dXr[j] = dx[j];
}

// ------------------------------------------------------------------------------------------
// small helpers: row mean (bank centrality), vector sums, transpose-add
// ------------------------------------------------------------------------------------------
__global__ void row_mean_kernel(const float* __restrict__ X, int64_t ld, int cols, float* __restrict__ out) {
  __shared__ float red[32];
  const float* r = X + (int64_t)blockIdx.x * ld;
  float s = 0.f;
  for (int j = threadIdx.x; j < cols; j += blockDim.x) s += r[j];
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[blockIdx.x] = s / (float)cols;     // sum / size  (until_module.py:181)
}

// out[v] = scale[v] * sum_j in[v, j]   (deterministic, one block per vector)
__global__ void vec_sums_kernel(const float* __restrict__ in, int64_t len, const float* __restrict__ scale,
                                float* __restrict__ out) {
  __shared__ float red[32];
  const float* r = in + (int64_t)blockIdx.x * len;
  float s = 0.f;
  for (int64_t j = threadIdx.x; j < len; j += blockDim.x) s += r[j];
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[blockIdx.x] = s * (scale ? scale[blockIdx.x] : 1.f);
}

// out[i, j] = alpha * a[i, j] + beta * b[j, i]      (tiled through shared memory; b may be null)
__global__ void transpose_add_kernel(const float* __restrict__ a, int64_t lda, const float* __restrict__ b,
                                     int64_t ldb, float* __restrict__ out, int64_t ldo, int rows, int cols,
                                     float alpha, float beta) {
  __shared__ float tile[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  if (b) {
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
      int bi = bx + r, bj = by + threadIdx.x;     // b[bi, bj] -> out[bj, bi]
      if (bi < cols && bj < rows) tile[r][threadIdx.x] = b[(int64_t)bi * ldb + bj];
    }
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    int oi = by + r, oj = bx + threadIdx.x;
    if (oi < rows && oj < cols) {
      float v = a ? alpha * a[(int64_t)oi * lda + oj] : 0.f;
      if (b) v += beta * tile[threadIdx.x][r];
      out[(int64_t)oi * ldo + oj] = v;
    }
  }
}

}  // namespace nr

using namespace nr;

static size_t row_smem(int B) { return (size_t)3 * B * sizeof(float); }

static int check_row_args(const char* fn, int64_t rows, int64_t B, int64_t row0, int k, int flags,
                          const void* X, const void* G, const void* cbank, const void* sk_u,
                          const void* sk_v) {
  NR_CHECK_ARG(rows > 0 && B > 0 && row0 >= 0 && row0 + rows <= B, "%s: bad rows/B/row0 (%lld,%lld,%lld)", fn,
               (long long)rows, (long long)B, (long long)row0);
  NR_CHECK_ARG(X != nullptr, "%s: X is null", fn);
  NR_CHECK_ARG(B <= NR_MAX_ROW_B, "%s: B=%lld exceeds the %d-column shared-memory row limit", fn, (long long)B,
               NR_MAX_ROW_B);
  if (flags & (NR_LOSS_KL | NR_LOSS_UNIFORM)) NR_CHECK_ARG(G != nullptr, "%s: G is null", fn);
  if (flags & NR_LOSS_UNIFORM) NR_CHECK_ARG(sk_u && sk_v, "%s: Sinkhorn duals are null", fn);
  if (flags & NR_LOSS_NEIGHBOR) {
    NR_CHECK_ARG(cbank != nullptr, "%s: cbank is null", fn);
    NR_CHECK_ARG(k >= 1 && k <= NR_MAX_K, "%s: num_neighbors=%d out of [1,%d]", fn, k, NR_MAX_K);
    // the reference needs B >= k+2: k neighbours + diagonal + >=1 non-neighbour
    NR_CHECK_ARG(B >= k + 2, "%s: B=%lld < num_neighbors+2=%d", fn, (long long)B, k + 2);
  }
  return 0;
}

extern "C" int nr_row_losses_fwd(const float* X, int64_t ldx, const float* G, int64_t ldg, const float* cbank,
                                 const float* w, const float* sk_u, const float* sk_v, int64_t rows, int64_t B,
                                 int64_t row0, const float* logit_scale, int k, float tau_nbr, float tau_uni,
                                 float beta, int flags, float* row_out, int32_t* nbr_idx, float* saved,
                                 void* stream) {
  if (int e = check_row_args("nr_row_losses_fwd", rows, B, row0, k, flags, X, G, cbank, sk_u, sk_v)) return e;
  NR_CHECK_ARG(row_out && saved, "nr_row_losses_fwd: output pointers are null");
  if (flags & NR_LOSS_NEIGHBOR) NR_CHECK_ARG(nbr_idx != nullptr, "nr_row_losses_fwd: nbr_idx is null");
  RowArgs a{X, ldx, G, ldg, cbank, w, sk_u, sk_v, (int)rows, (int)B, (int)row0, logit_scale, k, tau_nbr,
            tau_uni, beta, flags};
  if (B <= 1024) {        // the row fits the registers of one warp: no shared-memory staging, no block barriers
    const unsigned grid = (unsigned)((rows + ROWW_WARPS - 1) / ROWW_WARPS);
    cudaStream_t st = (cudaStream_t)stream;
    if (B <= 128) row_losses_fwd_warp_kernel<4><<<grid, ROWW_WARPS * 32, 0, st>>>(a, row_out, nbr_idx, saved);
    else if (B <= 256) row_losses_fwd_warp_kernel<8><<<grid, ROWW_WARPS * 32, 0, st>>>(a, row_out, nbr_idx, saved);
    else if (B <= 512) row_losses_fwd_warp_kernel<16><<<grid, ROWW_WARPS * 32, 0, st>>>(a, row_out, nbr_idx, saved);
    else row_losses_fwd_warp_kernel<32><<<grid, ROWW_WARPS * 32, 0, st>>>(a, row_out, nbr_idx, saved);
    NR_CHECK_LAUNCH("nr_row_losses_fwd(warp)");
    return 0;
  }
  size_t smem = row_smem((int)B);
  if (row_big(B)) {
    auto kern = row_losses_fwd_kernel<ROW_THREADS_BIG>;
    NR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    kern<<<(unsigned)rows, ROW_THREADS_BIG, smem, (cudaStream_t)stream>>>(a, row_out, nbr_idx, saved);
  } else {
    auto kern = row_losses_fwd_kernel<ROW_THREADS_SMALL>;
    if (smem > 48 * 1024) {
      NR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      NR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    kern<<<(unsigned)rows, ROW_THREADS_SMALL, smem, (cudaStream_t)stream>>>(a, row_out, nbr_idx, saved);
  }
  NR_CHECK_LAUNCH("nr_row_losses_fwd");
  return 0;
}

extern "C" int nr_row_losses_bwd(const float* X, int64_t ldx, const float* G, int64_t ldg, const float* cbank,
                                 const float* w, const float* sk_u, const float* sk_v, int64_t rows, int64_t B,
                                 int64_t row0, const float* logit_scale, int k, float tau_nbr, float tau_uni,
                                 float beta, int flags, const int32_t* nbr_idx, const float* saved,
                                 const float* gscale, float* dX, int64_t lddx, float* dG, int64_t lddg, float* dc,
                                 float* dw, float* dls, void* stream) {
  if (int e = check_row_args("nr_row_losses_bwd", rows, B, row0, k, flags, X, G, cbank, sk_u, sk_v)) return e;
  NR_CHECK_ARG(saved && gscale && dX, "nr_row_losses_bwd: saved/gscale/dX is null");
  if (flags & NR_LOSS_NEIGHBOR) NR_CHECK_ARG(nbr_idx != nullptr, "nr_row_losses_bwd: nbr_idx is null");
  RowArgs a{X, ldx, G, ldg, cbank, w, sk_u, sk_v, (int)rows, (int)B, (int)row0, logit_scale, k, tau_nbr,
            tau_uni, beta, flags};
  size_t smem = row_smem((int)B);
  if (row_big(B)) {
    auto kern = row_losses_bwd_kernel<ROW_THREADS_BIG>;
    NR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    NR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    kern<<<(unsigned)rows, ROW_THREADS_BIG, smem, (cudaStream_t)stream>>>(a, nbr_idx, saved, gscale, dX, lddx, dG, lddg, dc,
                                                                          dw, dls);
  } else {
    auto kern = row_losses_bwd_kernel<ROW_THREADS_SMALL>;
    if (smem > 48 * 1024) {
      NR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      NR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    }
    kern<<<(unsigned)rows, ROW_THREADS_SMALL, smem, (cudaStream_t)stream>>>(a, nbr_idx, saved, gscale, dX, lddx, dG, lddg,
                                                                            dc, dw, dls);
  }
  NR_CHECK_LAUNCH("nr_row_losses_bwd");
  return 0;
}

extern "C" int nr_row_mean(const float* X, int64_t ld, int64_t rows, int64_t cols, float* out, void* stream) {
  NR_CHECK_ARG(X && out && rows > 0 && cols > 0, "nr_row_mean: bad arguments");
  row_mean_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>(X, ld, (int)cols, out);
  NR_CHECK_LAUNCH("nr_row_mean");
  return 0;
}

extern "C" int nr_vec_sums(const float* in, int64_t nvec, int64_t len, const float* scale, float* out,
                           void* stream) {
  NR_CHECK_ARG(in && out && nvec > 0 && len > 0, "nr_vec_sums: bad arguments");
  vec_sums_kernel<<<(unsigned)nvec, 256, 0, (cudaStream_t)stream>>>(in, len, scale, out);
  NR_CHECK_LAUNCH("nr_vec_sums");
  return 0;
}

extern "C" int nr_transpose_add(const float* a, int64_t lda, const float* b, int64_t ldb, float* out, int64_t ldo,
                                int64_t rows, int64_t cols, float alpha, float beta, void* stream) {
  NR_CHECK_ARG(out && rows > 0 && cols > 0 && (a || b), "nr_transpose_add: bad arguments");
  dim3 grid((unsigned)((cols + 31) / 32), (unsigned)((rows + 31) / 32)), block(32, 8);
  transpose_add_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(a, lda, b, ldb, out, ldo, (int)rows, (int)cols,
                                                                 alpha, beta);
  NR_CHECK_LAUNCH("nr_transpose_add");
  return 0;
}
