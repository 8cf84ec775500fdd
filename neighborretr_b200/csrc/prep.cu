// Token preparation (L2 normalisation + bf16 operand copy + column sums) and the centrality
// weights built on it.  All kernels here are HBM-bound streaming passes: one warp per token row,
// 128-bit coalesced loads, warp-shuffle reductions, no atomics (per-CTA partials instead).
//
// Reference: F.normalize in local_level (NeighborRetr/models/modeling.py:495-496) and
// compute_centrality_weights (modeling.py:403-430).  The mean over all B*N tokens of
// <g_a, t_j> is computed as <g_a, mean_j t_j>: a column mean plus a GEMV instead of the
// reference's [B,D]x[D,B*N] GEMM.
// Algorithmic bytes: prep = rows*d*4 read + rows*d*(4+2) written.
#include "common.cuh"
#include "nrhead_internal.h"

namespace nr {

constexpr int PREP_WARPS = 8;
constexpr int PREP_MAXQ = 8;          // d <= 128*PREP_MAXQ
constexpr int PREP_MAX_BLOCKS = 296;  // 2 CTAs per SM

__global__ void __launch_bounds__(PREP_WARPS * 32)
prep_tokens_kernel(const float* __restrict__ x, int rows, int d, float* __restrict__ xn_f32,
                   __nv_bfloat16* __restrict__ xn_bf16, float* __restrict__ inv_norm,
                   float* __restrict__ partials, const int64_t* __restrict__ mask, int split) {
  __shared__ float colsm[PREP_WARPS][128 * PREP_MAXQ];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float4 acc[PREP_MAXQ];
#pragma unroll
  for (int q = 0; q < PREP_MAXQ; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int row = blockIdx.x * PREP_WARPS + warp; row < rows; row += gridDim.x * PREP_WARPS) {
    const float* xr = x + (int64_t)row * d;
    float4 v[PREP_MAXQ];
    float ss = 0.f;
#pragma unroll
    for (int q = 0; q < PREP_MAXQ; ++q) {
      int c = q * 128 + lane * 4;
      if (c < d) {
        v[q] = *reinterpret_cast<const float4*>(xr + c);
        ss += v[q].x * v[q].x + v[q].y * v[q].y + v[q].z * v[q].z + v[q].w * v[q].w;
      }
    }
    ss = warp_sum(ss);
    const float denom = fmaxf(sqrtf(ss), 1e-12f);    // F.normalize eps
    if (lane == 0 && inv_norm) inv_norm[row] = 1.0f / denom;
    // masked tokens become zero rows of the bf16 operand copy: their token pairs are then exactly 0, which is
    // what the reference's mask multiplies produce (modeling.py:500-501); fp32 copy and column sums keep them
    const bool live = mask ? (mask[row] != 0) : true;
#pragma unroll
    for (int q = 0; q < PREP_MAXQ; ++q) {
      int c = q * 128 + lane * 4;
      if (c < d) {
        float4 n = make_float4(v[q].x / denom, v[q].y / denom, v[q].z / denom, v[q].w / denom);
        if (xn_f32) *reinterpret_cast<float4*>(xn_f32 + (int64_t)row * d + c) = n;
        if (xn_bf16 && !split) {
          __nv_bfloat162 lo = __floats2bfloat162_rn(n.x, n.y), hi = __floats2bfloat162_rn(n.z, n.w);
          uint2 pk;
          pk.x = live ? *reinterpret_cast<uint32_t*>(&lo) : 0u;
          pk.y = live ? *reinterpret_cast<uint32_t*>(&hi) : 0u;
          *reinterpret_cast<uint2*>(xn_bf16 + (int64_t)row * d + c) = pk;
        } else if (xn_bf16) {
          // split operand [rows, 3d]: n = hi + lo (+ 2^-17 |n|), hi = bf16(n), lo = bf16(n - hi).
          // role 1 (X side): [hi | lo | hi], role 2 (Y side): [hi | hi | lo], so that the K = 3d contraction of an
          // X-role row with a Y-role row is  hi.hi + lo.hi + hi.lo  (the lo.lo term, ~2^-18, is dropped)
          const float f[4] = {n.x, n.y, n.z, n.w};
          __nv_bfloat16 h[4], l[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            h[e] = __float2bfloat16_rn(live ? f[e] : 0.f);
            l[e] = __float2bfloat16_rn(live ? f[e] - __bfloat162float(h[e]) : 0.f);
          }
          __nv_bfloat16* o = xn_bf16 + (int64_t)row * 3 * d + c;
          const uint2 ph = *reinterpret_cast<uint2*>(h), pl = *reinterpret_cast<uint2*>(l);
          *reinterpret_cast<uint2*>(o) = ph;
          *reinterpret_cast<uint2*>(o + d) = (split == 1) ? pl : ph;
          *reinterpret_cast<uint2*>(o + 2 * d) = (split == 1) ? ph : pl;
        }
        acc[q].x += n.x; acc[q].y += n.y; acc[q].z += n.z; acc[q].w += n.w;
      }
    }
  }
  if (partials) {
#pragma unroll
    for (int q = 0; q < PREP_MAXQ; ++q) {
      int c = q * 128 + lane * 4;
      if (c < d) *reinterpret_cast<float4*>(&colsm[warp][c]) = acc[q];
    }
    __syncthreads();
    for (int c = threadIdx.x; c < d; c += blockDim.x) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < PREP_WARPS; ++w) s += colsm[w][c];
      partials[(int64_t)blockIdx.x * d + c] = s;
    }
  }
}

__global__ void __launch_bounds__(PREP_WARPS * 32)
prep_tokens_bwd_kernel(const float* __restrict__ xn, const float* __restrict__ inv_norm,
                       const float* __restrict__ dxn, const float* __restrict__ add_vec,
                       const int64_t* __restrict__ mask, int rows, int d, float* __restrict__ dx, int accumulate) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x * PREP_WARPS + warp;
  if (row >= rows) return;
  const float* nr_ = xn + (int64_t)row * d;
  const bool live = mask ? (mask[row] != 0) : true;     // masked tokens take no max-sim gradient (dxn), only add_vec
  float4 n[PREP_MAXQ], t[PREP_MAXQ];
  float dot = 0.f;
#pragma unroll
  for (int q = 0; q < PREP_MAXQ; ++q) {
    int c = q * 128 + lane * 4;
    if (c < d) {
      n[q] = *reinterpret_cast<const float4*>(nr_ + c);
      t[q] = (dxn && live) ? *reinterpret_cast<const float4*>(dxn + (int64_t)row * d + c)
                           : make_float4(0.f, 0.f, 0.f, 0.f);
      if (add_vec) {
        float4 a = *reinterpret_cast<const float4*>(add_vec + c);
        t[q].x += a.x; t[q].y += a.y; t[q].z += a.z; t[q].w += a.w;
      }
      dot += n[q].x * t[q].x + n[q].y * t[q].y + n[q].z * t[q].z + n[q].w * t[q].w;
    }
  }
  dot = warp_sum(dot);
  const float inv = inv_norm[row];
  if (inv >= 1e12f) dot = 0.f;     // ||x|| below eps: the clamp has no gradient
#pragma unroll
  for (int q = 0; q < PREP_MAXQ; ++q) {
    int c = q * 128 + lane * 4;
    if (c < d) {
      float4 o = make_float4((t[q].x - n[q].x * dot) * inv, (t[q].y - n[q].y * dot) * inv,
                             (t[q].z - n[q].z * dot) * inv, (t[q].w - n[q].w * dot) * inv);
      float4* p = reinterpret_cast<float4*>(dx + (int64_t)row * d + c);
      if (accumulate) { float4 old = *p; o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w; }
      *p = o;
    }
  }
}

// block (32, 8): 32 columns x 8 interleaved partial rows, combined through shared memory
__global__ void mean_vec_kernel(const float* __restrict__ partials, int n_partials, int d, float inv_rows,
                                float* __restrict__ mean_vec) {
  __shared__ float sm[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < d)
    for (int p = threadIdx.y; p < n_partials; p += 8) s += partials[(int64_t)p * d + c];
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < d) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += sm[q][threadIdx.x];
    mean_vec[c] = t * inv_rows;
  }
}

// one warp per global-feature row: gn = normalize(g), w = exp(cs * <gn, mean>)
__global__ void __launch_bounds__(PREP_WARPS * 32)
centrality_rows_kernel(const float* __restrict__ g, const float* __restrict__ mean_vec, int B, int d, float cs,
                       float* __restrict__ gn, float* __restrict__ ginv, float* __restrict__ w) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x * PREP_WARPS + warp;
  if (row >= B) return;
  const float* gr = g + (int64_t)row * d;
  float ss = 0.f, dot = 0.f;
  for (int c = lane; c < d; c += 32) { float v = gr[c]; ss += v * v; dot += v * mean_vec[c]; }
  ss = warp_sum(ss);
  dot = warp_sum(dot);
  const float denom = fmaxf(sqrtf(ss), 1e-12f);
  for (int c = lane; c < d; c += 32) gn[(int64_t)row * d + c] = gr[c] / denom;
  if (lane == 0) {
    ginv[row] = 1.f / denom;
    w[row] = expf(cs * (dot / denom));       // exp(centrality * scale), modeling.py:427-428
  }
}

__global__ void __launch_bounds__(PREP_WARPS * 32)
centrality_rows_bwd_kernel(const float* __restrict__ mean_vec, const float* __restrict__ gn,
                           const float* __restrict__ ginv, const float* __restrict__ w,
                           const float* __restrict__ dw, int B, int d, float cs, float* __restrict__ dg,
                           int accumulate) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x * PREP_WARPS + warp;
  if (row >= B) return;
  const float dcv = cs * w[row] * dw[row];
  const float* nrow = gn + (int64_t)row * d;
  float dot = 0.f;
  for (int c = lane; c < d; c += 32) dot += nrow[c] * mean_vec[c];
  dot = warp_sum(dot) * dcv;               // <gn, dgn>, dgn = dc * mean
  const float inv = ginv[row];
  if (inv >= 1e12f) dot = 0.f;
  for (int c = lane; c < d; c += 32) {
    float o = (dcv * mean_vec[c] - nrow[c] * dot) * inv;
    float* p = dg + (int64_t)row * d + c;
    *p = accumulate ? (*p + o) : o;
  }
}

// dmean[c] = (1/rows_total) * sum_a cs*w[a]*dw[a]*gn[a,c]
__global__ void centrality_dmean_kernel(const float* __restrict__ gn, const float* __restrict__ w,
                                        const float* __restrict__ dw, int B, int d, float cs, float inv_rows,
                                        float* __restrict__ dmean) {
  __shared__ float sm[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float s = 0.f;
  if (c < d)
    for (int a = threadIdx.y; a < B; a += 8) s += cs * w[a] * dw[a] * gn[(int64_t)a * d + c];
  sm[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < d) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) t += sm[q][threadIdx.x];
    dmean[c] = t * inv_rows;
  }
}

// both parts of the centrality backward in ONE launch: blocks [0, nrb) = the per-row gradient w.r.t. the global
// features (centrality_rows_bwd_kernel), the remaining blocks = the gradient w.r.t. the token mean (32 columns each)
__global__ void __launch_bounds__(PREP_WARPS * 32)
centrality_bwd_kernel(const float* __restrict__ mean_vec, const float* __restrict__ gn, const float* __restrict__ ginv,
                      const float* __restrict__ w, const float* __restrict__ dw, int B, int d, float cs, float inv_rows,
                      float* __restrict__ dg, int accumulate, float* __restrict__ dmean, int nrb) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if ((int)blockIdx.x < nrb) {
    const int row = blockIdx.x * PREP_WARPS + warp;
    if (row >= B) return;
    const float dcv = cs * w[row] * dw[row];
    const float* nrow = gn + (int64_t)row * d;
    float dot = 0.f;
    for (int c = lane; c < d; c += 32) dot += nrow[c] * mean_vec[c];
    dot = warp_sum(dot) * dcv;               // <gn, dgn>, dgn = dc * mean
    const float inv = ginv[row];
    if (inv >= 1e12f) dot = 0.f;
    for (int c = lane; c < d; c += 32) {
      const float o = (dcv * mean_vec[c] - nrow[c] * dot) * inv;
      float* p = dg + (int64_t)row * d + c;
      *p = accumulate ? (*p + o) : o;
    }
  } else {
    __shared__ float sm[PREP_WARPS][33];
    const int c = ((int)blockIdx.x - nrb) * 32 + lane;
    float s = 0.f;
    if (c < d)
      for (int a = warp; a < B; a += PREP_WARPS) s += cs * w[a] * dw[a] * gn[(int64_t)a * d + c];
    sm[warp][lane] = s;
    __syncthreads();
    if (warp == 0 && c < d) {
      float t = 0.f;
#pragma unroll
      for (int q = 0; q < PREP_WARPS; ++q) t += sm[q][lane];
      dmean[c] = t * inv_rows;
    }
  }
}

}  // namespace nr

using namespace nr;

extern "C" int64_t nr_prep_partials(int64_t rows) {
  int64_t nb = (rows + PREP_WARPS - 1) / PREP_WARPS;
  return nb < PREP_MAX_BLOCKS ? (nb < 1 ? 1 : nb) : PREP_MAX_BLOCKS;
}

extern "C" int nr_prep_tokens(const float* x, int64_t rows, int64_t d, float* xn_f32, void* xn_bf16,
                              float* inv_norm, float* colsum_partials, const int64_t* mask, void* stream) {
  NR_CHECK_ARG(x && rows > 0, "nr_prep_tokens: bad arguments");
  NR_CHECK_ARG(d > 0 && d % 4 == 0 && d <= 128 * PREP_MAXQ, "nr_prep_tokens: d=%lld must be a multiple of 4, <= %d",
               (long long)d, 128 * PREP_MAXQ);
  int grid = (int)nr_prep_partials(rows);
  prep_tokens_kernel<<<grid, PREP_WARPS * 32, 0, (cudaStream_t)stream>>>(
      x, (int)rows, (int)d, xn_f32, (__nv_bfloat16*)xn_bf16, inv_norm, colsum_partials, mask, 0);
  NR_CHECK_LAUNCH("nr_prep_tokens");
  return 0;
}

extern "C" int nr_prep_tokens_split(const float* x, int64_t rows, int64_t d, float* xn_f32, void* xs_bf16, int role,
                                    float* inv_norm, float* colsum_partials, const int64_t* mask, void* stream) {
  NR_CHECK_ARG(x && xs_bf16 && rows > 0 && (role == 1 || role == 2), "nr_prep_tokens_split: bad arguments (role 1 = X, 2 = Y)");
  NR_CHECK_ARG(d > 0 && d % 4 == 0 && d <= 128 * PREP_MAXQ, "nr_prep_tokens_split: d=%lld must be a multiple of 4, <= %d",
               (long long)d, 128 * PREP_MAXQ);
  int grid = (int)nr_prep_partials(rows);
  prep_tokens_kernel<<<grid, PREP_WARPS * 32, 0, (cudaStream_t)stream>>>(
      x, (int)rows, (int)d, xn_f32, (__nv_bfloat16*)xs_bf16, inv_norm, colsum_partials, mask, role);
  NR_CHECK_LAUNCH("nr_prep_tokens_split");
  return 0;
}

extern "C" int nr_prep_tokens_bwd(const float* xn_f32, const float* inv_norm, const float* dxn,
                                  const float* add_vec, const int64_t* mask, int64_t rows, int64_t d, float* dx,
                                  int accumulate, void* stream) {
  NR_CHECK_ARG(xn_f32 && inv_norm && dx && rows > 0 && (dxn || add_vec), "nr_prep_tokens_bwd: bad arguments");
  NR_CHECK_ARG(d > 0 && d % 4 == 0 && d <= 128 * PREP_MAXQ, "nr_prep_tokens_bwd: unsupported d=%lld", (long long)d);
  int grid = (int)((rows + PREP_WARPS - 1) / PREP_WARPS);
  prep_tokens_bwd_kernel<<<grid, PREP_WARPS * 32, 0, (cudaStream_t)stream>>>(xn_f32, inv_norm, dxn, add_vec, mask,
                                                                            (int)rows, (int)d, dx, accumulate);
  NR_CHECK_LAUNCH("nr_prep_tokens_bwd");
  return 0;
}

extern "C" int nr_centrality_fwd(const float* colsum_partials, int64_t n_partials, int64_t rows_total,
                                 const float* g, int64_t B, int64_t d, float cs, float* mean_vec, float* gn,
                                 float* ginv, float* w, void* stream) {
  NR_CHECK_ARG(colsum_partials && g && mean_vec && gn && ginv && w && n_partials > 0 && rows_total > 0 && B > 0 &&
                   d > 0,
               "nr_centrality_fwd: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  mean_vec_kernel<<<(unsigned)((d + 31) / 32), dim3(32, 8), 0, s>>>(colsum_partials, (int)n_partials, (int)d,
                                                                   1.0f / (float)rows_total, mean_vec);
  NR_CHECK_LAUNCH("nr_centrality_fwd(mean)");
  centrality_rows_kernel<<<(unsigned)((B + PREP_WARPS - 1) / PREP_WARPS), PREP_WARPS * 32, 0, s>>>(
      g, mean_vec, (int)B, (int)d, cs, gn, ginv, w);
  NR_CHECK_LAUNCH("nr_centrality_fwd(rows)");
  return 0;
}

extern "C" int nr_centrality_bwd(const float* mean_vec, const float* gn, const float* ginv, const float* w,
                                 const float* dw, int64_t B, int64_t d, float cs, int64_t rows_total, float* dg,
                                 int accumulate, float* dmean, void* stream) {
  NR_CHECK_ARG(mean_vec && gn && ginv && w && dw && B > 0 && d > 0 && rows_total > 0,
               "nr_centrality_bwd: bad arguments");
  cudaStream_t s = (cudaStream_t)stream;
  if (dg && dmean) {
    const int nrb = (int)((B + PREP_WARPS - 1) / PREP_WARPS);
    centrality_bwd_kernel<<<(unsigned)(nrb + (d + 31) / 32), PREP_WARPS * 32, 0, s>>>(
        mean_vec, gn, ginv, w, dw, (int)B, (int)d, cs, 1.0f / (float)rows_total, dg, accumulate, dmean, nrb);
    NR_CHECK_LAUNCH("nr_centrality_bwd");
    return 0;
  }
  if (dg) {
    centrality_rows_bwd_kernel<<<(unsigned)((B + PREP_WARPS - 1) / PREP_WARPS), PREP_WARPS * 32, 0, s>>>(
        mean_vec, gn, ginv, w, dw, (int)B, (int)d, cs, dg, accumulate);
    NR_CHECK_LAUNCH("nr_centrality_bwd(rows)");
  }
  if (dmean) {
    centrality_dmean_kernel<<<(unsigned)((d + 31) / 32), dim3(32, 8), 0, s>>>(gn, w, dw, (int)B, (int)d, cs,
                                                                             1.0f / (float)rows_total, dmean);
    NR_CHECK_LAUNCH("nr_centrality_bwd(dmean)");
  }
  return 0;
}

// ---- global similarity for one global token per sample: G = gT gV^T and its transpose (modeling.py:516-539 with
// Gt = Gv = 1, where the softmax over a single token is 1).  Exact fp32 FMA; 32x32 output tile per CTA, 2x2
// outputs per thread, k-chunks of 32 staged in shared memory.  At B=128 the library picks a 32x32x16 SIMT
// kernel that takes 67 us for these 17 MFLOP; this one is latency-sized (16 CTAs x 16 k-steps).
namespace nr {
__global__ void __launch_bounds__(256)
gram_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, int Ra, int Rb, int d,
                float* __restrict__ out, float* __restrict__ outT) {
  // k-chunks of 128 (float4 loads, the next chunk's loads in flight during the FMAs): 4 dependent global round trips
  // for d = 512 instead of 16 — this kernel heads the longest chain of the forward (G -> Sinkhorn)
  constexpr int KC = 128;
  __shared__ __align__(16) float As[32][KC + 4], Bs[32][KC + 4];
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[2][2] = {{0.f, 0.f}, {0.f, 0.f}};
  float4 ra[4], rb[4];
  const bool vec = (d % 4 == 0);
  auto gload = [&](int k0) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int e = it * 256 + threadIdx.x, r = e >> 5, c = (e & 31) * 4;
      const int k = k0 + c;
      if (vec && k + 3 < d) {
        ra[it] = (i0 + r < Ra) ? *reinterpret_cast<const float4*>(a + (int64_t)(i0 + r) * d + k) : make_float4(0.f, 0.f, 0.f, 0.f);
        rb[it] = (j0 + r < Rb) ? *reinterpret_cast<const float4*>(b + (int64_t)(j0 + r) * d + k) : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
        float t[4], u[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          t[q] = (i0 + r < Ra && k + q < d) ? a[(int64_t)(i0 + r) * d + k + q] : 0.f;
          u[q] = (j0 + r < Rb && k + q < d) ? b[(int64_t)(j0 + r) * d + k + q] : 0.f;
        }
        ra[it] = make_float4(t[0], t[1], t[2], t[3]);
        rb[it] = make_float4(u[0], u[1], u[2], u[3]);
      }
    }
  };
  auto sstore = [&]() {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int e = it * 256 + threadIdx.x, r = e >> 5, c = (e & 31) * 4;
      *reinterpret_cast<float4*>(&As[r][c]) = ra[it];
      *reinterpret_cast<float4*>(&Bs[r][c]) = rb[it];
    }
  };
  gload(0);
  for (int k0 = 0; k0 < d; k0 += KC) {
    sstore();
    __syncthreads();
    if (k0 + KC < d) gload(k0 + KC);
#pragma unroll 8
    for (int k = 0; k < KC; k += 4) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[ty * 2][k]), a1 = *reinterpret_cast<const float4*>(&As[ty * 2 + 1][k]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[tx * 2][k]), b1 = *reinterpret_cast<const float4*>(&Bs[tx * 2 + 1][k]);
      acc[0][0] = fmaf(a0.x, b0.x, acc[0][0]); acc[0][0] = fmaf(a0.y, b0.y, acc[0][0]);
      acc[0][0] = fmaf(a0.z, b0.z, acc[0][0]); acc[0][0] = fmaf(a0.w, b0.w, acc[0][0]);
      acc[0][1] = fmaf(a0.x, b1.x, acc[0][1]); acc[0][1] = fmaf(a0.y, b1.y, acc[0][1]);
      acc[0][1] = fmaf(a0.z, b1.z, acc[0][1]); acc[0][1] = fmaf(a0.w, b1.w, acc[0][1]);
      acc[1][0] = fmaf(a1.x, b0.x, acc[1][0]); acc[1][0] = fmaf(a1.y, b0.y, acc[1][0]);
      acc[1][0] = fmaf(a1.z, b0.z, acc[1][0]); acc[1][0] = fmaf(a1.w, b0.w, acc[1][0]);
      acc[1][1] = fmaf(a1.x, b1.x, acc[1][1]); acc[1][1] = fmaf(a1.y, b1.y, acc[1][1]);
      acc[1][1] = fmaf(a1.z, b1.z, acc[1][1]); acc[1][1] = fmaf(a1.w, b1.w, acc[1][1]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int p = 0; p < 2; ++p)
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int i = i0 + ty * 2 + p, j = j0 + tx * 2 + q;
      if (i < Ra && j < Rb) {
        out[(int64_t)i * Rb + j] = acc[p][q];
        if (outT) outT[(int64_t)j * Ra + i] = acc[p][q];
      }
    }
}
}  // namespace nr

namespace nr {
// large batches (B >= 256): 64x64 output tile, 4x4 outputs per thread (8 float4 shared loads per 64 FMAs), k-chunks
// of 64 with the next chunk's global loads in flight.  At B = 1024 the 32x32 kernel above took 105 us in front of
// the Sinkhorn chain.
__global__ void __launch_bounds__(256)
gram_f32_big_kernel(const float* __restrict__ a, const float* __restrict__ b, int Ra, int Rb, int d,
                    float* __restrict__ out, float* __restrict__ outT) {
  constexpr int KC = 64;
  __shared__ __align__(16) float As[64][KC + 4], Bs[64][KC + 4];
  const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[p][q] = 0.f;
  float4 ra[4], rb[4];
  auto gload = [&](int k0) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int e = it * 256 + threadIdx.x, r = e >> 4, c = (e & 15) * 4;      // 64 rows x 16 float4
      const int k = k0 + c;
      const bool kin = k + 3 < d;
      ra[it] = (i0 + r < Ra && kin) ? *reinterpret_cast<const float4*>(a + (int64_t)(i0 + r) * d + k) : make_float4(0.f, 0.f, 0.f, 0.f);
      rb[it] = (j0 + r < Rb && kin) ? *reinterpret_cast<const float4*>(b + (int64_t)(j0 + r) * d + k) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  gload(0);
  for (int k0 = 0; k0 < d; k0 += KC) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int e = it * 256 + threadIdx.x, r = e >> 4, c = (e & 15) * 4;
      *reinterpret_cast<float4*>(&As[r][c]) = ra[it];
      *reinterpret_cast<float4*>(&Bs[r][c]) = rb[it];
    }
    __syncthreads();
    if (k0 + KC < d) gload(k0 + KC);
#pragma unroll 4
    for (int k = 0; k < KC; k += 4) {
      float4 av[4], bv[4];
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        av[p] = *reinterpret_cast<const float4*>(&As[ty * 4 + p][k]);
        bv[p] = *reinterpret_cast<const float4*>(&Bs[tx * 4 + p][k]);
      }
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          acc[p][q] = fmaf(av[p].x, bv[q].x, acc[p][q]); acc[p][q] = fmaf(av[p].y, bv[q].y, acc[p][q]);
          acc[p][q] = fmaf(av[p].z, bv[q].z, acc[p][q]); acc[p][q] = fmaf(av[p].w, bv[q].w, acc[p][q]);
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int i = i0 + ty * 4 + p, j = j0 + tx * 4 + q;
      if (i < Ra && j < Rb) {
        out[(int64_t)i * Rb + j] = acc[p][q];
        if (outT) outT[(int64_t)j * Ra + i] = acc[p][q];
      }
    }
}
}  // namespace nr

extern "C" int nr_gram_f32(const float* a, const float* b, int64_t Ra, int64_t Rb, int64_t d, float* out, float* outT,
                           void* stream) {
  NR_CHECK_ARG(a && b && out && Ra > 0 && Rb > 0 && d > 0, "nr_gram_f32: bad arguments");
  if (Ra >= 256 && Rb >= 256 && d % 4 == 0) {
    dim3 gridb((unsigned)((Rb + 63) / 64), (unsigned)((Ra + 63) / 64));
    nr::gram_f32_big_kernel<<<gridb, 256, 0, (cudaStream_t)stream>>>(a, b, (int)Ra, (int)Rb, (int)d, out, outT);
    NR_CHECK_LAUNCH("nr_gram_f32");
    return 0;
  }
  dim3 grid((unsigned)((Rb + 31) / 32), (unsigned)((Ra + 31) / 32));
  nr::gram_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(a, b, (int)Ra, (int)Rb, (int)d, out, outT);
  NR_CHECK_LAUNCH("nr_gram_f32");
  return 0;
}

// ---- token-weight MLP backward, hidden layer (reference modeling.py:148-153: Linear-ReLU-Linear(2D,1)) --------
// One streaming pass over the post-ReLU hidden activations h [T,H]:
//   dh[t,j] = dlogit[t] * w2[j] * (h[t,j] > 0)      (input of the dW1 / dx library GEMMs)
//   db1[j]  = sum_t dh[t,j],   dw2[j] = sum_t dlogit[t] * h[t,j]     (per-CTA partials, summed deterministically)
// replaces four ATen elementwise/reduction passes over [T,H]; HBM-bound: 8*T*H bytes.
namespace nr {
constexpr int MLPB_ROWS = 32;      // rows per CTA

// hidden activations are fp32 (fp32 / TF32 first layer) or bf16 (bf16 first layer): 4 consecutive elements per access
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ld4(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&u.x), b = *reinterpret_cast<const __nv_bfloat162*>(&u.y);
  const float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
  return make_float4(fa.x, fa.y, fb.x, fb.y);
}
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void st4(__nv_bfloat16* p, float4 v) {
  const __nv_bfloat162 a = __floats2bfloat162_rn(v.x, v.y), b = __floats2bfloat162_rn(v.z, v.w);
  uint2 u;
  u.x = *reinterpret_cast<const uint32_t*>(&a);
  u.y = *reinterpret_cast<const uint32_t*>(&b);
  *reinterpret_cast<uint2*>(p) = u;
}

// Second layer + masked softmax over the tokens of a sample (reference modeling.py:485-492):
//   logit[t] = <h[t,:], w2> + b2;  logit[masked] = -9e15;  w[r,:] = softmax_n(logit[r,:])
// One CTA per sample, one warp per token row (float4 loads of the post-ReLU activations), softmax by warp 0.
// Samples [0,Ra) take their mask from mask_a, samples [Ra,R) from mask_b (batch and bank tokens of one modality
// share the launch).  HBM-bound: 4*T*H bytes read.
template <typename HT>
__global__ void __launch_bounds__(256)
token_weights_fwd_kernel(const HT* __restrict__ h, const float* __restrict__ w2, const float* __restrict__ b2,
                         const int64_t* __restrict__ mask_a, const int64_t* __restrict__ mask_b, int Ra, int N, int H,
                         float* __restrict__ w) {
  __shared__ float logit[NR_MAX_TOKENS];
  const int r = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t* mrow = r < Ra ? (mask_a ? mask_a + (int64_t)r * N : nullptr)
                               : (mask_b ? mask_b + (int64_t)(r - Ra) * N : nullptr);
  const float bias = b2[0];
  for (int n = warp; n < N; n += 8) {
    const HT* hr = h + ((int64_t)r * N + n) * H;
    float s = 0.f;
    for (int c = lane * 4; c < H; c += 128) {
      const float4 hv = ld4(hr + c);
      const float4 wv = *reinterpret_cast<const float4*>(w2 + c);
      s += hv.x * wv.x + hv.y * wv.y + hv.z * wv.z + hv.w * wv.w;
    }
    s = warp_sum(s);
    if (lane == 0) logit[n] = (mrow && mrow[n] == 0) ? -9e15f : s + bias;
  }
  __syncthreads();
  if (warp == 0) {
    float mx = NR_NEG_INF;
    for (int n = lane; n < N; n += 32) mx = fmaxf(mx, logit[n]);
    mx = warp_max(mx);
    float sum = 0.f;
    for (int n = lane; n < N; n += 32) sum += expf(logit[n] - mx);
    sum = warp_sum(sum);
    for (int n = lane; n < N; n += 32) w[(int64_t)r * N + n] = expf(logit[n] - mx) / sum;
  }
}

// Backward of the same node down to the hidden layer, one streaming pass over h [T,H]:
//   dlogit[t] = w[t] * (dw[t] - sum_m w[r,m] dw[r,m])          (softmax; masked tokens have w = 0)
//   dh[t,j] = dlogit[t] * w2[j] * (h[t,j] > 0),  partials of db1[j] = sum_t dh[t,j], dw2[j] = sum_t dlogit[t] h[t,j],
//   db2 = sum_t dlogit[t]  (row 2H of the partials).   dw of samples [0,Ra) comes from dw_a, the rest from dw_b
//   (nullable = no gradient reached those weights).
template <typename HT>
__global__ void __launch_bounds__(256)
token_weights_bwd_kernel(const HT* __restrict__ h, const float* __restrict__ w, const float* __restrict__ dw_a,
                         const float* __restrict__ dw_b, int Ra, int N, const float* __restrict__ w2, int T, int H,
                         HT* __restrict__ dh, float* __restrict__ partials, int nchunks) {
  __shared__ float dl[MLPB_ROWS];
  const int chunk = blockIdx.x, t0 = chunk * MLPB_ROWS, t1 = min(T, t0 + MLPB_ROWS);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = warp; i < t1 - t0; i += 8) {
    const int t = t0 + i, r = t / N, n = t - r * N;
    const float* dwr = r < Ra ? (dw_a ? dw_a + (int64_t)r * N : nullptr) : (dw_b ? dw_b + (int64_t)(r - Ra) * N : nullptr);
    float v = 0.f;
    if (dwr) {
      const float* wr = w + (int64_t)r * N;
      float s = 0.f;
      for (int m = lane; m < N; m += 32) s += wr[m] * dwr[m];
      s = warp_sum(s);
      v = wr[n] * (dwr[n] - s);
    }
    if (lane == 0) dl[i] = v;
  }
  __syncthreads();
  if (warp == 0) {
    float s = (lane < t1 - t0) ? dl[lane] : 0.f;
    s = warp_sum(s);
    if (lane == 0) partials[(int64_t)(2 * H) * nchunks + chunk] = s;
  }
  for (int c = threadIdx.x * 4; c < H; c += 256 * 4) {
    const float4 wv = *reinterpret_cast<const float4*>(w2 + c);
    float4 sb = make_float4(0.f, 0.f, 0.f, 0.f), sw = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int tb = t0; tb < t1; tb += 8) {                 // 8 independent 16-byte loads in flight per thread
      float4 hv[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int t = min(tb + i, t1 - 1);
        hv[i] = ld4(h + (int64_t)t * H + c);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (tb + i < t1) {
          const float d = dl[tb + i - t0];
          float4 o;
          o.x = hv[i].x > 0.f ? d * wv.x : 0.f; o.y = hv[i].y > 0.f ? d * wv.y : 0.f;
          o.z = hv[i].z > 0.f ? d * wv.z : 0.f; o.w = hv[i].w > 0.f ? d * wv.w : 0.f;
          st4(dh + (int64_t)(tb + i) * H + c, o);
          sb.x += o.x; sb.y += o.y; sb.z += o.z; sb.w += o.w;
          sw.x += d * hv[i].x; sw.y += d * hv[i].y; sw.z += d * hv[i].z; sw.w += d * hv[i].w;
        }
      }
    }
    float* pb = partials + (int64_t)c * nchunks + chunk;
    pb[0] = sb.x; pb[nchunks] = sb.y; pb[2 * (int64_t)nchunks] = sb.z; pb[3 * (int64_t)nchunks] = sb.w;
    float* pw = partials + (int64_t)(H + c) * nchunks + chunk;
    pw[0] = sw.x; pw[nchunks] = sw.y; pw[2 * (int64_t)nchunks] = sw.z; pw[3 * (int64_t)nchunks] = sw.w;
  }
}

}  // namespace nr

extern "C" int64_t nr_mlp_chunks(int64_t T) { return (T + nr::MLPB_ROWS - 1) / nr::MLPB_ROWS; }


extern "C" int nr_token_weights_fwd(const void* h, int h_bf16, const float* w2, const float* b2, const int64_t* mask_a,
                                    const int64_t* mask_b, int64_t Ra, int64_t R, int64_t N, int64_t H, float* w,
                                    void* stream) {
  NR_CHECK_ARG(h && w2 && b2 && w && R > 0 && Ra >= 0 && Ra <= R && N > 0 && N <= NR_MAX_TOKENS && H > 0 && H % 4 == 0,
               "nr_token_weights_fwd: bad arguments");
  if (h_bf16)
    nr::token_weights_fwd_kernel<__nv_bfloat16><<<(unsigned)R, 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)h, w2, b2, mask_a, mask_b, (int)Ra, (int)N, (int)H, w);
  else
    nr::token_weights_fwd_kernel<float><<<(unsigned)R, 256, 0, (cudaStream_t)stream>>>((const float*)h, w2, b2, mask_a,
                                                                                      mask_b, (int)Ra, (int)N, (int)H, w);
  NR_CHECK_LAUNCH("nr_token_weights_fwd");
  return 0;
}

extern "C" int nr_token_weights_bwd(const void* h, int h_bf16, const float* w, const float* dw_a, const float* dw_b,
                                    int64_t Ra, int64_t R, int64_t N, const float* w2, int64_t H, void* dh,
                                    float* partials, void* stream) {
  NR_CHECK_ARG(h && w && w2 && dh && partials && R > 0 && Ra >= 0 && Ra <= R && N > 0 && H > 0 && H % 4 == 0,
               "nr_token_weights_bwd: bad arguments");
  const int64_t T = R * N;
  const int nchunks = (int)nr_mlp_chunks(T);
  if (h_bf16)
    nr::token_weights_bwd_kernel<__nv_bfloat16><<<nchunks, 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)h, w, dw_a, dw_b, (int)Ra, (int)N, w2, (int)T, (int)H, (__nv_bfloat16*)dh, partials, nchunks);
  else
    nr::token_weights_bwd_kernel<float><<<nchunks, 256, 0, (cudaStream_t)stream>>>(
        (const float*)h, w, dw_a, dw_b, (int)Ra, (int)N, w2, (int)T, (int)H, (float*)dh, partials, nchunks);
  NR_CHECK_LAUNCH("nr_token_weights_bwd");
  return 0;
}

// ---- persistent prepared memory bank (reference NeighborRetr/models/modeling.py:222-249) ----------
// The reference rebuilds the bank every step with five torch.cat calls; a head that re-prepares it would also
// re-normalise, re-cast and re-transpose all M rows although only the newest B changed.  Here the bank is a RING:
// raw fp32 rows, masks, the normalised bf16 operand copy (plain or split), its transposed copy and the raw bf16 copy
// the weight MLP multiplies all live in place, and a step writes only its B new samples at slots (head + j) mod M.
// `head` lives in device memory, so the same launch replays inside a CUDA graph.  Reference order is recovered as
// row i = slot (head + i) mod M (newest first).  HBM-bound: B*N*D*(4 read + 4 + 2 + 2 + 2 written) bytes.
namespace nr {

__global__ void bank_advance_kernel(int* __restrict__ head, int n_new, int M, const int64_t* __restrict__ new_ind,
                                    int64_t* __restrict__ ring_ind) {
  __shared__ int h;
  if (threadIdx.x == 0) {
    int v = *head - n_new;
    v %= M;
    if (v < 0) v += M;
    h = v;
  }
  __syncthreads();
  if (ring_ind)
    for (int j = threadIdx.x; j < n_new; j += blockDim.x) ring_ind[(h + j) % M] = new_ind[j];
  __syncthreads();
  if (threadIdx.x == 0) *head = h;
}

constexpr int BI_ROWS = 32;           // new token rows per CTA (= lanes of the transposed write)

struct BankSide {
  const float* x; const int64_t* mask; int rows, N;
  float* ring_feat; int64_t* ring_mask; __nv_bfloat16* ring_raw; __nv_bfloat16* ring_xn; __nv_bfloat16* ring_xnT;
  int64_t ld; int split, blk0;
};
struct BankInsertArgs { BankSide s[2]; int nsides, d, M; const int* head; };

// One CTA = 32 consecutive new token rows of one modality: each warp normalises 4 rows (float4 loads, warp-shuffle
// norm) and stores the row-major outputs directly; the bf16 hi (and lo) values are also staged in shared memory so
// that the transposed copy is written with lane = token: 32 consecutive 2-byte elements per dim instead of one
// 2-byte store per (token, dim) scattered over `ld`-strided rows.
__global__ void __launch_bounds__(256) bank_insert_kernel(const BankInsertArgs a) {
  extern __shared__ __align__(16) uint8_t bi_smem[];
  int si = 0;
  while (si + 1 < a.nsides && (int)blockIdx.x >= a.s[si + 1].blk0) ++si;
  const BankSide& S = a.s[si];
  const int d = a.d, M = a.M, N = S.N;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row0 = ((int)blockIdx.x - S.blk0) * BI_ROWS;
  const int sld = d + 2;                                          // padded row stride (elements): conflict-free columns
  __nv_bfloat16* hi_s = reinterpret_cast<__nv_bfloat16*>(bi_smem);
  __nv_bfloat16* lo_s = hi_s + BI_ROWS * sld;
  __shared__ int64_t drow_s[BI_ROWS];
  const int head = *a.head;
  for (int rr = warp; rr < BI_ROWS; rr += 8) {
    const int row = row0 + rr;
    if (row >= S.rows) { if (lane == 0) drow_s[rr] = -1; continue; }
    const int j = row / N, n = row - j * N;
    const int64_t drow = (int64_t)((head + j) % M) * N + n;
    if (lane == 0) drow_s[rr] = drow;
    const float* xr = S.x + (int64_t)row * d;
    float4 v[PREP_MAXQ];
    float ss = 0.f;
#pragma unroll
    for (int q = 0; q < PREP_MAXQ; ++q) {
      const int c = q * 128 + lane * 4;
      if (c < d) {
        v[q] = *reinterpret_cast<const float4*>(xr + c);
        ss += v[q].x * v[q].x + v[q].y * v[q].y + v[q].z * v[q].z + v[q].w * v[q].w;
      }
    }
    ss = warp_sum(ss);
    const float denom = fmaxf(sqrtf(ss), 1e-12f);
    const bool live = S.mask ? (S.mask[row] != 0) : true;
    if (S.ring_mask && lane == 0) S.ring_mask[drow] = S.mask ? S.mask[row] : 1;
    const int kd = S.split ? 3 * d : d;
#pragma unroll
    for (int q = 0; q < PREP_MAXQ; ++q) {
      const int c = q * 128 + lane * 4;
      if (c >= d) continue;
      if (S.ring_feat) *reinterpret_cast<float4*>(S.ring_feat + drow * d + c) = v[q];
      const float f[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
      if (S.ring_raw) {
        __nv_bfloat16 r4[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) r4[e] = __float2bfloat16_rn(f[e]);
        *reinterpret_cast<uint2*>(S.ring_raw + drow * d + c) = *reinterpret_cast<uint2*>(r4);
      }
      __nv_bfloat16 h[4], l[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float nv = live ? f[e] / denom : 0.f;
        h[e] = __float2bfloat16_rn(nv);
        l[e] = __float2bfloat16_rn(live ? nv - __bfloat162float(h[e]) : 0.f);
        hi_s[rr * sld + c + e] = h[e];
        if (S.split) lo_s[rr * sld + c + e] = l[e];
      }
      if (S.ring_xn) {
        const uint2 ph = *reinterpret_cast<uint2*>(h), pl = *reinterpret_cast<uint2*>(l);
        __nv_bfloat16* o = S.ring_xn + drow * kd + c;
        *reinterpret_cast<uint2*>(o) = ph;
        if (S.split) {
          *reinterpret_cast<uint2*>(o + d) = (S.split == 1) ? pl : ph;
          *reinterpret_cast<uint2*>(o + 2 * d) = (S.split == 1) ? ph : pl;
        }
      }
    }
  }
  if (!S.ring_xnT) return;
  __syncthreads();
  // transposed copy: lane = token (consecutive ring rows except across a sample that wraps), warps stride over dims
  const int64_t drow = drow_s[lane];
  if (drow < 0) return;
  for (int c = warp; c < d; c += 8) {
    const __nv_bfloat16 h = hi_s[lane * sld + c];
    S.ring_xnT[(int64_t)c * S.ld + drow] = h;
    if (S.split) {
      const __nv_bfloat16 l = lo_s[lane * sld + c];
      S.ring_xnT[(int64_t)(d + c) * S.ld + drow] = (S.split == 1) ? l : h;
      S.ring_xnT[(int64_t)(2 * d + c) * S.ld + drow] = (S.split == 1) ? h : l;
    }
  }
}

// The late half of a split insert: raw fp32 rows -> bf16 at the ring slots (head + j) mod M.  One thread = four
// consecutive elements; 12 bytes of traffic per element, nothing else.
__global__ void __launch_bounds__(256) bank_insert_raw_kernel(const BankInsertArgs a) {
  const int d4 = a.d >> 2, M = a.M;
  const int head = *a.head;
  for (int si = 0; si < a.nsides; ++si) {
    const BankSide& S = a.s[si];
    const int64_t units = (int64_t)S.rows * d4;
    for (int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; u < units; u += (int64_t)gridDim.x * blockDim.x) {
      const int row = (int)(u / d4), c = (int)(u - (int64_t)row * d4) * 4;
      const int j = row / S.N, n = row - j * S.N;
      const int64_t drow = (int64_t)((head + j) % M) * S.N + n;
      const float4 v = *reinterpret_cast<const float4*>(S.x + (int64_t)row * a.d + c);
      __nv_bfloat16 r4[4] = {__float2bfloat16_rn(v.x), __float2bfloat16_rn(v.y), __float2bfloat16_rn(v.z),
                             __float2bfloat16_rn(v.w)};
      *reinterpret_cast<uint2*>(S.ring_raw + drow * a.d + c) = *reinterpret_cast<uint2*>(r4);
    }
  }
}

}  // namespace nr

extern "C" int nr_bank_advance(int* head, int64_t n_new, int64_t M, const int64_t* new_ind, int64_t* ring_ind,
                               void* stream) {
  NR_CHECK_ARG(head && n_new > 0 && M > 0 && n_new <= M && (!ring_ind || new_ind), "nr_bank_advance: bad arguments");
  nr::bank_advance_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(head, (int)n_new, (int)M, new_ind, ring_ind);
  NR_CHECK_LAUNCH("nr_bank_advance");
  return 0;
}

static int bank_side_fill(nr::BankSide& S, const float* new_feat, const int64_t* new_mask, int64_t n_new, int64_t N,
                          int64_t M, float* ring_feat, int64_t* ring_mask, void* ring_raw_bf16, void* ring_xn_bf16,
                          int split_role, void* ring_xnT_bf16, int64_t ld, int& blocks) {
  NR_CHECK_ARG(new_feat && n_new > 0 && n_new <= M && N > 0, "nr_bank_insert: bad arguments");
  NR_CHECK_ARG(split_role >= 0 && split_role <= 2 && (!ring_xnT_bf16 || ld >= M * N), "nr_bank_insert: bad split role or ld");
  S.x = new_feat; S.mask = new_mask; S.rows = (int)(n_new * N); S.N = (int)N;
  S.ring_feat = ring_feat; S.ring_mask = ring_mask; S.ring_raw = (__nv_bfloat16*)ring_raw_bf16;
  S.ring_xn = (__nv_bfloat16*)ring_xn_bf16; S.ring_xnT = (__nv_bfloat16*)ring_xnT_bf16; S.ld = ld; S.split = split_role;
  S.blk0 = blocks;
  blocks += (S.rows + nr::BI_ROWS - 1) / nr::BI_ROWS;
  return 0;
}

static int bank_insert_launch(nr::BankInsertArgs& a, int blocks, cudaStream_t stream);

// only the raw bf16 rows (the weight-MLP operand) are requested: a plain cast into the ring slots, no normalisation
static bool bank_raw_only(const nr::BankInsertArgs& a) {
  for (int i = 0; i < a.nsides; ++i)
    if (!a.s[i].ring_raw || a.s[i].ring_feat || a.s[i].ring_mask || a.s[i].ring_xn || a.s[i].ring_xnT) return false;
  return true;
}

static int bank_insert_launch(nr::BankInsertArgs& a, int blocks, cudaStream_t stream) {
  NR_CHECK_ARG(a.d > 0 && a.d % 4 == 0 && a.d <= 128 * nr::PREP_MAXQ, "nr_bank_insert: d=%d must be a multiple of 4, <= %d", a.d,
               128 * nr::PREP_MAXQ);
  if (bank_raw_only(a)) {
    int64_t units = 0;
    for (int i = 0; i < a.nsides; ++i) units += (int64_t)a.s[i].rows * (a.d / 4);
    const int grid = (int)((units + 255) / 256 < 148 * 8 ? (units + 255) / 256 : 148 * 8);
    nr::bank_insert_raw_kernel<<<grid, 256, 0, stream>>>(a);
    NR_CHECK_LAUNCH("nr_bank_insert(raw)");
    return 0;
  }
  const size_t smem = (size_t)2 * nr::BI_ROWS * (a.d + 2) * sizeof(__nv_bfloat16);
  static bool attr_set = false;
  if (!attr_set) {
    NR_CUDA(cudaFuncSetAttribute(nr::bank_insert_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_set = true;
  }
  nr::bank_insert_kernel<<<blocks, 256, smem, stream>>>(a);
  NR_CHECK_LAUNCH("nr_bank_insert");
  return 0;
}

extern "C" int nr_bank_insert(const float* new_feat, const int64_t* new_mask, int64_t n_new, int64_t N, int64_t d,
                              int64_t M, const int* head, float* ring_feat, int64_t* ring_mask, void* ring_raw_bf16,
                              void* ring_xn_bf16, int split_role, void* ring_xnT_bf16, int64_t ld, void* stream) {
  NR_CHECK_ARG(head, "nr_bank_insert: head is null");
  nr::BankInsertArgs a{};
  int blocks = 0;
  if (int e = bank_side_fill(a.s[0], new_feat, new_mask, n_new, N, M, ring_feat, ring_mask, ring_raw_bf16, ring_xn_bf16,
                             split_role, ring_xnT_bf16, ld, blocks))
    return e;
  a.nsides = 1; a.d = (int)d; a.M = (int)M; a.head = head;
  return bank_insert_launch(a, blocks, (cudaStream_t)stream);
}

/* both modalities of a step in one launch */
extern "C" int nr_bank_insert_pair(const nr_bank_side* sides, int n_sides, int64_t n_new, int64_t d, int64_t M,
                                   const int* head, void* stream) {
  NR_CHECK_ARG(sides && n_sides >= 1 && n_sides <= 2 && head, "nr_bank_insert_pair: 1 or 2 sides, head required");
  nr::BankInsertArgs a{};
  int blocks = 0;
  for (int i = 0; i < n_sides; ++i)
    if (int e = bank_side_fill(a.s[i], sides[i].new_feat, sides[i].new_mask, n_new, sides[i].N, M, sides[i].ring_feat,
                               sides[i].ring_mask, sides[i].ring_raw_bf16, sides[i].ring_xn_bf16, sides[i].split_role,
                               sides[i].ring_xnT_bf16, sides[i].ld, blocks))
      return e;
  a.nsides = n_sides; a.d = (int)d; a.M = (int)M; a.head = head;
  return bank_insert_launch(a, blocks, (cudaStream_t)stream);
}
