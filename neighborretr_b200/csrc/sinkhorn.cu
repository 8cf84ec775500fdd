// Log-space Sinkhorn of the uniform-regularisation loss, both directions (G and G^T) advanced by
// ONE persistent cooperative kernel: every CTA keeps its slab of rows of G and of G^T resident in
// shared memory for all iterations, a warp owns a row log-sum-exp, and the four dual vectors are
// exchanged through L2 between half-iterations with a device-wide barrier.
//
// Reference: UniformRegularizationLoss.sinkhorn_algorithm, NeighborRetr/models/until_module.py:222-251
//   u = nu - LSE_row(G + v);  v = nu - LSE_col(G + u);  nu = -log(2B);  50 iterations, no grad.
// The reference issues 2*50 logsumexp launches per direction (200 per step); here it is one launch.
// Chain 2 is the same recursion on G^T, so both chains need "rows of G" and "rows of G^T" in every
// half-iteration and share the resident slabs.
// Roofline: resident => latency/barrier bound at B <= ~1500; beyond that 100 * 2 * 4 * B^2 bytes of L2/HBM.
#include <cooperative_groups.h>
#include <stdlib.h>
#include <string.h>
#include "common.cuh"
#include "nrhead_internal.h"

namespace nr {

constexpr int SK_THREADS = 256;
constexpr int SK_WARPS = SK_THREADS / 32;

__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(counter) : "memory");
    } while (v < target);
    __threadfence();
  }
  __syncthreads();
}

// Grid variant (any B): every CTA keeps its slab of rows of K = exp(G - max G) and of K^T in shared memory (or
// streams them from L2 when they do not fit), the four scaling vectors live in global memory and are staged
// into shared memory once per half-iteration (one coalesced L2 read instead of a dependent load per element).
// Scaling domain alpha = 1/(K beta), beta = 1/(K^T alpha) when max G - min G <= 30, log domain otherwise.
__device__ __forceinline__ float warp_dot(const float* __restrict__ row, const float* __restrict__ vec, int B, int lane) {
  float s = 0.f;
  for (int j = lane; j < B; j += 32) s = fmaf(row[j], vec[j], s);
  return warp_sum(s);
}
__device__ __forceinline__ float warp_lse(const float* __restrict__ row, const float* __restrict__ vec, int B, int lane) {
  float m = NR_NEG_INF;
  for (int j = lane; j < B; j += 32) m = fmaxf(m, row[j] + vec[j]);
  m = warp_max(m);
  float s = 0.f;
  for (int j = lane; j < B; j += 32) s += expf(row[j] + vec[j] - m);
  s = warp_sum(s);
  return m + logf(s);
}

__global__ void __launch_bounds__(SK_THREADS)
sinkhorn_kernel(const float* __restrict__ G, const float* __restrict__ GT, int B, int iters, int rows_per_cta,
                int resident, float* u1, float* v1, float* u2, float* v2, unsigned int* counter, float* gstat) {
  extern __shared__ float sm[];
  __shared__ float red[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r0 = blockIdx.x * rows_per_cta;
  const int nr = max(0, min(rows_per_cta, B - r0));
  float* vs = sm;                                   // [2][B] staged source vectors (chain 1, chain 2)
  float* Gs = sm + 2 * (size_t)B;
  float* GTs = Gs + (size_t)rows_per_cta * B;
  // global max / min of G: per-CTA partials through global memory, combined after the first barrier
  float mx = NR_NEG_INF, mn = INFINITY;
  for (int e = tid; e < nr * B; e += SK_THREADS) {
    const float g = G[(size_t)r0 * B + e];
    mx = fmaxf(mx, g); mn = fminf(mn, g);
    if (resident) { Gs[e] = g; GTs[e] = GT[(size_t)r0 * B + e]; }
  }
  mx = block_max(mx, red);
  mn = -block_max(-mn, red);
  if (tid == 0) { gstat[2 * blockIdx.x] = mx; gstat[2 * blockIdx.x + 1] = mn; }
  unsigned int bar = 0;
  grid_barrier(counter, (++bar) * gridDim.x);
  mx = NR_NEG_INF; mn = INFINITY;
  for (int c = tid; c < (int)gridDim.x; c += SK_THREADS) {
    mx = fmaxf(mx, __ldcg(gstat + 2 * c)); mn = fminf(mn, __ldcg(gstat + 2 * c + 1));
  }
  mx = block_max(mx, red);
  mn = -block_max(-mn, red);
  const float nu = -logf(2.0f * (float)B);
  const bool scaling = resident && (mx - mn) <= 30.f;
  if (scaling)
    for (int e = tid; e < nr * B; e += SK_THREADS) { Gs[e] = expf(Gs[e] - mx); GTs[e] = expf(GTs[e] - mx); }
  float* vecs[4] = {u1, v1, u2, v2};
  for (int r = tid; r < nr; r += SK_THREADS) {
    const float init = scaling ? 1.f : 0.f;
    u1[r0 + r] = init; v1[r0 + r] = init; u2[r0 + r] = init; v2[r0 + r] = init;
  }
  grid_barrier(counter, (++bar) * gridDim.x);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      // half 0: (u1 <- rows of G with v1, u2 <- rows of G^T with v2); half 1: (v1 <- G^T with u1, v2 <- G with u2)
      const float* s1 = vecs[half ? 0 : 1];
      const float* s2 = vecs[half ? 2 : 3];
      for (int j = tid; j < B; j += SK_THREADS) { vs[j] = __ldcg(s1 + j); vs[B + j] = __ldcg(s2 + j); }
      __syncthreads();
      for (int w = warp; w < 2 * nr; w += SK_WARPS) {
        const int r = w >> 1, chain = w & 1;
        const bool useT = (chain ^ half) != 0;
        const float* row = resident ? ((useT ? GTs : Gs) + (size_t)r * B) : ((useT ? GT : G) + (size_t)(r0 + r) * B);
        float val;
        if (scaling) val = 1.0f / warp_dot(row, vs + chain * B, B, lane);
        else val = nu - warp_lse(row, vs + chain * B, B, lane);
        if (lane == 0) vecs[chain * 2 + (half ? 1 : 0)][r0 + r] = val;
      }
      grid_barrier(counter, (++bar) * gridDim.x);
    }
  }
  if (scaling) {        // back to the log duals: u = (nu - max G) + log alpha, v = log beta
    for (int r = tid; r < nr; r += SK_THREADS) {
      u1[r0 + r] = (nu - mx) + logf(u1[r0 + r]); v1[r0 + r] = logf(v1[r0 + r]);
      u2[r0 + r] = (nu - mx) + logf(u2[r0 + r]); v2[r0 + r] = logf(v2[r0 + r]);
    }
  }
}

// ---- streaming variant for batches whose slabs do not fit shared memory (B = 8192: 2 x 256 MB) -------------------
// In the scaling form (range of G <= 30, as above) an iteration is four matrix-vector products with
// K = exp(G - max G) and K^T; K and K^T are written ONCE into the workspace and every half-iteration streams them
// with 16-byte loads, 4 in flight per lane, 32 warps per SM — HBM-bound: 100 x 2 x 4 B^2 bytes.  (The log-domain
// kernel above re-evaluates exp(G + v) per element per half-iteration with one 4-byte load in flight per lane:
// measured 1.07 TB/s at B = 8192.)  Falls back to the log-domain loops when the range check fails.
constexpr int SKS_THREADS = 1024;
constexpr int SKS_WARPS = SKS_THREADS / 32;

__device__ __forceinline__ float warp_dot4(const float* __restrict__ row, const float* __restrict__ vec, int B, int lane) {
  float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
  int j = lane * 4;
  for (; j + 3 * 128 < B; j += 4 * 128) {                 // 4 independent 16-byte loads per lane in flight
    const float4 a0 = __ldcs(reinterpret_cast<const float4*>(row + j));
    const float4 a1 = __ldcs(reinterpret_cast<const float4*>(row + j + 128));
    const float4 a2 = __ldcs(reinterpret_cast<const float4*>(row + j + 256));
    const float4 a3 = __ldcs(reinterpret_cast<const float4*>(row + j + 384));
    const float4 b0 = *reinterpret_cast<const float4*>(vec + j), b1 = *reinterpret_cast<const float4*>(vec + j + 128);
    const float4 b2 = *reinterpret_cast<const float4*>(vec + j + 256), b3 = *reinterpret_cast<const float4*>(vec + j + 384);
    s0 += a0.x * b0.x + a0.y * b0.y + a0.z * b0.z + a0.w * b0.w;
    s1 += a1.x * b1.x + a1.y * b1.y + a1.z * b1.z + a1.w * b1.w;
    s2 += a2.x * b2.x + a2.y * b2.y + a2.z * b2.z + a2.w * b2.w;
    s3 += a3.x * b3.x + a3.y * b3.y + a3.z * b3.z + a3.w * b3.w;
  }
  for (; j < B; j += 128) {
    const float4 a0 = __ldcs(reinterpret_cast<const float4*>(row + j));
    const float4 b0 = *reinterpret_cast<const float4*>(vec + j);
    s0 += a0.x * b0.x + a0.y * b0.y + a0.z * b0.z + a0.w * b0.w;
  }
  return warp_sum((s0 + s1) + (s2 + s3));
}

__global__ void __launch_bounds__(SKS_THREADS)
sinkhorn_stream_kernel(const float* __restrict__ G, const float* __restrict__ GT, int B, int iters, int rows_per_cta,
                       float* u1, float* v1, float* u2, float* v2, unsigned int* counter, float* gstat,
                       float* __restrict__ K, float* __restrict__ KT) {
  extern __shared__ float sm[];                      // [2][B] staged source vectors
  __shared__ float red[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r0 = blockIdx.x * rows_per_cta;
  const int nr = max(0, min(rows_per_cta, B - r0));
  float* vs = sm;
  float mx = NR_NEG_INF, mn = INFINITY;
  for (size_t e = (size_t)tid * 4; e < (size_t)nr * B; e += (size_t)SKS_THREADS * 4) {
    const float4 g = *reinterpret_cast<const float4*>(G + (size_t)r0 * B + e);
    mx = fmaxf(fmaxf(mx, fmaxf(g.x, g.y)), fmaxf(g.z, g.w));
    mn = fminf(fminf(mn, fminf(g.x, g.y)), fminf(g.z, g.w));
  }
  mx = block_max(mx, red);
  mn = -block_max(-mn, red);
  if (tid == 0) { gstat[2 * blockIdx.x] = mx; gstat[2 * blockIdx.x + 1] = mn; }
  unsigned int bar = 0;
  grid_barrier(counter, (++bar) * gridDim.x);
  mx = NR_NEG_INF; mn = INFINITY;
  for (int c = tid; c < (int)gridDim.x; c += SKS_THREADS) {
    mx = fmaxf(mx, __ldcg(gstat + 2 * c)); mn = fminf(mn, __ldcg(gstat + 2 * c + 1));
  }
  mx = block_max(mx, red);
  mn = -block_max(-mn, red);
  const float nu = -logf(2.0f * (float)B);
  const bool scaling = (mx - mn) <= 30.f;
  if (scaling) {
    for (size_t e = (size_t)tid * 4; e < (size_t)nr * B; e += (size_t)SKS_THREADS * 4) {
      const size_t o = (size_t)r0 * B + e;
      const float4 g = *reinterpret_cast<const float4*>(G + o), t = *reinterpret_cast<const float4*>(GT + o);
      *reinterpret_cast<float4*>(K + o) = make_float4(expf(g.x - mx), expf(g.y - mx), expf(g.z - mx), expf(g.w - mx));
      *reinterpret_cast<float4*>(KT + o) = make_float4(expf(t.x - mx), expf(t.y - mx), expf(t.z - mx), expf(t.w - mx));
    }
  }
  float* vecs[4] = {u1, v1, u2, v2};
  for (int r = tid; r < nr; r += SKS_THREADS) {
    const float init = scaling ? 1.f : 0.f;
    u1[r0 + r] = init; v1[r0 + r] = init; u2[r0 + r] = init; v2[r0 + r] = init;
  }
  grid_barrier(counter, (++bar) * gridDim.x);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const float* s1 = vecs[half ? 0 : 1];
      const float* s2 = vecs[half ? 2 : 3];
      for (int j = tid; j < B; j += SKS_THREADS) { vs[j] = __ldcg(s1 + j); vs[B + j] = __ldcg(s2 + j); }
      __syncthreads();
      for (int w = warp; w < 2 * nr; w += SKS_WARPS) {
        const int r = w >> 1, chain = w & 1;
        const bool useT = (chain ^ half) != 0;
        float val;
        if (scaling) val = 1.0f / warp_dot4((useT ? KT : K) + (size_t)(r0 + r) * B, vs + chain * B, B, lane);
        else val = nu - warp_lse((useT ? GT : G) + (size_t)(r0 + r) * B, vs + chain * B, B, lane);
        if (lane == 0) vecs[chain * 2 + (half ? 1 : 0)][r0 + r] = val;
      }
      grid_barrier(counter, (++bar) * gridDim.x);
    }
  }
  if (scaling) {
    for (int r = tid; r < nr; r += SKS_THREADS) {
      u1[r0 + r] = (nu - mx) + logf(u1[r0 + r]); v1[r0 + r] = logf(v1[r0 + r]);
      u2[r0 + r] = (nu - mx) + logf(u2[r0 + r]); v2[r0 + r] = logf(v2[r0 + r]);
    }
  }
}

// ---- grid variant with data-carrying exchange (the default for 128 < B when the slabs are resident) ---------
// Same slabs and arithmetic as sinkhorn_kernel, but the half-iterations are not separated by a device-wide barrier:
// every scaling/dual value is published as one aligned 8-byte {value, epoch} word and its consumers spin on the
// word itself until the epoch of the previous half-iteration shows up.  One L2 round trip per half-iteration
// instead of three (barrier atomic, barrier poll, vector load).  No slot is rewritten before every CTA has
// consumed it: a CTA writes epoch e+2 into a slot only after it has read all values of epoch e+1, which exist
// only after every CTA has read all values of epoch e.  The tag area is zeroed per launch (epochs start at 1).
constexpr int SKT_THREADS = 512;
constexpr int SKT_WARPS = SKT_THREADS / 32;

__device__ __forceinline__ float wait_tagged(const uint2* slot, unsigned int epoch) {
  uint2 w;
  do {
    asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(w.x), "=r"(w.y) : "l"(slot) : "memory");
  } while (w.y != epoch);
  return __uint_as_float(w.x);
}
__device__ __forceinline__ void publish_tagged(uint2* slot, float v, unsigned int epoch) {
  asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(slot), "r"(__float_as_uint(v)), "r"(epoch) : "memory");
}

__global__ void __launch_bounds__(SKT_THREADS)
sinkhorn_tag_kernel(const float* __restrict__ G, const float* __restrict__ GT, int B, int iters, int rows_per_cta,
                    float* u1, float* v1, float* u2, float* v2, unsigned int* counter, float* gstat, uint2* tv) {
  extern __shared__ float sm[];
  __shared__ float red[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r0 = blockIdx.x * rows_per_cta;
  const int nr = max(0, min(rows_per_cta, B - r0));
  float* vs = sm;                                   // [2][B] staged source vectors (chain 1, chain 2)
  float* mine = sm + 2 * (size_t)B;                 // [4][rows_per_cta] this CTA's latest u1, v1, u2, v2
  float* Gs = mine + 4 * (size_t)rows_per_cta;
  float* GTs = Gs + (size_t)rows_per_cta * B;
  float mx = NR_NEG_INF, mn = INFINITY;
  for (int e = tid; e < nr * B; e += SKT_THREADS) {
    const float g = G[(size_t)r0 * B + e];
    mx = fmaxf(mx, g); mn = fminf(mn, g);
    Gs[e] = g; GTs[e] = GT[(size_t)r0 * B + e];
  }
  mx = block_max(mx, red);
  mn = -block_max(-mn, red);
  if (tid == 0) { gstat[2 * blockIdx.x] = mx; gstat[2 * blockIdx.x + 1] = mn; }
  grid_barrier(counter, gridDim.x);                 // the only device-wide barrier: global max / min of G
  mx = NR_NEG_INF; mn = INFINITY;
  for (int c = tid; c < (int)gridDim.x; c += SKT_THREADS) {
    mx = fmaxf(mx, __ldcg(gstat + 2 * c)); mn = fminf(mn, __ldcg(gstat + 2 * c + 1));
  }
  mx = block_max(mx, red);
  mn = -block_max(-mn, red);
  const float nu = -logf(2.0f * (float)B);
  const bool scaling = (mx - mn) <= 30.f;
  if (scaling)
    for (int e = tid; e < nr * B; e += SKT_THREADS) { Gs[e] = expf(Gs[e] - mx); GTs[e] = expf(GTs[e] - mx); }
  const float init = scaling ? 1.f : 0.f;
  for (int e = tid; e < 4 * rows_per_cta; e += SKT_THREADS) mine[e] = init;
  __syncthreads();
  unsigned int epoch = 0;                           // epoch of the values the next half-iteration consumes
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      // half 0: (u1 <- rows of G with v1, u2 <- rows of G^T with v2); half 1: (v1 <- G^T with u1, v2 <- G with u2)
      const uint2* s1 = tv + (size_t)(half ? 0 : 1) * B;
      const uint2* s2 = tv + (size_t)(half ? 2 : 3) * B;
      if (epoch == 0) {
        for (int j = tid; j < B; j += SKT_THREADS) { vs[j] = init; vs[B + j] = init; }
      } else {
        for (int j = tid; j < B; j += SKT_THREADS) { vs[j] = wait_tagged(s1 + j, epoch); vs[B + j] = wait_tagged(s2 + j, epoch); }
      }
      __syncthreads();
      ++epoch;
      for (int w = warp; w < 2 * nr; w += SKT_WARPS) {
        const int r = w >> 1, chain = w & 1;
        const bool useT = (chain ^ half) != 0;
        const float* row = (useT ? GTs : Gs) + (size_t)r * B;
        float val;
        if (scaling) val = 1.0f / warp_dot(row, vs + chain * B, B, lane);
        else val = nu - warp_lse(row, vs + chain * B, B, lane);
        if (lane == 0) {
          const int which = chain * 2 + (half ? 1 : 0);
          publish_tagged(tv + (size_t)which * B + r0 + r, val, epoch);
          mine[which * rows_per_cta + r] = val;
        }
      }
      __syncthreads();                              // vs is rewritten by the next half-iteration
    }
  }
  // own rows of the duals; scaling domain -> log duals: u = (nu - max G) + log alpha, v = log beta
  for (int r = tid; r < nr; r += SKT_THREADS) {
    const float a1 = mine[0 * rows_per_cta + r], b1 = mine[1 * rows_per_cta + r];
    const float a2 = mine[2 * rows_per_cta + r], b2 = mine[3 * rows_per_cta + r];
    u1[r0 + r] = scaling ? (nu - mx) + logf(a1) : a1; v1[r0 + r] = scaling ? logf(b1) : b1;
    u2[r0 + r] = scaling ? (nu - mx) + logf(a2) : a2; v2[r0 + r] = scaling ? logf(b2) : b2;
  }
}

// ---- single-cluster variant (B <= ~600): the whole problem lives in the shared memory of one thread-block
// cluster (<= 16 CTAs); the dual vectors are replicated in every CTA's shared memory and updated with DSMEM
// stores, and half-iterations are separated by the hardware cluster barrier instead of a global-memory one.
namespace cg = cooperative_groups;

__device__ __forceinline__ float warp_row_lse_smem(const float* __restrict__ row, const float* __restrict__ vec,
                                                   int B, int lane) {
  float m = NR_NEG_INF;
  for (int j = lane; j < B; j += 32) m = fmaxf(m, row[j] + vec[j]);
  m = warp_max(m);
  float s = 0.f;
  for (int j = lane; j < B; j += 32) s += expf(row[j] + vec[j] - m);
  s = warp_sum(s);
  return m + logf(s);
}

__global__ void __launch_bounds__(SK_THREADS)
sinkhorn_cluster_kernel(const float* __restrict__ G, const float* __restrict__ GT, int B, int iters,
                        int rows_per_cta, float* u1, float* v1, float* u2, float* v2) {
  cg::cluster_group cluster = cg::this_cluster();
  extern __shared__ float sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int rank = (int)cluster.block_rank(), ncta = (int)cluster.num_blocks();
  const int r0 = rank * rows_per_cta;
  const int nr = max(0, min(rows_per_cta, B - r0));
  float* Gs = sm;
  float* GTs = Gs + (size_t)rows_per_cta * B;
  float* vec = GTs + (size_t)rows_per_cta * B;      // [4][B]: u1, v1, u2, v2 (replica of the full vectors)
  for (int e = tid; e < nr * B; e += SK_THREADS) {
    Gs[e] = G[(size_t)r0 * B + e];
    GTs[e] = GT[(size_t)r0 * B + e];
  }
  for (int e = tid; e < 4 * B; e += SK_THREADS) vec[e] = 0.f;
  const float nu = -logf(2.0f * (float)B);
  cluster.sync();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      // half 0: u1 <- rows of G with v1, u2 <- rows of GT with v2;  half 1: v1 <- rows of GT with u1, v2 <- rows of G with u2
      for (int w = warp; w < 2 * nr; w += SK_WARPS) {
        const int r = w >> 1, chain = w & 1;
        const float* row = ((chain ^ half) ? GTs : Gs) + (size_t)r * B;
        const int src = chain * 2 + (half ? 0 : 1), dst = chain * 2 + (half ? 1 : 0);
        const float val = nu - warp_row_lse_smem(row, vec + src * B, B, lane);
        for (int c = lane; c < ncta; c += 32) {
          float* remote = cluster.map_shared_rank(vec, c);
          remote[dst * B + r0 + r] = val;
        }
      }
      cluster.sync();
    }
  }
  for (int r = tid; r < nr; r += SK_THREADS) {
    u1[r0 + r] = vec[0 * B + r0 + r];
    v1[r0 + r] = vec[1 * B + r0 + r];
    u2[r0 + r] = vec[2 * B + r0 + r];
    v2[r0 + r] = vec[3 * B + r0 + r];
  }
}

// ---- single-CTA variant (B <= 168): K = exp(G - max G) and its transpose stay in shared memory and the
// recursion runs in the SCALING domain, alpha = 1/(K beta), beta = 1/(K^T alpha): no transcendental and no
// cross-CTA barrier inside the 100 half-iterations, only 4 FMAs + a warp reduction per row and __syncthreads.
// With alpha = e^u / e^(nu - gmax) and beta = e^v this is exactly u = nu - LSE_row(G + v), v = nu - LSE_col(G + u)
// (until_module.py:248-250).  The scaling domain needs a bounded dynamic range: when max G - min G > 30 the same
// kernel runs the log-domain recursion instead (identical to the reference in every regime).
constexpr int SKC_THREADS = 1024;
constexpr int SKC_WARPS = SKC_THREADS / 32;

__global__ void __launch_bounds__(SKC_THREADS)
sinkhorn_cta_kernel(const float* __restrict__ G, const float* __restrict__ GT, int B, int iters, float* u1, float* v1,
                    float* u2, float* v2) {
  extern __shared__ float sm[];
  __shared__ float red[32];
  float* K = sm;                       // [B][B]  rows of G
  float* KT = sm + (size_t)B * B;      // [B][B]  rows of G^T
  float* vec = KT + (size_t)B * B;     // [4][B]  chain 1: (alpha|u, beta|v), chain 2: (alpha|u, beta|v)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float mx = NR_NEG_INF, mn = INFINITY;
  for (int e = tid; e < B * B; e += SKC_THREADS) {
    float g = G[e];
    K[e] = g;
    KT[e] = GT[e];
    mx = fmaxf(mx, g);
    mn = fminf(mn, g);
  }
  mx = block_max(mx, red);
  mn = -block_max(-mn, red);
  const float nu = -logf(2.0f * (float)B);
  const bool scaling = (mx - mn) <= 30.f;
  if (scaling) {
    for (int e = tid; e < B * B; e += SKC_THREADS) { K[e] = expf(K[e] - mx); KT[e] = expf(KT[e] - mx); }
    for (int e = tid; e < 4 * B; e += SKC_THREADS) vec[e] = 1.f;       // beta = e^0
  } else {
    for (int e = tid; e < 4 * B; e += SKC_THREADS) vec[e] = 0.f;       // u = v = 0
  }
  __syncthreads();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      for (int w = warp; w < 2 * B; w += SKC_WARPS) {
        const int chain = w >= B, r = w - chain * B;
        const float* row = ((chain ^ half) ? KT : K) + (size_t)r * B;
        const float* src = vec + (chain * 2 + (half ? 0 : 1)) * B;
        float* dst = vec + (chain * 2 + (half ? 1 : 0)) * B;
        if (scaling) {
          float s = 0.f;
          for (int j = lane; j < B; j += 32) s = fmaf(row[j], src[j], s);
          s = warp_sum(s);
          if (lane == 0) dst[r] = 1.0f / s;
        } else {
          float m = NR_NEG_INF;
          for (int j = lane; j < B; j += 32) m = fmaxf(m, row[j] + src[j]);
          m = warp_max(m);
          float s = 0.f;
          for (int j = lane; j < B; j += 32) s += expf(row[j] + src[j] - m);
          s = warp_sum(s);
          if (lane == 0) dst[r] = nu - (m + logf(s));
        }
      }
      __syncthreads();
    }
  }
  for (int r = tid; r < B; r += SKC_THREADS) {
    if (scaling) {
      u1[r] = (nu - mx) + logf(vec[0 * B + r]); v1[r] = logf(vec[1 * B + r]);
      u2[r] = (nu - mx) + logf(vec[2 * B + r]); v2[r] = logf(vec[3 * B + r]);
    } else {
      u1[r] = vec[0 * B + r]; v1[r] = vec[1 * B + r]; u2[r] = vec[2 * B + r]; v2[r] = vec[3 * B + r];
    }
  }
}

// ---- register-resident single-CTA variant (B <= 128): warp w owns rows {w, w+32, ...} of K = exp(G - max G) and
// of K^T, NB x NB values each, IN REGISTERS for all 100 half-iterations; only the four length-B scaling vectors
// live in shared memory.  A half-iteration is 2*NB*NB FMAs + 2*NB warp reductions per thread and one
// __syncthreads: no shared/L2 matrix traffic at all, no cross-CTA barrier.
template <int NB>
__global__ void __launch_bounds__(1024)
sinkhorn_reg_kernel(const float* __restrict__ G, const float* __restrict__ GT, int B, int iters, float* u1, float* v1,
                    float* u2, float* v2) {
  __shared__ float vec[4][32 * NB];
  __shared__ float red[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float k[NB][NB], kt[NB][NB];
  float mx = NR_NEG_INF, mn = INFINITY;
#pragma unroll
  for (int i = 0; i < NB; ++i)
#pragma unroll
    for (int j = 0; j < NB; ++j) {
      const int r = warp + 32 * i, c = lane + 32 * j;
      const bool ok = r < B && c < B;
      k[i][j] = ok ? G[(size_t)r * B + c] : NR_NEG_INF;
      kt[i][j] = ok ? GT[(size_t)r * B + c] : NR_NEG_INF;
      if (ok) { mx = fmaxf(mx, k[i][j]); mn = fminf(mn, k[i][j]); }
    }
  mx = block_max(mx, red);
  mn = -block_max(-mn, red);
  const float nu = -logf(2.0f * (float)B);
  const bool scaling = (mx - mn) <= 30.f;
  if (scaling) {
#pragma unroll
    for (int i = 0; i < NB; ++i)
#pragma unroll
      for (int j = 0; j < NB; ++j) { k[i][j] = expf(k[i][j] - mx); kt[i][j] = expf(kt[i][j] - mx); }   // exp(-inf) = 0
  }
  for (int e = tid; e < 4 * 32 * NB; e += 1024) (&vec[0][0])[e] = scaling ? 1.f : 0.f;
  __syncthreads();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      // half 0: alpha1 <- rows of K with beta1, alpha2 <- rows of K^T with beta2
      // half 1: beta1  <- rows of K^T with alpha1, beta2 <- rows of K with alpha2
      const float* s1 = vec[half ? 0 : 1];
      const float* s2 = vec[half ? 2 : 3];
      float a1[NB], a2[NB];
#pragma unroll
      for (int j = 0; j < NB; ++j) { a1[j] = s1[lane + 32 * j]; a2[j] = s2[lane + 32 * j]; }
      if (scaling) {
        float V[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) V[i] = 0.f;
#pragma unroll
        for (int i = 0; i < NB; ++i) {
          float x1 = 0.f, x2 = 0.f;
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            x1 = fmaf(half ? kt[i][j] : k[i][j], a1[j], x1);
            x2 = fmaf(half ? k[i][j] : kt[i][j], a2[j], x2);
          }
          V[i] = x1; V[4 + i] = x2;
        }
        // transposed warp reduction of the 8 partial sums: 9 shuffles instead of 40; lane l ends up with the
        // total of value (l >> 2)
        const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
        float W[4], X[2], Y;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float send = b4 ? V[i] : V[i + 4], keep = b4 ? V[i + 4] : V[i];
          W[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const float send = b3 ? W[i] : W[i + 2], keep = b3 ? W[i + 2] : W[i];
          X[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
        }
        {
          const float send = b2 ? X[0] : X[1], keep = b2 ? X[1] : X[0];
          Y = keep + __shfl_xor_sync(0xffffffffu, send, 4);
        }
        Y += __shfl_xor_sync(0xffffffffu, Y, 2);
        Y += __shfl_xor_sync(0xffffffffu, Y, 1);
        if ((lane & 3) == 0) {
          const int idx = lane >> 2, i = idx & 3, r = warp + 32 * i;
          if (i < NB && r < B) vec[(idx < 4) ? (half ? 1 : 0) : (half ? 3 : 2)][r] = 1.0f / Y;
        }
      } else {
        float r1[NB], r2[NB];
#pragma unroll
        for (int i = 0; i < NB; ++i) {
          float m1 = NR_NEG_INF, m2 = NR_NEG_INF;
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            m1 = fmaxf(m1, (half ? kt[i][j] : k[i][j]) + a1[j]);
            m2 = fmaxf(m2, (half ? k[i][j] : kt[i][j]) + a2[j]);
          }
          m1 = warp_max(m1); m2 = warp_max(m2);
          float x1 = 0.f, x2 = 0.f;
#pragma unroll
          for (int j = 0; j < NB; ++j) {
            x1 += expf((half ? kt[i][j] : k[i][j]) + a1[j] - m1);
            x2 += expf((half ? k[i][j] : kt[i][j]) + a2[j] - m2);
          }
          x1 = warp_sum(x1); x2 = warp_sum(x2);
          r1[i] = nu - (m1 + logf(x1)); r2[i] = nu - (m2 + logf(x2));
        }
        if (lane == 0) {
#pragma unroll
          for (int i = 0; i < NB; ++i) {
            if (warp + 32 * i < B) {
              vec[half ? 1 : 0][warp + 32 * i] = r1[i];
              vec[half ? 3 : 2][warp + 32 * i] = r2[i];
            }
          }
        }
      }
      __syncthreads();
    }
  }
  for (int r = tid; r < B; r += 1024) {
    if (scaling) {
      u1[r] = (nu - mx) + logf(vec[0][r]); v1[r] = logf(vec[1][r]);
      u2[r] = (nu - mx) + logf(vec[2][r]); v2[r] = logf(vec[3][r]);
    } else {
      u1[r] = vec[0][r]; v1[r] = vec[1][r]; u2[r] = vec[2][r]; v2[r] = vec[3][r];
    }
  }
}

}  // namespace nr

using namespace nr;

static int sinkhorn_cluster_launch(const float* G, const float* GT, int B, int iters, float* u1, float* v1,
                                   float* u2, float* v2, cudaStream_t s, bool* launched) {
  *launched = false;
  for (int cs = 16; cs >= 8; cs >>= 1) {
    int rows_per_cta = (B + cs - 1) / cs;
    size_t smem = ((size_t)2 * rows_per_cta * B + (size_t)4 * B) * sizeof(float);
    if (smem > 200 * 1024) continue;
    if (cs > 8 && cudaFuncSetAttribute(sinkhorn_cluster_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) !=
                      cudaSuccess) {
      cudaGetLastError();
      continue;
    }
    if (smem > 48 * 1024)
      NR_CUDA(cudaFuncSetAttribute(sinkhorn_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(cs);
    cfg.blockDim = dim3(SK_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nclusters = 0;
    if (cudaOccupancyMaxActiveClusters(&nclusters, sinkhorn_cluster_kernel, &cfg) != cudaSuccess || nclusters < 1) {
      cudaGetLastError();
      continue;
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, sinkhorn_cluster_kernel, G, GT, B, iters, rows_per_cta, u1, v1, u2, v2);
    if (e != cudaSuccess) {
      cudaGetLastError();
      continue;
    }
    *launched = true;
    return 0;
  }
  return 0;
}

// [0,256): barrier counter; [256, 256+8K): per-CTA max/min; then 4*B tagged {value, epoch} words
static size_t sinkhorn_base_bytes(int64_t B) { return 256 + 2 * 1024 * sizeof(float) + (size_t)4 * (size_t)B * 8; }
// batches whose slabs do not stay in shared memory (B > ~1900) stream exp(G - max G) and its transpose from the
// workspace instead of re-evaluating exp(G + v) in every one of the 100 half-iterations: 2 * B^2 floats more
static size_t sinkhorn_stream_offset(int64_t B) { return (sinkhorn_base_bytes(B) + 255) / 256 * 256; }
extern "C" size_t nr_sinkhorn_workspace_bytes(int64_t B) {
  return B > 1536 ? sinkhorn_stream_offset(B) + (size_t)2 * (size_t)B * (size_t)B * sizeof(float) : sinkhorn_base_bytes(B);
}

static int sinkhorn_impl(const float* G, const float* GT, int64_t B, int iters, float* u1, float* v1, float* u2,
                         float* v2, void* workspace, size_t workspace_bytes, int rows_hint, void* stream);

extern "C" int nr_sinkhorn(const float* G, const float* GT, int64_t B, int iters, float* u1, float* v1, float* u2,
                           float* v2, void* workspace, size_t workspace_bytes, void* stream) {
  return sinkhorn_impl(G, GT, B, iters, u1, v1, u2, v2, workspace, workspace_bytes, 0, stream);
}

/* rows_per_cta > 0: matrix rows per CTA of the multi-CTA variants.  The chain is latency-bound (its CTAs mostly wait
 * for each other), so inside a head step — where the token-pair contraction wants every SM the chain does not hold —
 * fewer, larger CTAs make the STEP faster although the chain alone gets slower (measured at 2 ranks: B = 256 with 16
 * rows per CTA 705 us per step vs 744 us with 4; B = 1024 with 16 rows 2.21 ms vs 2.40 ms with 8). */
extern "C" int nr_sinkhorn_ex(const float* G, const float* GT, int64_t B, int iters, float* u1, float* v1, float* u2,
                              float* v2, void* workspace, size_t workspace_bytes, int rows_per_cta, void* stream) {
  return sinkhorn_impl(G, GT, B, iters, u1, v1, u2, v2, workspace, workspace_bytes, rows_per_cta, stream);
}

static int sinkhorn_impl(const float* G, const float* GT, int64_t B, int iters, float* u1, float* v1, float* u2,
                         float* v2, void* workspace, size_t workspace_bytes, int rows_hint, void* stream) {
  NR_CHECK_ARG(G && GT && u1 && v1 && u2 && v2 && workspace && B > 0 && iters >= 0, "nr_sinkhorn: bad arguments");
  NR_CHECK_ARG(workspace_bytes >= nr_sinkhorn_workspace_bytes(B), "nr_sinkhorn: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  {
    const size_t smem1 = ((size_t)2 * B * B + (size_t)4 * B) * sizeof(float);
    const char* var = getenv("NR_SINKHORN_VARIANT");      // tuning knob: "cta" | "cluster" | "grid"
    const bool want_cta = var && !strcmp(var, "cta");
    if (B <= 128 && (!var || !strcmp(var, "reg"))) {
      const int nb = (int)((B + 31) / 32);
      switch (nb) {
        case 1: sinkhorn_reg_kernel<1><<<1, 1024, 0, s>>>(G, GT, (int)B, iters, u1, v1, u2, v2); break;
        case 2: sinkhorn_reg_kernel<2><<<1, 1024, 0, s>>>(G, GT, (int)B, iters, u1, v1, u2, v2); break;
        case 3: sinkhorn_reg_kernel<3><<<1, 1024, 0, s>>>(G, GT, (int)B, iters, u1, v1, u2, v2); break;
        default: sinkhorn_reg_kernel<4><<<1, 1024, 0, s>>>(G, GT, (int)B, iters, u1, v1, u2, v2); break;
      }
      NR_CHECK_LAUNCH("nr_sinkhorn(reg)");
      return 0;
    }
    if (smem1 <= 225 * 1024 && iters >= 1 && want_cta) {
      if (smem1 > 48 * 1024)
        NR_CUDA(cudaFuncSetAttribute(sinkhorn_cta_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
      sinkhorn_cta_kernel<<<1, SKC_THREADS, smem1, s>>>(G, GT, (int)B, iters, u1, v1, u2, v2);
      NR_CHECK_LAUNCH("nr_sinkhorn(cta)");
      return 0;
    }
  }
  const char* var2 = getenv("NR_SINKHORN_VARIANT");
  if (var2 && !strcmp(var2, "cluster")) {
    bool launched = false;
    if (int e = sinkhorn_cluster_launch(G, GT, (int)B, iters, u1, v1, u2, v2, s, &launched)) return e;
    if (launched) return 0;
  }
  int dev = 0, sms = 0, coop = 0;
  NR_CUDA(cudaGetDevice(&dev));
  NR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  NR_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  NR_CHECK_ARG(coop, "nr_sinkhorn: device lacks cooperative launch");
  // rows per CTA, measured on B200 with the data-carrying exchange (us at 4 / 8 / 16 rows): B=256 142 / 148 / 172,
  // B=512 335 / 185 / 217, B=1024 319 / 265 / 336 -> 4 rows up to B=383, 8 rows beyond (more if the grid would not
  // be co-resident, see below)
  int rows_per_cta = B < 384 ? 4 : 8;
  if (rows_hint > 0) rows_per_cta = rows_hint;
  if (const char* rv = getenv("NR_SINKHORN_ROWS")) { int r2 = atoi(rv); if (r2 >= 1) rows_per_cta = r2; }   // tuning knob
  while (rows_per_cta > 4 && ((size_t)2 * rows_per_cta * B + (size_t)2 * B) * sizeof(float) > 200 * 1024 &&
         (B + rows_per_cta - 2) / (rows_per_cta - 1) <= sms)
    --rows_per_cta;
  int grid = (int)((B + rows_per_cta - 1) / rows_per_cta);
  if (grid > sms) {
    grid = sms;
    rows_per_cta = (int)((B + grid - 1) / grid);
    grid = (int)((B + rows_per_cta - 1) / rows_per_cta);
  }
  size_t smem = ((size_t)2 * rows_per_cta * B + (size_t)2 * B) * sizeof(float);
  int resident = smem <= 200 * 1024;
  int Bi = (int)B;
  unsigned int* counter = (unsigned int*)workspace;
  float* gstat = (float*)((char*)workspace + 256);
  const char* var3 = getenv("NR_SINKHORN_VARIANT");
  if (resident && grid <= 1024 && !(var3 && !strcmp(var3, "grid"))) {
    // data-carrying exchange instead of device-wide barriers (sinkhorn_tag_kernel)
    uint2* tv = (uint2*)((char*)workspace + 256 + 2 * 1024 * sizeof(float));
    const size_t smem_t = smem + (size_t)4 * rows_per_cta * sizeof(float);
    if (smem_t > 48 * 1024)
      NR_CUDA(cudaFuncSetAttribute(sinkhorn_tag_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_t));
    NR_CUDA(cudaMemsetAsync(workspace, 0, sinkhorn_base_bytes(B), s));              // counter and every epoch tag
    void* targs[] = {(void*)&G, (void*)&GT, (void*)&Bi, (void*)&iters, (void*)&rows_per_cta, (void*)&u1, (void*)&v1,
                     (void*)&u2, (void*)&v2, (void*)&counter, (void*)&gstat, (void*)&tv};
    NR_CUDA(cudaLaunchCooperativeKernel((void*)sinkhorn_tag_kernel, dim3(grid), dim3(SKT_THREADS), targs, smem_t, s));
    return 0;
  }
  if (!resident && B > 1536 && B % 4 == 0 && !(var3 && !strcmp(var3, "grid"))) {
    // slabs do not fit shared memory: stream exp(G - max) and its transpose from the workspace
    float* K = (float*)((char*)workspace + sinkhorn_stream_offset(B));
    float* KT = K + (size_t)B * (size_t)B;
    const size_t smem_s = (size_t)2 * B * sizeof(float);
    if (smem_s > 48 * 1024)
      NR_CUDA(cudaFuncSetAttribute(sinkhorn_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_s));
    int gs = sms;
    int rows_s = (int)((B + gs - 1) / gs);
    gs = (int)((B + rows_s - 1) / rows_s);
    NR_CUDA(cudaMemsetAsync(workspace, 0, 256, s));
    void* sargs[] = {(void*)&G, (void*)&GT, (void*)&Bi, (void*)&iters, (void*)&rows_s, (void*)&u1, (void*)&v1, (void*)&u2,
                     (void*)&v2, (void*)&counter, (void*)&gstat, (void*)&K, (void*)&KT};
    NR_CUDA(cudaLaunchCooperativeKernel((void*)sinkhorn_stream_kernel, dim3(gs), dim3(SKS_THREADS), sargs, smem_s, s));
    return 0;
  }
  if (!resident) smem = (size_t)2 * B * sizeof(float);
  if (smem > 48 * 1024)
    NR_CUDA(cudaFuncSetAttribute(sinkhorn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  NR_CUDA(cudaMemsetAsync(workspace, 0, 256, s));
  void* args[] = {(void*)&G, (void*)&GT, (void*)&Bi, (void*)&iters, (void*)&rows_per_cta, (void*)&resident,
                  (void*)&u1, (void*)&v1, (void*)&u2, (void*)&v2, (void*)&counter, (void*)&gstat};
  NR_CUDA(cudaLaunchCooperativeKernel((void*)sinkhorn_kernel, dim3(grid), dim3(SK_THREADS), args, smem, s));
  return 0;
}
