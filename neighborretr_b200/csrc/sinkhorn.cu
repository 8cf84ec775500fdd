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

// LSE over j of (row[j] + vec[j]); vec is read through L2 (written by other CTAs this launch)
__device__ __forceinline__ float warp_row_lse(const float* __restrict__ row, const float* vec, int B, int lane) {
  float m = NR_NEG_INF;
  for (int j = lane; j < B; j += 32) m = fmaxf(m, row[j] + __ldcg(vec + j));
  m = warp_max(m);
  float s = 0.f;
  for (int j = lane; j < B; j += 32) s += expf(row[j] + __ldcg(vec + j) - m);
  s = warp_sum(s);
  return m + logf(s);
}

__global__ void __launch_bounds__(SK_THREADS)
sinkhorn_kernel(const float* __restrict__ G, const float* __restrict__ GT, int B, int iters, int rows_per_cta,
                int resident, float* u1, float* v1, float* u2, float* v2, unsigned int* counter) {
  extern __shared__ float sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int r0 = blockIdx.x * rows_per_cta;
  const int nr = max(0, min(rows_per_cta, B - r0));
  float* Gs = sm;
  float* GTs = sm + (size_t)rows_per_cta * B;
  if (resident) {
    for (int e = tid; e < nr * B; e += SK_THREADS) {
      Gs[e] = G[(size_t)r0 * B + e];
      GTs[e] = GT[(size_t)r0 * B + e];
    }
  }
  // duals start at zero (until_module.py:243)
  for (int r = tid; r < nr; r += SK_THREADS) { u1[r0 + r] = 0.f; v1[r0 + r] = 0.f; u2[r0 + r] = 0.f; v2[r0 + r] = 0.f; }
  const float nu = -logf(2.0f * (float)B);
  unsigned int bar = 0;
  grid_barrier(counter, (++bar) * gridDim.x);
  for (int it = 0; it < iters; ++it) {
    // half-iteration A: u1 = nu - LSE_b(G[r,b] + v1[b]);   u2 = nu - LSE_b(GT[r,b] + v2[b])
    for (int w = warp; w < 2 * nr; w += SK_WARPS) {
      const int r = w >> 1, chain = w & 1;
      const float* row = resident ? ((chain ? GTs : Gs) + (size_t)r * B) : ((chain ? GT : G) + (size_t)(r0 + r) * B);
      float l = warp_row_lse(row, chain ? v2 : v1, B, lane);
      if (lane == 0) (chain ? u2 : u1)[r0 + r] = nu - l;
    }
    grid_barrier(counter, (++bar) * gridDim.x);
    // half-iteration B: v1[r] = nu - LSE_a(G[a,r] + u1[a]) = rows of GT;   v2[r] = rows of G with u2
    for (int w = warp; w < 2 * nr; w += SK_WARPS) {
      const int r = w >> 1, chain = w & 1;
      const float* row = resident ? ((chain ? Gs : GTs) + (size_t)r * B) : ((chain ? G : GT) + (size_t)(r0 + r) * B);
      float l = warp_row_lse(row, chain ? u2 : u1, B, lane);
      if (lane == 0) (chain ? v2 : v1)[r0 + r] = nu - l;
    }
    grid_barrier(counter, (++bar) * gridDim.x);
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

extern "C" size_t nr_sinkhorn_workspace_bytes(int64_t B) { (void)B; return 256; }

extern "C" int nr_sinkhorn(const float* G, const float* GT, int64_t B, int iters, float* u1, float* v1, float* u2,
                           float* v2, void* workspace, size_t workspace_bytes, void* stream) {
  NR_CHECK_ARG(G && GT && u1 && v1 && u2 && v2 && workspace && B > 0 && iters >= 0, "nr_sinkhorn: bad arguments");
  NR_CHECK_ARG(workspace_bytes >= 256, "nr_sinkhorn: workspace too small");
  cudaStream_t s = (cudaStream_t)stream;
  {
    bool launched = false;
    if (int e = sinkhorn_cluster_launch(G, GT, (int)B, iters, u1, v1, u2, v2, s, &launched)) return e;
    if (launched) return 0;
  }
  int dev = 0, sms = 0, coop = 0;
  NR_CUDA(cudaGetDevice(&dev));
  NR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  NR_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
  NR_CHECK_ARG(coop, "nr_sinkhorn: device lacks cooperative launch");
  // one warp per (row, chain): aim at SK_WARPS/2 rows per CTA, capped by the SM count
  int rows_per_cta = SK_WARPS / 2;
  int grid = (int)((B + rows_per_cta - 1) / rows_per_cta);
  if (grid > sms) {
    grid = sms;
    rows_per_cta = (int)((B + grid - 1) / grid);
    grid = (int)((B + rows_per_cta - 1) / rows_per_cta);
  }
  size_t smem = (size_t)2 * rows_per_cta * B * sizeof(float);
  int resident = smem <= 200 * 1024;
  if (!resident) smem = 0;
  if (smem > 48 * 1024)
    NR_CUDA(cudaFuncSetAttribute(sinkhorn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  NR_CUDA(cudaMemsetAsync(workspace, 0, 256, s));
  int Bi = (int)B;
  unsigned int* counter = (unsigned int*)workspace;
  void* args[] = {(void*)&G, (void*)&GT, (void*)&Bi, (void*)&iters, (void*)&rows_per_cta, (void*)&resident,
                  (void*)&u1, (void*)&v1, (void*)&u2, (void*)&v2, (void*)&counter};
  NR_CUDA(cudaLaunchCooperativeKernel((void*)sinkhorn_kernel, dim3(grid), dim3(SK_THREADS), args, smem, s));
  return 0;
}
