// Evaluation ranking and memory-bank FIFO kernels (HBM-bound, integer-exact).
//
// Reference: RetrievalMetrics.compute_metrics, NeighborRetr/utils/metrics.py:38-79 — the rank of the
// positive is found by a full np.sort of every row; the no-tie identity cols[i] = #{j: S[i,j] > S[i,i]}
// and its tie expansion turn that into one counting pass over the row:
// 4*Q*N bytes read once, comparison-only, so ranks are bit-exact for the same fp32 matrix.
// Memory bank: NeighborRetr.update_memory_bank, NeighborRetr/models/modeling.py:222-249.
#include "common.cuh"
#include "nrhead_internal.h"

namespace nr {

__global__ void __launch_bounds__(256)
rank_count_kernel(const float* __restrict__ S, int64_t lds, int N, const float* __restrict__ diag,
                  int64_t diag_col0, int32_t* __restrict__ gt, int32_t* __restrict__ eq) {
  __shared__ int red_g[8], red_e[8];
  const int q = blockIdx.x, tid = threadIdx.x;
  const float* row = S + (int64_t)q * lds;
  const float sd = diag ? diag[q] : row[diag_col0 + q];
  int g = 0, e = 0;
  const int n4 = ((reinterpret_cast<uintptr_t>(row) & 15) == 0) ? (N / 4) : 0;
  for (int j = tid; j < n4; j += 256) {
    float4 v = reinterpret_cast<const float4*>(row)[j];
    g += (v.x > sd) + (v.y > sd) + (v.z > sd) + (v.w > sd);
    e += (v.x == sd) + (v.y == sd) + (v.z == sd) + (v.w == sd);
  }
  for (int j = n4 * 4 + tid; j < N; j += 256) {
    float v = row[j];
    g += v > sd;
    e += v == sd;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    g += __shfl_xor_sync(0xffffffffu, g, o);
    e += __shfl_xor_sync(0xffffffffu, e, o);
  }
  if ((tid & 31) == 0) { red_g[tid >> 5] = g; red_e[tid >> 5] = e; }
  __syncthreads();
  if (tid == 0) {
    int G = 0, E = 0;
    for (int w = 0; w < 8; ++w) { G += red_g[w]; E += red_e[w]; }
    gt[q] += G;      // one CTA per q and stream-ordered launches: plain accumulate
    eq[q] += E;
  }
}

// per-row top-k by k rounds of block arg-max over a shared-memory copy (ties -> lower column)
__global__ void __launch_bounds__(256)
topk_rows_kernel(const float* __restrict__ S, int64_t lds, int N, int k, int32_t col_offset,
                 float* __restrict__ vals, int32_t* __restrict__ idx) {
  extern __shared__ __align__(16) float rowsm[];
  __shared__ unsigned long long red64[32];
  const int q = blockIdx.x, tid = threadIdx.x;
  const float* row = S + (int64_t)q * lds;
  // every thread stages its columns (j = tid, tid + 256, ...) and keeps the best key among them in a register; a
  // selection round is one block reduction of those keys, and only the owner of the winner rescans its columns
  // (k full passes over the staged row made this kernel 0.10 of the HBM stream)
  unsigned long long mine = 0ull;
  stage_row<256>(row, rowsm, N, tid);
  __syncthreads();
  for (int j = tid; j < N; j += 256) {
    const float v = rowsm[j];
    if (v != NR_NEG_INF) {
      const unsigned long long key = argmax_key(v, (uint32_t)j);
      mine = key > mine ? key : mine;
    }
  }
  for (int r = 0; r < k; ++r) {
    const unsigned long long best = block_max_u64(mine, red64);
    if (best == 0ull) {                 // fewer than k finite entries
      if (tid == 0) { vals[(int64_t)q * k + r] = NR_NEG_INF; idx[(int64_t)q * k + r] = -1; }
      continue;
    }
    const int j = (int)key_index(best);
    if (tid == 0) { vals[(int64_t)q * k + r] = argmax_key_value(best); idx[(int64_t)q * k + r] = j + col_offset; }
    if ((j & 255) == tid) {             // only the owner ever reads its columns again
      rowsm[j] = NR_NEG_INF;
      mine = 0ull;
      for (int jj = tid; jj < N; jj += 256) {
        const float v = rowsm[jj];
        if (v != NR_NEG_INF) {
          const unsigned long long key = argmax_key(v, (uint32_t)jj);
          mine = key > mine ? key : mine;
        }
      }
    }
  }
}

// merge W lists [W,Q,k] (global column ids) into the global top-k; ties -> lower global column
__global__ void __launch_bounds__(256)
topk_merge_kernel(const float* __restrict__ vals, const int32_t* __restrict__ idx, int W, int Q, int k,
                  float* __restrict__ out_vals, int32_t* __restrict__ out_idx) {
  extern __shared__ unsigned char mergesm[];
  __shared__ unsigned long long red64[32];
  const int q = blockIdx.x, tid = threadIdx.x, n = W * k;
  float* cv = reinterpret_cast<float*>(mergesm);
  int32_t* ci = reinterpret_cast<int32_t*>(cv + n);
  for (int e = tid; e < n; e += 256) {
    int w = e / k, r = e % k;
    cv[e] = vals[((int64_t)w * Q + q) * k + r];
    ci[e] = idx[((int64_t)w * Q + q) * k + r];
  }
  __syncthreads();
  for (int r = 0; r < k; ++r) {
    unsigned long long best = 0ull;
    int bpos = -1;
    for (int e = tid; e < n; e += 256) {
      if (ci[e] >= 0) {
        unsigned long long key = argmax_key(cv[e], (uint32_t)ci[e]);
        if (key > best) { best = key; bpos = e; }
      }
    }
    unsigned long long gbest = block_max_u64(best, red64);
    if (gbest == 0ull) {
      if (tid == 0) { out_vals[(int64_t)q * k + r] = NR_NEG_INF; out_idx[(int64_t)q * k + r] = -1; }
      continue;
    }
    if (best == gbest && bpos >= 0) {    // global column ids are unique, so exactly one owner
      out_vals[(int64_t)q * k + r] = cv[bpos];
      out_idx[(int64_t)q * k + r] = ci[bpos];
      ci[bpos] = -1;
    }
    __syncthreads();
  }
}

__global__ void fifo_update_kernel(const uint4* __restrict__ nw, int64_t n_new_units, const uint4* __restrict__ old,
                                   uint4* __restrict__ out, int64_t total_units) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total_units;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = i < n_new_units ? nw[i] : old[i - n_new_units];
}
__global__ void fifo_update_bytes_kernel(const unsigned char* __restrict__ nw, int64_t n_new_bytes,
                                         const unsigned char* __restrict__ old, unsigned char* __restrict__ out,
                                         int64_t total) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = i < n_new_bytes ? nw[i] : old[i - n_new_bytes];
}

}  // namespace nr

using namespace nr;

extern "C" int nr_rank_count(const float* S, int64_t lds, int64_t Q, int64_t N, const float* diag,
                             int64_t diag_col0, int32_t* gt, int32_t* eq, void* stream) {
  NR_CHECK_ARG(S && gt && eq && Q > 0 && N > 0, "nr_rank_count: bad arguments");
  if (!diag) NR_CHECK_ARG(diag_col0 >= 0 && diag_col0 + Q <= N, "nr_rank_count: diagonal outside the block");
  rank_count_kernel<<<(unsigned)Q, 256, 0, (cudaStream_t)stream>>>(S, lds, (int)N, diag, diag_col0, gt, eq);
  NR_CHECK_LAUNCH("nr_rank_count");
  return 0;
}

extern "C" int nr_topk_rows(const float* S, int64_t lds, int64_t Q, int64_t N, int k, int32_t col_offset,
                            float* vals, int32_t* idx, void* stream) {
  NR_CHECK_ARG(S && vals && idx && Q > 0 && N > 0 && k > 0, "nr_topk_rows: bad arguments");
  size_t smem = (size_t)N * sizeof(float);
  NR_CHECK_ARG(smem <= 200 * 1024, "nr_topk_rows: N=%lld too large for a shared-memory row", (long long)N);
  if (smem > 48 * 1024)
    NR_CUDA(cudaFuncSetAttribute(topk_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  topk_rows_kernel<<<(unsigned)Q, 256, smem, (cudaStream_t)stream>>>(S, lds, (int)N, k, col_offset, vals, idx);
  NR_CHECK_LAUNCH("nr_topk_rows");
  return 0;
}

extern "C" int nr_topk_merge(const float* vals, const int32_t* idx, int64_t W, int64_t Q, int k, float* out_vals,
                             int32_t* out_idx, void* stream) {
  NR_CHECK_ARG(vals && idx && out_vals && out_idx && W > 0 && Q > 0 && k > 0, "nr_topk_merge: bad arguments");
  size_t smem = (size_t)W * k * 8;
  NR_CHECK_ARG(smem <= 48 * 1024, "nr_topk_merge: W*k too large");
  topk_merge_kernel<<<(unsigned)Q, 256, smem, (cudaStream_t)stream>>>(vals, idx, (int)W, (int)Q, k, out_vals, out_idx);
  NR_CHECK_LAUNCH("nr_topk_merge");
  return 0;
}

extern "C" int nr_fifo_update(const void* new_rows, int64_t n_new, const void* old_rows, int64_t n_old, void* out,
                              int64_t capacity, int64_t row_bytes, void* stream) {
  NR_CHECK_ARG(out && capacity > 0 && row_bytes > 0 && n_new >= 0 && n_old >= 0, "nr_fifo_update: bad arguments");
  NR_CHECK_ARG(n_new + n_old >= capacity, "nr_fifo_update: fewer rows (%lld) than capacity (%lld)",
               (long long)(n_new + n_old), (long long)capacity);
  NR_CHECK_ARG(out != old_rows && out != new_rows, "nr_fifo_update: output must not alias an input");
  int64_t take_new = n_new < capacity ? n_new : capacity;
  int64_t total = capacity * row_bytes, nb = take_new * row_bytes;
  bool vec = (row_bytes % 16 == 0) && ((((uintptr_t)new_rows | (uintptr_t)old_rows | (uintptr_t)out) & 15) == 0);
  if (vec) {
    int64_t units = total / 16;
    int grid = (int)((units + 255) / 256 < 2368 ? (units + 255) / 256 : 2368);
    fifo_update_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)new_rows, nb / 16, (const uint4*)old_rows,
                                                               (uint4*)out, units);
  } else {
    int grid = (int)((total + 255) / 256 < 2368 ? (total + 255) / 256 : 2368);
    fifo_update_bytes_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const unsigned char*)new_rows, nb,
                                                                     (const unsigned char*)old_rows,
                                                                     (unsigned char*)out, total);
  }
  NR_CHECK_LAUNCH("nr_fifo_update");
  return 0;
}
