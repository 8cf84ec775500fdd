// Multi-sentence evaluation kernels (several captions per video, MSVD-style test sets): HBM-bound,
// comparison-only, integer-exact.
//
// Reference: RetrievalMetrics.tensor_text_to_video_metrics (NeighborRetr/utils/metrics.py:81-122) ranks the
// video of every caption by a double argsort over a -inf padded [V, maxlen, V] copy of the similarity matrix;
// RetrievalMetrics.tensor_video_to_text_sim (:124-145) takes, for every video, the maximum over the captions
// of each caption group of that padded copy; the padding itself is NeighborRetr/training/evaluator.py:216-239.
// Both follow from the un-padded matrix S [T, V] (captions of a video contiguous) in one pass:
//   rank(t) = #{j : S[t,j] > S[t,c]} + #{j < c : S[t,j] == S[t,c]},  c = video of caption t
//             (descending stable argsort: equal scores keep their column order; NaN sorts in front)
//   M[j,i]  = max over the captions t of video i of S[t,j]            (NaN -> -inf as at :139)
// Bytes: 4*T*V read once by each kernel, 4*V*G written by the second.
#include "common.cuh"
#include "nrhead_internal.h"

namespace nr {

// counts of one row over the columns {first, first+stride, ...}: scalar head up to the first 16-byte boundary of the
// row (rows of a matrix whose width is not a multiple of 4 start at any 4-byte offset), float4 body, scalar tail
template <int STRIDE>
__device__ __forceinline__ void count_row(const float* __restrict__ row, int N, float sd, int64_t c, int first,
                                          int& g, int& e) {
  int head = (int)(((16u - (unsigned)(reinterpret_cast<uintptr_t>(row) & 15)) & 15) >> 2);
  head = head < N ? head : N;
  const int n4 = (N - head) / 4;
  const float4* body = reinterpret_cast<const float4*>(row + head);
#pragma unroll 4
  for (int j = first; j < n4; j += STRIDE) {
    float4 v = body[j];
    const int64_t j0 = head + 4 * (int64_t)j;
    g += (v.x > sd || v.x != v.x) + (v.y > sd || v.y != v.y) + (v.z > sd || v.z != v.z) + (v.w > sd || v.w != v.w);
    e += (v.x == sd && j0 < c) + (v.y == sd && j0 + 1 < c) + (v.z == sd && j0 + 2 < c) + (v.w == sd && j0 + 3 < c);
  }
  // head columns [0, head) and tail columns [head + 4*n4, N): at most 3 + 3 elements
  const int tail0 = head + 4 * n4;
  for (int j = first; j < head + (N - tail0); j += STRIDE) {
    const int col = j < head ? j : tail0 + (j - head);
    float v = row[col];
    g += (v > sd || v != v);
    e += (v == sd && col < c);
  }
}

__device__ __forceinline__ float positive_score(const float* __restrict__ row, int N, const float* diag, int64_t q,
                                                int64_t c) {
  if (diag) return diag[q];
  return (c >= 0 && c < N) ? row[c] : __int_as_float(0x7fc00000);   // outside the block: NaN -> invalid, counts 0
}

__device__ __forceinline__ void store_counts(int64_t q, float sd, int G, int E, int32_t* gt, int32_t* eq_before,
                                             int32_t* valid) {
  const bool ok = (sd == sd) && (fabsf(sd) != INFINITY);   // metrics.py:107-109: neither inf nor NaN
  if (ok) {                   // one writer per q and stream-ordered launches: plain accumulate (column shards)
    gt[q] += G;
    eq_before[q] += E;
  }
  if (valid) valid[q] = ok ? 1 : 0;
}

// few long rows: one CTA per caption row
__global__ void __launch_bounds__(256)
rank_target_kernel(const float* __restrict__ S, int64_t lds, int N, const int32_t* __restrict__ target,
                   const float* __restrict__ diag, int64_t col_offset, int32_t* __restrict__ gt,
                   int32_t* __restrict__ eq_before, int32_t* __restrict__ valid) {
  __shared__ int red_g[8], red_e[8];
  const int q = blockIdx.x, tid = threadIdx.x;
  const float* row = S + (int64_t)q * lds;
  const int64_t c = (int64_t)target[q] - col_offset;       // positive's column inside this block of columns
  const float sd = positive_score(row, N, diag, q, c);
  int g = 0, e = 0;
  count_row<256>(row, N, sd, c, tid, g, e);
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
    store_counts(q, sd, G, E, gt, eq_before, valid);
  }
}

// short rows, or many rows (the usual test sets: thousands of captions x a few hundred to a few thousand videos):
// one WARP per caption row, 8 rows per CTA, shuffles only — measured 243 us for 98k x 4096 (6.6 TB/s) against
// 411 us with a 256-thread CTA per row, which spends its time on launch and barrier overhead
__global__ void __launch_bounds__(256)
rank_target_warp_kernel(const float* __restrict__ S, int64_t lds, int64_t Q, int N,
                        const int32_t* __restrict__ target, const float* __restrict__ diag, int64_t col_offset,
                        int32_t* __restrict__ gt, int32_t* __restrict__ eq_before, int32_t* __restrict__ valid) {
  const int lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (q >= Q) return;                                      // warp-uniform; no block-level barrier below
  const float* row = S + q * lds;
  const int64_t c = (int64_t)target[q] - col_offset;
  const float sd = positive_score(row, N, diag, q, c);
  int g = 0, e = 0;
  count_row<32>(row, N, sd, c, lane, g, e);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    g += __shfl_xor_sync(0xffffffffu, g, o);
    e += __shfl_xor_sync(0xffffffffu, e, o);
  }
  if (lane == 0) store_counts(q, sd, g, e, gt, eq_before, valid);
}

// one warp per caption group, 128 columns per CTA (4 per lane), 8 groups per CTA; the [8 x 128] tile of maxima
// is written through shared memory so that the 8 groups of a column leave as one 32-byte segment
__global__ void __launch_bounds__(256)
group_max_t_kernel(const float* __restrict__ S, int64_t lds, int V, const int32_t* __restrict__ group_start, int G,
                   float* __restrict__ out, int64_t ldo) {
  __shared__ float tile[8][129];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int j0 = blockIdx.x * 128, i0 = blockIdx.y * 8, i = i0 + warp;
  float m0 = NR_NEG_INF, m1 = NR_NEG_INF, m2 = NR_NEG_INF, m3 = NR_NEG_INF;
  if (i < G) {
    const int t0 = group_start[i], t1 = group_start[i + 1];
    const int jl = j0 + lane * 4;
    const bool vec = ((lds & 3) == 0) && ((reinterpret_cast<uintptr_t>(S) & 15) == 0) && (jl + 3 < V);
    if (vec) {
      const float* p = S + (int64_t)t0 * lds + jl;
#pragma unroll 4
      for (int t = t0; t < t1; ++t, p += lds) {
        float4 v = *reinterpret_cast<const float4*>(p);
        m0 = v.x > m0 ? v.x : m0;        // NaN compares false: skipped, i.e. treated as -inf
        m1 = v.y > m1 ? v.y : m1;
        m2 = v.z > m2 ? v.z : m2;
        m3 = v.w > m3 ? v.w : m3;
      }
    } else if (jl < V) {
      // rows not 16-byte aligned (matrix width not a multiple of 4) or the last partial float4 of a row: the same
      // 4 consecutive columns per lane with scalar loads.  Measured on 98k x 4097: 356 us (4.7 TB/s); an interleaved
      // mapping (lane l -> columns l, l+32, l+64, l+96) was slower (574 us), so this one stays
      const float* p = S + (int64_t)t0 * lds + jl;
      const int nc = V - jl;              // 1..3 real columns (or >= 4 on an unaligned matrix)
      for (int t = t0; t < t1; ++t, p += lds) {
        float a = p[0];
        m0 = a > m0 ? a : m0;
        if (nc > 1) { float b = p[1]; m1 = b > m1 ? b : m1; }
        if (nc > 2) { float c = p[2]; m2 = c > m2 ? c : m2; }
        if (nc > 3) { float d = p[3]; m3 = d > m3 ? d : m3; }
      }
    }
  }
  tile[warp][lane * 4 + 0] = m0;
  tile[warp][lane * 4 + 1] = m1;
  tile[warp][lane * 4 + 2] = m2;
  tile[warp][lane * 4 + 3] = m3;
  __syncthreads();
  for (int e = threadIdx.x; e < 8 * 128; e += 256) {
    const int jj = e >> 3, ii = e & 7;
    if (j0 + jj < V && i0 + ii < G) out[(int64_t)(j0 + jj) * ldo + i0 + ii] = tile[ii][jj];
  }
}

}  // namespace nr

using namespace nr;

extern "C" int nr_rank_count_target(const float* S, int64_t lds, int64_t Q, int64_t N, const int32_t* target,
                                    const float* diag, int64_t col_offset, int32_t* gt, int32_t* eq_before,
                                    int32_t* valid, void* stream) {
  NR_CHECK_ARG(S && target && gt && eq_before && Q > 0 && N > 0 && lds >= N, "nr_rank_count_target: bad arguments");
  NR_CHECK_ARG(Q <= 2147483647LL && N <= 2147483647LL, "nr_rank_count_target: sizes exceed int32");
  // a warp per row needs enough rows to fill the machine (148 SMs x 64 warps); few long rows take a CTA per row
  if (N <= 2048 || Q >= 4736)
    rank_target_warp_kernel<<<(unsigned)((Q + 7) / 8), 256, 0, (cudaStream_t)stream>>>(S, lds, Q, (int)N, target, diag,
                                                                                      col_offset, gt, eq_before, valid);
  else
    rank_target_kernel<<<(unsigned)Q, 256, 0, (cudaStream_t)stream>>>(S, lds, (int)N, target, diag, col_offset, gt,
                                                                     eq_before, valid);
  NR_CHECK_LAUNCH("nr_rank_count_target");
  return 0;
}

extern "C" int nr_group_max_t(const float* S, int64_t lds, int64_t T, int64_t V, const int32_t* group_start,
                              int64_t G, float* out, int64_t ldo, void* stream) {
  NR_CHECK_ARG(S && group_start && out && T > 0 && V > 0 && G > 0 && lds >= V && ldo >= G,
               "nr_group_max_t: bad arguments");
  NR_CHECK_ARG(T <= 2147483647LL && V <= 2147483647LL, "nr_group_max_t: sizes exceed int32");
  const int64_t gy = (G + 7) / 8;
  NR_CHECK_ARG(gy <= 65535, "nr_group_max_t: more than 524280 caption groups");
  dim3 grid((unsigned)((V + 127) / 128), (unsigned)gy);
  group_max_t_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(S, lds, (int)V, group_start, (int)G, out, ldo);
  NR_CHECK_LAUNCH("nr_group_max_t");
  return 0;
}
