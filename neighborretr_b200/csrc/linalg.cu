// Small exact-fp32 products of the head that are NOT token contractions (CUDA cores, latency-sized):
//   nr_matmul_f32   out[M,N] (+)= op(A)[M,K] X[K,N]   — the global-feature gradients dgT = dG gV, dgV = dG^T gT
//                   (autograd of reference NeighborRetr/models/modeling.py:516-539 for one global token per sample,
//                   where G = gT gV^T), 2*B*B*D flops;
//   nr_matvec_small out[r] = sum_c A[r,c] (x[c] + x2[c])  or its transpose — the 5x4 loss combination
//                   [total, centrality, uniform, neighbor, kl] = M54 (sums_dir1 + sums_dir2) (modeling.py:353-358)
//                   and its backward.
#include "common.cuh"
#include "nrhead_internal.h"

namespace nr {

// 64x64 output tile per CTA, 4x4 outputs per thread, k-chunks of 16 through shared memory
__global__ void __launch_bounds__(256)
matmul_f32_kernel(const float* __restrict__ A, int64_t lda, int transA, const float* __restrict__ X, int64_t ldx, int M,
                  int K, int N, float* __restrict__ out, int64_t ldo, int accumulate) {
  __shared__ float As[16][68], Xs[16][68];
  const int i0 = blockIdx.y * 64, j0 = blockIdx.x * 64, tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  float acc[4][4];
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < 4; ++q) acc[p][q] = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    if (transA) {             // A stored [K, M]
      const int kk = tid >> 4, m4 = (tid & 15) * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int m = i0 + m4 + q, k = k0 + kk;
        As[kk][m4 + q] = (m < M && k < K) ? A[(int64_t)k * lda + m] : 0.f;
      }
    } else {                  // A stored [M, K]
      const int r = tid >> 2, c4 = (tid & 3) * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int m = i0 + r, k = k0 + c4 + q;
        As[c4 + q][r] = (m < M && k < K) ? A[(int64_t)m * lda + k] : 0.f;
      }
    }
    {
      const int kk = tid >> 4, n4 = (tid & 15) * 4;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int n = j0 + n4 + q, k = k0 + kk;
        Xs[kk][n4 + q] = (n < N && k < K) ? X[(int64_t)k * ldx + n] : 0.f;
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Xs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[p][q] = fmaf(av[p], bv[q], acc[p][q]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int p = 0; p < 4; ++p)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int m = i0 + ty * 4 + p, n = j0 + tx * 4 + q;
      if (m < M && n < N) {
        float* o = out + (int64_t)m * ldo + n;
        *o = accumulate ? *o + acc[p][q] : acc[p][q];
      }
    }
}

__global__ void matvec_small_kernel(const float* __restrict__ A, int rows, int cols, int trans, const float* __restrict__ x,
                                    const float* __restrict__ x2, float* __restrict__ out) {
  const int r = threadIdx.x;
  const int nout = trans ? cols : rows, nin = trans ? rows : cols;
  if (r >= nout) return;
  float s = 0.f;
  for (int c = 0; c < nin; ++c) {
    const float a = trans ? A[c * cols + r] : A[r * cols + c];
    s = fmaf(a, x[c] + (x2 ? x2[c] : 0.f), s);
  }
  out[r] = s;
}

}  // namespace nr

extern "C" int nr_matmul_f32(const float* A, int64_t lda, int transA, const float* X, int64_t ldx, int64_t M, int64_t K,
                             int64_t N, float* out, int64_t ldo, int accumulate, void* stream) {
  NR_CHECK_ARG(A && X && out && M > 0 && K > 0 && N > 0, "nr_matmul_f32: bad arguments");
  dim3 grid((unsigned)((N + 63) / 64), (unsigned)((M + 63) / 64));
  nr::matmul_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(A, lda, transA, X, ldx, (int)M, (int)K, (int)N, out, ldo,
                                                              accumulate);
  NR_CHECK_LAUNCH("nr_matmul_f32");
  return 0;
}

extern "C" int nr_matvec_small(const float* A, int64_t rows, int64_t cols, int trans, const float* x, const float* x2,
                               float* out, void* stream) {
  NR_CHECK_ARG(A && x && out && rows > 0 && cols > 0 && rows <= 32 && cols <= 32, "nr_matvec_small: bad arguments (<= 32 x 32)");
  nr::matvec_small_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(A, (int)rows, (int)cols, trans, x, x2, out);
  NR_CHECK_LAUNCH("nr_matvec_small");
  return 0;
}
