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

// 32x64 output tile per CTA (256 threads, 2x4 outputs each), k-chunks of 32 staged in shared memory with the next
// chunk's global loads in flight during the FMAs (one barrier per chunk).  These products are tiny (B = 128: 17
// MFLOP) and sit on the backward's critical path, so the kernel is sized for latency: many small CTAs.
__global__ void __launch_bounds__(256)
matmul_f32_kernel(const float* __restrict__ A, int64_t lda, int transA, const float* __restrict__ X, int64_t ldx, int M,
                  int K, int N, float* __restrict__ out, int64_t ldo, int accumulate) {
  __shared__ float As[2][32][33];      // [k][m]
  __shared__ float Xs[2][32][68];      // [k][n]
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 64, tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;               // outputs: rows ty*2..+1, columns tx*4..+3
  float acc[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
  float ra[4], rx[8];
  auto gload = [&](int k0) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {                    // A chunk: 32 (m) x 32 (k)
      const int e = it * 256 + tid;
      int m, k;
      if (transA) { k = e >> 5; m = e & 31; } else { m = e >> 5; k = e & 31; }
      const bool ok = (i0 + m < M) && (k0 + k < K);
      ra[it] = ok ? (transA ? A[(int64_t)(k0 + k) * lda + i0 + m] : A[(int64_t)(i0 + m) * lda + k0 + k]) : 0.f;
    }
#pragma unroll
    for (int it = 0; it < 8; ++it) {                    // X chunk: 32 (k) x 64 (n)
      const int e = it * 256 + tid, k = e >> 6, n = e & 63;
      rx[it] = (k0 + k < K && j0 + n < N) ? X[(int64_t)(k0 + k) * ldx + j0 + n] : 0.f;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int e = it * 256 + tid;
      int m, k;
      if (transA) { k = e >> 5; m = e & 31; } else { m = e >> 5; k = e & 31; }
      As[buf][k][m] = ra[it];
    }
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int e = it * 256 + tid;
      Xs[buf][e >> 6][e & 63] = rx[it];
    }
  };
  gload(0);
  sstore(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < K; k0 += 32) {
    const bool more = k0 + 32 < K;
    if (more) gload(k0 + 32);
#pragma unroll
    for (int k = 0; k < 32; ++k) {
      const float a0 = As[buf][k][ty * 2], a1 = As[buf][k][ty * 2 + 1];
      const float4 b = *reinterpret_cast<const float4*>(&Xs[buf][k][tx * 4]);
      acc[0][0] = fmaf(a0, b.x, acc[0][0]); acc[0][1] = fmaf(a0, b.y, acc[0][1]);
      acc[0][2] = fmaf(a0, b.z, acc[0][2]); acc[0][3] = fmaf(a0, b.w, acc[0][3]);
      acc[1][0] = fmaf(a1, b.x, acc[1][0]); acc[1][1] = fmaf(a1, b.y, acc[1][1]);
      acc[1][2] = fmaf(a1, b.z, acc[1][2]); acc[1][3] = fmaf(a1, b.w, acc[1][3]);
    }
    if (more) sstore(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }
#pragma unroll
  for (int p = 0; p < 2; ++p)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int m = i0 + ty * 2 + p, n = j0 + tx * 4 + q;
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
  dim3 grid((unsigned)((N + 63) / 64), (unsigned)((M + 31) / 32));
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
