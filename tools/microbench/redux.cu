#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ int redux_max_s32(int v, unsigned mask) {
  int r; asm volatile("redux.sync.max.s32 %0, %1, %2;" : "=r"(r) : "r"(v), "r"(mask)); return r;
}
__device__ __forceinline__ float redux_max_f32(float v, unsigned mask) {
  float r; asm volatile("redux.sync.max.f32 %0, %1, %2;" : "=f"(r) : "f"(v), "r"(mask)); return r;
}
template <int MODE>
__global__ void k(int* out, long long* cyc, int iters) {
  int lane = threadIdx.x & 31;
  int v = threadIdx.x * 2654435761u + blockIdx.x;
  float f = __int_as_float((v & 0x007fffff) | 0x3f800000);
  unsigned m_lo = 0x00ffffffu, m_hi = 0xff000000u;
  unsigned mymask = lane < 24 ? m_lo : m_hi;
  int acc = 0; float facc = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      if (MODE == 0) { acc += redux_max_s32(v + u + i, 0xffffffffu); }
      if (MODE == 1) { acc += redux_max_s32(v + u + i, mymask); }            // two segments, divergent masks
      if (MODE == 2) { facc += redux_max_f32(f + (float)(u + i), 0xffffffffu); }
      if (MODE == 3) { acc += __shfl_xor_sync(0xffffffffu, v + u + i, 16); }
      if (MODE == 4) { acc += __ballot_sync(0xffffffffu, (v + u + i) & 1); }
      if (MODE == 5) { facc += redux_max_f32(f + (float)(u + i), mymask); }
      if (MODE == 6) { // two-seg via full-warp: each half with predicate select
        int a = redux_max_s32(lane < 24 ? v + u + i : INT_MIN, 0xffffffffu);
        int b = redux_max_s32(lane < 24 ? INT_MIN : v + u + i, 0xffffffffu);
        acc += a ^ b; }
    }
  }
  long long t1 = clock64();
  __syncthreads();
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[MODE] = t1 - t0;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc + (int)facc;
}
int main() {
  int* out; long long* cyc; cudaMalloc(&out, 148 * 1024 * 4); cudaMallocManaged(&cyc, 64);
  const int iters = 1000;
  for (int nw : {4, 8, 16}) {
    k<0><<<148, nw * 32>>>(out, cyc, iters); k<1><<<148, nw * 32>>>(out, cyc, iters); k<2><<<148, nw * 32>>>(out, cyc, iters);
    k<3><<<148, nw * 32>>>(out, cyc, iters); k<4><<<148, nw * 32>>>(out, cyc, iters); k<5><<<148, nw * 32>>>(out, cyc, iters);
    k<6><<<148, nw * 32>>>(out, cyc, iters);
    cudaError_t e = cudaDeviceSynchronize();
    printf("warps/CTA %d (%s)\n", nw, cudaGetErrorString(e));
    const char* names[] = {"redux.s32 full", "redux.s32 2seg divergent", "redux.f32 full", "shfl", "ballot", "redux.f32 2seg", "redux.s32 2x full select"};
    for (int m = 0; m < 7; ++m) printf("  %-28s %.2f cyc per warp-instr per SM (=> %.2f per warp)\n", names[m], (double)cyc[m] / (iters * 16.0 * nw) , (double)cyc[m] / (iters * 16.0));
  }
  return 0;
}
