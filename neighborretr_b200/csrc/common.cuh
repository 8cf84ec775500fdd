// Shared device helpers for the NeighborRetr B200 retrieval-head kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <math.h>

#define NR_NEG_INF (-INFINITY)

namespace nr {

void set_error(const char* fmt, ...);

#define NR_CHECK_ARG(cond, ...)                         \
  do {                                                  \
    if (!(cond)) {                                      \
      nr::set_error(__VA_ARGS__);                       \
      return -1;                                        \
    }                                                   \
  } while (0)

#define NR_CHECK_LAUNCH(name)                                                   \
  do {                                                                          \
    cudaError_t e__ = cudaGetLastError();                                       \
    if (e__ != cudaSuccess) {                                                   \
      nr::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__));    \
      return -(int)e__ - 1000;                                                  \
    }                                                                           \
  } while (0)

#define NR_CUDA(call)                                                           \
  do {                                                                          \
    cudaError_t e__ = (call);                                                   \
    if (e__ != cudaSuccess) {                                                   \
      nr::set_error("%s failed: %s", #call, cudaGetErrorString(e__));           \
      return -(int)e__ - 1000;                                                  \
    }                                                                           \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ unsigned long long warp_max_u64(unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    unsigned long long t = __shfl_xor_sync(0xffffffffu, v, o);
    v = t > v ? t : v;
  }
  return v;
}

// Block-wide reductions; every thread gets the result.  `scratch` holds >= 32 elements.
// Safe to call back-to-back with the same scratch (leading barrier).
__device__ __forceinline__ float block_sum(float v, float* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  float r = (lane < nw) ? scratch[lane] : 0.f;
  return warp_sum(r);
}
__device__ __forceinline__ float block_max(float v, float* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  float r = (lane < nw) ? scratch[lane] : NR_NEG_INF;
  return warp_max(r);
}
__device__ __forceinline__ unsigned long long block_max_u64(unsigned long long v,
                                                            unsigned long long* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max_u64(v);
  __syncthreads();
  if (lane == 0) scratch[wid] = v;
  __syncthreads();
  unsigned long long r = (lane < nw) ? scratch[lane] : 0ull;
  return warp_max_u64(r);
}

// Order-preserving float -> uint32 (larger float => larger uint).
__device__ __forceinline__ uint32_t float_ord(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord_float(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}
// (value, index) key whose max is the largest value, ties towards the LOWER index.
__device__ __forceinline__ unsigned long long argmax_key(float v, uint32_t idx) {
  return ((unsigned long long)float_ord(v) << 32) | (unsigned long long)(0xffffffffu - idx);
}
// (value, index) key whose max is the SMALLEST value, ties towards the lower index.
__device__ __forceinline__ unsigned long long argmin_key(float v, uint32_t idx) {
  return ((unsigned long long)(~float_ord(v)) << 32) | (unsigned long long)(0xffffffffu - idx);
}
__device__ __forceinline__ uint32_t key_index(unsigned long long k) {
  return 0xffffffffu - (uint32_t)(k & 0xffffffffull);
}
__device__ __forceinline__ float argmax_key_value(unsigned long long k) {
  return ord_float((uint32_t)(k >> 32));
}
__device__ __forceinline__ float argmin_key_value(unsigned long long k) {
  return ord_float(~(uint32_t)(k >> 32));
}

// Arg-max over (ord, index) candidates — ord = float_ord(value), or ~float_ord(value) for an arg-min; (0, NR_NO_INDEX) =
// no candidate.  Largest ord wins, ties towards the LOWEST index: the order of argmax_key / argmin_key, but a warp
// reduces it with two redux.sync instructions instead of a 5-stage shuffle tree over 64-bit keys.
constexpr uint32_t NR_NO_INDEX = 0xffffffffu;
__device__ __forceinline__ void warp_argmax_ord(uint32_t& o, uint32_t& j) {
  const uint32_t m = __reduce_max_sync(0xffffffffu, o);
  j = __reduce_min_sync(0xffffffffu, o == m ? j : NR_NO_INDEX);
  o = m;
}
// block-wide; every thread gets the result.  `scratch` holds >= 32 uint2; safe back-to-back (leading barrier).
__device__ __forceinline__ void block_argmax_ord(uint32_t& o, uint32_t& j, uint2* scratch) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  warp_argmax_ord(o, j);
  __syncthreads();
  if (lane == 0) scratch[wid] = make_uint2(o, j);
  __syncthreads();
  const uint2 e = lane < nw ? scratch[lane] : make_uint2(0u, NR_NO_INDEX);
  o = e.x; j = e.y;
  warp_argmax_ord(o, j);
}

// Stage one row of n floats from global into shared memory with T threads: 16-byte loads, four per thread in flight
// before the first store (a scalar load-store loop keeps 2 x 128 B per warp in flight: with one or two CTAs per SM
// that is ~0.3 TB/s over the whole chip, measured on the B = 8192 row-loss kernels).  Falls back to scalar accesses
// when the source row or the destination is not 16-byte aligned.
template <int T>
__device__ __forceinline__ void stage_row(const float* __restrict__ src, float* dst, int n, int tid) {
  const bool vec = ((reinterpret_cast<uintptr_t>(src) | (uintptr_t)__cvta_generic_to_shared(dst)) & 15) == 0;
  int done = 0;
  if (vec) {
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    const int n4 = n >> 2;
    int j = tid;
    for (; j + 3 * T < n4; j += 4 * T) {
      const float4 a = __ldg(s4 + j), b = __ldg(s4 + j + T), c = __ldg(s4 + j + 2 * T), d = __ldg(s4 + j + 3 * T);
      d4[j] = a; d4[j + T] = b; d4[j + 2 * T] = c; d4[j + 3 * T] = d;
    }
    for (; j < n4; j += T) d4[j] = __ldg(s4 + j);
    done = n4 << 2;
  }
  for (int j = done + tid; j < n; j += T) dst[j] = src[j];
}

}  // namespace nr
