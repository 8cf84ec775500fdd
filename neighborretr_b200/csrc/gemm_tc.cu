// tcgen05 GEMM of the token-weight MLPs (reference NeighborRetr/models/modeling.py:137-153, used :485-492):
//   forward   h = relu(x W1^T + b1)  [T,1024]   with the second layer's dot product  logits[t] += <h[t,:], w2>
//             folded into the epilogue (the hidden activations are written once, as bf16, only when a backward
//             will need them; in evaluation they never leave the SM);
//   backward  dx  = dh W1            [T,512]    (A = dh K-major, B = W1 as stored: MN-major)
//             dW1 = dh^T x           [1024,512] (both operands as stored: MN-major; split-K over the tokens)
// One kernel: C[M,N] (+)= A[M,K] * B[N,K]^T with bf16 operands staged by TMA (128B swizzle) either K-major
// (rows of 64 k-elements) or MN-major (rows of 64 m/n-elements, one row per k: what a row-major [K, M] array
// gives without any transposed copy), fp32 accumulation in TMEM.
//
// Persistent, warp-specialised, one CTA per SM: warp 0 = TMA producer (4-stage mbarrier ring of 48 KB stages),
// warp 1 = single-thread tcgen05.mma issuer (M=128 x N=256 x K=16, two 256-column TMEM accumulators so that the
// next tile's MMAs run under this tile's epilogue), warps 2..9 = epilogue (the two warps of a TMEM lane quarter
// split the 256 columns).  Work items = (m-tile, n-tile, k-split) in a static round-robin over the grid; split-K
// partials are combined with red.global.add.f32 into a zero-initialised output.
// Roofline: tensor pipe; 2*M*N*K flops per problem; operands are read from HBM once (they fit L2).
#include "common.cuh"
#include "nrhead_internal.h"
#include "tc_common.cuh"

namespace nr {
using namespace tc;

typedef CUresult (*EncodeTiledFnG)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFnG gemm_get_encode() {
  static EncodeTiledFnG fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFnG)p;
  }
  return fn;
}

constexpr int G_BM = 128, G_BN = 256, G_BK = 64;
constexpr int G_STAGES = 4;
constexpr int G_A_BYTES = G_BM * 128;          // 16 KB
constexpr int G_B_BYTES = G_BN * 128;          // 32 KB
constexpr int G_STAGE_BYTES = G_A_BYTES + G_B_BYTES;
constexpr int G_THREADS = 64 + 256;            // TMA warp, MMA warp, 8 epilogue warps
constexpr int G_EPI = 256;

enum { G_EPI_STORE_F32 = 0, G_EPI_RED_F32 = 1, G_EPI_RELU_BF16 = 2 };

constexpr int G_MAX_PROB = 4;

struct GemmProb {
  int M, N, K;
  int a_mn, b_mn;                // operand stored MN-major ([K, M] / [K, N] row-major) instead of K-major
  int n_mt, n_nt, ksplit, kb_per_split, num_kb, item0;
  int bn;                        // MMA N of this problem (<= 256, multiple of 16)
  int epi;
  float* c_f32; int64_t ldc;     // G_EPI_STORE_F32 / G_EPI_RED_F32
  __nv_bfloat16* h_bf16;         // G_EPI_RELU_BF16: hidden activations [M, ldc] (nullable: evaluation)
  const float* bias; const float* w2; float* logits;
};

// Up to 4 independent problems share a launch (both modalities' forward GEMMs; dW1 and dx of both modalities): one
// item list, so the persistent grid never idles between them.
struct alignas(64) GemmArgs {
  CUtensorMap tma[G_MAX_PROB], tmb[G_MAX_PROB];
  GemmProb p[G_MAX_PROB];
  int nprob, n_items;
};

// MN-major operand tile: 64-element (128 B) rows, one per k; 8-row groups 1024 B apart (SBO), the next 64 m/n
// elements 64 rows = 8192 B further (LBO)
__device__ __forceinline__ uint64_t umma_desc_mnmajor_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);
  d |= (uint64_t)(8192 >> 4) << 16;              // LBO: stride between 64-element atoms along M/N
  d |= (uint64_t)(1024 >> 4) << 32;              // SBO: stride between 8-row groups along K
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ void red_add_f32(float* addr, float v) {
  asm volatile("red.relaxed.gpu.global.add.f32 [%0], %1;" ::"l"(addr), "f"(v) : "memory");
}

__device__ __forceinline__ void red_add_v4_f32(float* addr, float a, float b, float c, float d) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

__global__ void __launch_bounds__(G_THREADS, 1) gemm_bf16_tc_kernel(const __grid_constant__ GemmArgs a) {
  extern __shared__ __align__(1024) uint8_t g_smem_raw[];
  // keep the shared-space provenance of the pointer (offset add, no integer round trip): STS/LDS, not generic ST/LD
  uint8_t* smem = g_smem_raw + ((1024u - (smem_u32(g_smem_raw) & 1023u)) & 1023u);
  float* sbias = reinterpret_cast<float*>(smem + (size_t)G_STAGES * G_STAGE_BYTES);     // [2][256]
  float* sw2 = sbias + 2 * G_BN;                                                          // [2][256]
  uint64_t* bars = reinterpret_cast<uint64_t*>(sw2 + 2 * G_BN);
  uint64_t* full = bars;                    // [4] TMA -> MMA
  uint64_t* empty = bars + G_STAGES;        // [4] MMA -> TMA
  uint64_t* tfull = bars + 2 * G_STAGES;    // [2] MMA -> epilogue
  uint64_t* tempty = tfull + 2;             // [2] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int i = 0; i < a.nprob; ++i) { tma_prefetch_desc(&a.tma[i]); tma_prefetch_desc(&a.tmb[i]); }
    for (int s = 0; s < G_STAGES; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, G_EPI); }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // item -> (problem, m-tile, n-tile, k-block range); n fastest so that concurrently running CTAs share the A tile in L2
  auto decode = [&](int item, int& pi, int& mt, int& nt, int& kbA, int& kbB) {
    pi = 0;
    while (pi + 1 < a.nprob && item >= a.p[pi + 1].item0) ++pi;
    const GemmProb& P = a.p[pi];
    const int local = item - P.item0;
    const int ks = local % P.ksplit;
    const int t = local / P.ksplit;
    nt = t % P.n_nt;
    mt = t / P.n_nt;
    kbA = ks * P.kb_per_split;
    kbB = min(P.num_kb, kbA + P.kb_per_split);
  };

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        int pi, mt, nt, kbA, kbB;
        decode(item, pi, mt, nt, kbA, kbB);
        const GemmProb& P = a.p[pi];
        const int a_boxes = P.a_mn ? G_BM / 64 : 1;
        const int b_boxes = P.b_mn ? (P.bn + 63) / 64 : 1;
        // bytes landing per stage: full boxes (TMA zero-fills rows / columns beyond the tensor)
        const uint32_t tx = (uint32_t)(P.a_mn ? a_boxes * 64 * 128 : G_BM * 128) +
                            (uint32_t)(P.b_mn ? b_boxes * 64 * 128 : P.bn * 128);
        for (int kb = kbA; kb < kbB; ++kb) {
          mbar_wait(empty + stage, phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * G_STAGE_BYTES;
          uint8_t* sb = sa + G_A_BYTES;
          mbar_expect_tx(full + stage, tx);
          if (P.a_mn) {
            for (int i = 0; i < a_boxes; ++i) tma_load_2d(sa + i * 8192, &a.tma[pi], full + stage, mt * G_BM + i * 64, kb * G_BK);
          } else {
            tma_load_2d(sa, &a.tma[pi], full + stage, kb * G_BK, mt * G_BM);
          }
          if (P.b_mn) {
            for (int i = 0; i < b_boxes; ++i) tma_load_2d(sb + i * 8192, &a.tmb[pi], full + stage, nt * G_BN + i * 64, kb * G_BK);
          } else {
            tma_load_2d(sb, &a.tmb[pi], full + stage, kb * G_BK, nt * G_BN);
          }
          if (++stage == G_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int item = blockIdx.x; item < a.n_items; item += gridDim.x, ++it) {
        int pi, mt, nt, kbA, kbB;
        decode(item, pi, mt, nt, kbA, kbB);
        const GemmProb& P = a.p[pi];
        const uint32_t idesc = umma_idesc_bf16(G_BM, P.bn) | ((uint32_t)(P.a_mn ? 1 : 0) << 15) |
                               ((uint32_t)(P.b_mn ? 1 : 0) << 16);
        const uint64_t a_step = P.a_mn ? (uint64_t)(2048 >> 4) : 2ull;    // advance K by 16 elements
        const uint64_t b_step = P.b_mn ? (uint64_t)(2048 >> 4) : 2ull;
        const int acc = it & 1;
        mbar_wait(tempty + acc, (uint32_t)((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * G_BN);
        for (int kb = kbA; kb < kbB; ++kb) {
          mbar_wait(full + stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * G_STAGE_BYTES);
          const uint64_t adesc = P.a_mn ? umma_desc_mnmajor_sw128(sa) : umma_desc_kmajor_sw128(sa);
          const uint64_t bdesc = P.b_mn ? umma_desc_mnmajor_sw128(sa + G_A_BYTES) : umma_desc_kmajor_sw128(sa + G_A_BYTES);
#pragma unroll
          for (int k = 0; k < G_BK / 16; ++k)
            umma_bf16(tmem_d, adesc + a_step * (uint64_t)k, bdesc + b_step * (uint64_t)k, idesc, (kb > kbA || k > 0) ? 1u : 0u);
          umma_commit(empty + stage);
          if (++stage == G_STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull + acc);
      }
    }
  } else {
    // ===================== epilogue: 8 warps, two per TMEM lane quarter (column halves) =====================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = q * 32 + lane;                                   // accumulator row
    const int et = (int)threadIdx.x - 64;
    int it = 0;
    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x, ++it) {
      int pi, mt, nt, kbA, kbB;
      decode(item, pi, mt, nt, kbA, kbB);
      const GemmProb& P = a.p[pi];
      const int hc = ((P.bn / 2) + 15) / 16 * 16;                  // columns per half (multiple of 16)
      const int acc = it & 1;
      const int n0 = nt * G_BN;
      const int m = mt * G_BM + r;
      const bool row_ok = m < P.M;
      float* sb_ = sbias + acc * G_BN;
      float* sw_ = sw2 + acc * G_BN;
      if (P.epi == G_EPI_RELU_BF16) {
        // stage this tile's bias / second-layer weights while its MMAs are still running.  Buffer `acc` was last
        // read two items ago; the barrier at the end of the previous item orders those reads before these writes.
        for (int c = et; c < P.bn; c += G_EPI) {
          const bool ok = n0 + c < P.N;
          sb_[c] = ok ? P.bias[n0 + c] : 0.f;
          sw_[c] = ok ? P.w2[n0 + c] : 0.f;
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      mbar_wait(tfull + acc, (uint32_t)((it >> 1) & 1));
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * G_BN);
      const int c_lo = half * hc, c_hi = min(P.bn, c_lo + hc);
      float dot = 0.f;
      for (int c = c_lo; c < c_hi; c += 16) {
        uint32_t v[16];
        tmem_ld16(taddr + (uint32_t)c, v);
        tmem_ld_wait();
        reg_fence<16>(v);
        const int n = n0 + c;
        if (P.epi == G_EPI_RELU_BF16) {
          uint32_t pk[8];
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            const float h0 = fmaxf(__uint_as_float(v[e]) + sb_[c + e], 0.f);
            const float h1 = fmaxf(__uint_as_float(v[e + 1]) + sb_[c + e + 1], 0.f);
            dot = fmaf(h0, sw_[c + e], dot);
            dot = fmaf(h1, sw_[c + e + 1], dot);
            __nv_bfloat162 p2 = __floats2bfloat162_rn(h0, h1);
            pk[e >> 1] = *reinterpret_cast<uint32_t*>(&p2);
          }
          if (P.h_bf16 && row_ok) {
            __nv_bfloat16* dst = P.h_bf16 + (int64_t)m * P.ldc + n;
            if (n + 16 <= P.N) {
              *reinterpret_cast<uint4*>(dst) = make_uint4(pk[0], pk[1], pk[2], pk[3]);
              *reinterpret_cast<uint4*>(dst + 8) = make_uint4(pk[4], pk[5], pk[6], pk[7]);
            } else {
              for (int e = 0; e < 16 && n + e < P.N; ++e)
                dst[e] = reinterpret_cast<__nv_bfloat16*>(pk)[e];
            }
          }
        } else if (row_ok) {
          float* dst = P.c_f32 + (int64_t)m * P.ldc + n;
          if (P.epi == G_EPI_STORE_F32) {
            if (n + 16 <= P.N) {
#pragma unroll
              for (int e = 0; e < 16; e += 4)
                *reinterpret_cast<float4*>(dst + e) = make_float4(__uint_as_float(v[e]), __uint_as_float(v[e + 1]),
                                                                  __uint_as_float(v[e + 2]), __uint_as_float(v[e + 3]));
            } else {
              for (int e = 0; e < 16 && n + e < P.N; ++e) dst[e] = __uint_as_float(v[e]);
            }
          } else if (n + 16 <= P.N) {
#pragma unroll
            for (int e = 0; e < 16; e += 4)
              red_add_v4_f32(dst + e, __uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]),
                             __uint_as_float(v[e + 3]));
          } else {
            for (int e = 0; e < 16 && n + e < P.N; ++e) red_add_f32(dst + e, __uint_as_float(v[e]));
          }
        }
      }
      tc_fence_before();
      mbar_arrive(tempty + acc);
      if (P.epi == G_EPI_RELU_BF16) {
        if (row_ok) atomicAdd(P.logits + m, dot);
        asm volatile("bar.sync 1, 256;" ::: "memory");      // sbias / sw2 of this item fully read
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

// bf16 2-D tensor map over a row-major [rows, cols] array with leading dimension ld: boxes {64 columns, box_rows},
// 128B swizzle.  K-major operand: cols = K; MN-major operand: cols = M or N, rows = K.
static int gemm_tmap(CUtensorMap* m, const void* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
  EncodeTiledFnG enc = gemm_get_encode();
  NR_CHECK_ARG(enc != nullptr, "cuTensorMapEncodeTiled unavailable (driver too old?)");
  cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NR_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(gemm) failed (%d) rows=%lld cols=%lld ld=%lld box_rows=%d", (int)r,
               (long long)rows, (long long)cols, (long long)ld, box_rows);
  return 0;
}

struct GemmJob {
  const void* A; int64_t lda; int a_mn;
  const void* B; int64_t ldb; int b_mn;
  int M, N, K, epi, want_split;
  float* c_f32; int64_t ldc; __nv_bfloat16* h_bf16;
  const float* bias; const float* w2; float* logits;
};

static int gemm_launch(const GemmJob* jobs, int njobs, cudaStream_t stream, int reserve_sms = 0) {
  NR_CHECK_ARG(jobs && njobs >= 1 && njobs <= G_MAX_PROB, "nr_gemm: 1..%d problems per launch (got %d)", G_MAX_PROB, njobs);
  int dev = 0, sms = 0;
  NR_CUDA(cudaGetDevice(&dev));
  NR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  GemmArgs a{};
  a.nprob = njobs;
  int total_tiles = 0;
  for (int i = 0; i < njobs; ++i) total_tiles += ((jobs[i].M + G_BM - 1) / G_BM) * ((jobs[i].N + G_BN - 1) / G_BN);
  int items = 0;
  for (int i = 0; i < njobs; ++i) {
    const GemmJob& j = jobs[i];
    GemmProb& P = a.p[i];
    NR_CHECK_ARG(j.A && j.B && j.M > 0 && j.N > 0 && j.K > 0, "nr_gemm: bad arguments (problem %d)", i);
    NR_CHECK_ARG(((uintptr_t)j.A & 15) == 0 && ((uintptr_t)j.B & 15) == 0 && j.lda % 8 == 0 && j.ldb % 8 == 0,
                 "nr_gemm: operands must be 16-byte aligned with leading dimensions that are multiples of 8");
    NR_CHECK_ARG(j.N % 16 == 0, "nr_gemm: N=%d must be a multiple of 16", j.N);
    NR_CHECK_ARG(j.epi == G_EPI_RELU_BF16 || (((uintptr_t)j.c_f32 & 15) == 0 && j.ldc % 4 == 0),
                 "nr_gemm: fp32 output must be 16-byte aligned with ldc %% 4 == 0");
    P.M = j.M; P.N = j.N; P.K = j.K; P.a_mn = j.a_mn; P.b_mn = j.b_mn; P.epi = j.epi;
    P.c_f32 = j.c_f32; P.ldc = j.ldc; P.h_bf16 = j.h_bf16; P.bias = j.bias; P.w2 = j.w2; P.logits = j.logits;
    P.bn = j.N < G_BN ? j.N : G_BN;
    P.n_mt = (j.M + G_BM - 1) / G_BM;
    P.n_nt = (j.N + G_BN - 1) / G_BN;
    P.num_kb = (j.K + G_BK - 1) / G_BK;
    // split-K (red.add epilogue only): items of about 16 k-blocks when the launch would otherwise leave SMs idle or
    // a few long items would set its length
    int ks = 1;
    if (j.want_split && j.epi == G_EPI_RED_F32 && (total_tiles < 2 * sms || P.num_kb > 32)) {
      ks = (P.num_kb + 15) / 16;
      if (njobs == 1 && P.n_mt * P.n_nt * ks < sms) {            // a lone small problem: fill the GPU instead
        ks = sms / (P.n_mt * P.n_nt);
        if (ks > P.num_kb) ks = P.num_kb;
      }
      if (ks < 1) ks = 1;
    }
    P.kb_per_split = (P.num_kb + ks - 1) / ks;
    P.ksplit = (P.num_kb + P.kb_per_split - 1) / P.kb_per_split;
    P.item0 = items;
    items += P.n_mt * P.n_nt * P.ksplit;
    if (j.a_mn) { if (int e = gemm_tmap(&a.tma[i], j.A, j.K, j.M, j.lda, 64)) return e; }
    else { if (int e = gemm_tmap(&a.tma[i], j.A, j.M, j.K, j.lda, G_BM)) return e; }
    if (j.b_mn) { if (int e = gemm_tmap(&a.tmb[i], j.B, j.K, j.N, j.ldb, 64)) return e; }
    else { if (int e = gemm_tmap(&a.tmb[i], j.B, j.N, j.K, j.ldb, P.bn)) return e; }
  }
  a.n_items = items;
  const size_t smem = (size_t)G_STAGES * G_STAGE_BYTES + 4 * G_BN * sizeof(float) + 256 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    NR_CUDA(cudaFuncSetAttribute(gemm_bf16_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  int cap = sms - reserve_sms;
  if (cap < 1) cap = 1;
  const int grid = a.n_items < cap ? a.n_items : cap;
  gemm_bf16_tc_kernel<<<grid, G_THREADS, smem, stream>>>(a);
  NR_CHECK_LAUNCH("nr_gemm");
  return 0;
}

// fp32 -> bf16 copy (operand copies of the MLP inputs / W1); 8 elements per thread
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int64_t n) {
  const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) * 8;
  if (i + 8 <= n) {
    const float4 a = *reinterpret_cast<const float4*>(x + i), b = *reinterpret_cast<const float4*>(x + i + 4);
    __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
    *reinterpret_cast<uint4*>(y + i) = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                                                  *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
  } else {
    for (int64_t j = i; j < n; ++j) y[j] = __float2bfloat16_rn(x[j]);
  }
}

struct CastSegs { const float* src[8]; __nv_bfloat16* dst[8]; long long n[8]; int blk0[9]; int nseg; };
// several fp32 -> bf16 operand copies in one launch (both modalities' batch / bank tokens and W1)
__global__ void __launch_bounds__(256) cast_bf16_multi_kernel(const CastSegs c) {
  int sgm = 0;
  while (sgm + 1 < c.nseg && (int)blockIdx.x >= c.blk0[sgm + 1]) ++sgm;
  const float* __restrict__ x = c.src[sgm];
  __nv_bfloat16* __restrict__ y = c.dst[sgm];
  const int64_t n = c.n[sgm];
  const int64_t i = ((int64_t)(blockIdx.x - c.blk0[sgm]) * 256 + threadIdx.x) * 8;
  if (i + 8 <= n) {
    const float4 a = *reinterpret_cast<const float4*>(x + i), b = *reinterpret_cast<const float4*>(x + i + 4);
    __nv_bfloat162 p0 = __floats2bfloat162_rn(a.x, a.y), p1 = __floats2bfloat162_rn(a.z, a.w);
    __nv_bfloat162 p2 = __floats2bfloat162_rn(b.x, b.y), p3 = __floats2bfloat162_rn(b.z, b.w);
    *reinterpret_cast<uint4*>(y + i) = make_uint4(*reinterpret_cast<uint32_t*>(&p0), *reinterpret_cast<uint32_t*>(&p1),
                                                  *reinterpret_cast<uint32_t*>(&p2), *reinterpret_cast<uint32_t*>(&p3));
  } else {
    for (int64_t j = i; j < n; ++j) y[j] = __float2bfloat16_rn(x[j]);
  }
}

// masked softmax over the tokens of each sample from the second-layer logits (reference modeling.py:486-487,
// 491-492): one warp per sample, N <= 128 tokens
__device__ __forceinline__ void token_softmax_row(const float* __restrict__ logits, const float* __restrict__ b2,
                                                  const int64_t* __restrict__ mask_a, const int64_t* __restrict__ mask_b,
                                                  int Ra, int R, int N, float* __restrict__ w, int row, int lane) {
  if (row >= R) return;
  const int64_t* mk = row < Ra ? (mask_a ? mask_a + (int64_t)row * N : nullptr)
                               : (mask_b ? mask_b + (int64_t)(row - Ra) * N : nullptr);
  const float bias = b2[0];
  float v[4];
  float mx = NR_NEG_INF;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = lane + 32 * i;
    v[i] = NR_NEG_INF;
    if (t < N) {
      v[i] = logits[(int64_t)row * N + t] + bias;
      if (mk && mk[t] == 0) v[i] = -9e15f;                  // masked_fill_ value of the reference
      mx = fmaxf(mx, v[i]);
    }
  }
  mx = warp_max(mx);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = lane + 32 * i;
    v[i] = (t < N) ? expf(v[i] - mx) : 0.f;
    s += v[i];
  }
  s = warp_sum(s);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int t = lane + 32 * i;
    if (t < N) w[(int64_t)row * N + t] = v[i] / s;
  }
}

__global__ void __launch_bounds__(256) token_softmax_kernel(const float* __restrict__ logits, const float* __restrict__ b2,
                                                            const int64_t* __restrict__ mask_a, const int64_t* __restrict__ mask_b,
                                                            int Ra, int R, int N, float* __restrict__ w) {
  token_softmax_row(logits, b2, mask_a, mask_b, Ra, R, N, w, blockIdx.x * 8 + (threadIdx.x >> 5), threadIdx.x & 31);
}

// both modalities of a step in one launch (their token counts differ): blocks [0, blk1) serve side 0, the rest side 1
struct SoftmaxSide { const float* logits; const float* b2; const int64_t* mask_a; const int64_t* mask_b; int Ra, R, N; float* w; };
struct SoftmaxPairArgs { SoftmaxSide s[2]; int blk1; };
__global__ void __launch_bounds__(256) token_softmax_pair_kernel(const SoftmaxPairArgs a) {
  const int si = (int)blockIdx.x >= a.blk1 ? 1 : 0;
  const SoftmaxSide& S = a.s[si];
  const int blk = (int)blockIdx.x - (si ? a.blk1 : 0);
  token_softmax_row(S.logits, S.b2, S.mask_a, S.mask_b, S.Ra, S.R, S.N, S.w, blk * 8 + (threadIdx.x >> 5), threadIdx.x & 31);
}

}  // namespace nr

using namespace nr;

extern "C" int nr_cast_bf16(const float* x, void* y, int64_t n, void* stream) {
  NR_CHECK_ARG(x && y && n > 0 && ((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0, "nr_cast_bf16: bad arguments");
  const int64_t blocks = (n + 2047) / 2048;
  cast_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(x, (__nv_bfloat16*)y, n);
  NR_CHECK_LAUNCH("nr_cast_bf16");
  return 0;
}

static GemmJob job_fwd(const void* x_bf16, int64_t T, int64_t D, const void* w1_bf16, int64_t H, const float* b1,
                       const float* w2, void* h_bf16, float* logits) {
  GemmJob j{};
  j.A = x_bf16; j.lda = D; j.a_mn = 0; j.B = w1_bf16; j.ldb = D; j.b_mn = 0;
  j.M = (int)T; j.N = (int)H; j.K = (int)D; j.epi = G_EPI_RELU_BF16;
  j.h_bf16 = (__nv_bfloat16*)h_bf16; j.ldc = H; j.bias = b1; j.w2 = w2; j.logits = logits;
  return j;
}
static GemmJob job_dx(const void* dh_bf16, int64_t T, int64_t H, const void* w1_bf16, int64_t D, float* dx, int accumulate) {
  GemmJob j{};
  j.A = dh_bf16; j.lda = H; j.a_mn = 0; j.B = w1_bf16; j.ldb = D; j.b_mn = 1;   // B[n=d][k=h] = W1[h][d]: n contiguous
  j.M = (int)T; j.N = (int)D; j.K = (int)H; j.epi = accumulate ? G_EPI_RED_F32 : G_EPI_STORE_F32; j.want_split = accumulate;
  j.c_f32 = dx; j.ldc = D;
  return j;
}
static GemmJob job_dw1(const void* dh_bf16, int64_t T, int64_t H, const void* x_bf16, int64_t D, float* dw1) {
  GemmJob j{};
  j.A = dh_bf16; j.lda = H; j.a_mn = 1; j.B = x_bf16; j.ldb = D; j.b_mn = 1;    // both operands as stored
  j.M = (int)H; j.N = (int)D; j.K = (int)T; j.epi = G_EPI_RED_F32; j.want_split = 1;
  j.c_f32 = dw1; j.ldc = D;
  return j;
}

extern "C" int nr_cast_bf16_multi(const float* const* src, void* const* dst, const int64_t* n, int nseg, void* stream) {
  NR_CHECK_ARG(src && dst && n && nseg >= 1 && nseg <= 8, "nr_cast_bf16_multi: 1..8 segments");
  CastSegs c{};
  c.nseg = nseg;
  int blocks = 0;
  for (int i = 0; i < nseg; ++i) {
    NR_CHECK_ARG(src[i] && dst[i] && n[i] > 0 && ((uintptr_t)src[i] & 15) == 0 && ((uintptr_t)dst[i] & 15) == 0,
                 "nr_cast_bf16_multi: segment %d: null, empty or misaligned", i);
    c.src[i] = src[i]; c.dst[i] = (__nv_bfloat16*)dst[i]; c.n[i] = n[i];
    c.blk0[i] = blocks;
    blocks += (int)((n[i] + 2047) / 2048);
  }
  c.blk0[nseg] = blocks;
  cast_bf16_multi_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(c);
  NR_CHECK_LAUNCH("nr_cast_bf16_multi");
  return 0;
}

extern "C" int nr_mlp_fwd(const void* x_bf16, int64_t T, int64_t D, const void* w1_bf16, int64_t H, const float* b1,
                          const float* w2, void* h_bf16, float* logits, void* stream) {
  NR_CHECK_ARG(b1 && w2 && logits, "nr_mlp_fwd: bad arguments");
  GemmJob j = job_fwd(x_bf16, T, D, w1_bf16, H, b1, w2, h_bf16, logits);
  return gemm_launch(&j, 1, (cudaStream_t)stream);
}

extern "C" int nr_mlp_bwd_dx(const void* dh_bf16, int64_t T, int64_t H, const void* w1_bf16, int64_t D, float* dx,
                             int accumulate, void* stream) {
  NR_CHECK_ARG(dx, "nr_mlp_bwd_dx: bad arguments");
  GemmJob j = job_dx(dh_bf16, T, H, w1_bf16, D, dx, accumulate);
  return gemm_launch(&j, 1, (cudaStream_t)stream);
}

extern "C" int nr_mlp_bwd_dw1(const void* dh_bf16, int64_t T, int64_t H, const void* x_bf16, int64_t D, float* dw1,
                              void* stream) {
  NR_CHECK_ARG(dw1, "nr_mlp_bwd_dw1: bad arguments");
  GemmJob j = job_dw1(dh_bf16, T, H, x_bf16, D, dw1);
  return gemm_launch(&j, 1, (cudaStream_t)stream);
}

/* Both modalities in ONE launch each way (nr_mlp_pair of include/nrhead.h). */
extern "C" int nr_mlp_fwd_pair(const nr_mlp_side* s, int n, int64_t D, int64_t H, int reserve_sms, void* stream) {
  NR_CHECK_ARG(s && n >= 1 && n <= 2, "nr_mlp_fwd_pair: 1 or 2 sides");
  GemmJob j[2];
  for (int i = 0; i < n; ++i) {
    NR_CHECK_ARG(s[i].b1 && s[i].w2 && s[i].logits, "nr_mlp_fwd_pair: side %d has a null pointer", i);
    j[i] = job_fwd(s[i].x_bf16, s[i].T, D, s[i].w1_bf16, H, s[i].b1, s[i].w2, s[i].h_bf16, s[i].logits);
  }
  return gemm_launch(j, n, (cudaStream_t)stream, reserve_sms < 0 ? 0 : reserve_sms);
}

extern "C" int nr_mlp_bwd_pair(const nr_mlp_side* s, int n, int64_t D, int64_t H, void* stream) {
  NR_CHECK_ARG(s && n >= 1 && n <= 2, "nr_mlp_bwd_pair: 1 or 2 sides");
  GemmJob j[4];
  int nj = 0;
  for (int i = 0; i < n; ++i) {
    // long split-K problems first: the static item order then ends with the short dx items
    if (s[i].dw1) j[nj++] = job_dw1(s[i].dh_bf16, s[i].T, H, s[i].x_bf16, D, s[i].dw1);
  }
  for (int i = 0; i < n; ++i)
    if (s[i].dx) j[nj++] = job_dx(s[i].dh_bf16, s[i].T_dx, H, s[i].w1_bf16, D, s[i].dx, 1);
  if (nj == 0) return 0;
  return gemm_launch(j, nj, (cudaStream_t)stream);
}

extern "C" int nr_token_softmax_pair(const nr_softmax_side* sides, int n_sides, void* stream) {
  NR_CHECK_ARG(sides && (n_sides == 1 || n_sides == 2), "nr_token_softmax_pair: 1 or 2 sides");
  SoftmaxPairArgs a{};
  int blocks = 0;
  for (int i = 0; i < n_sides; ++i) {
    const nr_softmax_side& q = sides[i];
    NR_CHECK_ARG(q.logits && q.b2 && q.w && q.R > 0 && q.N > 0 && q.N <= 128 && q.Ra >= 0 && q.Ra <= q.R,
                 "nr_token_softmax_pair: side %d: bad arguments", i);
    a.s[i] = SoftmaxSide{q.logits, q.b2, q.mask_a, q.mask_b, (int)q.Ra, (int)q.R, (int)q.N, q.w};
    if (i == 1) a.blk1 = blocks;
    blocks += (int)((q.R + 7) / 8);
  }
  if (n_sides == 1) a.blk1 = blocks;
  token_softmax_pair_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(a);
  NR_CHECK_LAUNCH("nr_token_softmax_pair");
  return 0;
}

extern "C" int nr_token_softmax(const float* logits, const float* b2, const int64_t* mask_a, const int64_t* mask_b,
                                int64_t Ra, int64_t R, int64_t N, float* w, void* stream) {
  NR_CHECK_ARG(logits && b2 && w && R > 0 && N > 0 && N <= 128 && Ra >= 0 && Ra <= R, "nr_token_softmax: bad arguments");
  token_softmax_kernel<<<(unsigned)((R + 7) / 8), 256, 0, (cudaStream_t)stream>>>(logits, b2, mask_a, mask_b, (int)Ra, (int)R,
                                                                                 (int)N, w);
  NR_CHECK_LAUNCH("nr_token_softmax");
  return 0;
}
