// Masked max-sim late interaction on the 5th-gen tensor cores (NR_PREC_BF16).
//
//   H[rx, ry] = sum_x wx[rx,x] * max_y ( <xn[rx,x,:], yn[ry,y,:]> * mx[rx,x] * my[ry,y] )
//
// (one direction of NeighborRetr.local_level, reference NeighborRetr/models/modeling.py:499-509; the
// reference materialises the 4-D [A,B,Nt,Nv] fp32 tensor and makes >= 7 passes over it.)
//
// Persistent warp-specialised kernel, one CTA per SM:
//   warp 0      TMA producer: {64 x rows} bf16 boxes of X and Y tokens, 128B swizzle, 4-stage mbarrier ring
//   warp 1      tcgen05.mma issuer (one elected lane), M=128 x N<=256 x K=16, fp32 accumulators in TMEM,
//               two accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1
//   warps 2..5  epilogue: tcgen05.ld one accumulator row per thread (row = one X token), masked running
//               max / arg-max over the Ny columns of each Y sample in registers, token-weighted partial to
//               shared memory, segmented sum over the Nx rows of each X sample -> H.  Only H, pmax and the
//               arg-max bytes leave the SM: the token-pair tensor never exists in HBM.
// A tile is SX whole X samples (SX*Nx <= 128 rows; rows up to 128 are don't-care) by SY whole Y samples
// (SY*Ny <= 256 columns), so every max / sum segment is tile-local.
// Roofline: tensor pipe.  Algorithmic flops per launch 2*Rx*Nx*Ry*Ny*D; executed: 128/(SX*Nx) more.
#include <stdlib.h>
#include "common.cuh"
#include "nrhead_internal.h"
#include "tc_common.cuh"

namespace nr {
using namespace tc;

constexpr int TC_THREADS = 192;            // backward kernel: TMA warp, MMA warp, 4 generator/epilogue warps
constexpr int TCF_THREADS = 320;           // forward kernel: TMA warp, MMA warp, 8 epilogue warps (2 per TMEM lane quarter)
constexpr int TCF_EPI = TCF_THREADS - 64;
constexpr int TC_BM = 128;
constexpr int TC_BK = 64;                  // bf16 elements per k-block = one 128-byte swizzle row
constexpr int TC_MAX_STAGES = 4;
constexpr int TC_A_BYTES = TC_BM * 128;    // 16 KB
constexpr int TC_ACC_COLS = 256;           // TMEM columns per accumulator stage

struct TcFwdArgs {
  const float* wx; const int64_t* mx; const int64_t* my;
  int Rx, Nx, Ry, Ny;
  float alpha;
  float* out; int64_t out_sr, out_sc; float* out2; int64_t out2_sr, out2_sc; int accumulate;
  float* pmax; uint8_t* ystar;
  int SX, SY, MU, UN;                      // samples per tile, used rows, UMMA N (multiple of 16)
  int n_mt, n_nt, num_kb, stages, b_bytes; // tiles, k-blocks, pipeline depth, bytes of one B stage
  int hp_ld;                               // leading dimension (floats) of the weighted-partial buffer
};

template <int NY>
__device__ __forceinline__ void sample_argmax(const uint32_t* v, uint64_t mb, bool full, float& best, int& bi) {
  float f[NY];
  if (full) {
#pragma unroll
    for (int y = 0; y < NY; ++y) f[y] = __uint_as_float(v[y]);
  } else {
#pragma unroll
    for (int y = 0; y < NY; ++y) f[y] = ((mb >> y) & 1ull) ? __uint_as_float(v[y]) : 0.f;   // masked pairs are exactly 0
  }
  best = TreeRed<NY>::fmax_(f);
  bi = TreeRed<NY>::first_eq(f, best, 0);
}

template <int NY>
__global__ void __launch_bounds__(TCF_THREADS, 1)
maxsim_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmx, const __grid_constant__ CUtensorMap tmy,
                     const TcFwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x (A | B)] [hp 2 x 128 x hp_ld floats] [cmask 2 x 256] [barriers]
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int stage_bytes = TC_A_BYTES + a.b_bytes;
  float* hp = reinterpret_cast<float*>(smem + (size_t)a.stages * stage_bytes);
  uint64_t* cmask = reinterpret_cast<uint64_t*>(hp + 2 * TC_BM * a.hp_ld);   // [2][64] bit y = my[ry, y]
  uint64_t* bars = cmask + 2 * 64;
  uint64_t* full = bars;                          // [stages]  TMA -> MMA
  uint64_t* empty = bars + TC_MAX_STAGES;         // [stages]  MMA -> TMA
  uint64_t* tfull = bars + 2 * TC_MAX_STAGES;     // [2]       MMA -> epilogue
  uint64_t* tempty = tfull + 2;                   // [2]       epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = a.n_mt * a.n_nt;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmx);
    tma_prefetch_desc(&tmy);
    for (int s = 0; s < a.stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, TCF_EPI); }
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

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const uint32_t tx_bytes = (uint32_t)(a.MU + a.SY * NY) * 128u;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int mt = tile / a.n_nt, nt = tile % a.n_nt;
        const int row_x = mt * a.SX * a.Nx, row_y = nt * a.SY * NY;
        for (int kb = 0; kb < a.num_kb; ++kb) {
          mbar_wait(empty + stage, phase ^ 1);
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          mbar_expect_tx(full + stage, tx_bytes);
          tma_load_2d(sa, &tmx, full + stage, kb * TC_BK, row_x);
          tma_load_2d(sa + TC_A_BYTES, &tmy, full + stage, kb * TC_BK, row_y);
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(TC_BM, a.UN);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(tempty + acc, acc_phase ^ 1);          // epilogue drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * TC_ACC_COLS);
        for (int kb = 0; kb < a.num_kb; ++kb) {
          mbar_wait(full + stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t adesc = umma_desc_kmajor_sw128(sa);
          const uint64_t bdesc = umma_desc_kmajor_sw128(sa + TC_A_BYTES);
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k)           // advance 32 B (16 bf16) inside the swizzle row
            umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
          umma_commit(empty + stage);                    // smem slot free when these MMAs retire
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(tfull + acc);                        // accumulator ready for the epilogue
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    const int q = warp & 3;                              // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;                    // the two warps of a quarter take alternate Y samples
    const int r = q * 32 + lane;                         // accumulator row = X token of the tile
    const int et = threadIdx.x - 64;                     // 0..255
    const uint64_t fullmask = NY == 64 ? ~0ull : ((1ull << NY) - 1ull);
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
      const int mt = tile / a.n_nt, nt = tile % a.n_nt;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int ry0 = nt * a.SY;
      const int sy_n = min(a.SY, a.Ry - ry0);
      uint64_t* cm = cmask + acc * 64;
      float* hpb = hp + (size_t)acc * TC_BM * a.hp_ld;
      // column masks of this tile, one bit per Y token
      if (et < sy_n) {
        uint64_t bits = fullmask;
        if (a.my) {
          bits = 0;
          const int64_t* mrow = a.my + (int64_t)(ry0 + et) * NY;
#pragma unroll
          for (int y = 0; y < NY; ++y) bits |= (uint64_t)(mrow[y] != 0) << y;
        }
        cm[et] = bits;
      }
      const int sx = r / a.Nx, x = r - sx * a.Nx;
      const int rx = mt * a.SX + sx;
      const bool row_ok = (r < a.MU) && (rx < a.Rx);
      const bool mxv = row_ok && (a.mx ? (a.mx[(int64_t)rx * a.Nx + x] != 0) : true);
      const float wxv = row_ok ? a.wx[(int64_t)rx * a.Nx + x] : 0.f;
      const int64_t obase = ((int64_t)rx * a.Ry + ry0) * a.Nx + x;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mbar_wait(tfull + acc, acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * TC_ACC_COLS);
      auto finish = [&](const uint32_t* v, int sy) {
        const uint64_t mb = cm[sy];
        float best; int bi;
        sample_argmax<NY>(v, mb, mb == fullmask, best, bi);
        // arg-max byte 255 = "no gradient": masked X token, or the max is a masked (exactly 0) pair
        if (!mxv) { best = 0.f; bi = 255; }
        else if (!((mb >> bi) & 1ull)) bi = 255;
        hpb[r * a.hp_ld + sy] = wxv * best;
        if (row_ok) {
          const int64_t o = obase + (int64_t)sy * a.Nx;
          if (a.pmax) a.pmax[o] = best;
          if (a.ystar) a.ystar[o] = (uint8_t)bi;
        }
      };
      if constexpr (NY <= 32) {
        // two register buffers: the TMEM load of the next sample is in flight while this one is reduced
        uint32_t va[NY], vb[NY];
        int sy = half;
        if (sy < sy_n) tmem_ld_cols<NY>(taddr + (uint32_t)(sy * NY), va);
        for (; sy < sy_n; sy += 4) {
          tmem_ld_wait();
          reg_fence<NY>(va);
          const int sy2 = sy + 2;
          if (sy2 < sy_n) tmem_ld_cols<NY>(taddr + (uint32_t)(sy2 * NY), vb);
          finish(va, sy);
          if (sy2 < sy_n) {
            tmem_ld_wait();
            reg_fence<NY>(vb);
            if (sy + 4 < sy_n) tmem_ld_cols<NY>(taddr + (uint32_t)((sy + 4) * NY), va);
            finish(vb, sy2);
          }
        }
      } else {
        for (int sy = half; sy < sy_n; sy += 2) {
          uint32_t v[NY];
          tmem_ld_cols<NY>(taddr + (uint32_t)(sy * NY), v);
          tmem_ld_wait();
          reg_fence<NY>(v);
          finish(v, sy);
        }
      }
      tc_fence_before();
      mbar_arrive(tempty + acc);                         // TMEM stage may be overwritten
      asm volatile("bar.sync 1, 256;" ::: "memory");
      // segmented sum over the Nx rows of each X sample
      for (int e = et; e < a.SX * sy_n; e += TCF_EPI) {
        const int s = e / sy_n, sy = e - s * sy_n;
        const int rxx = mt * a.SX + s;
        if (rxx < a.Rx) {
          float h = 0.f;
          for (int xx = 0; xx < a.Nx; ++xx) h += hpb[(s * a.Nx + xx) * a.hp_ld + sy];
          h *= a.alpha;
          const int ry = ry0 + sy;
          float* p = a.out + (int64_t)rxx * a.out_sr + (int64_t)ry * a.out_sc;
          *p = a.accumulate ? (*p + h) : h;
          if (a.out2) {
            float* p2 = a.out2 + (int64_t)rxx * a.out2_sr + (int64_t)ry * a.out2_sc;
            *p2 = a.accumulate ? (*p2 + h) : h;
          }
        }
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

// ---- host side ----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  return fn;
}

// 2-D bf16 row-major [rows, d] -> boxes of {64 elements, box_rows} with the 128B swizzle
int make_tmap_bf16(CUtensorMap* m, const void* base, int64_t rows, int64_t d, int box_rows) {
  EncodeTiledFn enc = get_encode();
  NR_CHECK_ARG(enc != nullptr, "cuTensorMapEncodeTiled unavailable (driver too old?)");
  cuuint64_t gdim[2] = {(cuuint64_t)d, (cuuint64_t)rows};
  cuuint64_t gstr[1] = {(cuuint64_t)d * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  if (const char* pv = getenv("NR_TMA_PROMO")) {        // measurement knob: 0 none, 64, 128, 256
    const int v = atoi(pv);
    promo = v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
          : v == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  }
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NR_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed (%d) rows=%lld d=%lld box_rows=%d", (int)r,
               (long long)rows, (long long)d, box_rows);
  return 0;
}

template <int NY>
static int launch_fwd(const CUtensorMap& tmx, const CUtensorMap& tmy, const TcFwdArgs& a, size_t smem, int grid,
                      cudaStream_t stream) {
  NR_CUDA(cudaFuncSetAttribute(maxsim_fwd_tc_kernel<NY>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  maxsim_fwd_tc_kernel<NY><<<grid, TCF_THREADS, smem, stream>>>(tmx, tmy, a);
  NR_CHECK_LAUNCH("nr_maxsim_fwd(bf16)");
  return 0;
}

}  // namespace nr

using namespace nr;

int nr_maxsim_fwd_tc(const void* xn, const void* yn, const float* wx, const int64_t* mx, const int64_t* my,
                     int64_t Rx, int64_t Nx, int64_t Ry, int64_t Ny, int64_t d, float alpha, float* out,
                     int64_t out_sr, int64_t out_sc, float* out2, int64_t out2_sr, int64_t out2_sc, int accumulate,
                     float* pmax, uint8_t* ystar, cudaStream_t stream) {
  NR_CHECK_ARG(d % TC_BK == 0, "nr_maxsim_fwd(bf16): d=%lld must be a multiple of %d", (long long)d, TC_BK);
  NR_CHECK_ARG(Ny % 4 == 0, "nr_maxsim_fwd(bf16): Ny=%lld must be a multiple of 4", (long long)Ny);
  NR_CHECK_ARG(((uintptr_t)xn & 15) == 0 && ((uintptr_t)yn & 15) == 0, "nr_maxsim_fwd(bf16): operands must be 16B aligned");
  TcFwdArgs a{};
  a.wx = wx; a.mx = mx; a.my = my;
  a.Rx = (int)Rx; a.Nx = (int)Nx; a.Ry = (int)Ry; a.Ny = (int)Ny;
  a.alpha = alpha;
  a.out = out; a.out_sr = out_sr; a.out_sc = out_sc; a.out2 = out2; a.out2_sr = out2_sr; a.out2_sc = out2_sc;
  a.accumulate = accumulate; a.pmax = pmax; a.ystar = ystar;
  a.SX = TC_BM / (int)Nx;
  if (a.SX > Rx) a.SX = (int)Rx;
  a.MU = a.SX * (int)Nx;
  a.SY = 256 / (int)Ny;
  if (a.SY > Ry) a.SY = (int)Ry;
  a.UN = (a.SY * (int)Ny + 15) / 16 * 16;
  a.n_mt = (int)((Rx + a.SX - 1) / a.SX);
  a.n_nt = (int)((Ry + a.SY - 1) / a.SY);
  a.num_kb = (int)(d / TC_BK);
  a.b_bytes = (a.UN * 128 + 1023) / 1024 * 1024;
  a.hp_ld = a.SY | 1;
  const size_t tail = (size_t)2 * TC_BM * a.hp_ld * sizeof(float) + 2 * 64 * 8 + 256;
  const size_t budget = 227 * 1024 - 1024;   // alignment slack
  int stages = (int)((budget - tail) / (size_t)(TC_A_BYTES + a.b_bytes));
  if (stages > TC_MAX_STAGES) stages = TC_MAX_STAGES;
  NR_CHECK_ARG(stages >= 2, "nr_maxsim_fwd(bf16): tile does not fit shared memory");
  a.stages = stages;
  const size_t smem = (size_t)stages * (TC_A_BYTES + a.b_bytes) + tail + 1024;
  CUtensorMap tmx, tmy;
  if (int e = make_tmap_bf16(&tmx, xn, Rx * Nx, d, a.MU)) return e;
  if (int e = make_tmap_bf16(&tmy, yn, Ry * Ny, d, a.SY * (int)Ny)) return e;
  int dev = 0, sms = 0;
  NR_CUDA(cudaGetDevice(&dev));
  NR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int grid = a.n_mt * a.n_nt < sms ? a.n_mt * a.n_nt : sms;
  switch (Ny) {
    case 4: return launch_fwd<4>(tmx, tmy, a, smem, grid, stream);
    case 8: return launch_fwd<8>(tmx, tmy, a, smem, grid, stream);
    case 12: return launch_fwd<12>(tmx, tmy, a, smem, grid, stream);
    case 16: return launch_fwd<16>(tmx, tmy, a, smem, grid, stream);
    case 24: return launch_fwd<24>(tmx, tmy, a, smem, grid, stream);
    case 32: return launch_fwd<32>(tmx, tmy, a, smem, grid, stream);
    case 48: return launch_fwd<48>(tmx, tmy, a, smem, grid, stream);
    case 64: return launch_fwd<64>(tmx, tmy, a, smem, grid, stream);
    default:
      nr::set_error("nr_maxsim_fwd(bf16): Ny=%lld has no tensor-core instantiation (4,8,12,16,24,32,48,64)",
                    (long long)Ny);
      return -3;
  }
}

namespace nr {
// =================================================================================================
// backward: dOut[out tokens, d] += C[out tokens, src tokens] * Src[src tokens, d]
//
// C is the (at most one non-zero per (token, partner sample)) routing matrix of the max: it is never
// stored — the generator warps build each [128 x 64] bf16 tile of it in shared memory (128B-swizzled,
// K-major, exactly the image a TMA load would have produced) from the saved arg-max bytes, dH, the token
// weights and the masks, and the tensor core multiplies it with TMA-staged tiles of the TRANSPOSED source
// tokens (srcT [d, tokens], K-major for this product).  fp32 accumulators for all d <= 512 columns of a
// 128-token output tile fill the whole TMEM (2 x 256 columns); split-K over the source tokens spreads
// the few output tiles over all SMs and partials are combined with red.global.add.v4.f32.
//   side 0 ("gather"):  out = X tokens, src = Y tokens,  C[(rx,x),(ry,y)] = g(rx,ry) wx mx my [y == y*(rx,ry,x)]
//   side 1 ("scatter"): out = Y tokens, src = X tokens,  C[(ry,y),(rx,x)] = same entry, transposed role
// =================================================================================================
struct TcBwdArgs {
  const float* wx; const int64_t* mx; const int64_t* my; const uint8_t* ystar;
  const float* dH; int64_t dh_sr, dh_sc; float dh_scale;
  int Rx, Nx, Ry, Ny, D;
  float* dst;
  int out_tokens, src_tokens, n_mt, num_kb, KS, kb_per_split, n_half, half_cols, stages;
};

constexpr int TCB_B_HALF_BYTES = 256 * 128;   // one d-half of a source k-block: 256 rows x 128 B

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

template <int SIDE>
__global__ void __launch_bounds__(TC_THREADS, 1)
maxsim_bwd_tc_kernel(const __grid_constant__ CUtensorMap tms, const TcBwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int stage_bytes = TC_A_BYTES + a.n_half * TCB_B_HALF_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)a.stages * stage_bytes);
  uint64_t* b_full = bars;            // [stages] TMA -> MMA
  uint64_t* a_full = bars + 4;        // [stages] generators -> MMA
  uint64_t* empty = bars + 8;         // [stages] MMA -> TMA + generators
  uint64_t* acc_full = bars + 12;     // MMA -> epilogue
  uint64_t* acc_empty = bars + 13;    // epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_items = a.n_mt * a.KS;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tms);
    for (int s = 0; s < a.stages; ++s) { mbar_init(b_full + s, 1); mbar_init(a_full + s, 128); mbar_init(empty + s, 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 128);
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

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      const uint32_t tx_bytes = (uint32_t)a.n_half * (uint32_t)a.half_cols * 128u;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
        const int ks = item % a.KS;
        const int kb0 = ks * a.kb_per_split, kb1 = min(a.num_kb, kb0 + a.kb_per_split);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(empty + stage, phase ^ 1);
          uint8_t* sb = smem + (size_t)stage * stage_bytes + TC_A_BYTES;
          mbar_expect_tx(b_full + stage, tx_bytes);
          for (int h = 0; h < a.n_half; ++h)
            tma_load_2d(sb + h * TCB_B_HALF_BYTES, &tms, b_full + stage, kb * TC_BK, h * 256);
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(TC_BM, a.half_cols);
      int stage = 0; uint32_t phase = 0;
      int it = 0;
      for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
        const int ks = item % a.KS;
        const int kb0 = ks * a.kb_per_split, kb1 = min(a.num_kb, kb0 + a.kb_per_split);
        mbar_wait(acc_empty, (uint32_t)(it & 1) ^ 1);
        tc_fence_after();
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(a_full + stage, phase);
          mbar_wait(b_full + stage, phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t adesc = umma_desc_kmajor_sw128(sa);
          for (int h = 0; h < a.n_half; ++h) {
            const uint64_t bdesc = umma_desc_kmajor_sw128(sa + TC_A_BYTES + h * TCB_B_HALF_BYTES);
#pragma unroll
            for (int k = 0; k < TC_BK / 16; ++k)
              umma_bf16(tmem_base + (uint32_t)(h * 256), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                        (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(empty + stage);
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
        umma_commit(acc_full);
      }
    }
  } else {
    // ===================== generators + epilogue (warps 2..5) =====================
    const int q = warp & 3;
    const int m = q * 32 + lane;                       // row of the output tile / TMEM lane
    int stage = 0; uint32_t phase = 0;
    int it = 0;
    const int No = SIDE == 0 ? a.Nx : a.Ny;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int mt = item / a.KS, ks = item % a.KS;
      const int kb0 = ks * a.kb_per_split, kb1 = min(a.num_kb, kb0 + a.kb_per_split);
      const int g = mt * TC_BM + m;                    // global output token
      const bool valid = g < a.out_tokens;
      const int ro = valid ? g / No : 0, o = valid ? g - ro * No : 0;   // (sample, token) of the output row
      float coefx = 0.f;                               // side 0: wx*scale of this X token
      if (SIDE == 0 && valid) coefx = a.wx[g] * a.dh_scale;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(empty + stage, phase ^ 1);
        uint8_t* sa = smem + (size_t)stage * stage_bytes;
        // cooperative zero fill of this warp's 32 rows (512 contiguous bytes per store instruction)
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<uint4*>(sa + (q * 32 + i * 4) * 128 + lane * 16) = make_uint4(0u, 0u, 0u, 0u);
        __syncwarp();
        const int t0 = kb * TC_BK;
        uint8_t* srow = sa + m * 128;
        if (SIDE == 0) {
          if (coefx != 0.f) {
            const int ry_lo = t0 / a.Ny, ry_hi = min(a.Ry - 1, (t0 + TC_BK - 1) / a.Ny);
            for (int rb = ry_lo; rb <= ry_hi; rb += 8) {          // 8 independent loads in flight
              int yv[8]; float gv[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int ry = min(rb + i, ry_hi);
                yv[i] = a.ystar[((int64_t)ro * a.Ry + ry) * a.Nx + o];
                gv[i] = a.dH[(int64_t)ro * a.dh_sr + (int64_t)ry * a.dh_sc];
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int ry = rb + i;
                const int t = ry * a.Ny + yv[i] - t0;
                if (ry <= ry_hi && yv[i] != 255 && t >= 0 && t < TC_BK)
                  *reinterpret_cast<__nv_bfloat16*>(srow + (((t >> 3) ^ (m & 7)) << 4) + (t & 7) * 2) =
                      __float2bfloat16_rn(gv[i] * coefx);
              }
            }
          }
        } else {
          // pair scatter: thread = (source token j, every other partner sample of this output tile); the single
          // non-zero of pair (j, ry) lands in row (ry, y*) of the tile
          asm volatile("bar.sync 1, 128;" ::: "memory");       // all rows zeroed before foreign-row stores
          const int et = threadIdx.x - 64;
          const int j = et & 63;
          const int tsrc = t0 + j;
          if (tsrc < a.src_tokens) {
            const int rx = tsrc / a.Nx, x = tsrc - rx * a.Nx;
            const float cw = a.wx[tsrc] * a.dh_scale;
            const int row0 = mt * TC_BM;
            const int ry_lo = row0 / a.Ny, ry_hi = min(a.Ry - 1, (row0 + TC_BM - 1) / a.Ny);
            for (int rb = ry_lo + (et >> 6); rb <= ry_hi; rb += 16) {
              int yv[8]; float gv[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int ry = min(rb + 2 * i, ry_hi);
                yv[i] = a.ystar[((int64_t)rx * a.Ry + ry) * a.Nx + x];
                gv[i] = a.dH[(int64_t)rx * a.dh_sr + (int64_t)ry * a.dh_sc];
              }
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const int ry = rb + 2 * i;
                const int mm = ry * a.Ny + yv[i] - row0;
                if (ry <= ry_hi && yv[i] != 255 && mm >= 0 && mm < TC_BM)
                  *reinterpret_cast<__nv_bfloat16*>(sa + mm * 128 + (((j >> 3) ^ (mm & 7)) << 4) + (j & 7) * 2) =
                      __float2bfloat16_rn(gv[i] * cw);
              }
            }
          }
        }
        fence_proxy_async();                             // generic-proxy stores -> visible to the tensor core
        mbar_arrive(a_full + stage);
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
      }
      // ---- epilogue: TMEM -> red.global.add ----
      mbar_wait(acc_full, (uint32_t)(it & 1));
      tc_fence_after();
      if (kb1 > kb0) {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        float* drow = a.dst + (int64_t)g * a.D;
        for (int c = 0; c < a.D; c += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + (uint32_t)((c >> 8) * 256 + (c & 255)), v);
          tmem_ld_wait();
          reg_fence<16>(v);
          if (valid) {
#pragma unroll
            for (int e = 0; e < 16; e += 4)
              red_add_v4(drow + c + e, __uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]),
                         __uint_as_float(v[e + 3]));
          }
        }
      }
      tc_fence_before();
      mbar_arrive(acc_empty);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

// 2-D bf16 [d, ld] (transposed tokens) -> boxes {64 tokens, box_rows d-rows}, 128B swizzle
int make_tmap_srcT(CUtensorMap* m, const void* base, int64_t d, int64_t tokens, int64_t ld, int box_rows) {
  EncodeTiledFn enc = get_encode();
  NR_CHECK_ARG(enc != nullptr, "cuTensorMapEncodeTiled unavailable (driver too old?)");
  cuuint64_t gdim[2] = {(cuuint64_t)tokens, (cuuint64_t)d};
  cuuint64_t gstr[1] = {(cuuint64_t)ld * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  if (const char* pv = getenv("NR_TMA_PROMO")) {        // measurement knob: 0 none, 64, 128, 256
    const int v = atoi(pv);
    promo = v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 64 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
          : v == 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
  }
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), gdim, gstr, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  NR_CHECK_ARG(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(srcT) failed (%d) d=%lld tokens=%lld ld=%lld", (int)r,
               (long long)d, (long long)tokens, (long long)ld);
  return 0;
}

// tiled transpose of the bf16 operand copy: in [rows, d] -> out [d, ld]; 64x64 tiles, 8-byte accesses both ways
__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, int rows, int d, __nv_bfloat16* __restrict__ out,
                      int64_t ld) {
  __shared__ __nv_bfloat16 tile[64][68];
  const int r0 = blockIdx.x * 64, c0 = blockIdx.y * 64, tid = threadIdx.x;
  const bool vec_in = (d % 4 == 0), vec_out = (ld % 4 == 0);
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int e = it * 256 + tid, i = e >> 4, j = (e & 15) * 4;       // row i, columns j..j+3 of the tile
    const int r = r0 + i, c = c0 + j;
    __nv_bfloat16 v[4];
    if (r < rows && vec_in && c + 3 < d) {
      *reinterpret_cast<uint2*>(v) = *reinterpret_cast<const uint2*>(in + (int64_t)r * d + c);
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q) v[q] = (r < rows && c + q < d) ? in[(int64_t)r * d + c + q] : __float2bfloat16_rn(0.f);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) tile[i][j + q] = v[q];
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const int e = it * 256 + tid, i = e >> 4, j = (e & 15) * 4;       // out row (column c0+i), tokens r0+j..r0+j+3
    const int c = c0 + i, r = r0 + j;
    if (c >= d) continue;
    __nv_bfloat16 v[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) v[q] = tile[j + q][i];
    if (vec_out && r + 3 < rows) {
      *reinterpret_cast<uint2*>(out + (int64_t)c * ld + r) = *reinterpret_cast<uint2*>(v);
    } else {
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (r + q < rows) out[(int64_t)c * ld + r + q] = v[q];
    }
  }
}

}  // namespace nr

using namespace nr;

extern "C" int nr_transpose_tokens_bf16(const void* xn_bf16, int64_t rows, int64_t d, void* out, int64_t ld,
                                        void* stream) {
  NR_CHECK_ARG(xn_bf16 && out && rows > 0 && d > 0, "nr_transpose_tokens_bf16: bad arguments");
  NR_CHECK_ARG(ld >= rows && ld % 8 == 0, "nr_transpose_tokens_bf16: ld=%lld must be >= rows and a multiple of 8",
               (long long)ld);
  dim3 grid((unsigned)((rows + 63) / 64), (unsigned)((d + 63) / 64));
  transpose_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)xn_bf16, (int)rows, (int)d,
                                                                 (__nv_bfloat16*)out, ld);
  NR_CHECK_LAUNCH("nr_transpose_tokens_bf16");
  return 0;
}

// side 0: srcT = transposed Y tokens [d, src_ld]; side 1: srcT = transposed X tokens
int nr_maxsim_bwd_tc(int side, const void* srcT, int64_t src_ld, const float* wx, const int64_t* mx,
                     const int64_t* my, const uint8_t* ystar, const float* dH, int64_t dh_sr, int64_t dh_sc,
                     float dh_scale, int64_t Rx, int64_t Nx, int64_t Ry, int64_t Ny, int64_t d, float* dst,
                     cudaStream_t stream) {
  NR_CHECK_ARG(d % 16 == 0 && d <= 512 && (d <= 256 || d == 512),
               "nr_maxsim_bwd(bf16): d=%lld unsupported (multiple of 16 up to 256, or 512)", (long long)d);
  NR_CHECK_ARG(src_ld % 8 == 0 && ((uintptr_t)srcT & 15) == 0, "nr_maxsim_bwd(bf16): srcT must be 16B aligned, ld %% 8 == 0");
  TcBwdArgs a{};
  a.wx = wx; a.mx = mx; a.my = my; a.ystar = ystar; a.dH = dH; a.dh_sr = dh_sr; a.dh_sc = dh_sc; a.dh_scale = dh_scale;
  a.Rx = (int)Rx; a.Nx = (int)Nx; a.Ry = (int)Ry; a.Ny = (int)Ny; a.D = (int)d; a.dst = dst;
  a.out_tokens = side == 0 ? (int)(Rx * Nx) : (int)(Ry * Ny);
  a.src_tokens = side == 0 ? (int)(Ry * Ny) : (int)(Rx * Nx);
  a.n_mt = (a.out_tokens + TC_BM - 1) / TC_BM;
  a.num_kb = (a.src_tokens + TC_BK - 1) / TC_BK;
  a.n_half = d > 256 ? 2 : 1;
  a.half_cols = d > 256 ? 256 : (int)d;
  int dev = 0, sms = 0;
  NR_CUDA(cudaGetDevice(&dev));
  NR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int ks = (sms + a.n_mt / 2) / a.n_mt;           // split-K so that n_mt*KS ~ one wave
  if (ks < 1) ks = 1;
  if (ks > a.num_kb) ks = a.num_kb;
  a.kb_per_split = (a.num_kb + ks - 1) / ks;
  a.KS = (a.num_kb + a.kb_per_split - 1) / a.kb_per_split;
  a.stages = 2;
  const size_t stage_bytes = (size_t)TC_A_BYTES + (size_t)a.n_half * TCB_B_HALF_BYTES;
  const size_t smem = a.stages * stage_bytes + 256 + 1024;
  CUtensorMap tms;
  if (int e = make_tmap_srcT(&tms, srcT, d, a.src_tokens, src_ld, a.half_cols)) return e;
  const int items = a.n_mt * a.KS;
  const int grid = items < sms ? items : sms;
  if (side == 0) {
    NR_CUDA(cudaFuncSetAttribute(maxsim_bwd_tc_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    maxsim_bwd_tc_kernel<0><<<grid, TC_THREADS, smem, stream>>>(tms, a);
  } else {
    NR_CUDA(cudaFuncSetAttribute(maxsim_bwd_tc_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    maxsim_bwd_tc_kernel<1><<<grid, TC_THREADS, smem, stream>>>(tms, a);
  }
  NR_CHECK_LAUNCH("nr_maxsim_bwd(bf16)");
  return 0;
}
