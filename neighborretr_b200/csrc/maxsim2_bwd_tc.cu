// Backward of the fused two-direction max-sim (maxsim2_tc.cu) on the tensor cores.
//
// With S[rx,ry] = alpha (sum_x wx max_y R + sum_y wy max_x R) and g = dL/dS, the gradient w.r.t. the normalised
// tokens is ONE sparse routing matrix applied from either side:
//   C[(rx,x),(ry,y)] = g[rx,ry] * ( wx[rx,x] * [y == ystar[rx,ry,x]]  +  wy[ry,y] * [x == xstar[rx,ry,y]] )
//   dX = C * Y,   dY = C^T * X                    (<= Nx + Ny non-zeros per Nx x Ny block of C)
// (autograd of reference NeighborRetr/models/modeling.py:499-509, where it runs as a dense backward through the
// 4-D tensor).  C is never stored: the generator warps build each [128 out tokens x 64 source tokens] bf16 tile of
// it in shared memory, 128B-swizzled and K-major exactly as a TMA load would have left it — first the entries found
// from the out-token side (one per row and partner sample: race-free plain stores), then, after a barrier, the
// entries found from the source-token side (one per column and out sample: read-modify-write of distinct
// elements) — and tcgen05.mma multiplies it with TMA-staged tiles of the TRANSPOSED source tokens.  fp32
// accumulators for all d <= 512 columns of the 128-token output tile fill TMEM (2 x 256 columns); split-K over
// the source tokens spreads the few output tiles over all SMs, partials are combined with red.global.add.v4.f32.
// The one-direction kernels (maxsim_tc.cu) need 4 launches and twice the MMA work for the same two gradients.
#include "common.cuh"
#include "nrhead_internal.h"
#include "tc_common.cuh"
#include <stdlib.h>

namespace nr {
using namespace tc;

int make_tmap_srcT(CUtensorMap* m, const void* base, int64_t d, int64_t tokens, int64_t ld, int box_rows);

constexpr int B2_THREADS = 192;               // TMA warp, MMA warp, 4 generator/epilogue warps
constexpr int B2_BM = 128;
constexpr int B2_BK = 64;
constexpr int B2_A_BYTES = B2_BM * 128;       // 16 KB
constexpr int B2_B_HALF_BYTES = 256 * 128;    // one d-half of a source k-block: 256 rows x 128 B

// "out" side O (rows of C / of the result), "source" side S (columns of C / rows of the staged operand)
struct Tc2BwdArgs {
  const float* wO; const float* wS;           // token weights [Ro,No], [Rs,Ns]
  const uint8_t* starO; int64_t aO_o, aO_s;   // arg-max over source tokens per out token:  starO[ro*aO_o + rs*aO_s + o]
  const uint8_t* starS; int64_t aS_o, aS_s;   // arg-max over out tokens per source token:  starS[ro*aS_o + rs*aS_s + s]
  const float* dH; int64_t g_o, g_s; float scale;   // g(ro,rs) = dH[ro*g_o + rs*g_s] * scale
  int Ro, No, Rs, Ns, D;
  float* dst;
  int out_tokens, src_tokens, n_mt, num_kb, KS, kb_per_split, n_half, half_cols, stages;
  int debug;                                  // timing experiments only (NR_B2_DEBUG): 1 no red.add, 2 no generator, 4 no TMA
};

__device__ __forceinline__ void red_add_v4_(float* addr, float a, float b, float c, float d) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

__global__ void __launch_bounds__(B2_THREADS, 1)
maxsim2_bwd_tc_kernel(const __grid_constant__ CUtensorMap tms, const Tc2BwdArgs a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int stage_bytes = B2_A_BYTES + a.n_half * B2_B_HALF_BYTES;
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
          uint8_t* sb = smem + (size_t)stage * stage_bytes + B2_A_BYTES;
          if (a.debug & 4) { mbar_arrive(b_full + stage); }
          else {
          mbar_expect_tx(b_full + stage, tx_bytes);
          for (int h = 0; h < a.n_half; ++h)
            tma_load_2d(sb + h * B2_B_HALF_BYTES, &tms, b_full + stage, kb * B2_BK, h * 256);
          }
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(B2_BM, a.half_cols);
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
            const uint64_t bdesc = umma_desc_kmajor_sw128(sa + B2_A_BYTES + h * B2_B_HALF_BYTES);
#pragma unroll
            for (int k = 0; k < B2_BK / 16; ++k)
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
    const int et = threadIdx.x - 64;                   // 0..127
    int stage = 0; uint32_t phase = 0;
    int it = 0;
    const int No = a.No, Ns = a.Ns;
    for (int item = blockIdx.x; item < n_items; item += gridDim.x, ++it) {
      const int mt = item / a.KS, ks = item % a.KS;
      const int kb0 = ks * a.kb_per_split, kb1 = min(a.num_kb, kb0 + a.kb_per_split);
      const int row0 = mt * B2_BM;
      const int g = row0 + m;                          // global output token
      const bool valid = g < a.out_tokens;
      const int ro = valid ? g / No : 0, o = valid ? g - ro * No : 0;   // (sample, token) of the output row
      const float coefo = valid ? a.wO[g] * a.scale : 0.f;
      const int ro_lo = row0 / No, ro_hi = min(a.Ro - 1, (row0 + B2_BM - 1) / No);   // out samples touching this tile
      // Routing data (arg-max bytes, upstream gradients) of a k-block is loaded one k-block AHEAD into registers: the
      // dependent global loads would otherwise sit on the generator's critical path (2 stages cannot hide them).
      // Slots cover 8 partner samples per row / 8 out samples per column thread; longer ranges (tiny Ns / No) take
      // the direct-load remainder loops below.
      const int j = et & 63, jh = et >> 6;               // scatter side: source column and which half of the out samples
      const uint8_t* stO = a.starO + (int64_t)ro * a.aO_o + o;
      const float* gpO = a.dH + (int64_t)ro * a.g_o;
      struct Pre { int sv[8]; float gv[8]; int ov[8]; float hv[8]; float cw; };
      auto load_pre = [&](int kb, Pre& P) {
        const int t0 = kb * B2_BK;
        const int rs_lo = t0 / Ns, rs_hi = min(a.Rs - 1, (t0 + B2_BK - 1) / Ns);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rs = min(rs_lo + i, rs_hi);
          P.sv[i] = (coefo != 0.f) ? (int)stO[(int64_t)rs * a.aO_s] : 0;
          P.gv[i] = (coefo != 0.f) ? gpO[(int64_t)rs * a.g_s] : 0.f;
        }
        const int tsrc = min(t0 + j, a.src_tokens - 1);
        const int rs = tsrc / Ns, sidx = tsrc - rs * Ns;
        P.cw = (t0 + j < a.src_tokens) ? a.wS[tsrc] * a.scale : 0.f;
        const uint8_t* st = a.starS + (int64_t)rs * a.aS_s + sidx;
        const float* gp = a.dH + (int64_t)rs * a.g_s;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r2 = min(ro_lo + jh + 2 * i, ro_hi);
          P.ov[i] = (P.cw != 0.f) ? (int)st[(int64_t)r2 * a.aS_o] : 0;
          P.hv[i] = (P.cw != 0.f) ? gp[(int64_t)r2 * a.g_o] : 0.f;
        }
      };
      Pre cur, nxt;
      if (kb0 < kb1) load_pre(kb0, cur);
      for (int kb = kb0; kb < kb1; ++kb) {
        if (kb + 1 < kb1) load_pre(kb + 1, nxt);
        mbar_wait(empty + stage, phase ^ 1);
        uint8_t* sa = smem + (size_t)stage * stage_bytes;
        if (a.debug & 2) {
          fence_proxy_async();
          mbar_arrive(a_full + stage);
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
          continue;
        }
        // cooperative zero fill of this warp's 32 rows (512 contiguous bytes per store instruction)
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<uint4*>(sa + (q * 32 + i * 4) * 128 + lane * 16) = make_uint4(0u, 0u, 0u, 0u);
        __syncwarp();
        const int t0 = kb * B2_BK;
        // (1) out-token side: row m, one entry per source sample overlapping this k-block
        if (coefo != 0.f) {
          uint8_t* srow = sa + m * 128;
          const int rs_lo = t0 / Ns, rs_hi = min(a.Rs - 1, (t0 + B2_BK - 1) / Ns);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rs = rs_lo + i;
            const int t = rs * Ns + cur.sv[i] - t0;
            if (rs <= rs_hi && t >= 0 && t < B2_BK)
              *reinterpret_cast<__nv_bfloat16*>(srow + (((t >> 3) ^ (m & 7)) << 4) + (t & 7) * 2) =
                  __float2bfloat16_rn(cur.gv[i] * coefo);
          }
          for (int rs = rs_lo + 8; rs <= rs_hi; ++rs) {          // remainder (Ns < 10 only)
            const int t = rs * Ns + (int)stO[(int64_t)rs * a.aO_s] - t0;
            if (t >= 0 && t < B2_BK)
              *reinterpret_cast<__nv_bfloat16*>(srow + (((t >> 3) ^ (m & 7)) << 4) + (t & 7) * 2) =
                  __float2bfloat16_rn(gpO[(int64_t)rs * a.g_s] * coefo);
          }
        }
        asm volatile("bar.sync 1, 128;" ::: "memory");         // rows zeroed and side-(1) entries in place
        // (2) source-token side: column j, one entry per out sample of this tile; lands in row (ro, o*) and is
        //     ADDED to whatever side (1) put there (mutual arg-max pairs)
        if (cur.cw != 0.f) {
          auto put = [&](int r2, int ostar, float gval) {
            const int mm = r2 * No + ostar - row0;
            if (mm >= 0 && mm < B2_BM) {
              __nv_bfloat16* e =
                  reinterpret_cast<__nv_bfloat16*>(sa + mm * 128 + (((j >> 3) ^ (mm & 7)) << 4) + (j & 7) * 2);
              *e = __float2bfloat16_rn(__bfloat162float(*e) + gval * cur.cw);
            }
          };
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r2 = ro_lo + jh + 2 * i;
            if (r2 <= ro_hi) put(r2, cur.ov[i], cur.hv[i]);
          }
          if (ro_lo + jh + 16 <= ro_hi) {                        // remainder (No < 10 only)
            const int tsrc = t0 + j;
            const int rs = tsrc / Ns, sidx = tsrc - rs * Ns;
            const uint8_t* st = a.starS + (int64_t)rs * a.aS_s + sidx;
            const float* gp = a.dH + (int64_t)rs * a.g_s;
            for (int r2 = ro_lo + jh + 16; r2 <= ro_hi; r2 += 2)
              put(r2, (int)st[(int64_t)r2 * a.aS_o], gp[(int64_t)r2 * a.g_o]);
          }
        }
        fence_proxy_async();                             // generic-proxy stores -> visible to the tensor core
        mbar_arrive(a_full + stage);
        if (++stage == a.stages) { stage = 0; phase ^= 1; }
        cur = nxt;
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
          if (valid && !(a.debug & 1)) {
#pragma unroll
            for (int e = 0; e < 16; e += 4)
              red_add_v4_(drow + c + e, __uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]),
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

// dwx[rx,x] += scale * sum_ry g[rx,ry] pmax_x[rx,ry,x]   (blocks [0,Rx): one X sample each)
// dwy[ry,y] += scale * sum_rx g[rx,ry] pmax_y[rx,ry,y]   (remaining blocks: 256 columns x a slice of rx, atomics)
__global__ void __launch_bounds__(256)
maxsim2_bwd_w_kernel(const float* __restrict__ pmax_x, const float* __restrict__ pmax_y, const float* __restrict__ dH,
                     int64_t dh_sr, int64_t dh_sc, float scale, int Rx, int Nx, int Ry, int Ny, int nbx, int ncb,
                     int rsplit, float* __restrict__ dwx, float* __restrict__ dwy) {
  __shared__ float part[256];
  const int tid = threadIdx.x;
  if ((int)blockIdx.x < nbx) {
    const int rx = blockIdx.x;
    const int lanes = 256 / Nx;            // ry-lanes per x (Nx <= 128)
    const int x = tid % Nx, l = tid / Nx;
    float s = 0.f;
    if (l < lanes) {
#pragma unroll 8
      for (int ry = l; ry < Ry; ry += lanes)
        s += dH[(int64_t)rx * dh_sr + (int64_t)ry * dh_sc] * pmax_x[((int64_t)rx * Ry + ry) * Nx + x];
    }
    part[tid] = s;
    __syncthreads();
    if (tid < Nx) {
      float t = 0.f;
      for (int qq = 0; qq < lanes; ++qq) t += part[qq * Nx + tid];
      dwx[(int64_t)rx * Nx + tid] += t * scale;
    }
  } else {
    const int bb = blockIdx.x - nbx;
    const int cb = bb % ncb, sl = bb / ncb;
    const int c = cb * 256 + tid;
    const int cols = Ry * Ny;
    if (c >= cols) return;
    const int ry = c / Ny;
    const int per = (Rx + rsplit - 1) / rsplit;
    const int r0 = sl * per, r1 = min(Rx, r0 + per);
    float s = 0.f;
#pragma unroll 4
    for (int rx = r0; rx < r1; ++rx)
      s += dH[(int64_t)rx * dh_sr + (int64_t)ry * dh_sc] * pmax_y[(int64_t)rx * cols + c];
    atomicAdd(dwy + c, s * scale);
  }
}

}  // namespace nr

using namespace nr;

/* side 0: gradient w.r.t. the X tokens (srcT = transposed Y tokens), side 1: w.r.t. the Y tokens (srcT = X). */
extern "C" int nr_maxsim2_bwd(int side, const void* srcT, int64_t src_ld, const float* wx, const float* wy,
                              const uint8_t* ystar, const uint8_t* xstar, const float* dH, int64_t dh_sr,
                              int64_t dh_sc, float dh_scale, int64_t Rx, int64_t Nx, int64_t Ry, int64_t Ny, int64_t d,
                              float* dst, void* stream) {
  NR_CHECK_ARG(srcT && wx && wy && ystar && xstar && dH && dst, "nr_maxsim2_bwd: null pointer");
  NR_CHECK_ARG(Rx > 0 && Ry > 0 && Nx >= 1 && Nx <= NR_MAX_TOKENS && Ny >= 1 && Ny <= NR_MAX_TOKENS,
               "nr_maxsim2_bwd: bad sizes Rx=%lld Nx=%lld Ry=%lld Ny=%lld", (long long)Rx, (long long)Nx, (long long)Ry,
               (long long)Ny);
  NR_CHECK_ARG(d % 16 == 0 && d <= 512 && (d <= 256 || d == 512),
               "nr_maxsim2_bwd: d=%lld unsupported (multiple of 16 up to 256, or 512)", (long long)d);
  NR_CHECK_ARG(src_ld % 8 == 0 && ((uintptr_t)srcT & 15) == 0, "nr_maxsim2_bwd: srcT must be 16B aligned, ld %% 8 == 0");
  Tc2BwdArgs a{};
  if (side == 0) {
    a.wO = wx; a.wS = wy;
    a.starO = ystar; a.aO_o = Ry * Nx; a.aO_s = Nx;
    a.starS = xstar; a.aS_o = Ry * Ny; a.aS_s = Ny;
    a.g_o = dh_sr; a.g_s = dh_sc;
    a.Ro = (int)Rx; a.No = (int)Nx; a.Rs = (int)Ry; a.Ns = (int)Ny;
  } else {
    a.wO = wy; a.wS = wx;
    a.starO = xstar; a.aO_o = Ny; a.aO_s = Ry * Ny;
    a.starS = ystar; a.aS_o = Nx; a.aS_s = Ry * Nx;
    a.g_o = dh_sc; a.g_s = dh_sr;
    a.Ro = (int)Ry; a.No = (int)Ny; a.Rs = (int)Rx; a.Ns = (int)Nx;
  }
  a.dH = dH; a.scale = dh_scale; a.D = (int)d; a.dst = dst;
  a.out_tokens = a.Ro * a.No;
  a.src_tokens = a.Rs * a.Ns;
  a.n_mt = (a.out_tokens + B2_BM - 1) / B2_BM;
  a.num_kb = (a.src_tokens + B2_BK - 1) / B2_BK;
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
  if (const char* dbg = getenv("NR_B2_DEBUG")) a.debug = atoi(dbg);
  if (const char* ksv = getenv("NR_B2_KS")) {
    int k2 = atoi(ksv);
    if (k2 >= 1 && k2 <= a.num_kb) { a.kb_per_split = (a.num_kb + k2 - 1) / k2; a.KS = (a.num_kb + a.kb_per_split - 1) / a.kb_per_split; }
  }
  const size_t stage_bytes = (size_t)B2_A_BYTES + (size_t)a.n_half * B2_B_HALF_BYTES;
  const size_t smem = a.stages * stage_bytes + 256 + 1024;
  CUtensorMap tms;
  if (int e = make_tmap_srcT(&tms, srcT, d, a.src_tokens, src_ld, a.half_cols)) return e;
  const int items = a.n_mt * a.KS;
  const int grid = items < sms ? items : sms;
  NR_CUDA(cudaFuncSetAttribute(maxsim2_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  maxsim2_bwd_tc_kernel<<<grid, B2_THREADS, smem, (cudaStream_t)stream>>>(tms, a);
  NR_CHECK_LAUNCH("nr_maxsim2_bwd");
  return 0;
}

extern "C" int nr_maxsim2_bwd_w(const float* pmax_x, const float* pmax_y, const float* dH, int64_t dh_sr, int64_t dh_sc,
                                float dh_scale, int64_t Rx, int64_t Nx, int64_t Ry, int64_t Ny, float* dwx, float* dwy,
                                void* stream) {
  NR_CHECK_ARG(dH && Rx > 0 && Ry > 0 && Nx > 0 && Nx <= NR_MAX_TOKENS && Ny > 0 && Ny <= NR_MAX_TOKENS,
               "nr_maxsim2_bwd_w: bad arguments");
  NR_CHECK_ARG((!dwx || pmax_x) && (!dwy || pmax_y) && (dwx || dwy), "nr_maxsim2_bwd_w: missing pmax for a requested gradient");
  const int nbx = dwx ? (int)Rx : 0;
  const int ncb = dwy ? (int)((Ry * Ny + 255) / 256) : 0;
  int rsplit = 1;
  if (dwy) {
    rsplit = (296 + ncb - 1) / ncb;               // ~2 CTAs per SM in total
    if (rsplit > Rx) rsplit = (int)Rx;
    if (rsplit < 1) rsplit = 1;
  }
  const int grid = nbx + ncb * rsplit;
  maxsim2_bwd_w_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pmax_x, pmax_y, dH, dh_sr, dh_sc, dh_scale, (int)Rx,
                                                               (int)Nx, (int)Ry, (int)Ny, nbx, ncb, rsplit, dwx, dwy);
  NR_CHECK_LAUNCH("nr_maxsim2_bwd_w");
  return 0;
}
