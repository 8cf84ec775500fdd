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

constexpr int B2_THREADS = 320;               // TMA warp, MMA warp, 2 generator/epilogue groups of 4 warps
constexpr int B2_GROUP = 128;
constexpr int B2_BM = 128;
constexpr int B2_BK = 64;
constexpr int B2_A_BYTES = B2_BM * 128;       // 16 KB
constexpr int B2_B_HALF_BYTES = 256 * 128;    // one d-half of a source k-block: 256 rows x 128 B
constexpr int B2_MAX_JOBS = 12;
constexpr int B2_MAX_SIDES = 4;
constexpr int B2_MAX_SRC = 6;

// One job = one source operand S contributing to the gradient of one output operand O ("side"):
//   starO[ro*aO_o + rs*aO_s + o] arg-max over the source tokens for out token o of the pair (ro, rs)
//   starS[ro*aS_o + rs*aS_s + s] arg-max over the out tokens for source token s
//   g(ro, rs) = dH[ro*g_o + rs*g_s] * scale
struct B2Src {
  const float* wS;
  const uint8_t* starO; int64_t aO_o, aO_s;
  const uint8_t* starS; int64_t aS_o, aS_s;
  const float* dH; int64_t g_o, g_s; float scale;
  int Rs, Ns, src_tokens, num_kb, kb0;        // kb0: first k-block of this source in the side's concatenated K range
  int part;                                   // 0 plain bf16 tile; 2 / 3: hi / lo tile of the exact fp32 coefficient
  uint32_t ns_magic;                          // ceil(2^32 / Ns): t / Ns == umulhi(t, ns_magic) for t < 2^25
};
struct B2Side {
  const float* wO; float* dst;
  int Ro, No, out_tokens, n_mt, nsrc, kb_total, KS, kb_per_split, item0;
  uint32_t no_magic;                          // ceil(2^32 / No)
  int src[B2_MAX_SRC];                        // indices into srcs[] / tms[]
};
struct alignas(64) B2Args {
  CUtensorMap tms[B2_MAX_JOBS];
  B2Src srcs[B2_MAX_JOBS];
  B2Side sides[B2_MAX_SIDES];
  int nsides, n_items, D, n_half, half_cols, debug;
};

// t / n for 0 <= t < 2^25, 2 <= n <= 128 (exact up to n = 144), magic = ceil(2^32 / n) (0 stands for n = 1): one IMAD.HI instead of the
// ~25-instruction division
__device__ __forceinline__ int div_magic(int t, uint32_t magic) { return magic ? (int)__umulhi((uint32_t)t, magic) : t; }

// plain bf16 coefficient (part 0 / 2) or the remainder of its bf16 rounding (part 3)
__device__ __forceinline__ __nv_bfloat16 coef_part(float c, int part) {
  const __nv_bfloat16 hi = __float2bfloat16_rn(c);
  return part == 3 ? __float2bfloat16_rn(c - __bfloat162float(hi)) : hi;
}

__device__ __forceinline__ void red_add_v4_(float* addr, float a, float b, float c, float d) {
  asm volatile("red.relaxed.gpu.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d)
               : "memory");
}

// All gradient contractions of a step in ONE launch: an item = (side, 128-token output tile, K-split), where the K
// range of a side is the concatenation of its sources (e.g. text gradient: video tokens then bank-video tokens), so
// one accumulator pass and ONE red.add epilogue serve every pair that feeds the same output rows.
__global__ void __launch_bounds__(B2_THREADS, 1) maxsim2_bwd_tc_kernel(const __grid_constant__ B2Args a) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // offset arithmetic on the extern array (no integer round trip): the routing-tile stores stay STS with 32-bit addresses
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stage_bytes = B2_A_BYTES + a.n_half * B2_B_HALF_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)2 * stage_bytes);
  uint64_t* b_full = bars;            // [2] TMA -> MMA
  uint64_t* a_full = bars + 2;        // [2] generator group s -> MMA
  uint64_t* empty = bars + 4;         // [2] MMA -> TMA + generator group s
  uint64_t* acc_full = bars + 6;      // MMA -> epilogue
  uint64_t* acc_empty = bars + 7;     // epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 8);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < a.nsides; ++s)
      for (int q = 0; q < a.sides[s].nsrc; ++q) tma_prefetch_desc(&a.tms[a.sides[s].src[q]]);
    for (int s = 0; s < 2; ++s) { mbar_init(b_full + s, 1); mbar_init(a_full + s, B2_GROUP); mbar_init(empty + s, 1); }
    mbar_init(acc_full, 1);
    mbar_init(acc_empty, 2 * B2_GROUP);
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

  // item -> (side, output tile, concatenated k-block range)
  auto decode = [&](int item, int& sd, int& mt, int& kbA, int& kbB) {
    sd = 0;
    while (sd + 1 < a.nsides && item >= a.sides[sd + 1].item0) ++sd;
    const B2Side& S = a.sides[sd];
    const int local = item - S.item0;
    mt = local / S.KS;
    const int ks = local - mt * S.KS;
    kbA = ks * S.kb_per_split;
    kbB = min(S.kb_total, kbA + S.kb_per_split);
  };
  // concatenated k-block -> index of its source within the side
  auto which_src = [&](const B2Side& S, int kb) {
    int q = 0;
    while (q + 1 < S.nsrc && kb >= a.srcs[S.src[q + 1]].kb0) ++q;
    return q;
  };

  if (warp == 0) {
    if (lane == 0) {
      const uint32_t tx_bytes = (uint32_t)a.n_half * (uint32_t)a.half_cols * 128u;
      uint32_t c = 0;                                     // k-blocks issued by this CTA so far: stage = c & 1
      for (int item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        int sd, mt, kbA, kbB;
        decode(item, sd, mt, kbA, kbB);
        const B2Side& S = a.sides[sd];
        for (int kb = kbA; kb < kbB; ++kb, ++c) {
          const int stage = c & 1;
          const int ji = S.src[which_src(S, kb)];
          mbar_wait(empty + stage, ((c >> 1) & 1u) ^ 1u);
          uint8_t* sb = smem + (size_t)stage * stage_bytes + B2_A_BYTES;
          if (a.debug & 4) { mbar_arrive(b_full + stage); continue; }
          mbar_expect_tx(b_full + stage, tx_bytes);
          for (int h = 0; h < a.n_half; ++h)
            tma_load_2d(sb + h * B2_B_HALF_BYTES, &a.tms[ji], b_full + stage, (kb - a.srcs[ji].kb0) * B2_BK, h * 256);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(B2_BM, a.half_cols);
      uint32_t c = 0;
      int it = 0;
      for (int item = blockIdx.x; item < a.n_items; item += gridDim.x, ++it) {
        int sd, mt, kbA, kbB;
        decode(item, sd, mt, kbA, kbB);
        mbar_wait(acc_empty, (uint32_t)(it & 1) ^ 1);
        tc_fence_after();
        for (int kb = kbA; kb < kbB; ++kb, ++c) {
          const int stage = c & 1;
          const uint32_t par = (c >> 1) & 1u;
          mbar_wait(a_full + stage, par);
          mbar_wait(b_full + stage, par);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t adesc = umma_desc_kmajor_sw128(sa);
          for (int h = 0; h < a.n_half; ++h) {
            const uint64_t bdesc = umma_desc_kmajor_sw128(sa + B2_A_BYTES + h * B2_B_HALF_BYTES);
#pragma unroll
            for (int k = 0; k < B2_BK / 16; ++k)
              umma_bf16(tmem_base + (uint32_t)(h * 256), adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                        (kb > kbA || k > 0) ? 1u : 0u);
          }
          umma_commit(empty + stage);
        }
        umma_commit(acc_full);
      }
    }
  } else {
    // ===================== generator / epilogue groups (warps 2..5 = group 0 -> stage 0, 6..9 = group 1 -> stage 1)
    // Each group builds every other k-block (its own stage), so two routing tiles are under construction at any time
    // and the tensor core is fed at twice the single-group rate.
    const int grp = (warp - 2) >> 2;
    const int q = warp & 3;                            // TMEM lane quarter
    const int m = q * 32 + lane;                       // row of the output tile / TMEM lane
    const int et = (threadIdx.x - 64) & (B2_GROUP - 1);   // 0..127 within the group
    const int j = et & 63, jh = et >> 6;               // source-token side: column, and which half of the out samples
    uint8_t* const sa = smem + (size_t)grp * stage_bytes;
    uint32_t c = 0;
    int it = 0;
    struct Pre { int sv[8]; float gv[8]; int ov[8]; float hv[8]; float cw; int rs_lo, n_rs, rs_j, sidx_j; };
    for (int item = blockIdx.x; item < a.n_items; item += gridDim.x, ++it) {
      int sd, mt, kbA, kbB;
      decode(item, sd, mt, kbA, kbB);
      const B2Side& S = a.sides[sd];
      const int No = S.No;
      const int row0 = mt * B2_BM;
      const int g = row0 + m;                          // global output token
      const bool valid = g < S.out_tokens;
      const int ro = valid ? div_magic(g, S.no_magic) : 0, o = valid ? g - ro * No : 0;   // (sample, token) of the row
      const float wo = valid ? S.wO[g] : 0.f;
      const int ro_lo = div_magic(row0, S.no_magic);                   // out samples touching this tile
      const int ro_hi = min(S.Ro - 1, div_magic(row0 + B2_BM - 1, S.no_magic));
      // Routing data (arg-max bytes, upstream gradients) of a k-block is loaded into registers while the group's
      // previous k-block is being built.  Slots cover 8 partner samples per row / 8 out samples per column thread;
      // longer ranges (tiny Ns / No) take the direct-load remainder loops below.
      // (walking pointers and validity predicates instead of per-entry 64-bit index products and clamped indices)
      auto load_pre = [&](int kb, Pre& P) {
        const B2Src& J = a.srcs[S.src[which_src(S, kb)]];
        const int Ns = J.Ns;
        const int t0 = (kb - J.kb0) * B2_BK;
        const int rs_lo = div_magic(t0, J.ns_magic);
        const int n_rs = min(J.Rs - 1, div_magic(t0 + B2_BK - 1, J.ns_magic)) - rs_lo + 1;
        P.rs_lo = rs_lo; P.n_rs = n_rs;
        {
          const uint8_t* stO = J.starO + ((int64_t)ro * J.aO_o + (int64_t)rs_lo * J.aO_s + o);
          const float* gpO = J.dH + ((int64_t)ro * J.g_o + (int64_t)rs_lo * J.g_s);
          const int64_t so = J.aO_s, sg = J.g_s;
          const bool on = wo != 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const bool ok = on && i < n_rs;
            P.sv[i] = ok ? (int)*stO : 0;
            P.gv[i] = ok ? *gpO : 0.f;
            stO += so; gpO += sg;
          }
        }
        const bool cok = t0 + j < J.src_tokens;
        const int tsrc = cok ? t0 + j : J.src_tokens - 1;
        const int rs = div_magic(tsrc, J.ns_magic), sidx = tsrc - rs * Ns;
        P.rs_j = rs; P.sidx_j = sidx;
        P.cw = cok ? J.wS[tsrc] * J.scale : 0.f;
        {
          const int r2_0 = ro_lo + jh;
          const uint8_t* st = J.starS + ((int64_t)rs * J.aS_s + (int64_t)r2_0 * J.aS_o + sidx);
          const float* gp = J.dH + ((int64_t)rs * J.g_s + (int64_t)r2_0 * J.g_o);
          const int64_t so = 2 * J.aS_o, sg = 2 * J.g_o;
          const bool on = P.cw != 0.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const bool ok = on && r2_0 + 2 * i <= ro_hi;
            P.ov[i] = ok ? (int)*st : 0;
            P.hv[i] = ok ? *gp : 0.f;
            st += so; gp += sg;
          }
        }
      };
      // this group's k-blocks of the item: those whose CTA-wide counter has parity grp
      int kb = kbA + (int)((grp - (int)c) & 1);
      uint32_t cc = c + (uint32_t)(kb - kbA);
      Pre cur, nxt;
      if (kb < kbB) load_pre(kb, cur);
      for (; kb < kbB; kb += 2, cc += 2) {
        if (kb + 2 < kbB) load_pre(kb + 2, nxt);
        mbar_wait(empty + grp, ((cc >> 1) & 1u) ^ 1u);
        if (a.debug & 2) {
          fence_proxy_async();
          mbar_arrive(a_full + grp);
          continue;
        }
        const B2Src& J = a.srcs[S.src[which_src(S, kb)]];
        const int Ns = J.Ns;
        const float coefo = wo * J.scale;
        // cooperative zero fill of this warp's 32 rows (512 contiguous bytes per store instruction)
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<uint4*>(sa + (q * 32 + i * 4) * 128 + lane * 16) = make_uint4(0u, 0u, 0u, 0u);
        __syncwarp();
        const int t0 = (kb - J.kb0) * B2_BK;
        // (1) out-token side: row m, one entry per source sample overlapping this k-block
        if (coefo != 0.f) {
          uint8_t* srow = sa + m * 128;
          const int rs_lo = cur.rs_lo, rs_hi = rs_lo + cur.n_rs - 1;
          const int tb = rs_lo * Ns - t0;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int t = tb + i * Ns + cur.sv[i];
            if (i < cur.n_rs && (unsigned)t < (unsigned)B2_BK)
              *reinterpret_cast<__nv_bfloat16*>(srow + (((t >> 3) ^ (m & 7)) << 4) + (t & 7) * 2) =
                  coef_part(cur.gv[i] * coefo, J.part);
          }
          if (rs_lo + 8 <= rs_hi) {                                // remainder (Ns < 10 only)
            const uint8_t* stO = J.starO + (int64_t)ro * J.aO_o + o;
            const float* gpO = J.dH + (int64_t)ro * J.g_o;
            for (int rs = rs_lo + 8; rs <= rs_hi; ++rs) {
              const int t = rs * Ns + (int)stO[(int64_t)rs * J.aO_s] - t0;
              if (t >= 0 && t < B2_BK)
                *reinterpret_cast<__nv_bfloat16*>(srow + (((t >> 3) ^ (m & 7)) << 4) + (t & 7) * 2) =
                    coef_part(gpO[(int64_t)rs * J.g_s] * coefo, J.part);
            }
          }
        }
        if (grp == 0) asm volatile("bar.sync 1, 128;" ::: "memory");   // rows zeroed and side-(1) entries in place
        else asm volatile("bar.sync 2, 128;" ::: "memory");
        // (2) source-token side: column j, one entry per out sample of this tile; lands in row (ro, o*) and is
        //     ADDED to whatever side (1) put there (mutual arg-max pairs)
        if (cur.cw != 0.f) {
          const int rs_j = cur.rs_j, sidx_j = cur.sidx_j;   // cur.cw != 0 implies t0 + j < src_tokens
          auto put = [&](int r2, int ostar, float gval) {
            const int mm = r2 * No + ostar - row0;
            if ((unsigned)mm < (unsigned)B2_BM) {
              __nv_bfloat16* e =
                  reinterpret_cast<__nv_bfloat16*>(sa + mm * 128 + (((j >> 3) ^ (mm & 7)) << 4) + (j & 7) * 2);
              if (J.part == 0) {
                *e = __float2bfloat16_rn(__bfloat162float(*e) + gval * cur.cw);
              } else {
                // exact coefficient: the out-token-side contribution of this element (a mutual arg-max pair) is
                // recomputed in fp32 instead of being read back rounded, then the hi or lo part overwrites it
                const int go = r2 * No + ostar;
                const bool mutual = (int)J.starO[(int64_t)r2 * J.aO_o + (int64_t)rs_j * J.aO_s + ostar] == sidx_j;
                const float c1 = mutual ? gval * (S.wO[go] * J.scale) : 0.f;
                *e = coef_part(c1 + gval * cur.cw, J.part);
              }
            }
          };
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int r2 = ro_lo + jh + 2 * i;
            if (r2 <= ro_hi) put(r2, cur.ov[i], cur.hv[i]);
          }
          if (ro_lo + jh + 16 <= ro_hi) {                        // remainder (No < 10 only)
            const uint8_t* st = J.starS + (int64_t)rs_j * J.aS_s + sidx_j;
            const float* gp = J.dH + (int64_t)rs_j * J.g_s;
            for (int r2 = ro_lo + jh + 16; r2 <= ro_hi; r2 += 2)
              put(r2, (int)st[(int64_t)r2 * J.aS_o], gp[(int64_t)r2 * J.g_o]);
          }
        }
        fence_proxy_async();                             // generic-proxy stores -> visible to the tensor core
        mbar_arrive(a_full + grp);
        cur = nxt;
      }
      c += (uint32_t)(kbB - kbA);
      // ---- epilogue: TMEM -> red.global.add; group 0 takes the lower half of the d columns, group 1 the upper
      mbar_wait(acc_full, (uint32_t)(it & 1));
      tc_fence_after();
      {
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16);
        float* drow = S.dst + (int64_t)g * a.D;
        const int c_lo = grp * (a.D / 2), c_hi = c_lo + a.D / 2;
        for (int cidx = c_lo; cidx < c_hi; cidx += 16) {
          uint32_t v[16];
          tmem_ld16(taddr + (uint32_t)cidx, v);
          tmem_ld_wait();
          reg_fence<16>(v);
          if (valid && !(a.debug & 1)) {
#pragma unroll
            for (int e = 0; e < 16; e += 4)
              red_add_v4_(drow + cidx + e, __uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]),
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
// dwy[ry,y] += scale * sum_rx g[rx,ry] pmax_y[rx,ry,y]   (remaining blocks: 256 columns x a slice of rx)
// Up to 3 pairs (the batch pair and the two bank pairs of a step) share ONE launch; outputs shared between pairs
// (e.g. the text weights of the batch pair and of the text-vs-bank pair) are accumulated with atomics.
struct BwdWJob {
  const float* pmax_x; const float* pmax_y; const float* dH;
  int64_t dh_sr, dh_sc; float scale;
  int Rx, Nx, Ry, Ny, nbx, ncb, rsplit, blk0;
  float* dwx; float* dwy;
};
struct BwdWArgs { BwdWJob j[3]; int njobs; };

__global__ void __launch_bounds__(256) maxsim2_bwd_w_kernel(const BwdWArgs a) {
  __shared__ float part[256];
  int ji = 0;
  while (ji + 1 < a.njobs && (int)blockIdx.x >= a.j[ji + 1].blk0) ++ji;
  const BwdWJob& J = a.j[ji];
  const int blk = (int)blockIdx.x - J.blk0;
  const int tid = threadIdx.x;
  const int Nx = J.Nx, Ny = J.Ny, Rx = J.Rx, Ry = J.Ry;
  if (blk < J.nbx) {
    const int rx = blk;
    const int lanes = 256 / Nx;            // ry-lanes per x (Nx <= 128)
    const int x = tid % Nx, l = tid / Nx;
    float s = 0.f;
    if (l < lanes) {
#pragma unroll 8
      for (int ry = l; ry < Ry; ry += lanes)
        s += J.dH[(int64_t)rx * J.dh_sr + (int64_t)ry * J.dh_sc] * J.pmax_x[((int64_t)rx * Ry + ry) * Nx + x];
    }
    part[tid] = s;
    __syncthreads();
    if (tid < Nx) {
      float t = 0.f;
      for (int qq = 0; qq < lanes; ++qq) t += part[qq * Nx + tid];
      atomicAdd(J.dwx + (int64_t)rx * Nx + tid, t * J.scale);
    }
  } else {
    const int bb = blk - J.nbx;
    const int cb = bb % J.ncb, sl = bb / J.ncb;
    const int c = cb * 256 + tid;
    const int cols = Ry * Ny;
    if (c >= cols) return;
    const int ry = c / Ny;
    const int per = (Rx + J.rsplit - 1) / J.rsplit;
    const int r0 = sl * per, r1 = min(Rx, r0 + per);
    float s = 0.f;
#pragma unroll 4
    for (int rx = r0; rx < r1; ++rx)
      s += J.dH[(int64_t)rx * J.dh_sr + (int64_t)ry * J.dh_sc] * J.pmax_y[(int64_t)rx * cols + c];
    atomicAdd(J.dwy + c, s * J.scale);
  }
}

}  // namespace nr

using namespace nr;

/* All token-gradient contractions of a step in one launch; jobs that share `dst` are accumulated in one pass. */
extern "C" int nr_maxsim2_bwd(const nr_maxsim2_bwd_job* jobs, int njobs, int64_t Nx, int64_t Ny, int64_t d,
                              void* stream) {
  NR_CHECK_ARG(jobs && njobs >= 1 && njobs <= B2_MAX_JOBS, "nr_maxsim2_bwd: 1..%d jobs per launch (got %d)",
               B2_MAX_JOBS, njobs);
  NR_CHECK_ARG(Nx >= 1 && Nx <= NR_MAX_TOKENS && Ny >= 1 && Ny <= NR_MAX_TOKENS, "nr_maxsim2_bwd: bad Nx=%lld Ny=%lld",
               (long long)Nx, (long long)Ny);
  NR_CHECK_ARG(d % 32 == 0 && d <= 512 && (d <= 256 || d == 512),
               "nr_maxsim2_bwd: d=%lld unsupported (multiple of 32 up to 256, or 512)", (long long)d);
  B2Args a{};
  a.D = (int)d;
  a.n_half = d > 256 ? 2 : 1;
  a.half_cols = d > 256 ? 256 : (int)d;
  if (const char* dbg = getenv("NR_B2_DEBUG")) a.debug = atoi(dbg);
  for (int i = 0; i < njobs; ++i) {
    const nr_maxsim2_bwd_job& jb = jobs[i];
    NR_CHECK_ARG(jb.srcT && jb.wx && jb.wy && jb.ystar && jb.xstar && jb.dH && jb.dst && jb.Rx > 0 && jb.Ry > 0 &&
                     (jb.side == 0 || jb.side == 1),
                 "nr_maxsim2_bwd: job %d has a null pointer, an empty side or a bad side flag", i);
    NR_CHECK_ARG(jb.src_ld % 8 == 0 && ((uintptr_t)jb.srcT & 15) == 0 && ((uintptr_t)jb.dst & 15) == 0,
                 "nr_maxsim2_bwd: job %d: srcT / dst must be 16B aligned, ld %% 8 == 0", i);
    B2Src& J = a.srcs[i];
    const float* wO;
    int Ro, No;
    const int64_t Rx = jb.Rx, Ry = jb.Ry;
    if (jb.side == 0) {
      wO = jb.wx; J.wS = jb.wy;
      J.starO = jb.ystar; J.aO_o = Ry * Nx; J.aO_s = Nx;
      J.starS = jb.xstar; J.aS_o = Ry * Ny; J.aS_s = Ny;
      J.g_o = jb.dh_sr; J.g_s = jb.dh_sc;
      Ro = (int)Rx; No = (int)Nx; J.Rs = (int)Ry; J.Ns = (int)Ny;
    } else {
      wO = jb.wy; J.wS = jb.wx;
      J.starO = jb.xstar; J.aO_o = Ny; J.aO_s = Ry * Ny;
      J.starS = jb.ystar; J.aS_o = Nx; J.aS_s = Ry * Nx;
      J.g_o = jb.dh_sc; J.g_s = jb.dh_sr;
      Ro = (int)Ry; No = (int)Ny; J.Rs = (int)Rx; J.Ns = (int)Nx;
    }
    J.dH = jb.dH; J.scale = jb.dh_scale;
    NR_CHECK_ARG(jb.part == 0 || jb.part == 2 || jb.part == 3, "nr_maxsim2_bwd: job %d: part must be 0, 2 or 3", i);
    J.part = jb.part;
    J.src_tokens = J.Rs * J.Ns;
    J.ns_magic = (uint32_t)((((uint64_t)1 << 32) + (uint64_t)J.Ns - 1) / (uint64_t)J.Ns);
    NR_CHECK_ARG((int64_t)J.Rs * J.Ns < ((int64_t)1 << 25) && (int64_t)Ro * No < ((int64_t)1 << 25),
                 "nr_maxsim2_bwd: job %d: more than 2^25 tokens on one side", i);
    J.num_kb = (J.src_tokens + B2_BK - 1) / B2_BK;
    if (int e = make_tmap_srcT(&a.tms[i], jb.srcT, d, J.src_tokens, jb.src_ld, a.half_cols)) return e;
    // side = jobs with the same output rows
    int sd = -1;
    for (int s = 0; s < a.nsides; ++s)
      if (a.sides[s].dst == jb.dst && a.sides[s].Ro == Ro && a.sides[s].No == No && a.sides[s].wO == wO) sd = s;
    if (sd < 0) {
      NR_CHECK_ARG(a.nsides < B2_MAX_SIDES, "nr_maxsim2_bwd: more than %d distinct outputs in one launch", B2_MAX_SIDES);
      sd = a.nsides++;
      B2Side& S = a.sides[sd];
      S.wO = wO; S.dst = jb.dst; S.Ro = Ro; S.No = No; S.out_tokens = Ro * No;
      S.n_mt = (S.out_tokens + B2_BM - 1) / B2_BM;
      S.no_magic = (uint32_t)((((uint64_t)1 << 32) + (uint64_t)No - 1) / (uint64_t)No);
    }
    B2Side& S = a.sides[sd];
    NR_CHECK_ARG(S.nsrc < B2_MAX_SRC, "nr_maxsim2_bwd: more than %d sources for one output", B2_MAX_SRC);
    J.kb0 = S.kb_total;
    S.src[S.nsrc++] = i;
    S.kb_total += J.num_kb;
  }
  int dev = 0, sms = 0;
  NR_CUDA(cudaGetDevice(&dev));
  NR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  // split-K per side so that all items have about the same number of k-blocks.  Small launches (a step at b = 128 on
  // one GPU) fit ONE wave: per = work / SMs k-blocks per item.  Larger ones run several waves of items of <= ~40
  // k-blocks through the persistent grid instead of one wave of very long items whose longest (an unsplit side)
  // sets the kernel time (measured at 8 ranks: 463 us with one wave of up to 288 k-blocks per item).
  int64_t work = 0;
  int min_items = 0;
  for (int s = 0; s < a.nsides; ++s) { work += (int64_t)a.sides[s].n_mt * a.sides[s].kb_total; min_items += a.sides[s].n_mt; }
  int per = (int)((work + sms - 1) / sms);
  if (per < 1) per = 1;
  int waves = 1;
  if (per > 48) {
    waves = (per + 39) / 40;
    per = (int)((work + (int64_t)sms * waves - 1) / ((int64_t)sms * waves));
  }
  if (const char* pv = getenv("NR_B2_PER")) { int p2 = atoi(pv); if (p2 >= 1) per = p2; }
  for (;;) {
    int items = 0;
    for (int s = 0; s < a.nsides; ++s) {
      B2Side& S = a.sides[s];
      int ks = (S.kb_total + per / 2) / per;
      if (ks < 1) ks = 1;
      if (ks > S.kb_total) ks = S.kb_total;
      S.kb_per_split = (S.kb_total + ks - 1) / ks;
      S.KS = (S.kb_total + S.kb_per_split - 1) / S.kb_per_split;
      S.item0 = items;
      items += S.n_mt * S.KS;
    }
    a.n_items = items;
    if (items <= sms * waves || min_items > sms * waves || waves > 1 || getenv("NR_B2_PER")) break;
    ++per;
  }
  const size_t stage_bytes = (size_t)B2_A_BYTES + (size_t)a.n_half * B2_B_HALF_BYTES;
  const size_t smem = 2 * stage_bytes + 256 + 1024;
  const int grid = a.n_items < sms ? a.n_items : sms;
  NR_CUDA(cudaFuncSetAttribute(maxsim2_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  maxsim2_bwd_tc_kernel<<<grid, B2_THREADS, smem, (cudaStream_t)stream>>>(a);
  NR_CHECK_LAUNCH("nr_maxsim2_bwd");
  return 0;
}

static int bwd_w_fill(BwdWJob& J, const float* pmax_x, const float* pmax_y, const float* dH, int64_t dh_sr, int64_t dh_sc,
                      float dh_scale, int64_t Rx, int64_t Nx, int64_t Ry, int64_t Ny, float* dwx, float* dwy, int& blocks) {
  NR_CHECK_ARG(dH && Rx > 0 && Ry > 0 && Nx > 0 && Nx <= NR_MAX_TOKENS && Ny > 0 && Ny <= NR_MAX_TOKENS,
               "nr_maxsim2_bwd_w: bad arguments");
  NR_CHECK_ARG((!dwx || pmax_x) && (!dwy || pmax_y) && (dwx || dwy), "nr_maxsim2_bwd_w: missing pmax for a requested gradient");
  J.pmax_x = pmax_x; J.pmax_y = pmax_y; J.dH = dH; J.dh_sr = dh_sr; J.dh_sc = dh_sc; J.scale = dh_scale;
  J.Rx = (int)Rx; J.Nx = (int)Nx; J.Ry = (int)Ry; J.Ny = (int)Ny; J.dwx = dwx; J.dwy = dwy;
  J.nbx = dwx ? (int)Rx : 0;
  J.ncb = dwy ? (int)((Ry * Ny + 255) / 256) : 1;
  J.rsplit = 1;
  if (dwy) {
    J.rsplit = (296 + J.ncb - 1) / J.ncb;               // ~2 CTAs per SM in total
    if (J.rsplit > Rx) J.rsplit = (int)Rx;
    if (J.rsplit < 1) J.rsplit = 1;
  }
  J.blk0 = blocks;
  blocks += J.nbx + (dwy ? J.ncb * J.rsplit : 0);
  return 0;
}

extern "C" int nr_maxsim2_bwd_w(const float* pmax_x, const float* pmax_y, const float* dH, int64_t dh_sr, int64_t dh_sc,
                                float dh_scale, int64_t Rx, int64_t Nx, int64_t Ry, int64_t Ny, float* dwx, float* dwy,
                                void* stream) {
  BwdWArgs a{};
  int blocks = 0;
  if (int e = bwd_w_fill(a.j[0], pmax_x, pmax_y, dH, dh_sr, dh_sc, dh_scale, Rx, Nx, Ry, Ny, dwx, dwy, blocks)) return e;
  a.njobs = 1;
  maxsim2_bwd_w_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a);
  NR_CHECK_LAUNCH("nr_maxsim2_bwd_w");
  return 0;
}

/* up to 3 pairs in one launch (see nr_maxsim2_bwd_w_job in include/nrhead.h) */
extern "C" int nr_maxsim2_bwd_w_multi(const nr_maxsim2_bwd_w_job* jobs, int njobs, int64_t Nx, int64_t Ny, void* stream) {
  NR_CHECK_ARG(jobs && njobs >= 1 && njobs <= 3, "nr_maxsim2_bwd_w_multi: 1..3 jobs");
  BwdWArgs a{};
  int blocks = 0;
  for (int i = 0; i < njobs; ++i)
    if (int e = bwd_w_fill(a.j[i], jobs[i].pmax_x, jobs[i].pmax_y, jobs[i].dH, jobs[i].dh_sr, jobs[i].dh_sc, jobs[i].dh_scale,
                           jobs[i].Rx, Nx, jobs[i].Ry, Ny, jobs[i].dwx, jobs[i].dwy, blocks))
      return e;
  a.njobs = njobs;
  maxsim2_bwd_w_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(a);
  NR_CHECK_LAUNCH("nr_maxsim2_bwd_w_multi");
  return 0;
}
