// Both directions of the masked max-sim late interaction from ONE accumulator tile (NR_PREC_BF16).
//
//   S[rx, ry] = alpha * ( sum_x wx[rx,x] * max_y R[rx,ry,x,y]  +  sum_y wy[ry,y] * max_x R[rx,ry,x,y] )
//   R[rx,ry,x,y] = < xn[rx,x,:], yn[ry,y,:] >           (masked tokens are zero rows of the bf16 operands, so
//                                                         masked pairs are exactly 0 and take part in the max)
//
// = the whole of NeighborRetr.local_level below the token-weight MLPs (reference
// NeighborRetr/models/modeling.py:495-512: einsum, two mask multiplies, max over v, max over t, two weighted
// sums, average) — the reference materialises the 4-D [A,B,Nt,Nv] tensor and makes >= 7 passes over it.
// The one-direction kernel (maxsim_tc.cu) needs two launches and computes every token pair twice; here a tile's
// fp32 accumulator is read from TMEM once and reduced in BOTH directions:
//   row direction    (max over the Ny columns of a Y sample): one accumulator row per thread, in registers;
//   column direction (max over the Nx rows of an X sample):  rows live in different lanes -> (value, row) keys
//                    (order-preserving integer image of the fp32 value with its log2(GL) lowest mantissa bits
//                    replaced by the row index: 2^-20 RELATIVE resolution at every magnitude) are reduced by a
//                    halving butterfly over aligned groups of GL lanes (GL columns in, one column per lane out),
//                    the 128/GL group partials go to shared memory and are combined per X sample afterwards.
// Up to 4 independent (X, Y) problems with the same token counts share a launch (the batch pair and the two bank
// pairs of a head step), so the persistent grid sees one long tile list.
//
// Persistent warp-specialised kernel, one CTA per SM: warp 0 TMA producer, warp 1 tcgen05.mma issuer (M=128,
// N<=256, K=16 bf16 -> fp32 TMEM, two accumulator stages), warps 2..9 epilogue.
// Roofline: tensor pipe; algorithmic flops per problem 2*Rx*Nx*Ry*Ny*D, every token pair multiplied ONCE.
#include "common.cuh"
#include "nrhead_internal.h"
#include "tc_common.cuh"
#include <stdlib.h>

namespace nr {
using namespace tc;

int make_tmap_bf16(CUtensorMap* m, const void* base, int64_t rows, int64_t d, int box_rows);

// threads: TMA warp, MMA warp, then two epilogue sets of HS x 4 warps (HS = 1: 320 threads, HS = 2: 576 threads)
__host__ __device__ constexpr int t2_threads(int hs) { return 64 + 2 * 128 * hs; }
constexpr int T2_BM = 128;
constexpr int T2_BK = 64;                  // bf16 elements per k-block = one 128-byte swizzle row
constexpr int T2_MAX_STAGES = 6;
constexpr int T2_A_BYTES = T2_BM * 128;    // 16 KB
constexpr int T2_ACC_COLS = 256;           // TMEM columns per accumulator stage
constexpr int T2_MAX_PROB = 4;
constexpr int T2_RING = 4;                 // depth of the tile-index ring between the scheduler and its consumers

struct Tc2Prob {
  const float* wx; const float* wy;
  int Rx, Ry;
  float alpha; int tile0;
  float* out; int64_t out_sr, out_sc; float* out2; int64_t out2_sr, out2_sc;
  float* pmax_x; uint8_t* ystar; float* pmax_y; uint8_t* xstar;
  int n_mt, n_nt;
  // evaluation modes (T2_MODE_DIAG / T2_MODE_COUNT): X row rx is pair gx0 + rx, Y row ry is pair gy0 + ry; the
  // positive of a row is the row of the other side with the same pair id
  float* diag;                             // [pairs]: DIAG writes S of the positives it covers, COUNT reads it
  int* gt_x; int* eq_x; int* gt_y; int* eq_y;
  int gx0, gy0, dj;                        // dj: Y tiles one X box can need for its positives (DIAG tile list)
  int wn, n_me;                            // tile order: bands of wn Y tiles, inside a band X boxes (n_me of them) outermost
};

constexpr int T2_MODE_SIM = 0;             // write S (and what backward needs)
constexpr int T2_MODE_DIAG = 1;            // only the tiles that hold a positive pair: diag[pair] = S[positive]
constexpr int T2_MODE_COUNT = 2;           // every tile: rank counts against diag straight from the accumulator

struct alignas(64) Tc2Args {
  CUtensorMap tmx[T2_MAX_PROB];
  CUtensorMap tmy[T2_MAX_PROB];
  Tc2Prob p[T2_MAX_PROB];
  int nprob, n_tiles;
  int Nx, SX, MU, SY, UN, num_kb, stages, b_bytes, hp_ld, kg_ld;
  int mode, cnt_w;                         // cnt_w: ints per count vector in shared memory (COUNT), else 0
  int l2_hints;                            // 1: Y boxes evict_last, X boxes evict_first (banded order on large problems)
  int elect;                               // barrier releases by one elected thread per epilogue set
  unsigned int* tile_counter;              // zeroed per launch: dynamic tile scheduler
  unsigned long long* trace;               // debug (NR_TC2_TRACE=1): 16 globaltimer stamps per CTA, else nullptr
};

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define T2_TRACE(slot) do { if (a.trace) a.trace[(size_t)blockIdx.x * 16 + (slot)] = gtime(); } while (0)

__host__ __device__ constexpr int t2_gcd(int a, int b) { return b == 0 ? a : t2_gcd(b, a % b); }
__host__ __device__ constexpr int t2_lcm(int a, int b) { return a / t2_gcd(a, b) * b; }

// GL keys of one lane (GL accumulator columns of its row) -> the maximum over the GL lanes of its aligned
// group for ONE column: lane l ends up with column (l & (GL-1)).  3 (GL=8) / 2 (GL=4) halving exchange stages.
template <int GL>
__device__ __forceinline__ uint32_t group_colmax(uint32_t* k, int lane) {
#pragma unroll
  for (int w = GL / 2; w >= 1; w >>= 1) {
    const bool hi = (lane & w) != 0;
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const uint32_t send = hi ? k[i] : k[i + w];
      const uint32_t keep = hi ? k[i + w] : k[i];
      const uint32_t recv = __shfl_xor_sync(0xffffffffu, send, w);
      k[i] = keep > recv ? keep : recv;
    }
  }
  return k[0];
}

// DIAG tile list: entry `local` of a problem = (X box mt, j-th Y tile that holds positives of that box).  Returns
// false for entries with no such tile (the scheduler skips them).
__device__ __forceinline__ bool diag_tile(const Tc2Prob& P, int SX, int SY, int local, int& mt, int& nt) {
  mt = local / P.dj;
  const int j = local - mt * P.dj;
  const int off = P.gx0 - P.gy0;                         // positive of X row rx: Y row rx + off
  const int x0 = mt * SX;
  const int lo = max(x0 + off, 0);
  const int hi = min(min(x0 + SX, P.Rx) - 1 + off, P.Ry - 1);
  nt = lo / SY + j;
  return hi >= lo && nt <= hi / SY;
}

__device__ __forceinline__ uint32_t and_or(uint32_t a, uint32_t m, uint32_t c) {    // (a & m) | c
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xEA;" : "=r"(d) : "r"(a), "r"(m), "r"(c));
  return d;
}

template <int N>
__device__ __forceinline__ void set_barrier(int set) {      // named barrier 1 / 2: the N threads of one epilogue set
  if (set == 0) asm volatile("bar.sync 1, %0;" ::"n"(N) : "memory");
  else asm volatile("bar.sync 2, %0;" ::"n"(N) : "memory");
}

// CL2: CTA pairs (clusters of 2, the two SMs of a TPC) running tcgen05.mma.cta_group::2.  A pair works on two adjacent X
// boxes against ONE Y box: each CTA loads its own X box and HALF of the Y box into its own shared memory, the leader's
// MMA issuer multiplies M = 256 rows (128 per CTA, accumulators in each CTA's own TMEM) by the N columns whose halves
// the two CTAs hold.  Per tile an SM ingests its X box + half a Y box instead of X + Y: the L2 -> SM operand stream
// (ncu: 584 MB per MSR-VTT step launch with independent CTAs) drops by a third, and a stage is 31 KB instead of 46, so
// the ring gets deeper.  The leader CTA claims pair-tiles and publishes them into both CTAs' rings; every operand
// barrier that the MMA issuer waits on lives in the leader, the stage-free and accumulator-ready barriers are
// signalled in both CTAs by a multicast commit.  Opt-in: see the measurement quoted in maxsim2_fwd_impl.
// HS = 2: each epilogue set has 8 warps; the two warps that share a TMEM lane quarter split the accumulator
// columns of the tile (chunks [0, n/2) and [n/2, n)), which doubles the warps available to hide the latency of the
// shuffle / shared-memory / global-store chains of the reduction (ncu: 0.6 eligible warps per scheduler with HS = 1).
// XK = true: exact-order column keys (order-preserving integer image of the fp32 value, 2^-20 RELATIVE resolution at
// every magnitude) for the split-bf16 mode, whose products are good to ~3e-7; XK = false: keys are the fp32 bits of
// v + 2 (one FADD + one LOP3 per element, 2e-6 absolute resolution — far below the 3e-4 noise of bf16 operands).  The
// epilogue is issue-bound: the third instruction per element costs 8-10 % of the kernel (measured).
template <int NY, int GL, bool CL2, int HS, bool XK>
__global__ void __launch_bounds__(t2_threads(HS), 1) maxsim2_fwd_tc_kernel(const __grid_constant__ Tc2Args a) {
  constexpr int T2_SET = 128 * HS;         // threads of one epilogue set
  constexpr int CH = t2_lcm(NY, GL);       // accumulator columns per epilogue chunk: whole samples, whole groups
  constexpr int SPC = CH / NY;             // Y samples per chunk
  constexpr uint32_t LOWM = GL - 1;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages x (A | B)] then per epilogue set: [hp 128 x hp_ld f32] [keyG (128/GL) x kg_ld u32]
  // [colw SX x UN f32] [wy UN f32], then the barriers
  // (offset arithmetic on the extern array, not an integer round trip: the compiler keeps the shared address space and
  // emits 32-bit LDS / STS instead of generic accesses with 64-bit address arithmetic)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int stage_bytes = T2_A_BYTES + a.b_bytes;
  float* hp = reinterpret_cast<float*>(smem + (size_t)a.stages * stage_bytes);
  uint32_t* keyG = reinterpret_cast<uint32_t*>(hp + 2 * T2_BM * a.hp_ld);
  float* colw = reinterpret_cast<float*>(keyG + 2 * (T2_BM / GL) * a.kg_ld);
  float* wyst = colw + 2 * a.SX * a.UN;
  int* cnts = reinterpret_cast<int*>(wyst + 2 * a.UN);    // COUNT: per set {gt_x, eq_x, gt_y, eq_y} x cnt_w
  uint64_t* bars = reinterpret_cast<uint64_t*>(cnts + 2 * 4 * a.cnt_w);
  uint64_t* full = bars;                          // [stages]  TMA -> MMA
  uint64_t* empty = bars + T2_MAX_STAGES;         // [stages]  MMA -> TMA
  uint64_t* tfull = bars + 2 * T2_MAX_STAGES;     // [2]       MMA -> epilogue
  uint64_t* tempty = tfull + 2;                   // [2]       epilogue -> MMA
  uint64_t* rfull = tempty + 2;                   // [4]       tile scheduler (TMA warp) -> MMA + epilogue
  uint64_t* rempty = rfull + T2_RING;             // [4]       MMA + epilogue -> tile scheduler
  int* ring = reinterpret_cast<int*>(rempty + T2_RING);   // [4] tile index, -1 = no more tiles
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ring + T2_RING);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cr = CL2 ? cluster_ctarank() : 0u;       // rank in the CTA pair; 0 = leader

  if (warp == 0 && lane == 0) {
    for (int p = 0; p < a.nprob; ++p) { tma_prefetch_desc(&a.tmx[p]); tma_prefetch_desc(&a.tmy[p]); }
    for (int s = 0; s < a.stages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
    // CL2: the leader's issuer overwrites the accumulator stage in BOTH CTAs: both epilogue sets release it there
    // a.elect: ONE elected thread per epilogue set releases the accumulator stage and the ring slot (behind the set
    // barrier) instead of every thread of the set
    const int per_set = a.elect ? 1 : T2_SET;
    for (int s = 0; s < 2; ++s) { mbar_init(tfull + s, 1); mbar_init(tempty + s, CL2 ? 2 * per_set : per_set); }
    // ring: filled by the (leader's) scheduler; released by the MMA issuer and one epilogue set of every CTA of the
    // pair, plus the non-leader's TMA producer
    for (int s = 0; s < T2_RING; ++s) { mbar_init(rfull + s, 1); mbar_init(rempty + s, CL2 ? 2 * (1 + per_set) + 1 : 1 + per_set); }
    fence_barrier_init();
  }
  if (warp == 1) {
    if (CL2) { tmem_alloc2(tmem_slot, 512); tmem_relinquish2(); }
    else { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  }
  tc_fence_before();
  __syncthreads();
  if (CL2) cluster_sync_all();                            // the peer's barriers are initialised before any remote use
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) T2_TRACE(0);                      // set-up done
  // consumers of a ring slot: read the tile, release the slot on the LEADER's barrier
  auto ring_release = [&](int it) {
    const int slot = it & (T2_RING - 1);
    if (CL2) mbar_arrive_remote(rempty + slot, 0);
    else mbar_arrive(rempty + slot);
  };
  // release = false: the caller releases the slot later through ONE elected thread (the epilogue sets: 256 arrivals
  // per tile on one shared-memory word — cluster-scope ones for the peer of a pair — become one)
  auto ring_take = [&](int it, bool release = true) {
    const int slot = it & (T2_RING - 1);
    if (CL2) mbar_wait_cluster(rfull + slot, (uint32_t)(it / T2_RING) & 1u);
    else mbar_wait(rfull + slot, (uint32_t)(it / T2_RING) & 1u);
    const int tile = ring[slot];
    if (release) ring_release(it);
    return tile;
  };

  // tile index -> (problem, m-tile, n-tile); n fastest so that concurrently running CTAs share the X box in L2
  auto decode = [&](int tile, int& p, int& mt, int& nt) -> bool {
    p = 0;
    while (p + 1 < a.nprob && tile >= a.p[p + 1].tile0) ++p;
    const int local = tile - a.p[p].tile0;
    if (a.mode == T2_MODE_DIAG) return diag_tile(a.p[p], a.SX, a.SY, local, mt, nt);
    // Bands of wn Y tiles; inside a band the Y tile runs fastest, so the CTAs of a wave share a few X boxes and the
    // band's Y boxes stay in L2 while every X box passes by once.  (One band = the plain row-major order, which at
    // 8192 x 8192 re-read the Y operand from HBM for every X box: ncu 125 GB of DRAM reads for 0.3 GB of operands.)
    const Tc2Prob& Q = a.p[p];
    const int band_tiles = Q.n_me * Q.wn;
    const int bnd = local / band_tiles;
    const int rem = local - bnd * band_tiles;
    const int w = min(Q.wn, Q.n_nt - bnd * Q.wn);
    mt = rem / w;
    nt = bnd * Q.wn + rem - mt * w;
    if (CL2) mt = 2 * mt + (int)cr;                      // a pair-tile = two adjacent X boxes against one Y box
    return true;
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      // CL2: the leader's barrier collects the bytes of both CTAs (X box + half a Y box each)
      const int hb = a.SY * NY / 2;
      const uint32_t tx_bytes = CL2 ? 2u * (uint32_t)(a.MU + hb) * 128u : (uint32_t)(a.MU + a.SY * NY) * 128u;
      const uint64_t pol_x = a.l2_hints == 2 ? l2_policy_evict_last() : l2_policy_evict_first();
      const uint64_t pol_y = a.l2_hints == 2 ? l2_policy_evict_first() : l2_policy_evict_last();
      // Dynamic tile scheduler: tiles are claimed from a global counter, so a CTA whose SM was still busy with
      // another kernel of the step graph (the persistent grid is one CTA per SM) simply takes fewer tiles.
      for (int n = 0;; ++n) {
        const int slot = n & (T2_RING - 1);
        int tile;
        if (!CL2 || cr == 0) {
          if (CL2) mbar_wait_cluster(rempty + slot, ((uint32_t)(n / T2_RING) & 1u) ^ 1u);
          else mbar_wait(rempty + slot, ((uint32_t)(n / T2_RING) & 1u) ^ 1u);
          tile = (int)atomicAdd(a.tile_counter, 1u);
          if (a.mode == T2_MODE_DIAG) {                    // entries of the list without a positive are skipped HERE:
            int p_, mt_, nt_;                              // the MMA issuer and the epilogues never see them
            while (tile < a.n_tiles && !decode(tile, p_, mt_, nt_)) tile = (int)atomicAdd(a.tile_counter, 1u);
          }
          if (tile >= a.n_tiles) tile = -1;
          if (n == 0) T2_TRACE(1);                        // first claim
          if (tile < 0) { T2_TRACE(2); if (a.trace) a.trace[(size_t)blockIdx.x * 16 + 3] = (unsigned long long)n; }
          ring[slot] = tile;
          if (CL2) {
            st_remote_u32(ring + slot, 1, (uint32_t)tile);
            mbar_arrive_remote(rfull + slot, 1);
            mbar_arrive_remote(rfull + slot, 0);
          } else {
            mbar_arrive(rfull + slot);                     // release: the ring entry is visible to the waiters
          }
          if (tile < 0) {                                  // the other epilogue set reads the NEXT slot: end it too
            const int slot2 = (n + 1) & (T2_RING - 1);
            if (CL2) mbar_wait_cluster(rempty + slot2, ((uint32_t)((n + 1) / T2_RING) & 1u) ^ 1u);
            else mbar_wait(rempty + slot2, ((uint32_t)((n + 1) / T2_RING) & 1u) ^ 1u);
            ring[slot2] = -1;
            if (CL2) {
              st_remote_u32(ring + slot2, 1, (uint32_t)-1);
              mbar_arrive_remote(rfull + slot2, 1);
              mbar_arrive_remote(rfull + slot2, 0);
            } else {
              mbar_arrive(rfull + slot2);
            }
            // the workspace resets itself: the last claimer to finish (every claimer has made its final claim by
            // then) zeroes both counters, so no memset has to precede the next launch
            const unsigned int claimers = CL2 ? gridDim.x / 2 : gridDim.x;
            if (atomicAdd(a.tile_counter + 1, 1u) == claimers - 1) {
              atomicExch(a.tile_counter, 0u);
              atomicExch(a.tile_counter + 1, 0u);
              __threadfence();
            }
            break;
          }
        } else {                                           // non-leader producer: a consumer of the leader's ring
          tile = ring_take(n);
          if (tile < 0) { ring_take(n + 1); break; }       // the second end marker frees its slot like the others
        }
        int p, mt, nt;
        decode(tile, p, mt, nt);
        const int row_x = mt * a.MU, row_y = nt * a.SY * NY;
        for (int kb = 0; kb < a.num_kb; ++kb) {
          mbar_wait(empty + stage, phase ^ 1);             // CL2: the leader's commit frees the stage in both CTAs
          uint8_t* sa = smem + (size_t)stage * stage_bytes;
          if (CL2) {                                       // my X box and my half of the Y box, into MY shared memory
            if (cr == 0) mbar_expect_tx(full + stage, tx_bytes);
            tma_load_2d_cg2(sa, &a.tmx[p], full + stage, kb * T2_BK, row_x);
            tma_load_2d_cg2(sa + T2_A_BYTES, &a.tmy[p], full + stage, kb * T2_BK, row_y + (int)cr * hb);
          } else if (a.l2_hints) {
            mbar_expect_tx(full + stage, tx_bytes);
            tma_load_2d_hint(sa, &a.tmx[p], full + stage, kb * T2_BK, row_x, pol_x);
            tma_load_2d_hint(sa + T2_A_BYTES, &a.tmy[p], full + stage, kb * T2_BK, row_y, pol_y);
          } else {
            mbar_expect_tx(full + stage, tx_bytes);
            tma_load_2d(sa, &a.tmx[p], full + stage, kb * T2_BK, row_x);
            tma_load_2d(sa + T2_A_BYTES, &a.tmy[p], full + stage, kb * T2_BK, row_y);
          }
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16(CL2 ? 2 * T2_BM : T2_BM, a.UN);
      int stage = 0; uint32_t phase = 0;
      for (int it = 0;; ++it) {
        const int tile = ring_take(it);
        if (tile < 0) { if (CL2) ring_take(it + 1); break; }
        if (CL2 && cr != 0) continue;                    // the leader issues for the pair; the peer only frees ring slots
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        if (CL2) mbar_wait_cluster(tempty + acc, acc_phase ^ 1);   // both CTAs' epilogues drained this accumulator
        else mbar_wait(tempty + acc, acc_phase ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(acc * T2_ACC_COLS);
        for (int kb = 0; kb < a.num_kb; ++kb) {
          mbar_wait(full + stage, phase);
          tc_fence_after();
          if (it == 0 && kb == 0) T2_TRACE(4);             // first operands landed
          const uint32_t sa = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t adesc = umma_desc_kmajor_sw128(sa);
          const uint64_t bdesc = umma_desc_kmajor_sw128(sa + T2_A_BYTES);
#pragma unroll
          for (int k = 0; k < T2_BK / 16; ++k) {         // advance 32 B (16 bf16) inside the swizzle row
            if (CL2) umma_bf16_cg2(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
            else umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kb | k) != 0);
          }
          if (CL2) umma_commit_cg2(empty + stage, (uint16_t)3);  // the stage is free in both CTAs
          else umma_commit(empty + stage);               // smem slot free when these MMAs retire
          if (++stage == a.stages) { stage = 0; phase ^= 1; }
        }
        if (CL2) umma_commit_cg2(tfull + acc, (uint16_t)3);      // accumulator ready for the epilogues of both CTAs
        else umma_commit(tfull + acc);                   // accumulator ready for the epilogue
        if (it == 0) T2_TRACE(5);                          // first tile issued
        T2_TRACE(6);                                       // last tile issued (overwritten per tile)
      }
    }
  } else {
    // ===================== epilogue: two ping-pong sets (warps 2..5 = set 0, 6..9 = set 1) =====================
    // Set s owns TMEM accumulator stage s, i.e. every other tile of this CTA, with its own staging buffers: while
    // one set is in the latency-bound tail of a tile (group combine, output), the other drains the next accumulator.
    const int set = (warp - 2) / (4 * HS);
    const int half = ((warp - 2) >> 2) & (HS - 1);       // which part of the tile's columns this warp reduces
    const int q = warp & 3;                              // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;                         // accumulator row = X token of the tile
    const int et = (int)threadIdx.x - 64 - set * T2_SET; // 0..T2_SET-1 within the set
    const int Nx = a.Nx;
    const int sx = r / Nx, x = r - sx * Nx;
    const uint32_t low = LOWM - ((uint32_t)lane & LOWM);
    const uint32_t himask = ~LOWM;
    // and_or(): (u & himask) | low as ONE explicit LOP3 per accumulator element (left to the compiler it becomes an
    // AND with the immediate plus a second LOP3 that recomputes `low` from the lane index)
    float* const hpb = hp + (size_t)set * T2_BM * a.hp_ld;
    uint32_t* const kgs = keyG + (size_t)set * (T2_BM / GL) * a.kg_ld;
    float* const cws = colw + (size_t)set * a.SX * a.UN;
    float* const wys = wyst + (size_t)set * a.UN;
    uint32_t* const kg_row = kgs + (r / GL) * a.kg_ld + (lane & (int)LOWM);
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(set * T2_ACC_COLS);
    int* const cset = cnts + (size_t)set * 4 * a.cnt_w;
    // COUNT: zeroed once; afterwards the thread that flushes an entry re-zeroes it (two set barriers before its next use)
    for (int i = et; i < 4 * a.cnt_w; i += T2_SET) cset[i] = 0;
    for (int it = set;; it += 2) {
      const int tile = ring_take(it, !a.elect);
      if (tile < 0) {                                      // end marker: nothing overwrites the ring any more
        if (a.elect && et == 0) ring_release(it);
        break;
      }
      int pi, mt, nt;
      decode(tile, pi, mt, nt);
      const Tc2Prob& P = a.p[pi];
      const uint32_t acc_phase = (it >> 1) & 1;
      const int ry0 = nt * a.SY;
      const int sy_n = min(a.SY, P.Ry - ry0);
      const int sx_n = max(0, min(a.SX, P.Rx - mt * a.SX));    // 0: the odd X box of the last pair-tile
      const int n_ch = (sy_n + SPC - 1) / SPC;
      const int ncols = sy_n * NY;
      const int rx = mt * a.SX + sx;
      const bool row_ok = (r < a.MU) && (rx < P.Rx);
      const float wxv = row_ok ? P.wx[(int64_t)rx * Nx + x] : 0.f;
      // per-tile base pointers: inside the tile every offset is a small 32-bit product
      const int64_t obase = ((int64_t)rx * P.Ry + ry0) * Nx + x;
      float* const pmx = (row_ok && P.pmax_x) ? P.pmax_x + obase : nullptr;
      uint8_t* const yst = (row_ok && P.ystar) ? P.ystar + obase : nullptr;
      float* const hp_row = hpb + r * a.hp_ld;
      // the Y token weights of this tile, staged while the MMAs of the tile are still running (the set's previous
      // tile finished with a set barrier after its last read of wys)
      for (int c = et; c < ncols; c += T2_SET) wys[c] = P.wy[(int64_t)ry0 * NY + c];
      mbar_wait(tfull + set, acc_phase);
      tc_fence_after();
      if (et == 0 && it == set) T2_TRACE(8 + 2 * set);     // first accumulator of this set ready
      auto process = [&](const uint32_t* v, int ch) {
        // row direction: max / arg-max over the NY columns of each Y sample of the chunk
#pragma unroll
        for (int s2 = 0; s2 < SPC; ++s2) {
          const int sy = ch * SPC + s2;
          float f[NY];
#pragma unroll
          for (int y = 0; y < NY; ++y) f[y] = __uint_as_float(v[s2 * NY + y]);
          const float best = TreeRed<NY>::fmax_(f);
          if (sy < sy_n) {
            hp_row[sy] = wxv * best;
            const int o = sy * Nx;
            if (pmx) pmx[o] = best;
            // the arg-max (a third of the row direction's instructions) only where backward will read it: evaluation
            // and the rank passes save nothing
            if (yst) yst[o] = (uint8_t)TreeRed<NY>::first_eq(f, best, 0);
          }
        }
        // column direction: group partial of (value, row) keys, one column per lane
#pragma unroll
        for (int g = 0; g < CH / GL; ++g) {
          uint32_t k[GL];
#pragma unroll
          for (int j = 0; j < GL; ++j) {
            // order-preserving integer image of the fp32 value (all exponents keep their full relative precision:
            // only the log2(GL) lowest mantissa bits give way to the row index), ties -> lower row
            const uint32_t u = v[g * GL + j];
            if constexpr (XK) k[j] = and_or(u ^ ((uint32_t)((int32_t)u >> 31) | 0x80000000u), himask, low);
            else k[j] = and_or(__float_as_uint(__uint_as_float(u) + 2.0f), himask, low);
          }
          kg_row[ch * CH + g * GL] = group_colmax<GL>(k, lane);
        }
      };
      const int ch_per = (n_ch + HS - 1) / HS;
      const int ch_lo = half * ch_per, ch_hi = min(n_ch, ch_lo + ch_per);
      if constexpr (CH <= 32 && HS == 1) {
        // two register buffers: the TMEM load of the next chunk is in flight while this one is reduced
        uint32_t va[CH], vb[CH];
        int ch = ch_lo;
        if (ch < ch_hi) tmem_ld_cols<CH>(taddr + (uint32_t)(ch * CH), va);
        for (; ch < ch_hi; ch += 2) {
          tmem_ld_wait();
          reg_fence<CH>(va);
          const int ch2 = ch + 1;
          if (ch2 < ch_hi) tmem_ld_cols<CH>(taddr + (uint32_t)(ch2 * CH), vb);
          process(va, ch);
          if (ch2 < ch_hi) {
            tmem_ld_wait();
            reg_fence<CH>(vb);
            if (ch + 2 < ch_hi) tmem_ld_cols<CH>(taddr + (uint32_t)((ch + 2) * CH), va);
            process(vb, ch2);
          }
        }
      } else {
        for (int ch = ch_lo; ch < ch_hi; ++ch) {
          uint32_t v[CH];
          tmem_ld_cols<CH>(taddr + (uint32_t)(ch * CH), v);
          tmem_ld_wait();
          reg_fence<CH>(v);
          process(v, ch);
        }
      }
      tc_fence_before();
      if (!a.elect) {
        if (CL2) mbar_arrive_remote(tempty + set, 0);    // TMEM stage may be overwritten (the leader issues for both)
        else mbar_arrive(tempty + set);
      }
      set_barrier<T2_SET>(set);                          // every thread of the set has drained its part of the stage
      if (a.elect && et == 0) {
        if (CL2) mbar_arrive_remote(tempty + set, 0);
        else mbar_arrive(tempty + set);
        ring_release(it);                                // and every thread of the set has read its ring entry
      }
      // F1: combine the Nx/GL group partials of each (X sample, column): value, arg-max row, weighted value.
      // Element e = s * ncols + c, e = et, et + T2_SET, ...: (s, c) advance without a division.
      {
        const int ng = Nx / GL;
        const int64_t ybase = ((int64_t)(mt * a.SX) * P.Ry + ry0) * NY;
        float* const pmy = P.pmax_y ? P.pmax_y + ybase : nullptr;
        uint8_t* const xst = P.xstar ? P.xstar + ybase : nullptr;
        const int srow = P.Ry * NY;                       // output stride between the X samples of the tile
        int s = 0, c = et;
        while (c >= ncols) { c -= ncols; ++s; }
        for (; s < sx_n;) {
          const uint32_t* kp = kgs + (s * ng) * a.kg_ld + c;
          uint32_t best = kp[0]; int bg = 0;
          for (int gi = 1; gi < ng; ++gi) {
            const uint32_t kk = kp[gi * a.kg_ld];
            if ((kk & ~LOWM) > (best & ~LOWM)) { best = kk; bg = gi; }   // equal values: the lower group stays
          }
          const uint32_t tb = best & ~LOWM;
          const float val = XK ? __uint_as_float((tb & 0x80000000u) ? (tb ^ 0x80000000u) : ~tb) : __uint_as_float(tb) - 2.0f;
          const int xs = bg * GL + (int)(LOWM - (best & LOWM));
          const int o = s * srow + c;
          if (pmy) pmy[o] = val;
          if (xst) xst[o] = (uint8_t)xs;
          cws[s * a.UN + c] = wys[c] * val;
          c += T2_SET;
          while (c >= ncols) { c -= ncols; ++s; }
        }
      }
      set_barrier<T2_SET>(set);
      // F2: S[rx, ry] = alpha * (sum over the Nx rows of hp + sum over the NY columns of colw)
      for (int e = et, s = 0, sy = et; e < sx_n * sy_n; e += T2_SET, sy += T2_SET) {
        while (sy >= sy_n) { sy -= sy_n; ++s; }           // e = s * sy_n + sy without a division
        const float* hrow = hpb + (s * Nx) * a.hp_ld + sy;
        float h0 = 0.f, h1 = 0.f, h2 = 0.f, h3 = 0.f;     // independent chains: the loads pipeline
        int xx = 0;
        for (; xx + 4 <= Nx; xx += 4) {
          h0 += hrow[(xx + 0) * a.hp_ld]; h1 += hrow[(xx + 1) * a.hp_ld];
          h2 += hrow[(xx + 2) * a.hp_ld]; h3 += hrow[(xx + 3) * a.hp_ld];
        }
        for (; xx < Nx; ++xx) h0 += hrow[xx * a.hp_ld];
        const float* cw = cws + s * a.UN + sy * NY;
#pragma unroll
        for (int y = 0; y < NY; y += 4) { h0 += cw[y]; h1 += cw[y + 1]; h2 += cw[y + 2]; h3 += cw[y + 3]; }
        const float h = ((h0 + h1) + (h2 + h3)) * P.alpha;
        const int rxx = mt * a.SX + s, ry = ry0 + sy;
        if (a.mode == T2_MODE_SIM) {
          P.out[(int64_t)rxx * P.out_sr + (int64_t)ry * P.out_sc] = h;
          if (P.out2) P.out2[(int64_t)rxx * P.out2_sr + (int64_t)ry * P.out2_sc] = h;
        } else {
          // evaluation without S (reference utils/metrics.py:58-66 locates the positive in the sorted row: the same
          // rank follows from #{> positive} and #{== positive})
          const int gx = P.gx0 + rxx, gy = P.gy0 + ry;
          if (a.mode == T2_MODE_DIAG) {
            if (gx == gy) P.diag[gx] = h;
          } else {
            const float dx = P.diag[gx], dy = P.diag[gy];
            const bool self = gx == gy;                     // the positive itself: equal by definition
            if (self || h == dx) atomicAdd(cset + a.cnt_w + s, 1);
            else if (h > dx) atomicAdd(cset + s, 1);
            if (self || h == dy) atomicAdd(cset + 3 * a.cnt_w + sy, 1);
            else if (h > dy) atomicAdd(cset + 2 * a.cnt_w + sy, 1);
          }
        }
      }
      set_barrier<T2_SET>(set);       // the set's staging buffers (hp, keyG, colw, wys) are free for its next tile
      if (a.mode == T2_MODE_COUNT) {  // one global atomic per (row of the tile, non-zero count)
        auto flush = [&](int* c, int* g) { const int v = *c; if (v) { atomicAdd(g, v); *c = 0; } };
        if (et < sx_n) {
          flush(cset + et, P.gt_x + mt * a.SX + et);
          flush(cset + a.cnt_w + et, P.eq_x + mt * a.SX + et);
        }
        if (et < sy_n) {
          flush(cset + 2 * a.cnt_w + et, P.gt_y + ry0 + et);
          flush(cset + 3 * a.cnt_w + et, P.eq_y + ry0 + et);
        }
      }
      if (et == 0) T2_TRACE(9 + 2 * set);                  // end of this set's latest tile (overwritten per tile)
    }
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) T2_TRACE(12);                      // CTA done
  if (CL2) cluster_sync_all();          // no CTA leaves while its peer may still multicast into it or signal its barriers
  if (warp == 1) {
    __syncwarp();
    if (CL2) tmem_dealloc2(tmem_base, 512);
    else tmem_dealloc(tmem_base, 512);
  }
}

template <int NY, int GL, bool CL2, int HS, bool XK>
static int launch_variant(const Tc2Args& a, size_t smem, int grid, cudaStream_t stream) {
  auto kern = maxsim2_fwd_tc_kernel<NY, GL, CL2, HS, XK>;
  NR_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (CL2) {                   // CTA pairs: thread-block clusters of 2 (the two SMs of a TPC)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(t2_threads(HS));
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    NR_CUDA(cudaLaunchKernelEx(&cfg, kern, a));
    return 0;
  }
  kern<<<grid, t2_threads(HS), smem, stream>>>(a);
  NR_CHECK_LAUNCH("nr_maxsim2_fwd");
  return 0;
}

// exact_keys: split-bf16 operands (exact-order column keys); pair: cta_group::2; halves == 1: 8 epilogue warps (debug)
template <int NY, int GL>
static int launch2(const Tc2Args& a, size_t smem, int grid, bool pair, int halves, bool exact_keys, cudaStream_t stream) {
  if (pair) {
    if (exact_keys) return launch_variant<NY, GL, true, 2, true>(a, smem, grid, stream);
    return launch_variant<NY, GL, true, 2, false>(a, smem, grid, stream);
  }
  if (exact_keys) return launch_variant<NY, GL, false, 2, true>(a, smem, grid, stream);
  if (halves == 2) return launch_variant<NY, GL, false, 2, false>(a, smem, grid, stream);
  return launch_variant<NY, GL, false, 1, false>(a, smem, grid, stream);
}

template <int GL>
static int dispatch_ny(int Ny, const Tc2Args& a, size_t smem, int grid, bool pair, int halves, bool exact_keys,
                       cudaStream_t stream) {
  switch (Ny) {
    case 4: return launch2<4, GL>(a, smem, grid, pair, halves, exact_keys, stream);
    case 8: return launch2<8, GL>(a, smem, grid, pair, halves, exact_keys, stream);
    case 12: return launch2<12, GL>(a, smem, grid, pair, halves, exact_keys, stream);
    case 16: return launch2<16, GL>(a, smem, grid, pair, halves, exact_keys, stream);
    case 24: return launch2<24, GL>(a, smem, grid, pair, halves, exact_keys, stream);
    case 32: return launch2<32, GL>(a, smem, grid, pair, halves, exact_keys, stream);
    case 48: return launch2<48, GL>(a, smem, grid, pair, halves, exact_keys, stream);
    case 64: return launch2<64, GL>(a, smem, grid, pair, halves, exact_keys, stream);
    default:
      nr::set_error("nr_maxsim2_fwd: Ny=%d has no tensor-core instantiation (4,8,12,16,24,32,48,64)", Ny);
      return -3;
  }
}

}  // namespace nr

using namespace nr;

extern "C" int nr_maxsim2_supported(int64_t Nx, int64_t Ny, int64_t d) {
  const bool ny_ok = Ny == 4 || Ny == 8 || Ny == 12 || Ny == 16 || Ny == 24 || Ny == 32 || Ny == 48 || Ny == 64;
  return (ny_ok && Nx % 4 == 0 && Nx >= 4 && Nx <= 128 && d % T2_BK == 0 && d > 0) ? 1 : 0;
}

// evaluation modes: per-problem extension of nr_maxsim2_problem (nullptr = T2_MODE_SIM)
struct RankExt { float* diag; int* gt_x; int* eq_x; int* gt_y; int* eq_y; int64_t gx0, gy0; };

static int maxsim2_fwd_impl(const nr_maxsim2_problem* probs, int nprob, int64_t Nx, int64_t Ny, int64_t d,
                           void* workspace, int flags, void* stream, int mode = T2_MODE_SIM,
                           const RankExt* ext = nullptr);

extern "C" int nr_maxsim2_fwd(const nr_maxsim2_problem* probs, int nprob, int64_t Nx, int64_t Ny, int64_t d,
                              void* workspace, void* stream) {
  return maxsim2_fwd_impl(probs, nprob, Nx, Ny, d, workspace, 0, stream);
}

/* flags & 1: the operands are split-bf16 (nr_prep_tokens_split): exact-order keys in the column direction */
extern "C" int nr_maxsim2_fwd_ex(const nr_maxsim2_problem* probs, int nprob, int64_t Nx, int64_t Ny, int64_t d,
                                 void* workspace, int flags, void* stream) {
  return maxsim2_fwd_impl(probs, nprob, Nx, Ny, d, workspace, flags, stream);
}

/* Evaluation ranks without the similarity matrix (reference utils/metrics.py:38-79 on the matrix of
 * training/evaluator.py:21-63).  mode 1: diag[pair] = S[positive pair] for the positives inside the block (only the
 * tiles that hold one are computed); mode 2: the whole block is contracted and every S value is compared with the
 * positives' scores in the epilogue — S is never written. */
extern "C" int nr_maxsim2_rank(const nr_maxsim2_rank_problem* q, int mode, int64_t Nx, int64_t Ny, int64_t d,
                               void* workspace, int flags, void* stream) {
  NR_CHECK_ARG(q && (mode == T2_MODE_DIAG || mode == T2_MODE_COUNT), "nr_maxsim2_rank: mode 1 (diag) or 2 (count)");
  NR_CHECK_ARG(q->diag && q->gx0 >= 0 && q->gy0 >= 0, "nr_maxsim2_rank: diag and non-negative pair offsets required");
  NR_CHECK_ARG(mode == T2_MODE_DIAG || (q->gt_x && q->eq_x && q->gt_y && q->eq_y),
               "nr_maxsim2_rank: count mode needs the four count vectors");
  NR_CHECK_ARG(q->gx0 + q->Rx < (1ll << 31) && q->gy0 + q->Ry < (1ll << 31), "nr_maxsim2_rank: pair ids must fit int32");
  nr_maxsim2_problem p{};
  p.x_bf16 = q->x_bf16; p.y_bf16 = q->y_bf16; p.wx = q->wx; p.wy = q->wy;
  p.Rx = q->Rx; p.Ry = q->Ry; p.alpha = q->alpha;
  RankExt e{q->diag, q->gt_x, q->eq_x, q->gt_y, q->eq_y, q->gx0, q->gy0};
  return maxsim2_fwd_impl(&p, 1, Nx, Ny, d, workspace, flags, stream, mode, &e);
}

static int maxsim2_fwd_impl(const nr_maxsim2_problem* probs, int nprob, int64_t Nx, int64_t Ny, int64_t d,
                           void* workspace, int flags, void* stream, int mode, const RankExt* ext) {
  NR_CHECK_ARG(workspace && ((uintptr_t)workspace & 3) == 0, "nr_maxsim2_fwd: workspace (>= 16 bytes, 4-byte aligned) required");
  NR_CHECK_ARG(probs && nprob >= 1 && nprob <= T2_MAX_PROB, "nr_maxsim2_fwd: 1..%d problems per launch (got %d)",
               T2_MAX_PROB, nprob);
  NR_CHECK_ARG(nr_maxsim2_supported(Nx, Ny, d),
               "nr_maxsim2_fwd: unsupported shape Nx=%lld (multiple of 4, <=128) Ny=%lld (4,8,12,16,24,32,48,64) "
               "d=%lld (multiple of %d)", (long long)Nx, (long long)Ny, (long long)d, T2_BK);
  const int GL = (Nx % 8 == 0) ? 8 : 4;
  const int CH = t2_lcm((int)Ny, GL), SPC = CH / (int)Ny;
  int64_t max_rx = 0, max_ry = 0;
  for (int i = 0; i < nprob; ++i) {
    const nr_maxsim2_problem& q = probs[i];
    NR_CHECK_ARG(q.x_bf16 && q.y_bf16 && q.wx && q.wy && (q.out || mode != T2_MODE_SIM) && q.Rx > 0 && q.Ry > 0,
                 "nr_maxsim2_fwd: problem %d has a null pointer or an empty side", i);
    NR_CHECK_ARG(((uintptr_t)q.x_bf16 & 15) == 0 && ((uintptr_t)q.y_bf16 & 15) == 0,
                 "nr_maxsim2_fwd: operands must be 16-byte aligned");
    if (q.Rx > max_rx) max_rx = q.Rx;
    if (q.Ry > max_ry) max_ry = q.Ry;
  }
  Tc2Args a{};
  a.nprob = nprob;
  a.mode = mode;
  a.Nx = (int)Nx;
  a.SX = T2_BM / (int)Nx;
  if (a.SX > max_rx) a.SX = (int)max_rx;
  a.MU = a.SX * (int)Nx;
  a.SY = (256 / (int)Ny) / SPC * SPC;
  if (a.SY > max_ry) a.SY = (int)((max_ry + SPC - 1) / SPC * SPC);
  a.UN = (a.SY * (int)Ny + 15) / 16 * 16;
  a.num_kb = (int)(d / T2_BK);
  a.hp_ld = a.SY | 1;
  a.kg_ld = a.UN;
  if (GL == 8) { while (a.kg_ld % 32 != 8 && a.kg_ld % 32 != 24) a.kg_ld += 8; }
  else { while (a.kg_ld % 8 != 4) a.kg_ld += 4; }
  int dev = 0, sms = 0;
  NR_CUDA(cudaGetDevice(&dev));
  NR_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int64_t est_tiles = 0;
  for (int i = 0; i < nprob; ++i)
    est_tiles += ((probs[i].Rx + a.SX - 1) / a.SX) * ((probs[i].Ry + a.SY - 1) / a.SY);
  // CTA pairs (cta_group::2): possible when the Y box splits into two swizzle-aligned halves.  Opt-in — NR_TC2_PAIR=1:
  // when there are more than two waves of tiles; =2: whenever the shape allows (tests/test_gpu_tc2.py runs the parity
  // cases this way).  Measured on B200 (profiles/r2_pair_vs_single.txt): parity-exact, a third less L2 -> SM traffic,
  // but 6-9 % SLOWER than independent CTAs (MSR-VTT step tiles 4.14 vs 3.72 us, b = 1024 593 vs 557 us, ActivityNet
  // tiles 469 vs 430 us): with two 256-column accumulator stages per SM the leader's issuer waits for the epilogues of
  // BOTH CTAs (two cluster-scope barrier hops per tile), and the epilogue, not the operand stream, paces this kernel.
  const bool pair_ok = (a.SY * (int)Ny) % 16 == 0 && a.UN == a.SY * (int)Ny;
  bool pair = false;
  if (mode != T2_MODE_SIM) {
    // the evaluation modes index tiles themselves (DIAG) and count per X box (COUNT): independent CTAs only
  } else if (const char* ev = getenv("NR_TC2_PAIR")) {
    const int v = atoi(ev);
    pair = v >= 2 ? pair_ok : (v == 1 && pair_ok && est_tiles >= 2 * sms);
  }
  a.b_bytes = ((pair ? a.UN / 2 : a.UN) * 128 + 1023) / 1024 * 1024;
  int tiles = 0;
  for (int i = 0; i < nprob; ++i) {
    const nr_maxsim2_problem& q = probs[i];
    Tc2Prob& P = a.p[i];
    P.wx = q.wx; P.wy = q.wy; P.Rx = (int)q.Rx; P.Ry = (int)q.Ry; P.alpha = q.alpha;
    P.out = q.out; P.out_sr = q.out_sr; P.out_sc = q.out_sc;
    P.out2 = q.out2; P.out2_sr = q.out2_sr; P.out2_sc = q.out2_sc;
    P.pmax_x = q.pmax_x; P.ystar = q.ystar; P.pmax_y = q.pmax_y; P.xstar = q.xstar;
    P.n_mt = (int)((q.Rx + a.SX - 1) / a.SX);
    P.n_nt = (int)((q.Ry + a.SY - 1) / a.SY);
    P.tile0 = tiles;
    P.dj = 1;
    if (ext) {
      const RankExt& e = ext[i];
      P.diag = e.diag; P.gt_x = e.gt_x; P.eq_x = e.eq_x; P.gt_y = e.gt_y; P.eq_y = e.eq_y;
      P.gx0 = (int)e.gx0; P.gy0 = (int)e.gy0;
      P.dj = (a.SX + a.SY - 2) / a.SY + 1;                       // SX consecutive Y rows touch at most this many Y tiles
    }
    P.n_me = pair ? (P.n_mt + 1) / 2 : P.n_mt;                   // pair-tiles: two adjacent X boxes x one Y box
    // Y tiles per band: the band's Y boxes (16 MB) must survive in L2 — which holds a line once per die — while the
    // X operand streams past; problems with fewer Y tiles are a single band
    const int64_t box_bytes = (int64_t)a.SY * Ny * d * 2;
    int64_t wn = (16ll << 20) / (box_bytes > 0 ? box_bytes : 1);
    if (const char* wv = getenv("NR_TC2_BAND")) { if (atoi(wv) > 0) wn = atoi(wv); }
    P.wn = (int)(wn < 1 ? 1 : (wn > P.n_nt ? P.n_nt : wn));
    if (mode == T2_MODE_DIAG) tiles += P.n_mt * P.dj;            // (X box, j-th Y tile with positives); empty entries are skipped
    else tiles += P.n_me * P.n_nt;
    if (int e = make_tmap_bf16(&a.tmx[i], q.x_bf16, q.Rx * Nx, d, a.MU)) return e;
    if (int e = make_tmap_bf16(&a.tmy[i], q.y_bf16, q.Ry * Ny, d, a.SY * (int)Ny / (pair ? 2 : 1))) return e;   // pair: half boxes
  }
  a.n_tiles = tiles;
  a.l2_hints = 0;
  if (const char* hv = getenv("NR_TC2_L2HINT")) a.l2_hints = atoi(hv);
  // measured A/B on one B200 (tools/gpu_r2u.sh): no difference for independent CTAs (b = 1024: 584 vs 584 us), SLOWER for
  // CTA pairs (617 -> 643 us: the release then waits for the slowest thread of the set plus a barrier, and the pair's
  // critical path is exactly accumulator release -> leader's next MMA) -> off; NR_TC2_ELECT=1 turns it on
  a.elect = 0;
  if (const char* ev2 = getenv("NR_TC2_ELECT")) a.elect = atoi(ev2) != 0;
  a.tile_counter = (unsigned int*)workspace;        // {next tile, finished claimers}: zero on entry, zero again on exit
  // debug timeline: the caller provides >= 16 + 8 * 16 * gridDim bytes of workspace when it sets NR_TC2_TRACE
  a.trace = nullptr;
  if (const char* tv = getenv("NR_TC2_TRACE"))
    if (atoi(tv) != 0) a.trace = (unsigned long long*)((uint8_t*)workspace + 16);
  a.cnt_w = mode == T2_MODE_COUNT ? ((a.SX > a.SY ? a.SX : a.SY) + 1) / 2 * 2 : 0;
  const size_t tail = (size_t)2 * T2_BM * a.hp_ld * 4 + (size_t)2 * (T2_BM / GL) * a.kg_ld * 4 +
                      (size_t)2 * a.SX * a.UN * 4 + (size_t)2 * a.UN * 4 + (size_t)2 * 4 * a.cnt_w * 4 + 512;
  const size_t budget = 227 * 1024 - 1024;   // alignment slack
  int stages = (int)((budget - tail) / (size_t)(T2_A_BYTES + a.b_bytes));
  if (stages > T2_MAX_STAGES) stages = T2_MAX_STAGES;
  NR_CHECK_ARG(stages >= 2, "nr_maxsim2_fwd: tile does not fit shared memory");
  a.stages = stages;
  const size_t smem = (size_t)stages * (T2_A_BYTES + a.b_bytes) + tail + 1024;
  int grid = tiles < sms ? tiles : sms;
  if (pair) grid = 2 * (tiles < sms / 2 ? tiles : sms / 2);
  // epilogue warps per set: 8 (two column halves per TMEM lane quarter) unless NR_TC2_HALVES=1
  int halves = 2;
  if (const char* hv = getenv("NR_TC2_HALVES")) halves = atoi(hv) == 1 ? 1 : 2;
  const bool exact_keys = (flags & 1) != 0;
  if (GL == 8) return dispatch_ny<8>((int)Ny, a, smem, grid, pair, halves, exact_keys, (cudaStream_t)stream);
  return dispatch_ny<4>((int)Ny, a, smem, grid, pair, halves, exact_keys, (cudaStream_t)stream);
}
