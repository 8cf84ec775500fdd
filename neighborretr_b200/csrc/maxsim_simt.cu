// Masked max-sim late interaction in exact fp32 on the CUDA cores (NR_PREC_FP32).
//
// This is the fp32-tolerance parity mode of the token-pair contraction
// (NeighborRetr/models/modeling.py:499-509): a register-tiled fp32 FMA GEMM per (X sample, chunk
// of Y samples) whose [Nx, CY*Ny] tile never leaves shared memory — the epilogue takes the masked
// row max / arg-max per Y sample and the token-weighted sum, and writes only H, pmax and ystar.
// The tensor-core (tcgen05) version of the same contract lives in maxsim_tc.cu.
//
// The backward kernels work from the saved arg-max (at most Nx non-zeros per (rx,ry) pair out of
// Nx*Ny): a gather-accumulate for the X side, a scatter-accumulate (per-CTA shared
// memory accumulators, no atomics) for the Y side, and a batched dot for the token weights.
#include "common.cuh"
#include "nrhead_internal.h"

namespace nr {

constexpr int MS_THREADS = 256;
constexpr int MS_KC = 16;          // k-chunk (floats) staged per iteration

struct MaxsimArgs {
  const float* xn; const float* yn; const float* wx; const int64_t* mx; const int64_t* my;
  int Rx, Nx, Ry, Ny, d;
  float alpha;
  float* out; int64_t out_sr, out_sc; float* out2; int64_t out2_sr, out2_sc; int accumulate;
  float* pmax; uint8_t* ystar;
  int CY, NXP, NYTP;               // Y samples per CTA, padded tile dims (multiples of 4)
};

// grid (ceil(Ry/CY), Rx); thread (tx,ty) owns a 4x4 micro-tile of the [NXP, NYTP] tile
__global__ void __launch_bounds__(MS_THREADS)
maxsim_fwd_simt_kernel(MaxsimArgs a) {
  extern __shared__ float sm[];
  const int NXP = a.NXP, NYTP = a.NYTP, d = a.d;
  float* Xs = sm;                         // [MS_KC][NXP]
  float* Ys = Xs + MS_KC * NXP;           // [MS_KC][NYTP]
  float* Rs = Ys + MS_KC * NYTP;          // [NXP][NYTP+1]
  float* Ps = Rs + NXP * (NYTP + 1);      // [Nx][CY]
  const int tid = threadIdx.x, rx = blockIdx.y, ry0 = blockIdx.x * a.CY;
  const int cy_n = min(a.CY, a.Ry - ry0);
  const int nyt = cy_n * a.Ny;            // live columns
  const int ntx = NYTP / 4, nty = NXP / 4;
  const int tx = tid % ntx, ty = tid / ntx;
  const bool active = ty < nty;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const float* xbase = a.xn + (int64_t)rx * a.Nx * d;
  const float* ybase = a.yn + (int64_t)ry0 * a.Ny * d;
  const int kq = MS_KC / 4;               // float4 per row per chunk
  for (int k0 = 0; k0 < d; k0 += MS_KC) {
    // stage transposed chunks: Xs[k][row], Ys[k][col]
    for (int e = tid; e < (NXP + NYTP) * kq; e += MS_THREADS) {
      int r = e / kq, q = e % kq;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      float* dst;
      if (r < NXP) {
        if (r < a.Nx && k0 + q * 4 < d) v = *reinterpret_cast<const float4*>(xbase + (int64_t)r * d + k0 + q * 4);
        dst = Xs + (q * 4) * NXP + r;
        dst[0] = v.x; dst[NXP] = v.y; dst[2 * NXP] = v.z; dst[3 * NXP] = v.w;
      } else {
        int c = r - NXP;
        if (c < nyt && k0 + q * 4 < d) v = *reinterpret_cast<const float4*>(ybase + (int64_t)c * d + k0 + q * 4);
        dst = Ys + (q * 4) * NYTP + c;
        dst[0] = v.x; dst[NYTP] = v.y; dst[2 * NYTP] = v.z; dst[3 * NYTP] = v.w;
      }
    }
    __syncthreads();
    if (active) {
#pragma unroll
      for (int k = 0; k < MS_KC; ++k) {
        float4 xv = *reinterpret_cast<const float4*>(Xs + k * NXP + ty * 4);
        float4 yv = *reinterpret_cast<const float4*>(Ys + k * NYTP + tx * 4);
        const float xa[4] = {xv.x, xv.y, xv.z, xv.w}, ya[4] = {yv.x, yv.y, yv.z, yv.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(xa[i], ya[j], acc[i][j]);
      }
    }
    __syncthreads();
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) Rs[(ty * 4 + i) * (NYTP + 1) + tx * 4 + j] = acc[i][j];
  }
  __syncthreads();
  // epilogue 1: masked max / arg-max over the Ny tokens of each Y sample (ties -> lower y)
  for (int e = tid; e < a.Nx * cy_n; e += MS_THREADS) {
    int x = e / cy_n, cy = e % cy_n, ry = ry0 + cy;
    const bool mxv = a.mx ? (a.mx[(int64_t)rx * a.Nx + x] != 0) : true;
    float best = NR_NEG_INF;
    int bi = 0;
    bool bmask = true;
    for (int y = 0; y < a.Ny; ++y) {
      const bool myv = a.my ? (a.my[(int64_t)ry * a.Ny + y] != 0) : true;
      float v = (mxv && myv) ? Rs[x * (NYTP + 1) + cy * a.Ny + y] : 0.f;   // masked pairs are exactly 0
      if (v > best) { best = v; bi = y; bmask = myv; }
    }
    if (!mxv || !bmask) bi = 255;     // arg-max byte 255 = "no gradient" (masked token / masked winning pair)
    Ps[x * a.CY + cy] = best * a.wx[(int64_t)rx * a.Nx + x];
    int64_t o = ((int64_t)rx * a.Ry + ry) * a.Nx + x;
    if (a.pmax) a.pmax[o] = best;
    if (a.ystar) a.ystar[o] = (uint8_t)bi;
  }
  __syncthreads();
  // epilogue 2: token-weighted sum over x
  for (int cy = tid; cy < cy_n; cy += MS_THREADS) {
    float h = 0.f;
    for (int x = 0; x < a.Nx; ++x) h += Ps[x * a.CY + cy];
    h *= a.alpha;
    int ry = ry0 + cy;
    float* p = a.out + (int64_t)rx * a.out_sr + (int64_t)ry * a.out_sc;
    *p = a.accumulate ? (*p + h) : h;
    if (a.out2) {
      float* p2 = a.out2 + (int64_t)rx * a.out2_sr + (int64_t)ry * a.out2_sc;
      *p2 = a.accumulate ? (*p2 + h) : h;
    }
  }
}

struct MaxsimBwdArgs {
  const float* src;                 // yn for bwd_x, xn for bwd_y
  const float* wx; const int64_t* mx; const int64_t* my; const uint8_t* ystar;
  const float* dH; int64_t dh_sr, dh_sc; float dh_scale;
  int Rx, Nx, Ry, Ny, d;
  float* dst;                       // dxn or dyn (accumulated)
};

constexpr int BX_THREADS = 128;
constexpr int BX_XT = 4;            // x tokens per CTA
constexpr int BX_MAXV = 2;          // float4 per thread: d <= 128*4*BX_MAXV = 1024

// grid (ceil(Nx/BX_XT), Rx): gather-accumulate dxn[rx, x, :]
__global__ void __launch_bounds__(BX_THREADS)
maxsim_bwd_x_simt_kernel(MaxsimBwdArgs a) {
  const int rx = blockIdx.y, x0 = blockIdx.x * BX_XT, tid = threadIdx.x, d = a.d;
  float4 acc[BX_XT][BX_MAXV];
  float coefx[BX_XT];
#pragma unroll
  for (int i = 0; i < BX_XT; ++i) {
    int x = x0 + i;
    bool ok = x < a.Nx && (a.mx ? a.mx[(int64_t)rx * a.Nx + x] != 0 : true);
    coefx[i] = ok ? a.wx[(int64_t)rx * a.Nx + x] : 0.f;   // (masked tokens also carry ystar == 255)
#pragma unroll
    for (int v = 0; v < BX_MAXV; ++v) acc[i][v] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  for (int ry = 0; ry < a.Ry; ++ry) {
    const float g = a.dH[(int64_t)rx * a.dh_sr + (int64_t)ry * a.dh_sc] * a.dh_scale;
    if (g == 0.f) continue;
    const uint8_t* ys = a.ystar + ((int64_t)rx * a.Ry + ry) * a.Nx;
#pragma unroll
    for (int i = 0; i < BX_XT; ++i) {
      if (coefx[i] == 0.f) continue;
      const int y = ys[x0 + i];
      if (y == 255) continue;
      const float cf = g * coefx[i];
      const float* src = a.src + ((int64_t)ry * a.Ny + y) * d;
#pragma unroll
      for (int v = 0; v < BX_MAXV; ++v) {
        int c = (v * BX_THREADS + tid) * 4;
        if (c < d) {
          float4 s = *reinterpret_cast<const float4*>(src + c);
          acc[i][v].x = fmaf(cf, s.x, acc[i][v].x); acc[i][v].y = fmaf(cf, s.y, acc[i][v].y);
          acc[i][v].z = fmaf(cf, s.z, acc[i][v].z); acc[i][v].w = fmaf(cf, s.w, acc[i][v].w);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < BX_XT; ++i) {
    int x = x0 + i;
    if (x >= a.Nx || coefx[i] == 0.f) continue;
    float* dst = a.dst + ((int64_t)rx * a.Nx + x) * d;
#pragma unroll
    for (int v = 0; v < BX_MAXV; ++v) {
      int c = (v * BX_THREADS + tid) * 4;
      if (c < d) {
        float4* p = reinterpret_cast<float4*>(dst + c);
        float4 o = *p;
        o.x += acc[i][v].x; o.y += acc[i][v].y; o.z += acc[i][v].z; o.w += acc[i][v].w;
        *p = o;
      }
    }
  }
}

// grid (Ry): scatter-accumulate dyn[ry, y, :] in shared memory accumulators [Ny][d]
__global__ void __launch_bounds__(BX_THREADS)
maxsim_bwd_y_simt_kernel(MaxsimBwdArgs a) {
  extern __shared__ float accsm[];
  const int ry = blockIdx.x, tid = threadIdx.x, d = a.d;
  for (int e = tid; e < a.Ny * d; e += BX_THREADS) accsm[e] = 0.f;
  __syncthreads();
  for (int rx = 0; rx < a.Rx; ++rx) {
    const float g = a.dH[(int64_t)rx * a.dh_sr + (int64_t)ry * a.dh_sc] * a.dh_scale;
    if (g == 0.f) continue;
    const uint8_t* ys = a.ystar + ((int64_t)rx * a.Ry + ry) * a.Nx;
    for (int x = 0; x < a.Nx; ++x) {
      const int y = ys[x];
      if (y == 255) continue;
      const float cf = g * a.wx[(int64_t)rx * a.Nx + x];
      const float* src = a.src + ((int64_t)rx * a.Nx + x) * d;
      float* dstrow = accsm + y * d;
      for (int c = tid * 4; c < d; c += BX_THREADS * 4) {
        float4 s = *reinterpret_cast<const float4*>(src + c);
        float4* p = reinterpret_cast<float4*>(dstrow + c);     // each thread owns its columns: no race
        float4 o = *p;
        o.x = fmaf(cf, s.x, o.x); o.y = fmaf(cf, s.y, o.y); o.z = fmaf(cf, s.z, o.z); o.w = fmaf(cf, s.w, o.w);
        *p = o;
      }
    }
  }
  __syncthreads();
  float* dst = a.dst + (int64_t)ry * a.Ny * d;
  for (int e = tid; e < a.Ny * d; e += BX_THREADS) dst[e] += accsm[e];
}

// grid (Rx), block 256: dwx[rx, x] += sum_ry dH[rx,ry] * pmax[rx,ry,x]
__global__ void __launch_bounds__(256)
maxsim_bwd_w_kernel(const float* __restrict__ pmax, const float* __restrict__ dH, int64_t dh_sr, int64_t dh_sc,
                    float dh_scale, int Rx, int Nx, int Ry, float* __restrict__ dwx) {
  __shared__ float part[256];
  const int rx = blockIdx.x, tid = threadIdx.x;
  const int lanes = 256 / Nx;            // ry-lanes per x (Nx <= 128)
  const int x = tid % Nx, l = tid / Nx;
  float s = 0.f;
  if (l < lanes) {
#pragma unroll 8
    for (int ry = l; ry < Ry; ry += lanes)
      s += dH[(int64_t)rx * dh_sr + (int64_t)ry * dh_sc] * pmax[((int64_t)rx * Ry + ry) * Nx + x];
  }
  part[tid] = s;
  __syncthreads();
  if (tid < Nx) {
    float t = 0.f;
    for (int q = 0; q < lanes; ++q) t += part[q * Nx + tid];
    dwx[(int64_t)rx * Nx + tid] += t * dh_scale;
  }
}

}  // namespace nr

using namespace nr;

int nr_maxsim_fwd_simt(const float* xn, const float* yn, const float* wx, const int64_t* mx, const int64_t* my,
                       int64_t Rx, int64_t Nx, int64_t Ry, int64_t Ny, int64_t d, float alpha, float* out,
                       int64_t out_sr, int64_t out_sc, float* out2, int64_t out2_sr, int64_t out2_sc,
                       int accumulate, float* pmax, uint8_t* ystar, cudaStream_t stream) {
  MaxsimArgs a{xn, yn, wx, mx, my, (int)Rx, (int)Nx, (int)Ry, (int)Ny, (int)d, alpha, out, out_sr, out_sc,
               out2, out2_sr, out2_sc, accumulate, pmax, ystar, 0, 0, 0};
  a.NXP = ((int)Nx + 3) / 4 * 4;
  int nty = a.NXP / 4;
  int max_cols = 4 * (MS_THREADS / nty);
  if (max_cols > 256) max_cols = 256;
  int cy = max_cols / (int)Ny;
  if (cy < 1) cy = 1;
  if (cy > Ry) cy = (int)Ry;
  a.CY = cy;
  a.NYTP = (cy * (int)Ny + 3) / 4 * 4;
  NR_CHECK_ARG((a.NXP / 4) * (a.NYTP / 4) <= MS_THREADS, "nr_maxsim_fwd(fp32): tile %dx%d too large", a.NXP,
               a.NYTP);
  size_t smem = sizeof(float) * ((size_t)MS_KC * (a.NXP + a.NYTP) + (size_t)a.NXP * (a.NYTP + 1) + (size_t)Nx * cy);
  if (smem > 48 * 1024)
    NR_CUDA(cudaFuncSetAttribute(maxsim_fwd_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  dim3 grid((unsigned)((Ry + cy - 1) / cy), (unsigned)Rx);
  maxsim_fwd_simt_kernel<<<grid, MS_THREADS, smem, stream>>>(a);
  NR_CHECK_LAUNCH("nr_maxsim_fwd(fp32)");
  return 0;
}

int nr_maxsim_bwd_x_simt(const float* yn, const float* wx, const int64_t* mx, const int64_t* my,
                         const uint8_t* ystar, const float* dH, int64_t dh_sr, int64_t dh_sc, float dh_scale,
                         int64_t Rx, int64_t Nx, int64_t Ry, int64_t Ny, int64_t d, float* dxn,
                         cudaStream_t stream) {
  NR_CHECK_ARG(d % 4 == 0 && d <= BX_THREADS * 4 * BX_MAXV, "nr_maxsim_bwd_x(fp32): unsupported d=%lld", (long long)d);
  MaxsimBwdArgs a{yn, wx, mx, my, ystar, dH, dh_sr, dh_sc, dh_scale, (int)Rx, (int)Nx, (int)Ry, (int)Ny, (int)d, dxn};
  dim3 grid((unsigned)((Nx + BX_XT - 1) / BX_XT), (unsigned)Rx);
  maxsim_bwd_x_simt_kernel<<<grid, BX_THREADS, 0, stream>>>(a);
  NR_CHECK_LAUNCH("nr_maxsim_bwd_x(fp32)");
  return 0;
}

int nr_maxsim_bwd_y_simt(const float* xn, const float* wx, const int64_t* mx, const int64_t* my,
                         const uint8_t* ystar, const float* dH, int64_t dh_sr, int64_t dh_sc, float dh_scale,
                         int64_t Rx, int64_t Nx, int64_t Ry, int64_t Ny, int64_t d, float* dyn,
                         cudaStream_t stream) {
  NR_CHECK_ARG(d % 4 == 0, "nr_maxsim_bwd_y(fp32): unsupported d=%lld", (long long)d);
  MaxsimBwdArgs a{xn, wx, mx, my, ystar, dH, dh_sr, dh_sc, dh_scale, (int)Rx, (int)Nx, (int)Ry, (int)Ny, (int)d, dyn};
  size_t smem = sizeof(float) * (size_t)Ny * d;
  NR_CHECK_ARG(smem <= 200 * 1024, "nr_maxsim_bwd_y(fp32): Ny*d too large for shared memory");
  if (smem > 48 * 1024)
    NR_CUDA(cudaFuncSetAttribute(maxsim_bwd_y_simt_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  maxsim_bwd_y_simt_kernel<<<(unsigned)Ry, BX_THREADS, smem, stream>>>(a);
  NR_CHECK_LAUNCH("nr_maxsim_bwd_y(fp32)");
  return 0;
}

extern "C" int nr_maxsim_bwd_w(const float* pmax, const float* dH, int64_t dh_sr, int64_t dh_sc, float dh_scale,
                               int64_t Rx, int64_t Nx, int64_t Ry, float* dwx, void* stream) {
  NR_CHECK_ARG(pmax && dH && dwx && Rx > 0 && Ry > 0 && Nx > 0 && Nx <= NR_MAX_TOKENS, "nr_maxsim_bwd_w: bad arguments");
  maxsim_bwd_w_kernel<<<(unsigned)Rx, 256, 0, (cudaStream_t)stream>>>(pmax, dH, dh_sr, dh_sc, dh_scale, (int)Rx,
                                                                     (int)Nx, (int)Ry, dwx);
  NR_CHECK_LAUNCH("nr_maxsim_bwd_w");
  return 0;
}
