// C-ABI dispatch of the max-sim entry points onto the fp32 CUDA-core kernels (maxsim_simt.cu) or the
// tcgen05 tensor-core kernels (maxsim_tc.cu).  No CPU path exists: an unsupported precision fails.
#include "common.cuh"
#include "nrhead_internal.h"

int nr_maxsim_fwd_simt(const float*, const float*, const float*, const int64_t*, const int64_t*, int64_t, int64_t,
                       int64_t, int64_t, int64_t, float, float*, int64_t, int64_t, float*, int64_t, int64_t, int,
                       float*, uint8_t*, cudaStream_t);
int nr_maxsim_bwd_x_simt(const float*, const float*, const int64_t*, const int64_t*, const uint8_t*, const float*,
                         int64_t, int64_t, float, int64_t, int64_t, int64_t, int64_t, int64_t, float*, cudaStream_t);
int nr_maxsim_bwd_y_simt(const float*, const float*, const int64_t*, const int64_t*, const uint8_t*, const float*,
                         int64_t, int64_t, float, int64_t, int64_t, int64_t, int64_t, int64_t, float*, cudaStream_t);
int nr_maxsim_fwd_tc(const void*, const void*, const float*, const int64_t*, const int64_t*, int64_t, int64_t,
                     int64_t, int64_t, int64_t, float, float*, int64_t, int64_t, float*, int64_t, int64_t, int,
                     float*, uint8_t*, cudaStream_t);
int nr_maxsim_bwd_tc(int side, const void*, int64_t, const float*, const int64_t*, const int64_t*, const uint8_t*,
                     const float*, int64_t, int64_t, float, int64_t, int64_t, int64_t, int64_t, int64_t, float*,
                     cudaStream_t);

static int check_dims(const char* fn, int64_t Rx, int64_t Nx, int64_t Ry, int64_t Ny, int64_t d) {
  NR_CHECK_ARG(Rx > 0 && Ry > 0 && d > 0, "%s: empty problem (Rx=%lld Ry=%lld d=%lld)", fn, (long long)Rx,
               (long long)Ry, (long long)d);
  NR_CHECK_ARG(Nx >= 1 && Nx <= NR_MAX_TOKENS && Ny >= 1 && Ny <= NR_MAX_TOKENS,
               "%s: tokens per sample must be in [1,%d] (Nx=%lld Ny=%lld)", fn, NR_MAX_TOKENS, (long long)Nx,
               (long long)Ny);
  NR_CHECK_ARG(d % 4 == 0, "%s: d=%lld must be a multiple of 4", fn, (long long)d);
  return 0;
}

extern "C" int nr_maxsim_fwd(int precision, const void* xn, const void* yn, const float* wx, const int64_t* mx,
                             const int64_t* my, int64_t Rx, int64_t Nx, int64_t Ry, int64_t Ny, int64_t d,
                             float alpha, float* out, int64_t out_sr, int64_t out_sc, float* out2, int64_t out2_sr,
                             int64_t out2_sc, int accumulate, float* pmax, uint8_t* ystar, void* stream) {
  if (int e = check_dims("nr_maxsim_fwd", Rx, Nx, Ry, Ny, d)) return e;
  NR_CHECK_ARG(xn && yn && wx && out, "nr_maxsim_fwd: null pointer");
  if (precision == NR_PREC_FP32)
    return nr_maxsim_fwd_simt((const float*)xn, (const float*)yn, wx, mx, my, Rx, Nx, Ry, Ny, d, alpha, out, out_sr,
                              out_sc, out2, out2_sr, out2_sc, accumulate, pmax, ystar, (cudaStream_t)stream);
  if (precision == NR_PREC_BF16)
    return nr_maxsim_fwd_tc(xn, yn, wx, mx, my, Rx, Nx, Ry, Ny, d, alpha, out, out_sr, out_sc, out2, out2_sr,
                            out2_sc, accumulate, pmax, ystar, (cudaStream_t)stream);
  nr::set_error("nr_maxsim_fwd: unknown precision %d", precision);
  return -1;
}

extern "C" int nr_maxsim_bwd_x(int precision, const void* yn, int64_t src_ld, const float* wx, const int64_t* mx, const int64_t* my,
                               const uint8_t* ystar, const float* dH, int64_t dh_sr, int64_t dh_sc, float dh_scale,
                               int64_t Rx, int64_t Nx, int64_t Ry, int64_t Ny, int64_t d, float* dxn, void* stream) {
  if (int e = check_dims("nr_maxsim_bwd_x", Rx, Nx, Ry, Ny, d)) return e;
  NR_CHECK_ARG(yn && wx && ystar && dH && dxn, "nr_maxsim_bwd_x: null pointer");
  if (precision == NR_PREC_FP32)
    return nr_maxsim_bwd_x_simt((const float*)yn, wx, mx, my, ystar, dH, dh_sr, dh_sc, dh_scale, Rx, Nx, Ry, Ny, d,
                                dxn, (cudaStream_t)stream);
  if (precision == NR_PREC_BF16)
    return nr_maxsim_bwd_tc(0, yn, src_ld, wx, mx, my, ystar, dH, dh_sr, dh_sc, dh_scale, Rx, Nx, Ry, Ny, d, dxn,
                            (cudaStream_t)stream);
  nr::set_error("nr_maxsim_bwd_x: unknown precision %d", precision);
  return -1;
}

extern "C" int nr_maxsim_bwd_y(int precision, const void* xn, int64_t src_ld, const float* wx, const int64_t* mx, const int64_t* my,
                               const uint8_t* ystar, const float* dH, int64_t dh_sr, int64_t dh_sc, float dh_scale,
                               int64_t Rx, int64_t Nx, int64_t Ry, int64_t Ny, int64_t d, float* dyn, void* stream) {
  if (int e = check_dims("nr_maxsim_bwd_y", Rx, Nx, Ry, Ny, d)) return e;
  NR_CHECK_ARG(xn && wx && ystar && dH && dyn, "nr_maxsim_bwd_y: null pointer");
  if (precision == NR_PREC_FP32)
    return nr_maxsim_bwd_y_simt((const float*)xn, wx, mx, my, ystar, dH, dh_sr, dh_sc, dh_scale, Rx, Nx, Ry, Ny, d,
                                dyn, (cudaStream_t)stream);
  if (precision == NR_PREC_BF16)
    return nr_maxsim_bwd_tc(1, xn, src_ld, wx, mx, my, ystar, dH, dh_sr, dh_sc, dh_scale, Rx, Nx, Ry, Ny, d, dyn,
                            (cudaStream_t)stream);
  nr::set_error("nr_maxsim_bwd_y: unknown precision %d", precision);
  return -1;
}
