/* nrhead.h — C ABI of the B200-native NeighborRetr retrieval head (libnrhead.so, sm_100a).
 *
 * The reference (zzezze/NeighborRetr) is pure Python/PyTorch and has no FFI of its own; the
 * boundary it exposes for this path is the Python module surface of NeighborRetr/models
 * (SURVEY.md §8(b)).  Each entry point below replaces the ATen op chain of one reference
 * function and cites it (paths relative to the reference repo).  The Python mirror of the
 * reference interface lives in neighborretr_b200/{modeling,until_module,metrics,evaluator}.py and
 * calls these through ctypes.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer to a contiguous row-major buffer owned by the caller;
 *    the library never allocates, frees or retains device memory;
 *  - functions only ENQUEUE work on `stream` (a cudaStream_t passed as void*), never synchronise;
 *  - return 0 on success, a negative code otherwise; nr_last_error() gives the message;
 *  - masks are int64 {0,1} exactly as the reference passes them (modeling.py:486);
 *  - "f32" = IEEE binary32, "bf16" = bfloat16, index outputs are int32 / uint8 as stated.
 */
#ifndef NRHEAD_H_
#define NRHEAD_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NR_ABI_VERSION 1

/* loss selection flags for nr_row_losses_* */
#define NR_LOSS_CENTRALITY 1
#define NR_LOSS_NEIGHBOR 2
#define NR_LOSS_KL 4
#define NR_LOSS_UNIFORM 8
#define NR_NSAVE 16        /* per-row scalars kept between nr_row_losses_fwd and _bwd */
#define NR_MAX_K 128       /* largest num_neighbors */
#define NR_MAX_ROW_B 16384 /* largest B (columns) of a row-loss block */
#define NR_MAX_TOKENS 128  /* largest Nt / Nv of the max-sim kernels */

/* arithmetic of the token-pair contraction */
#define NR_PREC_FP32 0 /* CUDA-core fp32 FMA: the exact mode used for fp32-tolerance parity   */
#define NR_PREC_BF16 1 /* tcgen05 kind::f16, bf16 operands, fp32 accumulation in TMEM          */

int nr_version(void);
const char* nr_last_error(void);
/* 1 if `nr_*` tensor-core entry points can run on the current device (compute capability 10.x) */
int nr_device_supported(void);

/* ---- token preparation: F.normalize(x, dim=-1) (modeling.py:495-496, :415-418) -------------
 * x [rows, d] f32 -> xn_f32 [rows, d] (nullable), xn_bf16 [rows, d] (nullable),
 * inv_norm [rows] = 1/max(||x||, 1e-12), colsum_partials [nr_prep_partials(rows), d] (nullable):
 * per-CTA partial column sums of the normalised rows (for the centrality mean, modeling.py:419-424).
 * mask [rows] int64 (nullable): tokens with mask == 0 become ZERO rows of xn_bf16 only (so that their token
 * pairs are exactly 0 in the tensor-core contraction, as after the reference's mask multiplies,
 * modeling.py:500-501); xn_f32, inv_norm and the column sums are unaffected (compute_centrality_weights
 * uses padded tokens too, modeling.py:415-424). */
int64_t nr_prep_partials(int64_t rows);
int nr_prep_tokens(const float* x, int64_t rows, int64_t d, float* xn_f32, void* xn_bf16, float* inv_norm,
                   float* colsum_partials, const int64_t* mask, void* stream);
/* Split-bf16 operand of the fp32-accurate tensor-core mode (NR_PREC_BF16X3; same reference lines): every
 * normalised value n is written as hi = bf16(n), lo = bf16(n - hi) into xs_bf16 [rows, 3d]:
 *   role 1 (X side of nr_maxsim2_fwd): [hi | lo | hi]      role 2 (Y side): [hi | hi | lo]
 * so that nr_maxsim2_fwd with d' = 3d contracts hi.hi + lo.hi + hi.lo (error ~2^-17 per product instead of 2^-9).
 * Its transposed copy (nr_transpose_tokens_bf16 with d' = 3d) holds hi^T and lo^T for nr_maxsim2_bwd. */
int nr_prep_tokens_split(const float* x, int64_t rows, int64_t d, float* xn_f32, void* xs_bf16, int role,
                         float* inv_norm, float* colsum_partials, const int64_t* mask, void* stream);
/* backward of the normalisation: dx = (dxn + add_vec - xn <xn, dxn + add_vec>) * inv_norm.
 * add_vec [d] (nullable) is a gradient broadcast to every row (centrality mean path).
 * mask [rows] int64 (nullable): rows with mask == 0 ignore dxn (a masked token has no max-sim gradient).
 * accumulate != 0: dx += ... */
int nr_prep_tokens_bwd(const float* xn_f32, const float* inv_norm, const float* dxn, const float* add_vec,
                       const int64_t* mask, int64_t rows, int64_t d, float* dx, int accumulate, void* stream);

/* ---- persistent prepared memory bank: ring insert (reference modeling.py:222-249) ------------------
 * The bank lives in place as a ring of M sample slots; reference row i (newest first) = slot (head + i) mod M with
 * `head` an int32 in DEVICE memory (the launches replay inside a CUDA graph).  A step first moves the head back by
 * its n_new <= M samples (nr_bank_advance, which also stores their dataset indices), then writes the samples of one
 * modality at slots (head + j) mod M into every buffer given (all nullable): raw fp32 rows [M,N,d], int64 masks
 * [M,N], raw bf16 rows (the weight MLP's operand), the L2-normalised bf16 operand copy [M*N, d] (split_role 0) or
 * its split form [M*N, 3d] (split_role 1 / 2, see nr_prep_tokens_split) with masked tokens zeroed, and its transposed
 * copy [d or 3d, ld]. */
int nr_bank_advance(int* head, int64_t n_new, int64_t M, const int64_t* new_ind, int64_t* ring_ind, void* stream);
int nr_bank_insert(const float* new_feat, const int64_t* new_mask, int64_t n_new, int64_t N, int64_t d, int64_t M,
                   const int* head, float* ring_feat, int64_t* ring_mask, void* ring_raw_bf16, void* ring_xn_bf16,
                   int split_role, void* ring_xnT_bf16, int64_t ld, void* stream);
/* the text and the video rows of a step in one launch */
typedef struct {
  const float* new_feat; const int64_t* new_mask; int64_t N;
  float* ring_feat; int64_t* ring_mask; void* ring_raw_bf16; void* ring_xn_bf16; int split_role; void* ring_xnT_bf16;
  int64_t ld;
} nr_bank_side;
int nr_bank_insert_pair(const nr_bank_side* sides, int n_sides, int64_t n_new, int64_t d, int64_t M, const int* head,
                        void* stream);

/* ---- small exact-fp32 products (CUDA cores) --------------------------------------------------------------------
 * nr_matmul_f32: out[M,N] (+)= op(A)[M,K] X[K,N]; transA != 0: A is stored [K, M].  The global-feature gradients
 *   dgT = dG gV, dgV = dG^T gT (autograd of modeling.py:516-539 with one global token per sample).
 * nr_matvec_small: out = A (x + x2) (trans == 0, A [rows, cols]) or A^T (x + x2); x2 nullable; rows, cols <= 32.
 *   The 5 x 4 combination of the raw row sums into [total, centrality, uniform, neighbor, kl] (modeling.py:353-358)
 *   and its backward. */
int nr_matmul_f32(const float* A, int64_t lda, int transA, const float* X, int64_t ldx, int64_t M, int64_t K, int64_t N,
                  float* out, int64_t ldo, int accumulate, void* stream);
int nr_matvec_small(const float* A, int64_t rows, int64_t cols, int trans, const float* x, const float* x2, float* out,
                    void* stream);

/* ---- tcgen05 GEMMs of the token-weight MLPs (reference modeling.py:137-153, used :485-492) -------------------------
 * Linear(D -> H) + ReLU + Linear(H -> 1) per token, then masked softmax over the tokens of a sample.  bf16
 * operands, fp32 accumulation; operands are consumed AS STORED (K-major or MN-major TMA tiles), no transposes.
 * nr_cast_bf16: fp32 -> bf16 operand copy (x, W1).
 * nr_mlp_fwd : h = relu(x W1^T + b1) written as bf16 [T, H] (h_bf16 nullable: evaluation keeps nothing), and
 *              logits[t] += <h[t,:], w2>  (fp32 h, before the bf16 rounding; logits must be zero on entry).
 * nr_token_softmax: w = softmax_over_tokens(mask_fill(logits + b2, -9e15)); rows [0,Ra) use mask_a, the rest
 *              mask_b (batch tokens and bank tokens of one modality share the buffers; masks nullable).
 * nr_mlp_bwd_dx : dx [T, D] (+)= dh [T, H] W1 [H, D]       (accumulate != 0: split-K with red.add into dx)
 * nr_mlp_bwd_dw1: dw1 [H, D] += dh^T [H, T] x [T, D]       (split-K over the tokens, red.add: zero dw1 first) */
int nr_cast_bf16(const float* x, void* y_bf16, int64_t n, void* stream);
/* up to 8 such copies in one launch (src[i] -> dst[i], n[i] elements; host arrays of pointers / sizes) */
int nr_cast_bf16_multi(const float* const* src, void* const* dst, const int64_t* n, int n_segments, void* stream);
int nr_mlp_fwd(const void* x_bf16, int64_t T, int64_t D, const void* w1_bf16, int64_t H, const float* b1,
               const float* w2, void* h_bf16, float* logits, void* stream);
int nr_token_softmax(const float* logits, const float* b2, const int64_t* mask_a, const int64_t* mask_b, int64_t Ra,
                     int64_t R, int64_t N, float* w, void* stream);
/* the softmaxes of both modalities (different token counts) in one launch */
typedef struct {
  const float* logits; const float* b2; const int64_t* mask_a; const int64_t* mask_b;
  int64_t Ra, R, N;
  float* w;
} nr_softmax_side;
int nr_token_softmax_pair(const nr_softmax_side* sides, int n_sides, void* stream);
int nr_mlp_bwd_dx(const void* dh_bf16, int64_t T, int64_t H, const void* w1_bf16, int64_t D, float* dx, int accumulate,
                  void* stream);
int nr_mlp_bwd_dw1(const void* dh_bf16, int64_t T, int64_t H, const void* x_bf16, int64_t D, float* dw1, void* stream);
/* The text and the video MLP in ONE launch each way (a persistent grid per modality would serialise on shared memory):
 * nr_mlp_fwd_pair = nr_mlp_fwd of every side; nr_mlp_bwd_pair = nr_mlp_bwd_dw1 over T tokens (dw1 nullable) and
 * nr_mlp_bwd_dx (accumulating; dx nullable, zero on entry) over the first T_dx tokens of every side. */
typedef struct {
  const void* x_bf16; const void* w1_bf16; int64_t T;
  const float* b1; const float* w2; void* h_bf16; float* logits;       /* forward */
  const void* dh_bf16; float* dw1; float* dx; int64_t T_dx;            /* backward */
} nr_mlp_side;
/* reserve_sms: CTAs the persistent grid leaves free (the forward runs next to the single-CTA Sinkhorn kernel, which
 * needs a whole SM's registers and would otherwise wait for the first GEMM CTA to retire) */
int nr_mlp_fwd_pair(const nr_mlp_side* sides, int n_sides, int64_t D, int64_t H, int reserve_sms, void* stream);
int nr_mlp_bwd_pair(const nr_mlp_side* sides, int n_sides, int64_t D, int64_t H, void* stream);

/* chunks of 32 token rows: column count of the partial-sum buffer of nr_token_weights_bwd */
int64_t nr_mlp_chunks(int64_t T);

/* ---- token weights (modeling.py:485-492): second layer + masked softmax over the tokens of each sample --------
 * h [R*N, H] post-ReLU hidden activations (first layer = library GEMM; f32, or bf16 when h_bf16 != 0 — then dh of
 * the backward is bf16 as well), w2 [H], b2 [1] device scalars.
 *   logit[r,n] = <h[r,n,:], w2> + b2,  masked tokens -9e15,  w [R,N] = softmax over n.
 * Samples [0,Ra) read mask_a [Ra,N], samples [Ra,R) read mask_b [R-Ra,N] (int64, nullable): the batch tokens and
 * the memory-bank tokens of one modality share the launch. */
int nr_token_weights_fwd(const void* h, int h_bf16, const float* w2, const float* b2, const int64_t* mask_a,
                         const int64_t* mask_b, int64_t Ra, int64_t R, int64_t N, int64_t H, float* w, void* stream);
/* backward down to the hidden layer in one pass over h: dw_a [Ra,N] / dw_b [R-Ra,N] (nullable = zero) are the
 * gradients of w; dh [R*N,H] = dlogit * w2 * (h > 0); partials [2H+1, nr_mlp_chunks(R*N)] whose row sums are
 * db1 (rows 0..H), dw2 (rows H..2H) and db2 (row 2H). */
int nr_token_weights_bwd(const void* h, int h_bf16, const float* w, const float* dw_a, const float* dw_b, int64_t Ra,
                         int64_t R, int64_t N, const float* w2, int64_t H, void* dh, float* partials, void* stream);

/* ---- masked max-sim late interaction: one direction of local_level (modeling.py:499-509) ----
 *   H[rx, ry] = sum_x wx[rx,x] * max_y ( <xn[rx,x,:], yn[ry,y,:]> * mx[rx,x] * my[ry,y] )
 * local_level's S = 1/2 (H(text,video,text_weight) + H(video,text,video_weight)^T).
 * xn [Rx,Nx,d], yn [Ry,Ny,d] are L2-normalised tokens (f32 for NR_PREC_FP32, bf16 for NR_PREC_BF16),
 * wx [Rx,Nx] f32, mx [Rx,Nx] / my [Ry,Ny] int64 (nullable = all ones).
 * out[rx*out_sr + ry*out_sc] = alpha*H (+ previous value if accumulate);  out2 likewise (nullable),
 * so S and S^T can be produced by the same launch.
 * pmax [Rx,Ry,Nx] f32 and ystar [Rx,Ry,Nx] u8 (both nullable) keep max / arg-max for backward; ystar == 255
 * marks "no gradient" (masked X token, or the max is a masked pair, which is exactly 0). */
int nr_maxsim_fwd(int precision, const void* xn, const void* yn, const float* wx, const int64_t* mx,
                  const int64_t* my, int64_t Rx, int64_t Nx, int64_t Ry, int64_t Ny, int64_t d, float alpha,
                  float* out, int64_t out_sr, int64_t out_sc, float* out2, int64_t out2_sr, int64_t out2_sc,
                  int accumulate, float* pmax, uint8_t* ystar, void* stream);
/* ---- both directions of local_level from one accumulator tile (modeling.py:495-512), tensor cores only ------
 *   out[rx,ry] = alpha * ( sum_x wx[rx,x] max_y R + sum_y wy[ry,y] max_x R ),  R[rx,ry,x,y] = <x tokens, y tokens>
 * x_bf16 [Rx*Nx, d], y_bf16 [Ry*Ny, d]: L2-normalised bf16 tokens with masked tokens zeroed (nr_prep_tokens with a
 * mask), so no mask arguments exist here.  wx [Rx,Nx], wy [Ry,Ny] f32 token weights (0 for masked tokens).
 * out / out2 (nullable) are written as out[rx*sr + ry*sc]: S and S^T from the same launch.
 * Saved for backward (all nullable): pmax_x [Rx,Ry,Nx] f32 / ystar [Rx,Ry,Nx] u8 = max / arg-max over y for every
 * x token; pmax_y [Rx,Ry,Ny] f32 / xstar [Rx,Ry,Ny] u8 = max / arg-max over x for every y token (pmax_y carries
 * 3 fewer mantissa bits: the column arg-max travels in the low bits of the value).  Ties -> lower index.
 * Up to 4 problems with the same (Nx, Ny, d) share one persistent launch.  workspace: >= 16 bytes of device memory
 * holding the counters of the dynamic tile scheduler: ZERO before the first launch that uses it; every launch leaves
 * it zeroed again (no memset between launches).  One workspace per concurrently used stream. */
typedef struct {
  const void* x_bf16; const void* y_bf16;
  const float* wx; const float* wy;
  int64_t Rx, Ry;
  float alpha;
  float* out; int64_t out_sr, out_sc;
  float* out2; int64_t out2_sr, out2_sc;
  float* pmax_x; uint8_t* ystar;
  float* pmax_y; uint8_t* xstar;
} nr_maxsim2_problem;
int nr_maxsim2_supported(int64_t Nx, int64_t Ny, int64_t d);
int nr_maxsim2_fwd(const nr_maxsim2_problem* problems, int n_problems, int64_t Nx, int64_t Ny, int64_t d,
                   void* workspace, void* stream);
/* flags & 1: split-bf16 operands (NR_PREC_BF16X3, d = 3 x feature width): the column-direction arg-max compares
 * exact-order keys (2^-20 relative resolution) instead of the fp32 bits of v + 2 (2e-6 absolute) */
int nr_maxsim2_fwd_ex(const nr_maxsim2_problem* problems, int n_problems, int64_t Nx, int64_t Ny, int64_t d,
                      void* workspace, int flags, void* stream);
/* ---- evaluation ranks straight from the accumulator: the similarity matrix is never written ----------------
 * (reference: training/evaluator.py:21-63 builds the [Nq,Ng] matrix tile by tile on the host, utils/metrics.py:58-66
 * sorts every row and locates the positive; at a 100k gallery that matrix is 37 GiB.)
 * One block of nr_maxsim2_fwd's problem: X rows are the pairs gx0 .. gx0+Rx-1, Y rows the pairs gy0 .. gy0+Ry-1 of a
 * square test set; the positive of a row is the row of the other side with the same pair id.  diag [>= max(gx0+Rx,
 * gy0+Ry)] f32 holds the positives' scores by pair id.
 *   mode 1: diag[g] = S[g-gx0, g-gy0] for every pair g inside both ranges; only the tiles that contain a positive
 *           are contracted (with the tile geometry of mode 2: the values are the ones mode 2 recomputes).
 *   mode 2: gt_x[rx] += #{ry : S[rx,ry] > diag[gx0+rx]},  eq_x[rx] += #{ry : S[rx,ry] == diag[gx0+rx]},
 *           gt_y[ry] += #{rx : S[rx,ry] > diag[gy0+ry]},  eq_y[ry] += #{rx : S[rx,ry] == diag[gy0+ry]}
 *           (int32, accumulate; the positive itself counts as equal) — the (g, e) pairs nr_rank_count produces.
 * workspace / flags as nr_maxsim2_fwd_ex. */
typedef struct {
  const void* x_bf16; const void* y_bf16;
  const float* wx; const float* wy;
  int64_t Rx, Ry;
  float alpha;
  int64_t gx0, gy0;
  float* diag;
  int32_t* gt_x; int32_t* eq_x; int32_t* gt_y; int32_t* eq_y;
} nr_maxsim2_rank_problem;
int nr_maxsim2_rank(const nr_maxsim2_rank_problem* problem, int mode, int64_t Nx, int64_t Ny, int64_t d,
                    void* workspace, int flags, void* stream);
/* backward of nr_maxsim2_fwd w.r.t. the normalised tokens.  For one pair with g[rx,ry] = dH[rx*dh_sr + ry*dh_sc] *
 * dh_scale (dh_scale carries alpha) and the routing matrix
 *   C[(rx,x),(ry,y)] = g[rx,ry] * ( wx[rx,x] [y == ystar[rx,ry,x]] + wy[ry,y] [x == xstar[rx,ry,y]] ),
 * side 0: dst [Rx*Nx, d] += C   * Y tokens   (srcT = transposed bf16 Y tokens [d, src_ld]),
 * side 1: dst [Ry*Ny, d] += C^T * X tokens   (srcT = transposed bf16 X tokens [d, src_ld]).
 * dst is fp32, zero- or partially-filled: split-K partials are combined with red.global.add.
 * Up to 12 such jobs share ONE launch; jobs with the same dst (e.g. the text gradient from the batch pair and from
 * the bank pair) are accumulated in the same pass over their concatenated source tokens. */
typedef struct {
  int side;
  const void* srcT; int64_t src_ld;
  const float* wx; const float* wy;
  const uint8_t* ystar; const uint8_t* xstar;
  const float* dH; int64_t dh_sr, dh_sc; float dh_scale;
  int64_t Rx, Ry;
  float* dst;
  /* 0: the routing tile as plain bf16 (NR_PREC_BF16).  NR_PREC_BF16X3: every routing coefficient c (the exact fp32
   * sum of its two possible contributions) is split as hi = bf16(c), lo = bf16(c - hi): 2 = this job multiplies
   * the hi tile, 3 = the lo tile.  A pair then takes three jobs on the same dst: (2, hi^T source), (3, hi^T source),
   * (2, lo^T source). */
  int part;
} nr_maxsim2_bwd_job;
int nr_maxsim2_bwd(const nr_maxsim2_bwd_job* jobs, int n_jobs, int64_t Nx, int64_t Ny, int64_t d, void* stream);
/* ... w.r.t. the token weights (either output nullable):
 *   dwx[rx,x] += sum_ry g[rx,ry] pmax_x[rx,ry,x],   dwy[ry,y] += sum_rx g[rx,ry] pmax_y[rx,ry,y] */
int nr_maxsim2_bwd_w(const float* pmax_x, const float* pmax_y, const float* dH, int64_t dh_sr, int64_t dh_sc,
                     float dh_scale, int64_t Rx, int64_t Nx, int64_t Ry, int64_t Ny, float* dwx, float* dwy,
                     void* stream);
/* the same for up to 3 pairs in ONE launch (the batch pair and the two bank pairs of a step); gradients that
 * several pairs share (the text weights of the batch pair and of the text-vs-bank pair, ...) accumulate atomically */
typedef struct {
  const float* pmax_x; const float* pmax_y; const float* dH; int64_t dh_sr, dh_sc; float dh_scale;
  int64_t Rx, Ry; float* dwx; float* dwy;
} nr_maxsim2_bwd_w_job;
int nr_maxsim2_bwd_w_multi(const nr_maxsim2_bwd_w_job* jobs, int n_jobs, int64_t Nx, int64_t Ny, void* stream);
/* bf16 operand copy [rows, d] -> transposed [d, ld] (ld >= rows, multiple of 8): the K-major source
 * operand of the tensor-core backward contractions. */
int nr_transpose_tokens_bf16(const void* xn_bf16, int64_t rows, int64_t d, void* out, int64_t ld, void* stream);
/* backward of nr_maxsim_fwd w.r.t. the X tokens ("gather"):
 *   dxn[rx,x,:] += mx*wx[rx,x] * sum_ry dH[rx,ry] * my[ry,y*] * yn[ry,y*,:]
 * dH is read as dH[rx*dh_sr + ry*dh_sc] * dh_scale.
 * NR_PREC_FP32: yn = f32 tokens [Ry,Ny,d], src_ld ignored.
 * NR_PREC_BF16: yn = TRANSPOSED bf16 tokens [d, src_ld] (nr_transpose_tokens_bf16); dxn must be zero- or
 * partially-filled fp32: partial products are combined with float atomics (red.global.add). */
int nr_maxsim_bwd_x(int precision, const void* yn, int64_t src_ld, const float* wx, const int64_t* mx, const int64_t* my,
                    const uint8_t* ystar, const float* dH, int64_t dh_sr, int64_t dh_sc, float dh_scale,
                    int64_t Rx, int64_t Nx, int64_t Ry, int64_t Ny, int64_t d, float* dxn, void* stream);
/* ... w.r.t. the Y tokens ("scatter"):
 *   dyn[ry,y,:] += my[ry,y] * sum_rx dH[rx,ry] * sum_{x: y*(rx,ry,x)=y} mx*wx[rx,x] * xn[rx,x,:] */
int nr_maxsim_bwd_y(int precision, const void* xn, int64_t src_ld, const float* wx, const int64_t* mx, const int64_t* my,
                    const uint8_t* ystar, const float* dH, int64_t dh_sr, int64_t dh_sc, float dh_scale,
                    int64_t Rx, int64_t Nx, int64_t Ry, int64_t Ny, int64_t d, float* dyn, void* stream);
/* ... w.r.t. the token weights: dwx[rx,x] += sum_ry dH[rx,ry] * pmax[rx,ry,x] */
int nr_maxsim_bwd_w(const float* pmax, const float* dH, int64_t dh_sr, int64_t dh_sc, float dh_scale, int64_t Rx,
                    int64_t Nx, int64_t Ry, float* dwx, void* stream);

/* ---- centrality weights (modeling.py:403-430) ---------------------------------------------
 * mean_vec [d] = (sum over colsum partials) / rows_total;  g [B,d] raw global features.
 * w[a] = exp(cs * <normalize(g_a), mean_vec>).  gn [B,d] and ginv [B] are saved for backward. */
int nr_centrality_fwd(const float* colsum_partials, int64_t n_partials, int64_t rows_total, const float* g,
                      int64_t B, int64_t d, float cs, float* mean_vec, float* gn, float* ginv, float* w,
                      void* stream);
/* dw [B] -> dg [B,d] (+= if accumulate) and dmean [d] (the vector to broadcast to every token row,
 * already divided by rows_total). */
int nr_centrality_bwd(const float* mean_vec, const float* gn, const float* ginv, const float* w, const float* dw,
                      int64_t B, int64_t d, float cs, int64_t rows_total, float* dg, int accumulate,
                      float* dmean, void* stream);

/* ---- global similarity with one global token per sample (modeling.py:516-539 at Gt = Gv = 1):
 * out [Ra,Rb] = a [Ra,d] b[Rb,d]^T in exact fp32, outT [Rb,Ra] (nullable) its transpose from the same launch. */
int nr_gram_f32(const float* a, const float* b, int64_t Ra, int64_t Rb, int64_t d, float* out, float* outT,
                void* stream);

/* ---- row-block losses (until_module.py:56-211, :263-291, :303-328, :339-359) ----------------
 * X [rows,B] rows row0..row0+rows of the local similarity (t2v) or of its transpose (v2t);
 * G likewise for the global similarity; cbank [B] bank centrality by column (until_module.py:181);
 * w [rows]; sk_u [rows], sk_v [B] Sinkhorn duals; logit_scale: device scalar (NULL = 1).
 * row_out [4,rows]: per-row terms {centrality, neighbour, kl, uniform} (not yet divided by B);
 * nbr_idx [rows,k] int32: the top-k neighbour columns, descending similarity, ties -> lower column. */
int nr_row_losses_fwd(const float* X, int64_t ldx, const float* G, int64_t ldg, const float* cbank,
                      const float* w, const float* sk_u, const float* sk_v, int64_t rows, int64_t B,
                      int64_t row0, const float* logit_scale, int k, float tau_nbr, float tau_uni, float beta,
                      int flags, float* row_out, int32_t* nbr_idx, float* saved, void* stream);
/* gscale [4] device: upstream multipliers of the four per-row terms.  Writes dX [rows,B], dG [rows,B]
 * (nullable); accumulates dc [B] (atomic), writes dw [rows], accumulates dls [1] (atomic). */
int nr_row_losses_bwd(const float* X, int64_t ldx, const float* G, int64_t ldg, const float* cbank,
                      const float* w, const float* sk_u, const float* sk_v, int64_t rows, int64_t B,
                      int64_t row0, const float* logit_scale, int k, float tau_nbr, float tau_uni, float beta,
                      int flags, const int32_t* nbr_idx, const float* saved, const float* gscale, float* dX,
                      int64_t lddx, float* dG, int64_t lddg, float* dc, float* dw, float* dls, void* stream);
int nr_row_mean(const float* X, int64_t ld, int64_t rows, int64_t cols, float* out, void* stream);
int nr_vec_sums(const float* in, int64_t nvec, int64_t len, const float* scale, float* out, void* stream);
int nr_transpose_add(const float* a, int64_t lda, const float* b, int64_t ldb, float* out, int64_t ldo,
                     int64_t rows, int64_t cols, float alpha, float beta, void* stream);

/* ---- log-space Sinkhorn (until_module.py:222-251), both directions in one cooperative launch --
 * G [B,B], GT [B,B] (= G^T, caller-provided).  Chain 1 runs on G, chain 2 on G^T; outputs the
 * duals u1,v1,u2,v2 [B].  workspace: nr_sinkhorn_workspace_bytes(B) bytes, zero-initialised by
 * the library on `stream`. */
/* (for B > 1536 the workspace also holds exp(G - max G) and its transpose, 2 * B * B floats: the batch no longer
 * fits shared memory and every half-iteration streams them from HBM) */
size_t nr_sinkhorn_workspace_bytes(int64_t B);
int nr_sinkhorn(const float* G, const float* GT, int64_t B, int iters, float* u1, float* v1, float* u2,
                float* v2, void* workspace, size_t workspace_bytes, void* stream);
/* as nr_sinkhorn with the rows per CTA of the multi-CTA variants chosen by the caller (0 = automatic): inside a head
 * step fewer, larger CTAs leave SMs to the concurrent token-pair contraction */
int nr_sinkhorn_ex(const float* G, const float* GT, int64_t B, int iters, float* u1, float* v1, float* u2, float* v2,
                   void* workspace, size_t workspace_bytes, int rows_per_cta, void* stream);

/* ---- memory-bank FIFO (modeling.py:222-249): bank = cat(new, old)[:capacity] -----------------
 * out [cap, row_bytes] <- new [n_new, row_bytes] followed by old [n_old, row_bytes], truncated. */
int nr_fifo_update(const void* new_rows, int64_t n_new, const void* old_rows, int64_t n_old, void* out,
                   int64_t capacity, int64_t row_bytes, void* stream);

/* ---- evaluation ranking (utils/metrics.py:38-79) ---------------------------------------------
 * S [Q, N] (row stride lds), diag_col0: column of row q's positive is diag_col0 + q, its score is
 * read from diag[q] if diag != NULL (sharded galleries) else from S.
 * gt[q] += #{j : S[q,j] > s_qq},  eq[q] += #{j : S[q,j] == s_qq}   (int32, accumulate). */
int nr_rank_count(const float* S, int64_t lds, int64_t Q, int64_t N, const float* diag, int64_t diag_col0,
                  int32_t* gt, int32_t* eq, void* stream);
/* per-row top-k (value desc, ties -> lower column): vals [Q,k] f32, idx [Q,k] int32 (+col_offset) */
int nr_topk_rows(const float* S, int64_t lds, int64_t Q, int64_t N, int k, int32_t col_offset, float* vals,
                 int32_t* idx, void* stream);
/* merge W per-shard top-k lists [W,Q,k] into the global top-k [Q,k] (ties -> lower global column) */
int nr_topk_merge(const float* vals, const int32_t* idx, int64_t W, int64_t Q, int k, float* out_vals,
                  int32_t* out_idx, void* stream);

/* ---- multi-sentence evaluation (utils/metrics.py:81-145, padding at training/evaluator.py:216-239) ----
 * S [Q, N] (row stride lds) holds one caption per row; target[q] = GLOBAL column of caption q's video, the
 * block covers global columns [col_offset, col_offset + N).  Score of the positive: diag[q] if diag != NULL
 * (column-sharded galleries) else S[q, target[q] - col_offset].
 * gt[q]        += #{j : S[q,j] > s or S[q,j] is NaN}           (torch's descending sort puts NaN first)
 * eq_before[q] += #{j : S[q,j] == s and col_offset + j < target[q]}   (stable order of equal scores)
 * so that rank(q) = gt + eq_before is the reference's double-argsort diagonal (metrics.py:97-103).
 * valid[q] (optional) = 1 unless s is +-inf or NaN (metrics.py:106-109); invalid rows add nothing. */
int nr_rank_count_target(const float* S, int64_t lds, int64_t Q, int64_t N, const int32_t* target,
                         const float* diag, int64_t col_offset, int32_t* gt, int32_t* eq_before, int32_t* valid,
                         void* stream);
/* video->text similarity of a multi-sentence test set (metrics.py:124-145): captions of video group i are the
 * rows [group_start[i], group_start[i+1]) of S [T, V]; out[j, i] = max over those rows of S[t, j] with NaN
 * read as -inf (out row stride ldo >= G; an empty group gives -inf).  group_start: int32 [G+1] on the device. */
int nr_group_max_t(const float* S, int64_t lds, int64_t T, int64_t V, const int32_t* group_start, int64_t G,
                   float* out, int64_t ldo, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* NRHEAD_H_ */
