/*
 * dsvg_b200.h -- C ABI of libdsvg_b200.so: the sm_100a kernels behind deepsvg_b200.SVGTransformer / SVGLoss.
 *
 * The reference (alexandre01/deepsvg) has NO native/FFI interface: its hot path is Python calling ATen.  Each entry
 * point below therefore names the reference Python call site(s) (file:line under the reference repo) whose
 * arithmetic it replaces.  Conventions:
 *   - plain C types only; every pointer is a CUDA device pointer unless the name ends in _host;
 *   - the caller (PyTorch) owns every buffer; the library never allocates user-visible memory, never
 *     synchronises the device, and launches on the `stream` argument (a cudaStream_t passed as void*);
 *   - return value 0 = success; anything else = failure with a message available from dsvg_last_error()
 *     (thread-local); no C++ exception crosses this boundary;
 *   - "act" tensors are bf16 with an optional second "lo" plane `lo_off` ELEMENTS after the first
 *     (lo_off = 0: fast single-plane bf16; lo_off != 0: parity mode, value = hi + lo, GEMMs run as bf16x3);
 *   - dropout is Philox4x32-10 keyed by (seed, site); p = 0 disables it (eval mode).
 */
#ifndef DSVG_B200_H
#define DSVG_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint16_t dsvg_bf16; /* raw bfloat16 bits */

/* ---- library --------------------------------------------------------------------------------------- */
const char* dsvg_last_error(void);
int dsvg_abi_version(void);
/* number of kernels this library has launched since load (per process); bench.py reports it as gpu_launches */
unsigned long long dsvg_launch_count(void);

/* ---- dense contractions on the tcgen05 tensor cores ------------------------------------------------- */
/* Fused epilogue of dsvg_linear, applied to the fp32 accumulator in this order:
 *   v = acc; v += bias[col]; if (col < scale_cols) v *= scale; if (relu) v = max(v,0);
 *   v *= dropout(p, seed, site, idx = row*N+col); v += rowvec[(row / rows_per_group)*rowvec_ld + col];
 *   v *= (mask[row*mask_ld+col] != 0) ? mask_scale : 0;  v += residual[row*res_ld+col];
 *   out_f32[row*out_f32_ld+col] = v;  out_act[row*out_act_ld+col] = v (hi[/lo] bf16).
 * NULL pointers skip their step. */
typedef struct dsvg_epilogue {
  const float* bias;
  int scale_cols;
  float scale;
  int relu;
  float drop_p;
  uint32_t drop_site;
  uint64_t seed;
  const float* rowvec;
  int rowvec_ld;
  int rows_per_group;
  const dsvg_bf16* mask;
  size_t mask_lo_off;
  int mask_ld;
  float mask_scale;
  const float* residual;
  int res_ld;
  float* out_f32;
  int out_f32_ld;
  dsvg_bf16* out_act;
  size_t out_lo_off;
  int out_act_ld;
} dsvg_epilogue;

/* Y[M,N] = epilogue(X[M,K] . W[N,K]^T).  X, W row-major bf16 act tensors (K contiguous; lda/ldb in elements,
 * multiples of 8).  Replaces every F.linear on the path: functional.py:92,249 (QKV / out-proj),
 * improved_transformer.py:52,131,139 (FFN, linear_global), basic_blocks.py:18-21,36-37,60-63 (heads, ResNet),
 * model.py:50,182-183,197 (embed_fcn, VAE/bottleneck) -- and, with W^T, their input gradients. */
int dsvg_linear(const dsvg_bf16* X, size_t x_lo_off, int lda, const dsvg_bf16* W, size_t w_lo_off, int ldb, int M,
                int N, int K, const dsvg_epilogue* ep, void* stream);

/* C[P,Q] (+)= alpha * A[M,P]^T . B[M,Q]   (contraction over the M rows; A, B row-major act tensors).
 * accumulate != 0: fp32 atomic add into C (split over M across CTAs), else C must be zero-filled by the caller
 * when more than one split is used -- the library always adds.  Weight gradients of every F.linear above
 * (autograd of the reference, loss.backward() at train.py:98). */
int dsvg_outer(const dsvg_bf16* A, size_t a_lo_off, int lda, const dsvg_bf16* B, size_t b_lo_off, int ldb, int M,
               int P, int Q, float alpha, float* C, int ldc, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DSVG_B200_H */
