// LayerNorm forward/backward (eps 1e-5, affine), optionally fused with the masked mean over a sequence
// ("pool": model.py:137 and :161 of the reference) -- HBM-bound warp-per-row kernels.
//
//   reference: nn.LayerNorm at improved_transformer.py:35-36,43,51,127,138 and transformer.py:185-186,239-240;
//              pooled means at model.py:137,161.
//
// Each lane owns the channels {128*v + 4*lane .. +3 : v < D/128}; rows are reduced with warp shuffles; the residual
// stream x and its gradient stay fp32, the normalised output feeding a tensor-core GEMM is a (split-)bf16 act tensor.
#include "../../include/dsvg_b200.h"
#include "common.cuh"

namespace dsvg {
extern unsigned long long g_launches;

constexpr int kLnWarps = 8;

template <int NV>
__device__ __forceinline__ void load_row(const float* p, int lane, float4 (&v)[NV]) {
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = *reinterpret_cast<const float4*>(p + 128 * i + 4 * lane);
}

template <int NV>
__device__ __forceinline__ void row_stats(const float4 (&v)[NV], float& mean, float& rstd, float eps) {
  constexpr float invD = 1.f / float(NV * 128);
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
  mean = warp_sum(s) * invD;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  rstd = rsqrtf(warp_sum(q) * invD + eps);
}

__device__ __forceinline__ void store_act4(bf16* p, size_t lo_off, size_t i, float4 x) {
  act_store2(p, lo_off, i, x.x, x.y);
  act_store2(p, lo_off, i + 2, x.z, x.w);
}
__device__ __forceinline__ float4 load_act4(const bf16* p, size_t lo_off, size_t i) {
  float2 a = act_load2(p, lo_off, i), b = act_load2(p, lo_off, i + 2);
  return make_float4(a.x, a.y, b.x, b.y);
}

// ---------------------------------------------------------------------------------------------------------
// forward: y = LN(x) as act, stats saved
// ---------------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(kLnWarps * 32)
ln_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
              bf16* __restrict__ y, size_t y_lo, float* __restrict__ mean_out, float* __restrict__ rstd_out, int M,
              float eps) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int D = NV * 128;
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * kLnWarps;
  float4 g[NV], b[NV];
  load_row<NV>(gamma, lane, g);
  load_row<NV>(beta, lane, b);
  for (int r = warp; r < M; r += nwarps) {
    float4 v[NV];
    load_row<NV>(x + size_t(r) * D, lane, v);
    float mean, rstd;
    row_stats<NV>(v, mean, rstd, eps);
    if (lane == 0) {
      mean_out[r] = mean;
      rstd_out[r] = rstd;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 o;
      o.x = (v[i].x - mean) * rstd * g[i].x + b[i].x;
      o.y = (v[i].y - mean) * rstd * g[i].y + b[i].y;
      o.z = (v[i].z - mean) * rstd * g[i].z + b[i].z;
      o.w = (v[i].w - mean) * rstd * g[i].w + b[i].w;
      store_act4(y, y_lo, size_t(r) * D + 128 * i + 4 * lane, o);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// forward with masked mean over each sequence of L rows: z[seq] = sum_s valid*LN(x[seq,s]) / sum_s valid
// ---------------------------------------------------------------------------------------------------------
template <int NV>
__global__ void __launch_bounds__(kLnWarps * 32)
ln_pool_fwd_kernel(const float* __restrict__ x, const float* __restrict__ gamma, const float* __restrict__ beta,
                   const uint8_t* __restrict__ valid, float* __restrict__ z, float* __restrict__ mean_out,
                   float* __restrict__ rstd_out, float* __restrict__ inv_cnt, int nseq, int L, float eps) {
  pdl_launch_dependents();
  pdl_wait();
  constexpr int D = NV * 128;
  const int lane = threadIdx.x & 31;
  const int warp = blockIdx.x * kLnWarps + (threadIdx.x >> 5);
  const int nwarps = gridDim.x * kLnWarps;
  float4 g[NV], b[NV];
  load_row<NV>(gamma, lane, g);
  load_row<NV>(beta, lane, b);
  for (int q = warp; q < nseq; q += nwarps) {
    float4 acc[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    int cnt = 0;
    for (int s = 0; s < L; ++s) {
      const int r = q * L + s;
      if (!valid[r]) {
        if (lane == 0) {
          mean_out[r] = 0.f;
          rstd_out[r] = 0.f;
        }
        continue;
      }
      ++cnt;
      float4 v[NV];
      load_row<NV>(x + size_t(r) * D, lane, v);
      float mean, rstd;
      row_stats<NV>(v, mean, rstd, eps);
      if (lane == 0) {
        mean_out[r] = mean;
        rstd_out[r] = rstd;
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        acc[i].x += (v[i].x - mean) * rstd * g[i].x + b[i].x;
        acc[i].y += (v[i].y - mean) * rstd * g[i].y + b[i].y;
        acc[i].z += (v[i].z - mean) * rstd * g[i].z + b[i].z;
        acc[i].w += (v[i].w - mean) * rstd * g[i].w + b[i].w;
      }
    }
    const float ic = 1.f / float(cnt);  // cnt == 0 -> inf -> 0 * inf = NaN, as the reference (SURVEY 8c hazard 3)
    if (lane == 0) inv_cnt[q] = ic;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 o = make_float4(acc[i].x * ic, acc[i].y * ic, acc[i].z * ic, acc[i].w * ic);
      *reinterpret_cast<float4*>(z + size_t(q) * D + 128 * i + 4 * lane) = o;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// backward.  dy comes either from an act tensor (mode 0) or from the pooled gradient dz[row / L] * valid * inv_cnt
// (mode 1).  dx_out = dx_in + LN'(dy); optional act copy of dx_out with a dropout mask applied (the operand of the
// preceding residual branch's backward GEMMs).
// ---------------------------------------------------------------------------------------------------------
struct LnBwdArgs {
  const float* x;
  const float* mean;
  const float* rstd;
  const float* gamma;
  const bf16* dy;
  size_t dy_lo;
  const float* dz;
  const uint8_t* valid;
  const float* inv_cnt;
  int L;
  const float* dx_in;
  float* dx_out;
  bf16* dact;
  size_t dact_lo;
  Dropout drop;
  float* dgamma;
  float* dbeta;
  int M;
};

template <int NV>
__global__ void __launch_bounds__(kLnWarps * 32) ln_bwd_kernel(LnBwdArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  drop_resolve(a.drop);
  constexpr int D = NV * 128;
  constexpr float invD = 1.f / float(D);
  __shared__ float red[kLnWarps][D];
  const int lane = threadIdx.x & 31;
  const int wib = threadIdx.x >> 5;
  const int warp = blockIdx.x * kLnWarps + wib;
  const int nwarps = gridDim.x * kLnWarps;
  float4 g[NV], dg[NV], db[NV];
  load_row<NV>(a.gamma, lane, g);
#pragma unroll
  for (int i = 0; i < NV; ++i) dg[i] = db[i] = make_float4(0.f, 0.f, 0.f, 0.f);

  for (int r = warp; r < a.M; r += nwarps) {
    float4 dxv[NV];
    bool active = true;
    float w = 1.f;
    if (a.dz != nullptr) {
      active = a.valid[r] != 0;
      w = a.inv_cnt[r / a.L];
    }
    if (active) {
      float4 v[NV], dy[NV];
      load_row<NV>(a.x + size_t(r) * D, lane, v);
      if (a.dz != nullptr) {
        load_row<NV>(a.dz + size_t(r / a.L) * D, lane, dy);
#pragma unroll
        for (int i = 0; i < NV; ++i) {
          dy[i].x *= w; dy[i].y *= w; dy[i].z *= w; dy[i].w *= w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) dy[i] = load_act4(a.dy, a.dy_lo, size_t(r) * D + 128 * i + 4 * lane);
      }
      const float mean = a.mean[r], rstd = a.rstd[r];
      float s1 = 0.f, s2 = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float4 xh = make_float4((v[i].x - mean) * rstd, (v[i].y - mean) * rstd, (v[i].z - mean) * rstd,
                                (v[i].w - mean) * rstd);
        float4 t = make_float4(dy[i].x * g[i].x, dy[i].y * g[i].y, dy[i].z * g[i].z, dy[i].w * g[i].w);
        s1 += (t.x + t.y) + (t.z + t.w);
        s2 += (t.x * xh.x + t.y * xh.y) + (t.z * xh.z + t.w * xh.w);
        dg[i].x += dy[i].x * xh.x; dg[i].y += dy[i].y * xh.y; dg[i].z += dy[i].z * xh.z; dg[i].w += dy[i].w * xh.w;
        db[i].x += dy[i].x; db[i].y += dy[i].y; db[i].z += dy[i].z; db[i].w += dy[i].w;
        v[i] = xh;
        dy[i] = t;
      }
      s1 = warp_sum(s1) * invD;
      s2 = warp_sum(s2) * invD;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        dxv[i].x = rstd * (dy[i].x - s1 - v[i].x * s2);
        dxv[i].y = rstd * (dy[i].y - s1 - v[i].y * s2);
        dxv[i].z = rstd * (dy[i].z - s1 - v[i].z * s2);
        dxv[i].w = rstd * (dy[i].w - s1 - v[i].w * s2);
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) dxv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const size_t idx = size_t(r) * D + 128 * i + 4 * lane;
      float4 o = dxv[i];
      if (a.dx_in != nullptr) {
        float4 p = *reinterpret_cast<const float4*>(a.dx_in + idx);
        o.x += p.x; o.y += p.y; o.z += p.z; o.w += p.w;
      }
      if (a.dx_out != nullptr) *reinterpret_cast<float4*>(a.dx_out + idx) = o;
      if (a.dact != nullptr) {
        float4 m = dropout_mult4(a.drop, idx);
        store_act4(a.dact, a.dact_lo, idx, make_float4(o.x * m.x, o.y * m.y, o.z * m.z, o.w * m.w));
      }
    }
  }
  // block reduction of the affine gradients, then one atomic per channel per block
  for (int pass = 0; pass < 2; ++pass) {
    float4* src = pass == 0 ? dg : db;
    float* dst = pass == 0 ? a.dgamma : a.dbeta;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NV; ++i) *reinterpret_cast<float4*>(&red[wib][128 * i + 4 * lane]) = src[i];
    __syncthreads();
    if (dst != nullptr) {
      for (int c = threadIdx.x; c < D; c += kLnWarps * 32) {
        float s = 0.f;
#pragma unroll
        for (int w2 = 0; w2 < kLnWarps; ++w2) s += red[w2][c];
        atomicAdd(dst + c, s);
      }
    }
  }
}

static int ln_grid(int rows) {
  int blocks = (rows + kLnWarps - 1) / kLnWarps;
  int cap = 148 * 8;
  return blocks < cap ? (blocks > 0 ? blocks : 1) : cap;
}

#define DSVG_LN_DISPATCH(D, CALL)                                                           \
  switch ((D) / 128) {                                                                      \
    case 1: { constexpr int NV = 1; CALL; } break;                                          \
    case 2: { constexpr int NV = 2; CALL; } break;                                          \
    case 4: { constexpr int NV = 4; CALL; } break;                                          \
    default: DSVG_CHECK(false, "LayerNorm: d_model %d unsupported (128, 256 or 512)", (D)); \
  }

}  // namespace dsvg
using namespace dsvg;

extern "C" int dsvg_ln_fwd(const float* x, const float* gamma, const float* beta, dsvg_bf16* y, size_t y_lo_off,
                           float* mean, float* rstd, int M, int D, void* stream) {
  DSVG_CHECK(x && gamma && beta && y && mean && rstd && M > 0, "dsvg_ln_fwd: bad arguments");
  DSVG_CHECK(D % 128 == 0, "dsvg_ln_fwd: d_model must be a multiple of 128");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DSVG_LN_DISPATCH(D, DSVG_CUDA(launch_k(ln_fwd_kernel<NV>, dim3(ln_grid(M)), dim3(kLnWarps * 32), 0, st, x, gamma, beta,
                                         reinterpret_cast<bf16*>(y), y_lo_off, mean, rstd, M, 1e-5f)));
  ++g_launches;
  return 0;
}

extern "C" int dsvg_ln_pool_fwd(const float* x, const float* gamma, const float* beta, const uint8_t* valid, float* z,
                                float* mean, float* rstd, float* inv_cnt, int nseq, int L, int D, void* stream) {
  DSVG_CHECK(x && gamma && beta && valid && z && mean && rstd && inv_cnt && nseq > 0 && L > 0,
             "dsvg_ln_pool_fwd: bad arguments");
  DSVG_CHECK(D % 128 == 0, "dsvg_ln_pool_fwd: d_model must be a multiple of 128");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DSVG_LN_DISPATCH(D, DSVG_CUDA(launch_k(ln_pool_fwd_kernel<NV>, dim3(ln_grid(nseq)), dim3(kLnWarps * 32), 0, st, x, gamma,
                                         beta, valid, z, mean, rstd, inv_cnt, nseq, L, 1e-5f)));
  ++g_launches;
  return 0;
}

extern "C" int dsvg_ln_bwd(const float* x, const float* mean, const float* rstd, const float* gamma,
                           const dsvg_bf16* dy, size_t dy_lo_off, const float* dz, const uint8_t* valid,
                           const float* inv_cnt, int L, const float* dx_in, float* dx_out, dsvg_bf16* dact,
                           size_t dact_lo_off, float drop_p, uint32_t drop_site, uint64_t seed, float* dgamma,
                           float* dbeta, int M, int D, void* stream) {
  DSVG_CHECK(x && mean && rstd && gamma && M > 0, "dsvg_ln_bwd: bad arguments");
  DSVG_CHECK((dy != nullptr) != (dz != nullptr), "dsvg_ln_bwd: exactly one of dy / dz must be given");
  DSVG_CHECK(dz == nullptr || (valid && inv_cnt && L > 0), "dsvg_ln_bwd: pooled mode needs valid/inv_cnt/L");
  DSVG_CHECK(dx_out || dact, "dsvg_ln_bwd: no output");
  DSVG_CHECK(D % 128 == 0, "dsvg_ln_bwd: d_model must be a multiple of 128");
  LnBwdArgs a{};
  a.x = x; a.mean = mean; a.rstd = rstd; a.gamma = gamma;
  a.dy = reinterpret_cast<const bf16*>(dy); a.dy_lo = dy_lo_off;
  a.dz = dz; a.valid = valid; a.inv_cnt = inv_cnt; a.L = L > 0 ? L : 1;
  a.dx_in = dx_in; a.dx_out = dx_out;
  a.dact = reinterpret_cast<bf16*>(dact); a.dact_lo = dact_lo_off;
  a.drop = make_dropout(drop_p, drop_site, seed);
  a.dgamma = dgamma; a.dbeta = dbeta; a.M = M;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DSVG_LN_DISPATCH(D, DSVG_CUDA(launch_k(ln_bwd_kernel<NV>, dim3(ln_grid(M)), dim3(kLnWarps * 32), 0, st, a)));
  ++g_launches;
  return 0;
}
