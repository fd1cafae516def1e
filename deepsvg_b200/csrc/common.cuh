// Shared device/host helpers: error reporting across the C ABI, the split-bf16 activation format,
// Philox4x32-10 for in-kernel dropout, warp reductions.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

namespace dsvg {

// ----------------------------------------------------------------------------------------------
// Error handling: no C++ exception crosses the ABI. Entry points return 0 or a non-zero code and leave a
// thread-local message that dsvg_last_error() hands to the host language.
// ----------------------------------------------------------------------------------------------
void set_error(const char* fmt, ...);
const char* last_error();

#define DSVG_CHECK(cond, ...)          \
  do {                                 \
    if (!(cond)) {                     \
      ::dsvg::set_error(__VA_ARGS__);  \
      return 1;                        \
    }                                  \
  } while (0)

#define DSVG_CUDA(expr)                                                                       \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) {                                                                  \
      ::dsvg::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return 2;                                                                               \
    }                                                                                         \
  } while (0)

#define DSVG_LAUNCH_CHECK() DSVG_CUDA(cudaGetLastError())

// ----------------------------------------------------------------------------------------------
// Activation format.  An activation tensor is one bf16 plane ("hi") in fast mode, or two planes in parity mode:
// plane 0 = hi = bf16(v), plane 1 = lo = bf16(v - hi), the second plane `lo_off` elements after the first.
// hi+lo carries ~16 mantissa bits, and the tensor-core GEMMs consume the planes as three products
// (hi*hi + hi*lo + lo*hi) -- "bf16x3".  lo_off == 0 means fast mode.
// ----------------------------------------------------------------------------------------------
using bf16 = __nv_bfloat16;

struct Act {
  bf16* p;
  size_t lo_off;  // 0 => single plane
};
struct CAct {
  const bf16* p;
  size_t lo_off;
};

__device__ __forceinline__ float act_load(const bf16* p, size_t lo_off, size_t i) {
  float v = __bfloat162float(p[i]);
  if (lo_off) v += __bfloat162float(p[i + lo_off]);
  return v;
}
__device__ __forceinline__ void act_store(bf16* p, size_t lo_off, size_t i, float v) {
  bf16 h = __float2bfloat16_rn(v);
  p[i] = h;
  if (lo_off) p[i + lo_off] = __float2bfloat16_rn(v - __bfloat162float(h));
}
// two adjacent elements (i even, pointers 4-byte aligned)
__device__ __forceinline__ float2 act_load2(const bf16* p, size_t lo_off, size_t i) {
  __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(p + i);
  float2 v = __bfloat1622float2(h);
  if (lo_off) {
    __nv_bfloat162 l = *reinterpret_cast<const __nv_bfloat162*>(p + i + lo_off);
    float2 w = __bfloat1622float2(l);
    v.x += w.x;
    v.y += w.y;
  }
  return v;
}
__device__ __forceinline__ void act_store2(bf16* p, size_t lo_off, size_t i, float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  *reinterpret_cast<__nv_bfloat162*>(p + i) = h;
  if (lo_off) {
    float2 hf = __bfloat1622float2(h);
    *reinterpret_cast<__nv_bfloat162*>(p + i + lo_off) = __floats2bfloat162_rn(a - hf.x, b - hf.y);
  }
}

// ----------------------------------------------------------------------------------------------
// Philox4x32-10.  A dropout site is (seed, site id); element i of the site's tensor uses word (i & 3) of
// philox(counter = (i >> 2, site), key = seed).  The backward pass regenerates the same words.
// ----------------------------------------------------------------------------------------------
struct Dropout {
  float p;             // drop probability; 0 disables
  uint32_t thr;        // keep iff word >= thr   (thr = p * 2^32, host-computed)
  float scale;         // 1 / (1 - p)
  uint32_t site;       // call-site id
  unsigned long long seed;
};
inline Dropout make_dropout(float p, uint32_t site, unsigned long long seed) {
  Dropout d;
  d.p = p;
  double t = double(p) * 4294967296.0;
  d.thr = t >= 4294967295.0 ? 0xFFFFFFFFu : (t <= 0.0 ? 0u : uint32_t(t));
  d.scale = p < 1.f ? 1.f / (1.f - p) : 0.f;
  d.site = site;
  d.seed = seed;
  return d;
}

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                               uint32_t k1) {
  constexpr uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, c0), lo0 = M0 * c0;
    uint32_t hi1 = __umulhi(M1, c2), lo1 = M1 * c2;
    uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  return make_uint4(c0, c1, c2, c3);
}
__device__ __forceinline__ uint4 dropout_words(const Dropout& d, unsigned long long idx4) {
  return philox4x32_10(uint32_t(idx4), uint32_t(idx4 >> 32), d.site, 0x5eedu, uint32_t(d.seed),
                       uint32_t(d.seed >> 32));
}
// multiplier (0 or 1/(1-p)) for a single element
__device__ __forceinline__ float dropout_mult(const Dropout& d, unsigned long long idx) {
  if (d.p <= 0.f) return 1.f;
  uint4 w = dropout_words(d, idx >> 2);
  uint32_t r = (idx & 3) == 0 ? w.x : (idx & 3) == 1 ? w.y : (idx & 3) == 2 ? w.z : w.w;
  return r >= d.thr ? d.scale : 0.f;
}
// multipliers for 4 consecutive elements starting at idx (idx % 4 == 0)
__device__ __forceinline__ float4 dropout_mult4(const Dropout& d, unsigned long long idx) {
  if (d.p <= 0.f) return make_float4(1.f, 1.f, 1.f, 1.f);
  uint4 w = dropout_words(d, idx >> 2);
  float s = d.scale;
  return make_float4(w.x >= d.thr ? s : 0.f, w.y >= d.thr ? s : 0.f, w.z >= d.thr ? s : 0.f, w.w >= d.thr ? s : 0.f);
}

// ----------------------------------------------------------------------------------------------
// warp reductions
// ----------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

inline int ceil_div(long long a, long long b) { return int((a + b - 1) / b); }

}  // namespace dsvg
