// tcgen05 / TMA / TMEM contractions for sm_100a.
//
//   dsvg_linear : Y[M,N] = epilogue(X[M,K] . W[N,K]^T)     both operands K-major   (forward + dgrad)
//   dsvg_outer  : C[P,Q] += alpha * A[M,P]^T . B[M,Q]      both operands MN-major  (wgrad, contraction over rows)
//
// Both are warp-specialised: warp 0 = TMA producer (one elected lane), warp 1 = tcgen05.mma issuer (one lane) and
// TMEM owner, warps 2.. = epilogue (8 or 16 warps: TMEM -> registers -> bf16 staging tile + bulk tensor store, or
// shared-memory transpose + coalesced global accesses / vector reductions).
// Operands land in shared memory through TMA with the 128-byte swizzle that the UMMA shared-memory descriptors
// name; accumulators live in TMEM (fp32).  dsvg_linear is persistent (static round-robin tile schedule) with two
// TMEM accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1.
//
// Parity mode ("bf16x3"): operands carry a second bf16 plane (lo = v - bf16(v)); the issuer runs three products
// per K step (hi*hi + hi*lo + lo*hi) into the same accumulator.  Same kernel, NPLANES = 2.
#include <cstdarg>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <unordered_map>

#include "../../include/dsvg_b200.h"
#include "common.cuh"
#include "ptx.cuh"

namespace dsvg {

// ------------------------------------------------------------------------------------------------
// error string + launch counter (shared by every translation unit)
// ------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = {0};
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }
unsigned long long g_launches = 0;
bool pdl_enabled() {
  static const bool on = [] { const char* e = getenv("DSVG_PDL"); return !(e && e[0] == '0'); }();
  return on;
}
static uint32_t g_outer_lbo = 0, g_outer_sbo = 0;  // debug override of the MN-major descriptor strides

// ------------------------------------------------------------------------------------------------
// TMA tensor maps (host).  cuTensorMapEncodeTiled is fetched through the runtime so libcuda is not a link-time
// dependency (the library must load on a CPU-only box for the symbol test).
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

struct MapKey {
  const void* ptr;
  uint64_t d0, d1, stride;
  uint32_t b0, b1;
  bool operator==(const MapKey& o) const {
    return ptr == o.ptr && d0 == o.d0 && d1 == o.d1 && stride == o.stride && b0 == o.b0 && b1 == o.b1;
  }
};
struct MapKeyHash {
  size_t operator()(const MapKey& k) const {
    size_t h = reinterpret_cast<size_t>(k.ptr);
    h ^= k.d0 * 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h ^= k.d1 * 0xC2B2AE3D27D4EB4Full + (h << 6) + (h >> 2);
    h ^= k.stride * 0x165667B19E3779F9ull + (h << 6) + (h >> 2);
    h ^= (uint64_t(k.b0) << 32 | k.b1) + (h << 6) + (h >> 2);
    return h;
  }
};
static std::unordered_map<MapKey, CUtensorMap, MapKeyHash> g_maps;
static std::mutex g_maps_mu;

// 2-D bf16 tensor, dim0 = contiguous (d0 elements), dim1 rows (d1) with `stride` elements between rows;
// box = b0 x b1 elements, 128-byte swizzle (b0 * 2 bytes must be 128).
static int make_map(CUtensorMap* out, const void* ptr, uint64_t d0, uint64_t d1, uint64_t stride, uint32_t b0,
                    uint32_t b1) {
  MapKey key{ptr, d0, d1, stride, b0, b1};
  {
    std::lock_guard<std::mutex> g(g_maps_mu);
    auto it = g_maps.find(key);
    if (it != g_maps.end()) {
      *out = it->second;
      return 0;
    }
  }
  EncodeTiledFn enc = get_encode();
  DSVG_CHECK(enc != nullptr, "cuTensorMapEncodeTiled not available (no CUDA driver?)");
  DSVG_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA operand not 16-byte aligned");
  DSVG_CHECK((stride * 2) % 16 == 0, "TMA operand row stride must be a multiple of 8 elements (got %llu)",
             (unsigned long long)stride);
  cuuint64_t dims[2] = {d0, d1};
  cuuint64_t strides[1] = {stride * 2};
  cuuint32_t box[2] = {b0, b1};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DSVG_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with CUresult %d (dims %llu x %llu, stride %llu)",
             int(r), (unsigned long long)d0, (unsigned long long)d1, (unsigned long long)stride);
  {
    std::lock_guard<std::mutex> g(g_maps_mu);
    if (g_maps.size() > 4096) g_maps.clear();
    g_maps.emplace(key, *out);
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// shared epilogue
// ------------------------------------------------------------------------------------------------
struct Epi {
  const float* acc_scale_dev;  // optional device scalar multiplying the accumulator first
  const float* bias;
  int scale_cols;
  float scale;
  int relu;
  Dropout drop;
  const float* rowvec;
  int rowvec_ld;
  int rows_per_group;
  uint32_t rpg_magic;  // ceil(2^32 / rows_per_group): row / rows_per_group == __umulhi(row, magic) while row * rpg < 2^32
  const bf16* mask;
  size_t mask_lo_off;
  int mask_ld;
  float mask_scale;
  const float* residual;
  int res_ld;
  float* out_f32;
  int out_f32_ld;
  bf16* out_act;
  size_t out_lo_off;
  int out_act_ld;
  int vec;   // 1: every pointer/stride satisfies the 4-wide vector path
  int mode;  // 0: generic run-time epilogue; k > 0: lean epilogue kLeanFeat[k - 1]; 8 / 9: fused LayerNorm (see below)
  // ---- fused LayerNorm (N == BN == 256: a CTA tile owns whole rows) ----
  // mode 8 (forward):  x1 = residual epilogue of mode 4 -> out_f32;  ln_out = bf16(LN(x1) * gamma + beta); stats saved
  // mode 9 (backward): dy = accumulator (dgrad into the LN output); dx_out = LN'(dy; x, mean, rstd, gamma) + dx_in;
  //                    dact = bf16(dropout_mask * dx_out); dgamma / dbeta accumulated
  const float* ln_gamma;
  const float* ln_beta;
  bf16* ln_out;         // mode 8: [M, 256] ; mode 9: dact or null
  float* ln_mean;       // mode 8: written ; mode 9: read
  float* ln_rstd;
  const float* ln_x;    // mode 9: the LayerNorm input (fp32 residual stream), row stride 256
  const float* ln_dx_in;
  float* ln_dx_out;
  float* ln_dgamma;
  float* ln_dbeta;
  long long* dbg;  // development trace (dsvg_debug_linear_trace): per-tile clock64 stamps of CTA 0, or null
};

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;         // 64 bf16 = 128 bytes = one swizzle span
constexpr int kStageRow = 36;       // fp32 staging row stride in words: 16-byte aligned rows, conflict-free v4 access
constexpr int kStageWarpBytes = 32 * kStageRow * 4;

__device__ __forceinline__ void st_shared_v4(uint32_t addr, float a, float b, float c, float d) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
__device__ __forceinline__ float4 ld_shared_v4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr)
               : "memory");
  return v;
}
__device__ __forceinline__ float ld_shared_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}

// scalar epilogue of one element (row, col): everything after the accumulator
__device__ __forceinline__ void epi_scalar(float x, long long row, int col, int N, const Epi& ep, float acc_scale) {
  x *= acc_scale;
  if (ep.bias != nullptr) x += __ldg(ep.bias + col);
  if (col < ep.scale_cols) x *= ep.scale;
  if (ep.relu) x = fmaxf(x, 0.f);
  if (ep.drop.p > 0.f) x *= dropout_mult(ep.drop, (unsigned long long)row * (unsigned long long)N + col);
  if (ep.rowvec != nullptr) x += __ldg(ep.rowvec + (long long)(int(row) / ep.rows_per_group) * ep.rowvec_ld + col);
  if (ep.mask != nullptr) {
    float m = act_load(ep.mask, ep.mask_lo_off, size_t(row) * ep.mask_ld + col);
    x = (m != 0.f) ? x * ep.mask_scale : 0.f;
  }
  if (ep.residual != nullptr) x += ep.residual[row * (long long)ep.res_ld + col];
  if (ep.out_f32 != nullptr) ep.out_f32[row * (long long)ep.out_f32_ld + col] = x;
  if (ep.out_act != nullptr) act_store(ep.out_act, ep.out_lo_off, size_t(row) * ep.out_act_ld + col, x);
}

__device__ __forceinline__ float4 ld_bf16x4(const bf16* p) {
  uint2 u = *reinterpret_cast<const uint2*>(p);
  float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u.x));
  float2 b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void st_act4(bf16* p, size_t lo_off, size_t i, float4 x) {
  __nv_bfloat162 h0 = __floats2bfloat162_rn(x.x, x.y), h1 = __floats2bfloat162_rn(x.z, x.w);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&h0);
  u.y = *reinterpret_cast<uint32_t*>(&h1);
  *reinterpret_cast<uint2*>(p + i) = u;
  if (lo_off) {
    float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
    __nv_bfloat162 l0 = __floats2bfloat162_rn(x.x - f0.x, x.y - f0.y), l1 = __floats2bfloat162_rn(x.z - f1.x, x.w - f1.y);
    u.x = *reinterpret_cast<uint32_t*>(&l0);
    u.y = *reinterpret_cast<uint32_t*>(&l1);
    *reinterpret_cast<uint2*>(p + i + lo_off) = u;
  }
}

// Operands of the vector epilogue that come from global memory; loaded for all 8 row-iterations of a chunk BEFORE any
// store is issued, so the (up to three) DRAM round trips of a chunk overlap instead of serialising per iteration.
struct EpiLoads {
  float4 res, rv;
  uint2 mhi, mlo;
};
__device__ __forceinline__ float4 bf16x4_to_f4(uint2 u) {
  float2 a = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u.x));
  float2 b = __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&u.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void epi_vec4_load(EpiLoads& l, long long row, int col, const Epi& ep) {
  if (ep.rowvec != nullptr)
    l.rv = __ldg(reinterpret_cast<const float4*>(ep.rowvec + (long long)(int(row) / ep.rows_per_group) * ep.rowvec_ld + col));
  if (ep.mask != nullptr) {
    const size_t mi = size_t(row) * ep.mask_ld + col;
    l.mhi = *reinterpret_cast<const uint2*>(ep.mask + mi);
    if (ep.mask_lo_off) l.mlo = *reinterpret_cast<const uint2*>(ep.mask + mi + ep.mask_lo_off);
  }
  if (ep.residual != nullptr) l.res = *reinterpret_cast<const float4*>(ep.residual + row * (long long)ep.res_ld + col);
}
// 4 consecutive columns of one row, all vector accesses aligned (host guarantees ep.vec preconditions)
__device__ __forceinline__ void epi_vec4(float4 x, long long row, int col, int N, const Epi& ep, const float4& bias4,
                                         float acc_scale, const EpiLoads& l) {
  x.x *= acc_scale; x.y *= acc_scale; x.z *= acc_scale; x.w *= acc_scale;
  x.x += bias4.x; x.y += bias4.y; x.z += bias4.z; x.w += bias4.w;
  if (ep.scale_cols > 0) {
    if (col + 0 < ep.scale_cols) x.x *= ep.scale;
    if (col + 1 < ep.scale_cols) x.y *= ep.scale;
    if (col + 2 < ep.scale_cols) x.z *= ep.scale;
    if (col + 3 < ep.scale_cols) x.w *= ep.scale;
  }
  if (ep.relu) {
    x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f);
  }
  if (ep.drop.p > 0.f) {
    float4 m = dropout_mult4(ep.drop, (unsigned long long)row * (unsigned long long)N + col);
    x.x *= m.x; x.y *= m.y; x.z *= m.z; x.w *= m.w;
  }
  if (ep.rowvec != nullptr) {
    x.x += l.rv.x; x.y += l.rv.y; x.z += l.rv.z; x.w += l.rv.w;
  }
  if (ep.mask != nullptr) {
    float4 m = bf16x4_to_f4(l.mhi);
    if (ep.mask_lo_off) {
      float4 lo = bf16x4_to_f4(l.mlo);
      m.x += lo.x; m.y += lo.y; m.z += lo.z; m.w += lo.w;
    }
    x.x = m.x != 0.f ? x.x * ep.mask_scale : 0.f;
    x.y = m.y != 0.f ? x.y * ep.mask_scale : 0.f;
    x.z = m.z != 0.f ? x.z * ep.mask_scale : 0.f;
    x.w = m.w != 0.f ? x.w * ep.mask_scale : 0.f;
  }
  if (ep.residual != nullptr) {
    x.x += l.res.x; x.y += l.res.y; x.z += l.res.z; x.w += l.res.w;
  }
  if (ep.out_f32 != nullptr) *reinterpret_cast<float4*>(ep.out_f32 + row * (long long)ep.out_f32_ld + col) = x;
  if (ep.out_act != nullptr) st_act4(ep.out_act, ep.out_lo_off, size_t(row) * ep.out_act_ld + col, x);
}

// Processes 32 accumulator columns held one-row-per-thread (straight out of tcgen05.ld), transposing them through
// this warp's private shared-memory tile so that every global access is a coalesced row segment.
//   v[j]       : accumulator of row (row0 + lane), column (col0 + j)
//   stage_addr : shared-space byte address of this warp's [32][36] fp32 staging tile
template <bool kOuter>
__device__ __forceinline__ void epilogue_chunk(uint32_t (&v)[32], uint32_t stage_addr, int lane, long long row0, int col0,
                                               int M, int N, const Epi& ep, float alpha, float* C, int ldc) {
  const float acc_scale = (!kOuter && ep.acc_scale_dev != nullptr) ? __ldg(ep.acc_scale_dev) : 1.f;
  if constexpr (!kOuter) {
    if (ep.vec == 2) {
      // direct path: thread = row, eight aligned 4-column groups straight from the TMEM registers (no smem round trip)
      const long long row = row0 + lane;
      if (row < M) {
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          const int col = col0 + 4 * q;
          if (col < N) {
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ep.bias != nullptr) b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + col));
            EpiLoads l;
            epi_vec4_load(l, row, col, ep);
            epi_vec4(make_float4(__uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                                 __uint_as_float(v[4 * q + 3])),
                     row, col, N, ep, b4, acc_scale, l);
          }
        }
      }
      return;
    }
  }
  {
    const uint32_t my = stage_addr + lane * (kStageRow * 4);
#pragma unroll
    for (int q = 0; q < 8; ++q)
      st_shared_v4(my + q * 16, __uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                   __uint_as_float(v[4 * q + 3]));
  }
  __syncwarp();
  if constexpr (kOuter) {
    const int col = col0 + lane;
#pragma unroll 4
    for (int r = 0; r < 32; ++r) {
      long long row = row0 + r;
      float x = ld_shared_f32(stage_addr + (r * kStageRow + lane) * 4);
      if (row < M && col < N) atomicAdd(C + row * (long long)ldc + col, x * alpha);
    }
  } else if (ep.vec) {
    const int cg = lane & 7, rsub = lane >> 3;
    const int col = col0 + 4 * cg;
    const bool col_ok = col < N;
    float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ep.bias != nullptr && col_ok) bias4 = __ldg(reinterpret_cast<const float4*>(ep.bias + col));
#pragma unroll
    for (int h = 0; h < 2; ++h) {   // two groups of four rows: 4 x (residual, mask, rowvec) loads in flight per group
      EpiLoads ld[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const long long row = row0 + 4 * (4 * h + i) + rsub;
        if (row < M && col_ok) epi_vec4_load(ld[i], row, col, ep);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int rl = 4 * (4 * h + i) + rsub;
        float4 x = ld_shared_v4(stage_addr + (rl * kStageRow + 4 * cg) * 4);
        const long long row = row0 + rl;
        if (row < M && col_ok) epi_vec4(x, row, col, N, ep, bias4, acc_scale, ld[i]);
      }
    }
  } else {
    // scalar path (row stride not 16-byte friendly: the fp32 logits).  Lane = column; the column's bias is loaded once.
    const int col = col0 + lane;
    const bool col_ok = col < N;
    const bool simple = ep.rowvec == nullptr && ep.mask == nullptr && ep.residual == nullptr && ep.drop.p <= 0.f &&
                        ep.out_act == nullptr && ep.scale_cols == 0;
    if (simple) {
      const float b = (ep.bias != nullptr && col_ok) ? __ldg(ep.bias + col) : 0.f;
#pragma unroll 8
      for (int r = 0; r < 32; ++r) {
        const long long row = row0 + r;
        float x = ld_shared_f32(stage_addr + (r * kStageRow + lane) * 4) * acc_scale + b;
        if (ep.relu) x = fmaxf(x, 0.f);
        if (row < M && col_ok) ep.out_f32[row * (long long)ep.out_f32_ld + col] = x;
      }
    } else {
#pragma unroll 4
      for (int r = 0; r < 32; ++r) {
        const long long row = row0 + r;
        float x = ld_shared_f32(stage_addr + (r * kStageRow + lane) * 4);
        if (row < M && col_ok) epi_scalar(x, row, col, N, ep, acc_scale);
      }
    }
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// Lean epilogues.  The generic path above decides every step at run time (~100 instructions per 4 outputs, which made
// the 8 epilogue warps issue-bound: ncu smsp__issue_active 40 %, 2.4 k instructions per 32x32 chunk).  The model's
// fast-mode GEMMs use only a handful of step combinations; each is compiled with its steps fixed (FEAT bit mask) and
// pointer arithmetic hoisted out of the row loop.  All of them require the aligned vector layout (ep.vec != 0) and
// single-plane act tensors.
// ------------------------------------------------------------------------------------------------
enum : uint32_t { F_BIAS = 1, F_SCALE = 2, F_RELU = 4, F_DROP = 8, F_ROWVEC = 16, F_MASK = 32, F_RES = 64,
                  F_OUTF = 128, F_OUTA = 256, F_ACCS = 512 };

// global-memory operands of one 32x32 chunk (8 row-iterations of this lane), fetched one chunk ahead of their use
template <uint32_t FEAT>
struct LeanPre {
  float4 res[(FEAT & F_RES) ? 8 : 1];
  uint2 msk[(FEAT & F_MASK) ? 8 : 1];
};
template <uint32_t FEAT>
__device__ __forceinline__ void lean_prefetch(LeanPre<FEAT>& p, int lane, int row0, int col0, int M, int N, const Epi& ep) {
  if constexpr ((FEAT & (F_RES | F_MASK)) != 0) {
    const int cg = lane & 7, rsub = lane >> 3;
    const int col = col0 + 4 * cg;
    if (col < N) {
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int row = row0 + rsub + 4 * it;
        if (row < M) {
          if constexpr (FEAT & F_RES)
            p.res[it] = ep.residual != nullptr ? *reinterpret_cast<const float4*>(ep.residual + size_t(row) * ep.res_ld + col)
                                               : make_float4(0.f, 0.f, 0.f, 0.f);
          if constexpr (FEAT & F_MASK) p.msk[it] = *reinterpret_cast<const uint2*>(ep.mask + size_t(row) * ep.mask_ld + col);
        }
      }
    }
  }
}

template <uint32_t FEAT>
__device__ __forceinline__ void epilogue_chunk_lean(uint32_t (&v)[32], uint32_t stage_addr, int lane, int row0, int col0,
                                                    int M, int N, const Epi& ep, const LeanPre<FEAT>& pre,
                                                    const float4& bias_in) {
  {
    const uint32_t my = stage_addr + lane * (kStageRow * 4);
#pragma unroll
    for (int q = 0; q < 8; ++q)
      st_shared_v4(my + q * 16, __uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                   __uint_as_float(v[4 * q + 3]));
  }
  __syncwarp();
  const int cg = lane & 7, rsub = lane >> 3;
  const int col = col0 + 4 * cg;
  if (col < N) {
    const float4 bias4 = bias_in;   // fetched by the caller at the top of the tile (off the chunk's critical path)
    float accs = 1.f;
    if constexpr (FEAT & F_ACCS) accs = __ldg(ep.acc_scale_dev);
    const int r_first = row0 + rsub;
    float* outf_p = nullptr;
    bf16* outa_p = nullptr;
    if constexpr (FEAT & F_OUTF) outf_p = ep.out_f32 + size_t(r_first) * ep.out_f32_ld + col;
    if constexpr (FEAT & F_OUTA) outa_p = ep.out_act + size_t(r_first) * ep.out_act_ld + col;
    const bool has_rv = (FEAT & F_ROWVEC) && ep.rowvec != nullptr;
    const bool has_drop = (FEAT & F_DROP) && ep.drop.p > 0.f;
    const uint32_t lds_base = stage_addr + (rsub * kStageRow + 4 * cg) * 4;
    // dropout quad index of (r_first, col); N % 4 == 0 and col % 4 == 0 on this path
    const unsigned long long quad0 = ((unsigned long long)r_first * (unsigned long long)N + (unsigned long long)col) >> 2;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const int row = r_first + 4 * it;
      float4 x = ld_shared_v4(lds_base + it * (4 * kStageRow * 4));
      if (row < M) {
        if constexpr (FEAT & F_ACCS) { x.x *= accs; x.y *= accs; x.z *= accs; x.w *= accs; }
        if constexpr (FEAT & F_BIAS) { x.x += bias4.x; x.y += bias4.y; x.z += bias4.z; x.w += bias4.w; }
        if constexpr (FEAT & F_SCALE) {
          if (col < ep.scale_cols) { x.x *= ep.scale; x.y *= ep.scale; x.z *= ep.scale; x.w *= ep.scale; }  // scale_cols % 4 == 0
        }
        if constexpr (FEAT & F_RELU) {
          x.x = fmaxf(x.x, 0.f); x.y = fmaxf(x.y, 0.f); x.z = fmaxf(x.z, 0.f); x.w = fmaxf(x.w, 0.f);
        }
        if constexpr (FEAT & F_DROP) {
          if (has_drop) {
            const unsigned long long quad = quad0 + (unsigned long long)it * (unsigned long long)N;   // row advances by 4
            const float4 m = dropout_quad_mult(ep.drop, uint32_t(quad), drop_hikey(ep.drop, quad));
            x.x *= m.x; x.y *= m.y; x.z *= m.z; x.w *= m.w;
          }
        }
        if constexpr (FEAT & F_ROWVEC) {
          if (has_rv) {
            const uint32_t grp = ep.rpg_magic ? __umulhi(uint32_t(row), ep.rpg_magic) : uint32_t(row / ep.rows_per_group);
            float4 rv = __ldg(reinterpret_cast<const float4*>(ep.rowvec + size_t(grp) * ep.rowvec_ld + col));
            x.x += rv.x; x.y += rv.y; x.z += rv.z; x.w += rv.w;
          }
        }
        if constexpr (FEAT & F_MASK) {
          float4 m = bf16x4_to_f4(pre.msk[it]);
          x.x = m.x != 0.f ? x.x * ep.mask_scale : 0.f;
          x.y = m.y != 0.f ? x.y * ep.mask_scale : 0.f;
          x.z = m.z != 0.f ? x.z * ep.mask_scale : 0.f;
          x.w = m.w != 0.f ? x.w * ep.mask_scale : 0.f;
        }
        if constexpr (FEAT & F_RES) {
          x.x += pre.res[it].x; x.y += pre.res[it].y; x.z += pre.res[it].z; x.w += pre.res[it].w;
        }
        if constexpr (FEAT & F_OUTF) *reinterpret_cast<float4*>(outf_p + size_t(4 * it) * ep.out_f32_ld) = x;
        if constexpr (FEAT & F_OUTA) {
          __nv_bfloat162 h0 = __floats2bfloat162_rn(x.x, x.y), h1 = __floats2bfloat162_rn(x.z, x.w);
          uint2 u;
          u.x = *reinterpret_cast<uint32_t*>(&h0);
          u.y = *reinterpret_cast<uint32_t*>(&h1);
          *reinterpret_cast<uint2*>(outa_p + size_t(4 * it) * ep.out_act_ld) = u;
          if (ep.out_lo_off != 0) {   // parity mode: lo plane = bf16(v - hi)  (warp-uniform branch)
            const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
            h0 = __floats2bfloat162_rn(x.x - f0.x, x.y - f0.y);
            h1 = __floats2bfloat162_rn(x.z - f1.x, x.w - f1.y);
            u.x = *reinterpret_cast<uint32_t*>(&h0);
            u.y = *reinterpret_cast<uint32_t*>(&h1);
            *reinterpret_cast<uint2*>(outa_p + ep.out_lo_off + size_t(4 * it) * ep.out_act_ld) = u;
          }
        }
      }
    }
  }
  __syncwarp();
}

// sum over the 8 lanes that share a row group (lane bits 0..2 = column group)
__device__ __forceinline__ float sum_cg(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  v += __shfl_xor_sync(0xffffffffu, v, 2);
  v += __shfl_xor_sync(0xffffffffu, v, 4);
  return v;
}
__device__ __forceinline__ void st_bf16x4(bf16* p, float a, float b, float c, float d) {
  __nv_bfloat162 h0 = __floats2bfloat162_rn(a, b), h1 = __floats2bfloat162_rn(c, d);
  uint2 u;
  u.x = *reinterpret_cast<uint32_t*>(&h0);
  u.y = *reinterpret_cast<uint32_t*>(&h1);
  *reinterpret_cast<uint2*>(p) = u;
}

// feature sets with a compiled lean epilogue (index = Epi::mode - 1)
constexpr uint32_t kLeanFeat[] = {
    F_OUTA,                                   // 1: plain dgrad
    F_BIAS | F_SCALE | F_OUTA,                // 2: QKV projection
    F_BIAS | F_RELU | F_DROP | F_OUTA,        // 3: FFN first linear
    F_BIAS | F_DROP | F_ROWVEC | F_RES | F_OUTF,  // 4: out-proj / FFN second linear into the fp32 residual stream
    F_MASK | F_OUTA,                          // 5: dgrad through ReLU (+dropout) mask
    F_ACCS | F_RES | F_OUTF,                  // 6: head dgrads accumulated in fp32
    F_BIAS | F_OUTF,                          // 7: fp32 rows of ANY alignment (the 2827-wide logits, N = 7 / 2 heads);
                                              //    chosen by pick_mode for ep.vec == 0, never by the feature search
};
constexpr int kNumLean = 6;                   // modes found by the feature search in pick_mode

// mode 7: one 32 x 32 chunk -> out_f32[row, col] = acc + bias[col]; lane = column, so every store instruction writes one
// 128-byte row segment whatever the row stride (no 16-byte alignment needed).
__device__ __forceinline__ void plain_rows_chunk(uint32_t (&v)[32], uint32_t stage_addr, int lane, long long row0, int col0,
                                                 int M, int N, const Epi& ep) {
  const uint32_t my = stage_addr + lane * (kStageRow * 4);
#pragma unroll
  for (int q = 0; q < 8; ++q)
    st_shared_v4(my + q * 16, __uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                 __uint_as_float(v[4 * q + 3]));
  __syncwarp();
  const int col = col0 + lane;
  if (col < N) {
    const float b = ep.bias != nullptr ? __ldg(ep.bias + col) : 0.f;
    float* out = ep.out_f32 + row0 * (long long)ep.out_f32_ld + col;
    const int rows = (M - row0) < 32 ? int(M - row0) : 32;
    if (rows == 32) {
#pragma unroll
      for (int r = 0; r < 32; ++r) out[(long long)r * ep.out_f32_ld] = ld_shared_f32(stage_addr + (r * kStageRow + lane) * 4) + b;
    } else {
      for (int r = 0; r < rows; ++r) out[(long long)r * ep.out_f32_ld] = ld_shared_f32(stage_addr + (r * kStageRow + lane) * 4) + b;
    }
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// TMA-store epilogue (act outputs): thread = accumulator row, 32 columns straight from TMEM; every step runs in
// registers, the bf16 result goes into a 128-byte-swizzled [128 x BN] tile in shared memory and leaves the SM as one
// bulk tensor store per 64-column box.  No shared-memory read-back, no per-thread global stores.
// ------------------------------------------------------------------------------------------------
template <uint32_t FEAT>
__device__ __forceinline__ void tma_out_chunk(uint32_t (&v)[32], uint8_t* out_tile, int r, long long grow, int gc0, int c,
                                              int M, int N, const Epi& ep, const uint4 (&mk)[4], uint32_t bias_sm) {
  float x[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) x[j] = __uint_as_float(v[j]);
  if constexpr (FEAT & F_BIAS) {
    // bias_sm holds bias[n0 .. n0 + BN) (zero past N), staged once per n-tile by the epilogue warps: broadcast LDS.128
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 b = ld_shared_v4(bias_sm + (c + 4 * q) * 4);
      x[4 * q] += b.x; x[4 * q + 1] += b.y; x[4 * q + 2] += b.z; x[4 * q + 3] += b.w;
    }
  }
  if constexpr (FEAT & F_SCALE) {
    if (gc0 < ep.scale_cols) {   // scale_cols % 32 == 0 (pick_mode): the whole chunk is scaled or none of it
#pragma unroll
      for (int j = 0; j < 32; ++j) x[j] *= ep.scale;
    }
  }
  if constexpr (FEAT & F_RELU) {
#pragma unroll
    for (int j = 0; j < 32; ++j) x[j] = fmaxf(x[j], 0.f);
  }
  if constexpr (FEAT & F_DROP) {
    if (ep.drop.p > 0.f) {
      // 32 consecutive elements starting at a multiple of 8 (N % 8 == 0, gc0 % 32 == 0): 8 quads whose low index
      // word cannot carry, so key' is computed once
      const unsigned long long qb = ((unsigned long long)grow * (unsigned long long)N + (unsigned long long)gc0) >> 2;
      const uint32_t hk = drop_hikey(ep.drop, qb), q0 = uint32_t(qb), thr = ep.drop.thr16;
      const float sc = ep.drop.scale;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t s1 = drop_stage1(q0 + q, hk);
        const uint32_t a = drop_fin_a(s1), b = drop_fin_b(s1);
        x[4 * q] = drop_keep_lo(a, thr) ? x[4 * q] * sc : 0.f;
        x[4 * q + 1] = drop_keep_hi(a, thr) ? x[4 * q + 1] * sc : 0.f;
        x[4 * q + 2] = drop_keep_lo(b, thr) ? x[4 * q + 2] * sc : 0.f;
        x[4 * q + 3] = drop_keep_hi(b, thr) ? x[4 * q + 3] * sc : 0.f;
      }
    }
  }
  if constexpr (FEAT & F_MASK) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const uint32_t w[4] = {mk[q].x, mk[q].y, mk[q].z, mk[q].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        // a bf16 is non-zero iff any of its 15 magnitude bits is set
        x[8 * q + 2 * e] = (w[e] & 0x7FFFu) ? x[8 * q + 2 * e] * ep.mask_scale : 0.f;
        x[8 * q + 2 * e + 1] = (w[e] & 0x7FFF0000u) ? x[8 * q + 2 * e + 1] * ep.mask_scale : 0.f;
      }
    }
  }
  // pack and store: 4 x 16 bytes at swizzled positions of row r in the 64-column box (c / 64)
  uint8_t* box = out_tile + (c >> 6) * (kBlockM * 128) + r * 128;
  const int j0 = (c & 63) >> 3;
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint4 u;
    __nv_bfloat162 t;
    t = __floats2bfloat162_rn(x[8 * q], x[8 * q + 1]); u.x = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(x[8 * q + 2], x[8 * q + 3]); u.y = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(x[8 * q + 4], x[8 * q + 5]); u.z = *reinterpret_cast<uint32_t*>(&t);
    t = __floats2bfloat162_rn(x[8 * q + 6], x[8 * q + 7]); u.w = *reinterpret_cast<uint32_t*>(&t);
    *reinterpret_cast<uint4*>(box + (((j0 + q) ^ (r & 7)) << 4)) = u;
  }
}
template <uint32_t FEAT>
__device__ __forceinline__ void tma_out_load_mask(uint4 (&mk)[4], long long grow, int gc0, int M, int N, const Epi& ep) {
  if constexpr (FEAT & F_MASK) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      mk[q] = make_uint4(0, 0, 0, 0);
      if (grow < M && gc0 + 8 * q < N)
        mk[q] = *reinterpret_cast<const uint4*>(ep.mask + size_t(grow) * ep.mask_ld + gc0 + 8 * q);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// dsvg_linear kernel
// ------------------------------------------------------------------------------------------------
// warp 0 TMA, warp 1 MMA, then E epilogue warps (E/4 per TMEM lane quarter).
__host__ __device__ constexpr int lin_epi_warps(int mode, int bn) {
  // the fp32-residual epilogues (modes 4, 6) are stall-bound chains of shared / global accesses with no single hot
  // spot (ncu: issue slots 29 % busy with 2 warps per scheduler): the 256-wide, one-CTA-per-SM kernel runs them 16 wide
  // mode 6 (head dgrads, K = 2827: 45 k-blocks per tile, a light epilogue): 8 warps leave room for a third operand stage
  return ((mode >= 3 && mode <= 9 && mode != 6) && bn == 256) ? 16 : 8;
}
// act-output lean modes write their bf16 tile through shared memory with one TMA store per 64-column box
// (single-plane tensors only: the two-plane parity mode sends the same feature sets through the per-warp staged epilogue,
// which writes the hi and lo planes with plain vector stores)
__host__ __device__ constexpr bool lin_tma_out(int mode, int nplanes = 1) {
  return nplanes == 1 && (mode == 1 || mode == 2 || mode == 3 || mode == 5);
}

__device__ __forceinline__ void trace_stamp(const Epi& ep, int it, int slot, int lane) {
  if (ep.dbg != nullptr && blockIdx.x == 0 && lane == 0 && it < 16) ep.dbg[it * 16 + slot] = clock64();
}

template <int BN, int NPLANES, int MODE = 0>
struct LinearCfg {
  static constexpr int kEpiWarps = lin_epi_warps(MODE, BN);
  static constexpr int kThreads = 64 + 32 * kEpiWarps;
  static constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KB
  static constexpr int kBBytes = BN * kBlockK * 2;       // 16/32 KB
  static constexpr int kStageBytes = NPLANES * (kABytes + kBBytes);
  static constexpr int kStagingBytes = lin_tma_out(MODE, NPLANES) ? BN * kBlockM * 2 : kEpiWarps * kStageWarpBytes;
  // BN = 128 single-plane kernels are sized so that TWO CTAs fit one SM (2 x 256 TMEM columns, <= 113 KB of shared
  // memory and <= 102 registers each): twice the TMA loads in flight and twice the epilogue warps per SM.
  static constexpr int kCtasPerSm = (BN == 128 && NPLANES == 1) ? 2 : 1;
  static constexpr int kBudget = (kCtasPerSm == 2 ? 110 : 212) * 1024;
  static constexpr int kStages = (kBudget - kStagingBytes) / kStageBytes > 4 ? 4 : (kBudget - kStagingBytes) / kStageBytes;
  static_assert(kStages >= 2, "linear: at least two pipeline stages must fit");
  static constexpr int kTmemCols = 2 * BN;  // two accumulator stages (256 or 512: powers of two)
  static constexpr int kBiasBytes = BN * 4;   // bias slice of the current n-tile (TMA-out epilogues)
  // fused LayerNorm: row partials [2 tile parities][128 rows][4 column quarters][2] + per-CTA dgamma / dbeta [2][256]
  static constexpr int kLnBytes = (MODE == 8 || MODE == 9) ? (2 * 128 * 4 * 2 * 4 + 2 * 256 * 4) : 0;
  static constexpr int kSmemBytes = 1024 /*align slack*/ + kStages * kStageBytes + kStagingBytes + 256 + kBiasBytes + kLnBytes;
};

template <int BN, int NPLANES, int MODE>
__global__ void __launch_bounds__((LinearCfg<BN, NPLANES, MODE>::kThreads), (LinearCfg<BN, NPLANES, MODE>::kCtasPerSm))
linear_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
              const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBlo,
              const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmM, int M, int N, int K,
              Epi ep) {
  using Cfg = LinearCfg<BN, NPLANES, MODE>;
  constexpr int kCols = Cfg::kEpiWarps * 8;   // accumulator columns drained per pass of all epilogue warps (64 or 128)
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tiles = smem;
  float* staging = reinterpret_cast<float*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes + Cfg::kStagingBytes);
  uint64_t* full_bar = bars;                       // [kStages]
  uint64_t* empty_bar = bars + Cfg::kStages;       // [kStages]
  uint64_t* tfull_bar = bars + 2 * Cfg::kStages;   // [2]
  uint64_t* tempty_bar = tfull_bar + 2;            // [2]
  uint64_t* mfull_bar = tempty_bar + 2;            // [2] mask boxes of pass 0 / 1 have landed (mode 5)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mfull_bar + 2);
  float* bias_sm = reinterpret_cast<float*>(smem + Cfg::kStages * Cfg::kStageBytes + Cfg::kStagingBytes + 256);
  float* ln_part = bias_sm + BN;              // modes 8 / 9 only (Cfg::kLnBytes)
  float* ln_acc = ln_part + 2 * 128 * 4 * 2;  // mode 9: [2][256] dgamma / dbeta of this CTA

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (NPLANES == 2) {
      tma_prefetch_desc(&tmAlo);
      tma_prefetch_desc(&tmBlo);
    }
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&mfull_bar[a], 1);
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], Cfg::kEpiWarps);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above touched only this CTA's shared memory / TMEM

  const int m_tiles = (M + kBlockM - 1) / kBlockM;
  const int n_tiles = (N + BN - 1) / BN;
  const int num_tiles = m_tiles * n_tiles;
  const int num_kb = (K + kBlockK - 1) / kBlockK;

  if (warp == 0) {
    // =================== TMA producer (warp-uniform control flow, one elected lane issues) ===================
    int stage = 0;
    uint32_t phase = 0;
    int pit = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++pit) {
      const int m0 = (tile / n_tiles) * kBlockM;
      const int n0 = (tile % n_tiles) * BN;
      trace_stamp(ep, pit, 0, lane);
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (kb == num_kb - 1) trace_stamp(ep, pit, 1, lane);
        if (elect_one()) {
          uint8_t* st = tiles + stage * Cfg::kStageBytes;
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
          tma_load_2d(st, &tmA, &full_bar[stage], kb * kBlockK, m0);
          tma_load_2d(st + Cfg::kABytes, &tmB, &full_bar[stage], kb * kBlockK, n0);
          if (NPLANES == 2) {
            tma_load_2d(st + Cfg::kABytes + Cfg::kBBytes, &tmAlo, &full_bar[stage], kb * kBlockK, m0);
            tma_load_2d(st + 2 * Cfg::kABytes + Cfg::kBBytes, &tmBlo, &full_bar[stage], kb * kBlockK, n0);
          }
        }
        __syncwarp();
        if (++stage == Cfg::kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // =================== MMA issuer (warp-uniform control flow, one elected lane issues) ===================
    constexpr uint32_t idesc = umma_idesc_bf16(kBlockM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      trace_stamp(ep, it, 2, lane);
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
      tc_fence_after();
      trace_stamp(ep, it, 3, lane);
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (kb == 0) trace_stamp(ep, it, 4, lane);
        if (kb == num_kb - 1) trace_stamp(ep, it, 5, lane);
        if (elect_one()) {
          const uint32_t a_hi = smem_u32(tiles + stage * Cfg::kStageBytes);
          const uint32_t b_hi = a_hi + Cfg::kABytes;
          const uint32_t a_lo = b_hi + Cfg::kBBytes;
          const uint32_t b_lo = a_lo + Cfg::kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            umma_f16(d_tmem, umma_smem_desc(a_hi + k * 32, 16, 1024), umma_smem_desc(b_hi + k * 32, 16, 1024), idesc,
                     (kb | k) != 0 ? 1u : 0u);
          if (NPLANES == 2) {
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              umma_f16(d_tmem, umma_smem_desc(a_hi + k * 32, 16, 1024), umma_smem_desc(b_lo + k * 32, 16, 1024), idesc, 1u);
#pragma unroll
            for (int k = 0; k < kBlockK / 16; ++k)
              umma_f16(d_tmem, umma_smem_desc(a_lo + k * 32, 16, 1024), umma_smem_desc(b_hi + k * 32, 16, 1024), idesc, 1u);
          }
          umma_commit(&empty_bar[stage]);                      // frees the smem slot once these MMAs retire
          if (kb == num_kb - 1) umma_commit(&tfull_bar[acc]);  // accumulator complete
        }
        __syncwarp();
        if (kb == num_kb - 1) trace_stamp(ep, it, 6, lane);
        if (++stage == Cfg::kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else {
    // =================== epilogue warps (2 .. 2 + kEpiWarps) ===================
    drop_resolve(ep.drop);
    if constexpr (MODE == 9) {
      for (int j = threadIdx.x - 64; j < 512; j += Cfg::kEpiWarps * 32) ln_acc[j] = 0.f;
      named_bar_sync(2, Cfg::kEpiWarps * 32);
    }
    const int quarter = warp & 3;         // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;     // which 32-column chunk of every kCols-wide pass this warp drains
    const uint32_t stage_buf = smem_u32(staging) + (warp - 2) * kStageWarpBytes;
    int it = 0;
    int bias_n0 = -1;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int m0 = (tile / n_tiles) * kBlockM;
      const int n0 = (tile % n_tiles) * BN;
      const long long row0 = (long long)m0 + quarter * 32;
      if (warp == 2) trace_stamp(ep, it, 8, lane);
      if constexpr (MODE == 0) {
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = half * 32; c < BN; c += kCols) {
          if (n0 + c >= N) break;
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(acc * BN + c), v);
          tmem_ld_wait();
          epilogue_chunk<false>(v, stage_buf, lane, row0, n0 + c, M, N, ep, 1.f, nullptr, 0);
        }
      } else if constexpr (MODE == 7) {
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        if (warp == 2) trace_stamp(ep, it, 10, lane);
#pragma unroll 1
        for (int c = half * 32; c < BN; c += kCols) {
          if (n0 + c >= N) break;
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(acc * BN + c), v);
          tmem_ld_wait();
          plain_rows_chunk(v, stage_buf, lane, row0, n0 + c, M, N, ep);
        }
      } else if constexpr (lin_tma_out(MODE, NPLANES)) {
        // Streamed TMA-store epilogue.  A "pass" = all epilogue warps draining kCols accumulator columns into their
        // 64-column boxes of the staging tile; each pass ends with ONE named barrier after which an elected thread
        // bulk-stores the pass's boxes as one group.  Box reuse: the boxes of pass ci were last stored kPasses groups
        // ago, so before the barrier of pass ci - 1 the storing thread waits until at most kPasses - 2 groups are still
        // reading shared memory; the TMA engine therefore streams tile t's boxes out while tile t + 1 is drained.
        constexpr uint32_t FEAT = kLeanFeat[MODE > 0 ? MODE - 1 : 0];
        constexpr int kPasses = BN / kCols;
        static_assert(kPasses >= 2, "streamed TMA-store epilogue needs at least two passes per tile");
        uint8_t* out_tile = reinterpret_cast<uint8_t*>(staging);
        const int r = quarter * 32 + lane;
        const long long grow = (long long)m0 + r;
        // Mask dgrad: the ReLU/dropout mask tile arrives by TMA INTO the staging boxes the output is about to overwrite
        // (same 128-byte swizzle, so a thread finds its row's mask at the very positions it will write).  The per-thread
        // global loads it replaces read 64 bytes per lane from 32 different rows: latency-bound at 19 % issue / 40 % DRAM.
        constexpr bool kTmaMask = (FEAT & F_MASK) != 0 && kPasses == 2;
        constexpr int kBoxesPerPass = kCols / 64;
        auto issue_mask = [&](int tile_, int pass) {           // storing thread only
          const int mm0 = (tile_ / n_tiles) * kBlockM, nn0 = (tile_ % n_tiles) * BN;
          uint32_t bytes = 0;
#pragma unroll
          for (int bx = pass * kBoxesPerPass; bx < (pass + 1) * kBoxesPerPass; ++bx)
            if (nn0 + 64 * bx < N) bytes += kBlockM * 128;
          mbar_arrive_expect_tx(&mfull_bar[pass], bytes);
#pragma unroll
          for (int bx = pass * kBoxesPerPass; bx < (pass + 1) * kBoxesPerPass; ++bx)
            if (nn0 + 64 * bx < N) tma_load_2d(out_tile + bx * (kBlockM * 128), &tmM, &mfull_bar[pass], nn0 + 64 * bx, mm0);
        };
        if constexpr (kTmaMask) {
          if (warp == 2 && lane == 0) {
            if (it == 0) {
              issue_mask(tile, 0);                              // nothing has used the staging tile yet
              issue_mask(tile, 1);
            } else {
              tma_store_wait_read_n<0>();                       // the previous tile's last store has drained its boxes
              issue_mask(tile, 1);                              // (pass 0 of this tile was requested during the previous tile)
            }
          }
        }
        uint4 mk[2][4];
        if constexpr (!kTmaMask) tma_out_load_mask<FEAT>(mk[0], grow, n0 + half * 32, M, N, ep);
        if constexpr (FEAT & F_BIAS) {
          if (n0 != bias_n0) {   // same decision in every epilogue warp (they walk the same tile sequence)
            const int et = threadIdx.x - 64;
            named_bar_sync(2, Cfg::kEpiWarps * 32);        // nobody still reads the previous slice
            for (int j = et; j < BN; j += Cfg::kEpiWarps * 32) bias_sm[j] = (n0 + j < N) ? __ldg(ep.bias + n0 + j) : 0.f;
            named_bar_sync(2, Cfg::kEpiWarps * 32);
            bias_n0 = n0;
          }
        }
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        if (warp == 2) trace_stamp(ep, it, 10, lane);
#pragma unroll
        for (int ci = 0; ci < kPasses; ++ci) {
          const int c = half * 32 + kCols * ci;
          if (n0 + c < N) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(acc * BN + c), v);
            if constexpr (kTmaMask) {
              mbar_wait(&mfull_bar[ci], uint32_t(it & 1));
              const uint8_t* box = out_tile + (c >> 6) * (kBlockM * 128) + r * 128;
              const int j0 = (c & 63) >> 3;
#pragma unroll
              for (int q = 0; q < 4; ++q) mk[ci & 1][q] = *reinterpret_cast<const uint4*>(box + (((j0 + q) ^ (r & 7)) << 4));
            } else {
              if (ci + 1 < kPasses) tma_out_load_mask<FEAT>(mk[(ci + 1) & 1], grow, n0 + c + kCols, M, N, ep);
            }
            tmem_ld_wait();
            tma_out_chunk<FEAT>(v, out_tile, r, grow, n0 + c, c, M, N, ep, mk[ci & 1], smem_u32(bias_sm));
          }
          if (ci == kPasses - 1) {
            tc_fence_before();
            __syncwarp();
            if (warp == 2) trace_stamp(ep, it, 11, lane);
            if (lane == 0) mbar_arrive(&tempty_bar[acc]);   // TMEM stage drained: the MMA warp may start tile it + 2
          }
          fence_proxy_async_smem();                          // generic-proxy smem writes -> visible to the TMA engine
          if (warp == 2 && lane == 0) {
            tma_store_wait_read_n<(kPasses - 2)>();   // next pass's boxes are free again
            if constexpr (kTmaMask) {
              // pass 0's store of this tile has drained: request the next tile's pass-0 mask into those boxes now, a whole
              // pass ahead of its use
              if (ci == 1 && tile + int(gridDim.x) < num_tiles) issue_mask(tile + int(gridDim.x), 0);
            }
          }
          named_bar_sync(1, Cfg::kEpiWarps * 32);
          if (warp == 2 && lane == 0) {
#pragma unroll
            for (int bx = ci * (kCols / 64); bx < (ci + 1) * (kCols / 64); ++bx)
              if (n0 + 64 * bx < N) tma_store_2d(&tmC, out_tile + bx * (kBlockM * 128), n0 + 64 * bx, m0);
            tma_store_commit();
          }
        }
        if (warp == 2) trace_stamp(ep, it, 12, lane);
        continue;
      } else if constexpr (MODE == 8) {
        // ---------------- residual epilogue (as mode 4) + LayerNorm of the finished rows ----------------
        static_assert(BN == 256 && kCols == 128, "fused LayerNorm needs whole rows per CTA tile");
        const int cg = lane & 7, rsub = lane >> 3;
        const int r_first = int(row0) + rsub;
        float* part = ln_part + (it & 1) * (128 * 4 * 2);
        const bool has_rv = ep.rowvec != nullptr, has_drop = ep.drop.p > 0.f, has_res = ep.residual != nullptr;
        float rs[8], rq[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) rs[i] = rq[i] = 0.f;
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
#pragma unroll
        for (int ci = 0; ci < 2; ++ci) {
          const int c = half * 32 + kCols * ci;
          const int col = n0 + c + 4 * cg;
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(acc * BN + c), v);
          const float4 bias4 = ep.bias != nullptr ? __ldg(reinterpret_cast<const float4*>(ep.bias + col)) : make_float4(0.f, 0.f, 0.f, 0.f);
          tmem_ld_wait();
          {
            const uint32_t my = stage_buf + lane * (kStageRow * 4);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              st_shared_v4(my + q * 16, __uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                           __uint_as_float(v[4 * q + 3]));
          }
          __syncwarp();
          const uint32_t lds_base = stage_buf + (rsub * kStageRow + 4 * cg) * 4;
          const unsigned long long quad0 = ((unsigned long long)r_first * 256ull + (unsigned long long)col) >> 2;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {                  // four rows at a time: their residual loads fly together
            float4 res[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int row = r_first + 4 * (4 * hh + k);
              res[k] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (row < M && has_res) res[k] = *reinterpret_cast<const float4*>(ep.residual + size_t(row) * ep.res_ld + col);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int i = 4 * hh + k;
              const int row = r_first + 4 * i;
              float4 x = ld_shared_v4(lds_base + i * (4 * kStageRow * 4));
              if (row < M) {
                x.x += bias4.x; x.y += bias4.y; x.z += bias4.z; x.w += bias4.w;
                if (has_drop) {
                  const unsigned long long quad = quad0 + (unsigned long long)i * 256ull;   // row advances by 4
                  const float4 m = dropout_quad_mult(ep.drop, uint32_t(quad), drop_hikey(ep.drop, quad));
                  x.x *= m.x; x.y *= m.y; x.z *= m.z; x.w *= m.w;
                }
                if (has_rv) {
                  const uint32_t grp = ep.rpg_magic ? __umulhi(uint32_t(row), ep.rpg_magic) : uint32_t(row / ep.rows_per_group);
                  const float4 rv = __ldg(reinterpret_cast<const float4*>(ep.rowvec + size_t(grp) * ep.rowvec_ld + col));
                  x.x += rv.x; x.y += rv.y; x.z += rv.z; x.w += rv.w;
                }
                x.x += res[k].x; x.y += res[k].y; x.z += res[k].z; x.w += res[k].w;
                *reinterpret_cast<float4*>(ep.out_f32 + size_t(row) * ep.out_f32_ld + col) = x;
                rs[i] += (x.x + x.y) + (x.z + x.w);
                rq[i] += (x.x * x.x + x.y * x.y) + (x.z * x.z + x.w * x.w);
              }
            }
          }
          __syncwarp();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty_bar[acc]);     // accumulator consumed: the MMA warp may reuse this TMEM stage
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          rs[i] = sum_cg(rs[i]);
          rq[i] = sum_cg(rq[i]);
        }
        if (cg == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float2* dst = reinterpret_cast<float2*>(part + ((quarter * 32 + rsub + 4 * i) * 4 + half) * 2);
            *dst = make_float2(rs[i], rq[i]);
          }
        }
        named_bar_sync(3 + quarter, 128);                 // the four warps that share these 32 rows
#pragma unroll
        for (int i = 0; i < 8; ++i) {                     // rs <- mean, rq <- rstd
          const float4* src = reinterpret_cast<const float4*>(part + (quarter * 32 + rsub + 4 * i) * 8);
          const float4 p0 = src[0], p1 = src[1];          // (s, q) of column quarters 0, 1 | 2, 3
          const float sm = (p0.x + p0.z) + (p1.x + p1.z), sq = (p0.y + p0.w) + (p1.y + p1.w);
          rs[i] = sm * (1.f / 256.f);
          rq[i] = rsqrtf(fmaxf(sq * (1.f / 256.f) - rs[i] * rs[i], 0.f) + 1e-5f);
        }
        if (half == 0 && cg == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = r_first + 4 * i;
            if (row < M) {
              ep.ln_mean[row] = rs[i];
              ep.ln_rstd[row] = rq[i];
            }
          }
        }
#pragma unroll
        for (int ci = 0; ci < 2; ++ci) {
          const int col = n0 + half * 32 + kCols * ci + 4 * cg;
          const float4 g4 = __ldg(reinterpret_cast<const float4*>(ep.ln_gamma + col));
          const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.ln_beta + col));
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {
            float4 xv[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {                 // this lane's own stores of phase 1 (L2 hits)
              const int row = r_first + 4 * (4 * hh + k);
              if (row < M) xv[k] = *reinterpret_cast<const float4*>(ep.out_f32 + size_t(row) * ep.out_f32_ld + col);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int i = 4 * hh + k;
              const int row = r_first + 4 * i;
              const float sc = rq[i], sh = -rs[i] * rq[i];   // (x - mean) rstd = x sc + sh
              if (row < M)
                st_bf16x4(ep.ln_out + size_t(row) * 256 + col, (xv[k].x * sc + sh) * g4.x + b4.x, (xv[k].y * sc + sh) * g4.y + b4.y,
                          (xv[k].z * sc + sh) * g4.z + b4.z, (xv[k].w * sc + sh) * g4.w + b4.w);
            }
          }
        }
        continue;
      } else if constexpr (MODE == 9) {
        // ---------------- dgrad into a LayerNorm output + the LayerNorm backward of the finished rows ----------------
        static_assert(BN == 256 && kCols == 128, "fused LayerNorm needs whole rows per CTA tile");
        const int cg = lane & 7, rsub = lane >> 3;
        float* part = ln_part + (it & 1) * (128 * 4 * 2);
        float mean[8], rstd[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int row = int(row0) + rsub + 4 * i;
          mean[i] = row < M ? __ldg(ep.ln_mean + row) : 0.f;
          rstd[i] = row < M ? __ldg(ep.ln_rstd + row) : 0.f;
        }
        float4 g4[2];
#pragma unroll
        for (int ci = 0; ci < 2; ++ci)
          g4[ci] = __ldg(reinterpret_cast<const float4*>(ep.ln_gamma + n0 + half * 32 + kCols * ci + 4 * cg));
        float s1[8], s2[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) s1[i] = s2[i] = 0.f;
        float4 xv[8];
        {
          const int col = n0 + half * 32 + 4 * cg;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = int(row0) + rsub + 4 * i;
            if (row < M) xv[i] = *reinterpret_cast<const float4*>(ep.ln_x + size_t(row) * 256 + col);
          }
        }
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        // phase 1: row sums  s1 = sum_j dy_j g_j,  s2 = sum_j dy_j g_j xhat_j;  column sums for dgamma / dbeta
#pragma unroll
        for (int ci = 0; ci < 2; ++ci) {
          const int c = half * 32 + kCols * ci;
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(acc * BN + c), v);
          tmem_ld_wait();
          {
            const uint32_t my = stage_buf + lane * (kStageRow * 4);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              st_shared_v4(my + q * 16, __uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                           __uint_as_float(v[4 * q + 3]));
          }
          __syncwarp();
          const uint32_t lds_base = stage_buf + (rsub * kStageRow + 4 * cg) * 4;
          float4 dgc = make_float4(0.f, 0.f, 0.f, 0.f), dbc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int row = int(row0) + rsub + 4 * i;
            const float4 dy = ld_shared_v4(lds_base + i * (4 * kStageRow * 4));
            if (row < M) {
              const float4 xh = make_float4((xv[i].x - mean[i]) * rstd[i], (xv[i].y - mean[i]) * rstd[i],
                                            (xv[i].z - mean[i]) * rstd[i], (xv[i].w - mean[i]) * rstd[i]);
              const float4 t = make_float4(dy.x * g4[ci].x, dy.y * g4[ci].y, dy.z * g4[ci].z, dy.w * g4[ci].w);
              s1[i] += (t.x + t.y) + (t.z + t.w);
              s2[i] += (t.x * xh.x + t.y * xh.y) + (t.z * xh.z + t.w * xh.w);
              dgc.x += dy.x * xh.x; dgc.y += dy.y * xh.y; dgc.z += dy.z * xh.z; dgc.w += dy.w * xh.w;
              dbc.x += dy.x; dbc.y += dy.y; dbc.z += dy.z; dbc.w += dy.w;
            }
          }
          __syncwarp();
          // the 4 row groups of this warp share the lane's columns: fold them, then one shared-memory atomic per column
          dgc.x += __shfl_xor_sync(0xffffffffu, dgc.x, 8); dgc.y += __shfl_xor_sync(0xffffffffu, dgc.y, 8);
          dgc.z += __shfl_xor_sync(0xffffffffu, dgc.z, 8); dgc.w += __shfl_xor_sync(0xffffffffu, dgc.w, 8);
          dbc.x += __shfl_xor_sync(0xffffffffu, dbc.x, 8); dbc.y += __shfl_xor_sync(0xffffffffu, dbc.y, 8);
          dbc.z += __shfl_xor_sync(0xffffffffu, dbc.z, 8); dbc.w += __shfl_xor_sync(0xffffffffu, dbc.w, 8);
          dgc.x += __shfl_xor_sync(0xffffffffu, dgc.x, 16); dgc.y += __shfl_xor_sync(0xffffffffu, dgc.y, 16);
          dgc.z += __shfl_xor_sync(0xffffffffu, dgc.z, 16); dgc.w += __shfl_xor_sync(0xffffffffu, dgc.w, 16);
          dbc.x += __shfl_xor_sync(0xffffffffu, dbc.x, 16); dbc.y += __shfl_xor_sync(0xffffffffu, dbc.y, 16);
          dbc.z += __shfl_xor_sync(0xffffffffu, dbc.z, 16); dbc.w += __shfl_xor_sync(0xffffffffu, dbc.w, 16);
          if (rsub == 0) {
            float* ag = ln_acc + c + 4 * cg;
            atomicAdd(ag + 0, dgc.x); atomicAdd(ag + 1, dgc.y); atomicAdd(ag + 2, dgc.z); atomicAdd(ag + 3, dgc.w);
            atomicAdd(ag + 256, dbc.x); atomicAdd(ag + 257, dbc.y); atomicAdd(ag + 258, dbc.z); atomicAdd(ag + 259, dbc.w);
          }
          if (ci == 0) {
            const int col = n0 + c + kCols + 4 * cg;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int row = int(row0) + rsub + 4 * i;
              if (row < M) xv[i] = *reinterpret_cast<const float4*>(ep.ln_x + size_t(row) * 256 + col);
            }
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          s1[i] = sum_cg(s1[i]);
          s2[i] = sum_cg(s2[i]);
        }
        if (cg == 0) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            float2* dst = reinterpret_cast<float2*>(part + ((quarter * 32 + rsub + 4 * i) * 4 + half) * 2);
            *dst = make_float2(s1[i], s2[i]);
          }
        }
        named_bar_sync(3 + quarter, 128);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4* src = reinterpret_cast<const float4*>(part + (quarter * 32 + rsub + 4 * i) * 8);
          const float4 p0 = src[0], p1 = src[1];
          const float c1 = ((p0.x + p0.z) + (p1.x + p1.z)) * (1.f / 256.f);
          const float c2 = ((p0.y + p0.w) + (p1.y + p1.w)) * (1.f / 256.f);
          // dx = r (dy g - c1 - xhat c2) = dy (r g) + x B + C   with  B = -r^2 c2,  C = -r c1 + mean r^2 c2
          const float r = rstd[i];
          s2[i] = -r * r * c2;
          s1[i] = -r * c1 - mean[i] * s2[i];
        }
        // phase 2: dx = rstd * (dy g - s1 - xhat s2) + dx_in  (second pass over the accumulator, x re-read from L2)
        const bool has_drop = ep.drop.p > 0.f;
#pragma unroll
        for (int ci = 1; ci >= 0; --ci) {                 // xv still holds the columns of pass 1
          const int c = half * 32 + kCols * ci;
          const int col = n0 + c + 4 * cg;
          if (ci == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int row = int(row0) + rsub + 4 * i;
              if (row < M) xv[i] = *reinterpret_cast<const float4*>(ep.ln_x + size_t(row) * 256 + col);
            }
          }
          uint32_t v[32];
          tmem_ld_32x32(tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(acc * BN + c), v);
          tmem_ld_wait();
          {
            const uint32_t my = stage_buf + lane * (kStageRow * 4);
#pragma unroll
            for (int q = 0; q < 8; ++q)
              st_shared_v4(my + q * 16, __uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                           __uint_as_float(v[4 * q + 3]));
          }
          __syncwarp();
          const uint32_t lds_base = stage_buf + (rsub * kStageRow + 4 * cg) * 4;
          const unsigned long long quad0 =
              ((unsigned long long)(int(row0) + rsub) * 256ull + (unsigned long long)col) >> 2;
#pragma unroll
          for (int hh = 0; hh < 2; ++hh) {                 // four rows at a time: their dx_in loads fly together
            float4 din[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const int row = int(row0) + rsub + 4 * (4 * hh + k);
              din[k] = make_float4(0.f, 0.f, 0.f, 0.f);
              if (row < M && ep.ln_dx_in != nullptr) din[k] = *reinterpret_cast<const float4*>(ep.ln_dx_in + size_t(row) * 256 + col);
            }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int i = 4 * hh + k;
            const int row = int(row0) + rsub + 4 * i;
            const float4 dy = ld_shared_v4(lds_base + i * (4 * kStageRow * 4));
            if (row < M) {
              const float r = rstd[i];
              float4 o;
              o.x = dy.x * (r * g4[ci].x) + xv[i].x * s2[i] + s1[i] + din[k].x;
              o.y = dy.y * (r * g4[ci].y) + xv[i].y * s2[i] + s1[i] + din[k].y;
              o.z = dy.z * (r * g4[ci].z) + xv[i].z * s2[i] + s1[i] + din[k].z;
              o.w = dy.w * (r * g4[ci].w) + xv[i].w * s2[i] + s1[i] + din[k].w;
              if (ep.ln_dx_out != nullptr) *reinterpret_cast<float4*>(ep.ln_dx_out + size_t(row) * 256 + col) = o;
              if (ep.ln_out != nullptr) {
                if (has_drop) {
                  const unsigned long long quad = quad0 + (unsigned long long)i * 256ull;   // row advances by 4: 4 * 256 / 4 quads
                  const float4 mk = dropout_quad_mult(ep.drop, uint32_t(quad), drop_hikey(ep.drop, quad));
                  o.x *= mk.x; o.y *= mk.y; o.z *= mk.z; o.w *= mk.w;
                }
                st_bf16x4(ep.ln_out + size_t(row) * 256 + col, o.x, o.y, o.z, o.w);
              }
            }
          }
          }
          __syncwarp();
        }
      } else {
        constexpr uint32_t FEAT = kLeanFeat[MODE > 0 ? MODE - 1 : 0];
        // residual / mask operands are fetched one chunk ahead: the first chunk's loads fly while the MMAs of this
        // tile are still running, the others while the previous chunk is being written out
        // (the 16-warp configuration has 96 registers per thread: one prefetch buffer, refilled right after its last
        // use; its four warps per scheduler cover the shorter prefetch distance)
        constexpr int kPre = Cfg::kEpiWarps > 8 ? 1 : 2;
        LeanPre<FEAT> pre[kPre];
        lean_prefetch<FEAT>(pre[0], lane, int(row0), n0 + half * 32, M, N, ep);
        // this lane's bias columns of every chunk of the tile: the shared-memory carve-out leaves L1 too small to keep
        // the bias vector resident next to the residual stream, so a per-chunk load paid L2 latency on the critical path
        float4 bias_pre[BN / kCols];
#pragma unroll
        for (int ci = 0; ci < BN / kCols; ++ci) {
          bias_pre[ci] = make_float4(0.f, 0.f, 0.f, 0.f);
          if constexpr (FEAT & F_BIAS) {
            const int bc = n0 + half * 32 + kCols * ci + 4 * (lane & 7);
            if (bc < N && ep.bias != nullptr) bias_pre[ci] = __ldg(reinterpret_cast<const float4*>(ep.bias + bc));
          }
        }
        mbar_wait(&tfull_bar[acc], acc_phase);
        tc_fence_after();
        if (warp == 2) trace_stamp(ep, it, 10, lane);
#pragma unroll
        for (int ci = 0; ci < BN / kCols; ++ci) {
          const int c = half * 32 + kCols * ci;
          if (n0 + c < N) {
            uint32_t v[32];
            tmem_ld_32x32(tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(acc * BN + c), v);
            if (kPre == 2 && ci + 1 < BN / kCols)
              lean_prefetch<FEAT>(pre[(ci + 1) & 1], lane, int(row0), n0 + c + kCols, M, N, ep);
            tmem_ld_wait();
            epilogue_chunk_lean<FEAT>(v, stage_buf, lane, int(row0), n0 + c, M, N, ep, pre[ci & (kPre - 1)], bias_pre[ci]);
            if (kPre == 1 && ci + 1 < BN / kCols) lean_prefetch<FEAT>(pre[0], lane, int(row0), n0 + c + kCols, M, N, ep);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (warp == 2) trace_stamp(ep, it, 11, lane);
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
    if constexpr (lin_tma_out(MODE, NPLANES)) {
      if (warp == 2 && lane == 0) tma_store_wait_all();   // bulk stores must complete before the CTA's smem goes away
    }
    if constexpr (MODE == 9) {
      named_bar_sync(2, Cfg::kEpiWarps * 32);             // every warp's shared-memory atomics have landed
      for (int j = threadIdx.x - 64; j < 512; j += Cfg::kEpiWarps * 32) {
        float* dst = j < 256 ? ep.ln_dgamma : ep.ln_dbeta;
        if (dst != nullptr) atomicAdd(dst + (j & 255), ln_acc[j]);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

// 16-byte vector reduction (REDG.E.ADD.F32x4): four consecutive fp32 gradient entries per instruction
__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// one 32 x 32 accumulator chunk of the weight-gradient tile -> C[row0.., col0..] += alpha * chunk.
// Transposed through shared memory so that 8 lanes cover one 128-byte row segment; vector path needs ldc % 4 == 0,
// Q % 4 == 0 and a 16-byte aligned C.
__device__ __forceinline__ void outer_chunk_vec(uint32_t (&v)[32], uint32_t stage_addr, int lane, int row0, int col0,
                                                int P, int Q, float alpha, float* C, int ldc) {
  const uint32_t my = stage_addr + lane * (kStageRow * 4);
#pragma unroll
  for (int q = 0; q < 8; ++q)
    st_shared_v4(my + q * 16, __uint_as_float(v[4 * q]), __uint_as_float(v[4 * q + 1]), __uint_as_float(v[4 * q + 2]),
                 __uint_as_float(v[4 * q + 3]));
  __syncwarp();
  const int cg = lane & 7, rsub = lane >> 3;
  const int col = col0 + 4 * cg;
  if (col < Q) {
    float* cp = C + size_t(row0 + rsub) * ldc + col;
    const uint32_t lds_base = stage_addr + (rsub * kStageRow + 4 * cg) * 4;
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      const float4 x = ld_shared_v4(lds_base + it * (4 * kStageRow * 4));
      if (row0 + rsub + 4 * it < P) red_add_v4(cp + size_t(4 * it) * ldc, x.x * alpha, x.y * alpha, x.z * alpha, x.w * alpha);
    }
  }
  __syncwarp();
}

// ------------------------------------------------------------------------------------------------
// dsvg_outer kernel (MN-major operands): one output tile [128 x BQ] per CTA, contraction over an M range
// ------------------------------------------------------------------------------------------------
// BP = 256 ("tall" tile, fast mode, BQ = 256): two [128 x 256] accumulators fill all 512 TMEM columns and share every B
// stage, so a CTA ingests 64 KB per 64-row block for 256 x 256 outputs instead of 48 KB for 128 x 256 -- a third less
// through the SM's L2 port, which is what bounds this kernel (profiles/README.md: 604 MB at 10.1 TB/s for 268 MB of
// operands).  No TMEM is left for the ones-MMA: the column sums (bias gradient) are accumulated by the epilogue warps, which
// are idle during the main loop, straight from the A tiles in shared memory.
template <int BQ, int NPLANES, int BP = 128>
struct OuterCfg {
  static_assert(BP == 128 || (BP == 256 && BQ == 256 && NPLANES == 1), "outer: tall tiles are single-plane 256 x 256 only");
  static constexpr int kBoxBytes = 64 * 64 * 2;                 // [64 rows(m) x 64 cols] = 8 KB
  static constexpr int kABytes = (BP / 64) * kBoxBytes;         // BP P columns
  static constexpr int kBBytes = (BQ / 64) * kBoxBytes;         // BQ Q columns
  static constexpr int kStageBytes = NPLANES * (kABytes + kBBytes);
  static constexpr int kStages = (NPLANES == 1) ? (BP == 256 ? 3 : 4) : 2;
  // BP = 128: BQ accumulator columns + 64 for the column-sum (bias) tile, power of 2; BP = 256: two accumulators
  static constexpr int kTmemCols = 2 * BQ;
  static constexpr int kOnesBytes = 4096;   // 16 K-rows x 128 B of bf16 1.0 (B operand of the column-sum MMA) + slack
  // 8 epilogue warps (two per TMEM lane quarter, each draining half of the BQ columns).  Their transposition buffers
  // alias the operand stages: when the accumulator barrier fires every TMA load has landed and every MMA has retired.
  static constexpr int kEpiWarps = 8;
  static constexpr int kThreads = 64 + 32 * kEpiWarps;
  static constexpr int kStagingBytes = 0;
  static_assert(kEpiWarps * kStageWarpBytes <= kStages * kStageBytes, "outer: staging must fit in the operand stages");
  static constexpr int kSmemBytes = 1024 + kStages * kStageBytes + kOnesBytes + kStagingBytes + 256;
};

// The whole CTA program; `tile_id` / `split_id` select the output tile and the M range (the block indices of the single-problem
// kernel, a table lookup in the grouped one).  The tensor maps live in kernel-parameter space (__grid_constant__) of the caller.
template <int BQ, int NPLANES, int BP>
__device__ __forceinline__ void outer_body(const CUtensorMap* tmA_p, const CUtensorMap* tmAlo_p, const CUtensorMap* tmB_p,
                                           const CUtensorMap* tmBlo_p, int M, int P, int Q, int mblk_per_split, float alpha,
                                           const float* alpha_dev, float* C, int ldc, float* colsum_out, uint32_t lbo,
                                           uint32_t sbo, int vec, int tile_id, int split_id) {
  using Cfg = OuterCfg<BQ, NPLANES, BP>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* tiles = smem;
  uint8_t* ones = smem + Cfg::kStages * Cfg::kStageBytes;  // 1024-aligned
  const int q_tiles = (Q + BQ - 1) / BQ;
  const bool do_colsum = colsum_out != nullptr && (tile_id % q_tiles) == 0;
  // tall tiles: the epilogue warps read every A stage too (column sums) and release it together with the MMA commit
  const bool smem_colsum = BP == 256 && do_colsum;
  float* staging = reinterpret_cast<float*>(tiles);   // epilogue only, after the main loop (see OuterCfg)
  uint64_t* bars = reinterpret_cast<uint64_t*>(ones + Cfg::kOnesBytes + Cfg::kStagingBytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + Cfg::kStages;
  uint64_t* tfull_bar = bars + 2 * Cfg::kStages;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tfull_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(tmA_p);
    tma_prefetch_desc(tmB_p);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], smem_colsum ? 1 + Cfg::kEpiWarps : 1);
    }
    mbar_init(tfull_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::kTmemCols);
    tmem_relinquish();
  }
  // bf16 1.0 everywhere: the layout of an all-ones operand is irrelevant, only the descriptor must be valid
  for (int i = threadIdx.x; i < Cfg::kOnesBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(ones)[i] = 0x3F803F80u;
  fence_proxy_async_smem();
  pdl_launch_dependents();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();   // everything above touched only this CTA's shared memory / TMEM

  const int p0 = (tile_id / q_tiles) * BP;
  const int q0 = (tile_id % q_tiles) * BQ;
  const int total_mblk = (M + 63) / 64;
  const int mb_begin = split_id * mblk_per_split;
  const int mb_end = min(total_mblk, mb_begin + mblk_per_split);
  const int num_mb = mb_end - mb_begin;  // >= 1 by construction of the grid

  if (warp == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int mb = mb_begin; mb < mb_end; ++mb) {
      mbar_wait(&empty_bar[stage], phase ^ 1);
      if (elect_one()) {
        uint8_t* st = tiles + stage * Cfg::kStageBytes;
        mbar_arrive_expect_tx(&full_bar[stage], Cfg::kStageBytes);
#pragma unroll
        for (int i = 0; i < BP / 64; ++i) tma_load_2d(st + i * Cfg::kBoxBytes, tmA_p, &full_bar[stage], p0 + 64 * i, mb * 64);
#pragma unroll
        for (int j = 0; j < BQ / 64; ++j)
          tma_load_2d(st + Cfg::kABytes + j * Cfg::kBoxBytes, tmB_p, &full_bar[stage], q0 + 64 * j, mb * 64);
        if (NPLANES == 2) {
          uint8_t* lo = st + Cfg::kABytes + Cfg::kBBytes;
#pragma unroll
          for (int i = 0; i < 2; ++i)
            tma_load_2d(lo + i * Cfg::kBoxBytes, tmAlo_p, &full_bar[stage], p0 + 64 * i, mb * 64);
#pragma unroll
          for (int j = 0; j < BQ / 64; ++j)
            tma_load_2d(lo + Cfg::kABytes + j * Cfg::kBoxBytes, tmBlo_p, &full_bar[stage], q0 + 64 * j, mb * 64);
        }
      }
      __syncwarp();
      if (++stage == Cfg::kStages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma_idesc_bf16(128, BQ, 1, 1);
    constexpr uint32_t idesc1 = umma_idesc_bf16(128, 64, 1, 1);  // N = 64: one full 128-byte swizzle atom of ones per K row
    const uint32_t ones_addr = smem_u32(ones);
    int stage = 0;
    uint32_t phase = 0;
    for (int i = 0; i < num_mb; ++i) {
      mbar_wait(&full_bar[stage], phase);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t a_hi = smem_u32(tiles + stage * Cfg::kStageBytes);
        const uint32_t b_hi = a_hi + Cfg::kABytes;
        const uint32_t a_lo = b_hi + Cfg::kBBytes;
        const uint32_t b_lo = a_lo + Cfg::kABytes;
#pragma unroll
        for (int k = 0; k < 4; ++k) {  // 16 rows (= 2 KB) of the 64-row block per UMMA
          umma_f16(tmem_base, umma_smem_desc(a_hi + k * 2048, lbo, sbo), umma_smem_desc(b_hi + k * 2048, lbo, sbo), idesc,
                   (i | k) != 0 ? 1u : 0u);
          if constexpr (BP == 256)   // second accumulator: P columns [128, 256) of the tile against the same B rows
            umma_f16(tmem_base + BQ, umma_smem_desc(a_hi + 2 * Cfg::kBoxBytes + k * 2048, lbo, sbo),
                     umma_smem_desc(b_hi + k * 2048, lbo, sbo), idesc, (i | k) != 0 ? 1u : 0u);
        }
        if (NPLANES == 2) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base, umma_smem_desc(a_hi + k * 2048, lbo, sbo), umma_smem_desc(b_lo + k * 2048, lbo, sbo), idesc, 1u);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_f16(tmem_base, umma_smem_desc(a_lo + k * 2048, lbo, sbo), umma_smem_desc(b_hi + k * 2048, lbo, sbo), idesc, 1u);
        }
        if (BP == 128 && do_colsum) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            umma_f16(tmem_base + BQ, umma_smem_desc(a_hi + k * 2048, lbo, sbo), umma_smem_desc(ones_addr, lbo, sbo), idesc1,
                     (i | k) != 0 ? 1u : 0u);
            if (NPLANES == 2)
              umma_f16(tmem_base + BQ, umma_smem_desc(a_lo + k * 2048, lbo, sbo), umma_smem_desc(ones_addr, lbo, sbo), idesc1, 1u);
          }
        }
        umma_commit(&empty_bar[stage]);
        if (i == num_mb - 1) umma_commit(tfull_bar);
      }
      __syncwarp();
      if (++stage == Cfg::kStages) {
        stage = 0;
        phase ^= 1;
      }
    }
  } else {
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;      // which half of the BQ accumulator columns this warp drains
    const uint32_t stage_buf = smem_u32(staging) + (warp - 2) * kStageWarpBytes;
    if (alpha_dev != nullptr) alpha *= __ldg(alpha_dev);
    const int ew = warp - 2;
    float cs[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if constexpr (BP == 256) {
      if (smem_colsum) {
        // warp ew owns rows ew, ew + 8, ... of every 64-row block; lane l owns the 16-byte chunk of P columns [8 l, 8 l + 8)
        // (box l >> 3, chunk l & 7, stored at chunk ^ (row & 7) by the 128-byte swizzle; row & 7 == ew for all of this warp's rows)
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t off = uint32_t(lane >> 3) * Cfg::kBoxBytes + uint32_t(ew) * 128u + (uint32_t((lane & 7) ^ ew) << 4);
        for (int i = 0; i < num_mb; ++i) {
          mbar_wait(&full_bar[stage], phase);
          const uint8_t* at = tiles + stage * Cfg::kStageBytes + off;
#pragma unroll
          for (int rr = 0; rr < 8; ++rr) {
            const uint4 u = *reinterpret_cast<const uint4*>(at + rr * 1024);
            const float4 lo4 = bf16x4_to_f4(make_uint2(u.x, u.y)), hi4 = bf16x4_to_f4(make_uint2(u.z, u.w));
            cs[0] += lo4.x; cs[1] += lo4.y; cs[2] += lo4.z; cs[3] += lo4.w;
            cs[4] += hi4.x; cs[5] += hi4.y; cs[6] += hi4.z; cs[7] += hi4.w;
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty_bar[stage]);
          if (++stage == Cfg::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        // the transposition buffers below alias stage 0: no warp may start draining while another still reads an A tile
        named_bar_sync(3, Cfg::kEpiWarps * 32);
      }
    }
    mbar_wait(tfull_bar, 0);
    tc_fence_after();
    Epi dummy{};
#pragma unroll 1
    for (int pa = 0; pa < BP / 128; ++pa) {
      const int pr0 = p0 + 128 * pa + quarter * 32;
      if (p0 + 128 * pa >= P) break;
#pragma unroll 1
      for (int c = half * (BQ / 2); c < (half + 1) * (BQ / 2); c += 32) {
        if (q0 + c >= Q) break;
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(pa * BQ + c), v);
        tmem_ld_wait();
        if (vec) outer_chunk_vec(v, stage_buf, lane, pr0, q0 + c, P, Q, alpha, C, ldc);
        else epilogue_chunk<true>(v, stage_buf, lane, (long long)pr0, q0 + c, P, Q, dummy, alpha, C, ldc);
      }
    }
    if constexpr (BP == 256) {
      if (smem_colsum) {
        // cross-warp reduction of the column sums through the B half of stage 0, past every warp's transposition buffer
        constexpr int kScrOff = (Cfg::kEpiWarps * kStageWarpBytes + 1023) & ~1023;
        static_assert(kScrOff >= Cfg::kABytes && kScrOff + Cfg::kEpiWarps * 256 * 4 <= Cfg::kStageBytes,
                      "outer: column-sum scratch must sit in the B half of stage 0");
        float* scr = reinterpret_cast<float*>(tiles + kScrOff);
        *reinterpret_cast<float4*>(scr + ew * 256 + lane * 8) = make_float4(cs[0], cs[1], cs[2], cs[3]);
        *reinterpret_cast<float4*>(scr + ew * 256 + lane * 8 + 4) = make_float4(cs[4], cs[5], cs[6], cs[7]);
        named_bar_sync(3, Cfg::kEpiWarps * 32);
        const int j = threadIdx.x - 64;
        float sum = 0.f;
#pragma unroll
        for (int w2 = 0; w2 < Cfg::kEpiWarps; ++w2) sum += scr[w2 * 256 + j];
        if (p0 + j < P) atomicAdd(colsum_out + p0 + j, sum * alpha);
      }
    }
    if (BP == 128 && do_colsum && half == 0) {  // column 0 of the [128 x 64] tile = sum over this CTA's rows of A[:, p]
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + (uint32_t(quarter * 32) << 16) + uint32_t(BQ), v);
      tmem_ld_wait();
      const int prow = p0 + quarter * 32 + lane;
      if (prow < P) atomicAdd(colsum_out + prow, __uint_as_float(v[0]) * alpha);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    __syncwarp();
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::kTmemCols);
  }
}

template <int BQ, int NPLANES, int BP = 128>
__global__ void __launch_bounds__((OuterCfg<BQ, NPLANES, BP>::kThreads), 1)
outer_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmAlo,
             const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmBlo, int M, int P, int Q,
             int mblk_per_split, float alpha, const float* alpha_dev, float* C, int ldc, float* colsum_out,
             uint32_t lbo, uint32_t sbo, int vec) {
  outer_body<BQ, NPLANES, BP>(&tmA, &tmAlo, &tmB, &tmBlo, M, P, Q, mblk_per_split, alpha, alpha_dev, C, ldc, colsum_out, lbo, sbo,
                              vec, int(blockIdx.x), int(blockIdx.y));
}

// Grouped launch: the weight gradients of one transformer block (QKV, out-proj, FFN1, FFN2: same row count M, different
// operands) share ONE wave of CTAs.  Launched one by one, each of them splits its M range 24-74 ways to fill the machine and pays
// 19 MB of red.add traffic and a pipeline ramp per launch for 0.25-0.8 MB of result; together they have 8 tall tiles, the M
// range is split ~18 ways, and every CTA streams >100 row blocks.  Tall 256 x 256 tiles throughout (see OuterCfg).
constexpr int kMaxGroup = 4;
struct OuterGroup {
  CUtensorMap a[kMaxGroup];
  CUtensorMap b[kMaxGroup];
  float* C[kMaxGroup];
  float* colsum[kMaxGroup];
  const float* alpha_dev[kMaxGroup];
  float alpha[kMaxGroup];
  int P[kMaxGroup], Q[kMaxGroup], ldc[kMaxGroup], vec[kMaxGroup];
  int cta_begin[kMaxGroup + 1];   // first CTA of each problem (tiles x splits CTAs per problem)
  int n, M, splits, mblk_per_split;
};
__global__ void __launch_bounds__((OuterCfg<256, 1, 256>::kThreads), 1)
outer_group_kernel(const __grid_constant__ OuterGroup g, uint32_t lbo, uint32_t sbo) {
  int p = 0;
  while (p + 1 < g.n && int(blockIdx.x) >= g.cta_begin[p + 1]) ++p;
  const int local = int(blockIdx.x) - g.cta_begin[p];
  outer_body<256, 1, 256>(&g.a[p], &g.a[p], &g.b[p], &g.b[p], g.M, g.P[p], g.Q[p], g.mblk_per_split, g.alpha[p], g.alpha_dev[p],
                          g.C[p], g.ldc[p], g.colsum[p], lbo, sbo, g.vec[p], local / g.splits, local % g.splits);
}

static int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int BN, int NPLANES, int MODE>
static int launch_linear_mode(const CUtensorMap& a, const CUtensorMap& alo, const CUtensorMap& b, const CUtensorMap& blo,
                              int M, int N, int K, const Epi& ep, cudaStream_t st) {
  CUtensorMap c = a, mk = a;
  if (lin_tma_out(MODE, NPLANES)) {
    if (make_map(&c, ep.out_act, N, M, ep.out_act_ld, 64, 128)) return 1;
    if (ep.mask != nullptr && make_map(&mk, ep.mask, N, M, ep.mask_ld, 64, 128)) return 1;
  }
  using Cfg = LinearCfg<BN, NPLANES, MODE>;
  static bool configured[kMaxDevices] = {};
  if (first_use_on_device(configured)) {
    DSVG_CUDA(cudaFuncSetAttribute(linear_kernel<BN, NPLANES, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   Cfg::kSmemBytes));
  }
  const int tiles = ceil_div(M, kBlockM) * ceil_div(N, BN);
  const int slots = sm_count() * Cfg::kCtasPerSm;
  const int grid = tiles < slots ? tiles : slots;
  DSVG_CUDA(launch_k(linear_kernel<BN, NPLANES, MODE>, dim3(grid), dim3(Cfg::kThreads), Cfg::kSmemBytes, st, a, alo, b, blo, c, mk,
                     M, N, K, ep));
  ++g_launches;
  return 0;
}
template <int BN>
static int launch_linear_fast(const CUtensorMap& a, const CUtensorMap& b, int M, int N, int K, const Epi& ep,
                              cudaStream_t st) {
  switch (ep.mode) {
    case 1: return launch_linear_mode<BN, 1, 1>(a, a, b, b, M, N, K, ep, st);
    case 2: return launch_linear_mode<BN, 1, 2>(a, a, b, b, M, N, K, ep, st);
    case 3: return launch_linear_mode<BN, 1, 3>(a, a, b, b, M, N, K, ep, st);
    case 4: return launch_linear_mode<BN, 1, 4>(a, a, b, b, M, N, K, ep, st);
    case 5: return launch_linear_mode<BN, 1, 5>(a, a, b, b, M, N, K, ep, st);
    case 6: return launch_linear_mode<BN, 1, 6>(a, a, b, b, M, N, K, ep, st);
    case 7: return launch_linear_mode<BN, 1, 7>(a, a, b, b, M, N, K, ep, st);
    default: return launch_linear_mode<BN, 1, 0>(a, a, b, b, M, N, K, ep, st);
  }
}

// parity mode (two bf16 planes per operand, 128-wide tiles): the same lean feature sets, staged per-warp epilogue
static int launch_linear_split(const CUtensorMap& a, const CUtensorMap& alo, const CUtensorMap& b, const CUtensorMap& blo, int M,
                               int N, int K, const Epi& ep, cudaStream_t st) {
  switch (ep.mode) {
    case 1: return launch_linear_mode<128, 2, 1>(a, alo, b, blo, M, N, K, ep, st);
    case 2: return launch_linear_mode<128, 2, 2>(a, alo, b, blo, M, N, K, ep, st);
    case 3: return launch_linear_mode<128, 2, 3>(a, alo, b, blo, M, N, K, ep, st);
    case 4: return launch_linear_mode<128, 2, 4>(a, alo, b, blo, M, N, K, ep, st);
    case 5: return launch_linear_mode<128, 2, 5>(a, alo, b, blo, M, N, K, ep, st);
    case 6: return launch_linear_mode<128, 2, 6>(a, alo, b, blo, M, N, K, ep, st);
    case 7: return launch_linear_mode<128, 2, 7>(a, alo, b, blo, M, N, K, ep, st);
    default: return launch_linear_mode<128, 2, 0>(a, alo, b, blo, M, N, K, ep, st);
  }
}

// which lean epilogue (if any) covers exactly the requested steps
static int pick_mode(const Epi& ep, bool split, int N) {
  static const bool off = [] { const char* e = getenv("DSVG_EPI"); return e && e[0] == 'g'; }();  // "generic"
  if (off) return 0;
  if (ep.vec == 0) {   // unaligned rows: only the plain "acc + bias -> fp32" head epilogue has a lean version
    static const bool no7 = [] { const char* e = getenv("DSVG_EPI"); return e && e[0] == '7'; }();  // A/B switch
    const bool plain = ep.out_f32 && !ep.out_act && !ep.acc_scale_dev && ep.scale_cols == 0 && !ep.relu &&
                       !(ep.drop.p > 0.f) && !ep.rowvec && !ep.mask && !ep.residual;
    return (plain && !no7) ? 7 : 0;
  }
  if (ep.vec != 1) return 0;
  if (!split && (ep.mask_lo_off != 0 || ep.out_lo_off != 0)) return 0;   // single-plane operands with two-plane outputs: generic
  uint32_t f = 0;
  if (ep.acc_scale_dev) f |= F_ACCS;
  if (ep.bias) f |= F_BIAS;
  if (ep.scale_cols > 0) f |= F_SCALE;
  if (ep.relu) f |= F_RELU;
  if (ep.drop.p > 0.f) f |= F_DROP;
  if (ep.rowvec) f |= F_ROWVEC;
  if (ep.mask) f |= F_MASK;
  if (ep.residual) f |= F_RES;
  if (ep.out_f32) f |= F_OUTF;
  if (ep.out_act) f |= F_OUTA;
  if ((f & F_SCALE) && ep.scale_cols % 4 != 0) return 0;
  for (int k = 0; k < kNumLean; ++k) {
    const uint32_t have = kLeanFeat[k];
    // steps that the lean code checks at run time (warp-uniform branches): dropout, row vector; and for the fp32 output
    // mode 4 also bias and residual, so that it covers the latent / group-level "global" linears and their dgrads
    const uint32_t optional = have & (F_DROP | F_ROWVEC | ((have & F_OUTF) && !(have & F_ACCS) ? (F_BIAS | F_RES) : 0u));
    if ((f & ~optional) == (have & ~optional) && (f & ~have) == 0) {
      if (lin_tma_out(k + 1, split ? 2 : 1)) {
        const bool ok = N % 8 == 0 && ep.out_act_ld % 8 == 0 && ep.scale_cols % 32 == 0 && (reinterpret_cast<uintptr_t>(ep.out_act) & 15) == 0 &&
                        (!ep.mask || (ep.mask_ld % 8 == 0 && (reinterpret_cast<uintptr_t>(ep.mask) & 15) == 0)) &&
                        (!ep.bias || (reinterpret_cast<uintptr_t>(ep.bias) & 15) == 0);
        if (!ok) return 0;
      }
      return k + 1;
    }
  }
  return 0;
}

template <int BQ, int NPLANES, int BP = 128>
static int launch_outer(const CUtensorMap& a, const CUtensorMap& alo, const CUtensorMap& b, const CUtensorMap& blo,
                        int M, int P, int Q, float alpha, const float* alpha_dev, float* C, int ldc, float* colsum_out,
                        cudaStream_t st) {
  using Cfg = OuterCfg<BQ, NPLANES, BP>;
  static bool configured[kMaxDevices] = {};
  if (first_use_on_device(configured)) {
    DSVG_CUDA((cudaFuncSetAttribute(outer_kernel<BQ, NPLANES, BP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    Cfg::kSmemBytes)));
  }
  const int out_tiles = ceil_div(P, BP) * ceil_div(Q, BQ);
  const int total_mblk = ceil_div(M, 64);
  int splits = sm_count() / out_tiles;  // one CTA per SM fits (shared memory): fill exactly one wave, no ragged tail
  if (splits > total_mblk / 4) splits = total_mblk / 4;       // keep >= 4 blocks of 64 rows per split
  if (splits < 1) splits = 1;
  const int per = ceil_div(total_mblk, splits);
  splits = ceil_div(total_mblk, per);
  dim3 grid(out_tiles, splits);
  DSVG_CUDA(launch_k(outer_kernel<BQ, NPLANES, BP>, grid, dim3(Cfg::kThreads), Cfg::kSmemBytes, st, a, alo, b, blo, M, P, Q, per, alpha,
                     alpha_dev, C, ldc, colsum_out, g_outer_lbo ? g_outer_lbo : uint32_t(Cfg::kBoxBytes),
                     g_outer_sbo ? g_outer_sbo : 1024u,
                     int(ldc % 4 == 0 && Q % 4 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0)));
  ++g_launches;
  return 0;
}

}  // namespace dsvg

using namespace dsvg;

extern "C" const char* dsvg_last_error(void) { return dsvg::last_error(); }
extern "C" int dsvg_abi_version(void) { return 5; }
extern "C" void dsvg_debug_outer_desc(unsigned lbo, unsigned sbo) {
  dsvg::g_outer_lbo = lbo;
  dsvg::g_outer_sbo = sbo;
}
extern "C" unsigned long long dsvg_launch_count(void) { return dsvg::g_launches; }
// development aid (not in the public header): subsequent dsvg_linear launches write per-tile clock stamps of CTA 0
static long long* g_linear_trace = nullptr;
extern "C" void dsvg_debug_linear_trace(long long* dev_buf_256) { g_linear_trace = dev_buf_256; }

static int fill_epi(Epi& ep, const dsvg_epilogue* e, int M, int N) {
  ep.acc_scale_dev = e->acc_scale_dev;
  ep.bias = e->bias;
  ep.scale_cols = e->scale_cols;
  ep.scale = e->scale;
  ep.relu = e->relu;
  ep.drop = make_dropout(e->drop_p, e->drop_site, e->seed);
  ep.rowvec = e->rowvec;
  ep.rowvec_ld = e->rowvec_ld;
  ep.rows_per_group = e->rows_per_group > 0 ? e->rows_per_group : 1;
  ep.rpg_magic = 0;
  if (ep.rows_per_group > 1 && (unsigned long long)M * ep.rows_per_group < (1ull << 32))
    ep.rpg_magic = uint32_t(((1ull << 32) + ep.rows_per_group - 1) / ep.rows_per_group);
  ep.mask = reinterpret_cast<const bf16*>(e->mask);
  ep.mask_lo_off = e->mask_lo_off;
  ep.mask_ld = e->mask_ld;
  ep.mask_scale = e->mask_scale;
  ep.residual = e->residual;
  ep.res_ld = e->res_ld;
  ep.out_f32 = e->out_f32;
  ep.out_f32_ld = e->out_f32_ld;
  ep.out_act = reinterpret_cast<bf16*>(e->out_act);
  ep.out_lo_off = e->out_lo_off;
  ep.out_act_ld = e->out_act_ld;
  ep.dbg = g_linear_trace;
  {
    auto al = [](const void* p, size_t a) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) % a) == 0; };
    bool v = (N % 4 == 0) && al(ep.bias, 16) && al(ep.rowvec, 16) && al(ep.residual, 16) && al(ep.out_f32, 16) &&
             al(ep.mask, 8) && al(ep.out_act, 8);
    v = v && (!ep.rowvec || ep.rowvec_ld % 4 == 0) && (!ep.residual || ep.res_ld % 4 == 0) &&
        (!ep.out_f32 || ep.out_f32_ld % 4 == 0) && (!ep.mask || (ep.mask_ld % 4 == 0 && ep.mask_lo_off % 4 == 0)) &&
        (!ep.out_act || (ep.out_act_ld % 4 == 0 && ep.out_lo_off % 4 == 0));
    static const int direct = [] { const char* e = getenv("DSVG_EPI"); return (e && e[0] == 'd') ? 1 : 0; }();  // staged (transposed) epilogue measured 2.2x faster than direct
    ep.vec = v ? (direct ? 2 : 1) : 0;
  }
  return 0;
}

extern "C" int dsvg_linear(const dsvg_bf16* X, size_t x_lo_off, int lda, const dsvg_bf16* W, size_t w_lo_off, int ldb,
                           int M, int N, int K, const dsvg_epilogue* e, void* stream) {
  DSVG_CHECK(X && W && e, "dsvg_linear: null pointer");
  DSVG_CHECK(M > 0 && N > 0 && K > 0, "dsvg_linear: bad shape %d x %d x %d", M, N, K);
  DSVG_CHECK(lda % 8 == 0 && ldb % 8 == 0, "dsvg_linear: lda/ldb must be multiples of 8 elements (TMA row stride)");
  DSVG_CHECK((x_lo_off == 0) == (w_lo_off == 0), "dsvg_linear: both operands must have the same number of planes");
  Epi ep{};
  if (fill_epi(ep, e, M, N)) return 1;
  DSVG_CHECK(ep.out_f32 || ep.out_act, "dsvg_linear: no output requested");
  const bool split = x_lo_off != 0;
  static const bool force128 = [] { const char* e = getenv("DSVG_BN128"); return e && e[0] == '1'; }();
  // 256-wide tiles for the big path-level GEMMs; the group-level ones (M = N_icons * 8 rows, a few dozen tiles) run the
  // 128-wide kernel: twice the CTAs and two CTAs per SM shorten their latency-bound critical path.  Parity mode always
  // uses the 128-wide tile (shared-memory budget of the split planes).
  const bool wide = (N > 128) && !split && !force128 && M > 16384;
  const uint32_t bn = wide ? 256 : 128;
  CUtensorMap a, alo, b, blo;
  const bf16* Xb = reinterpret_cast<const bf16*>(X);
  const bf16* Wb = reinterpret_cast<const bf16*>(W);
  if (make_map(&a, Xb, K, M, lda, 64, 128)) return 1;
  if (make_map(&b, Wb, K, N, ldb, 64, bn)) return 1;
  alo = a;
  blo = b;
  if (split) {
    if (make_map(&alo, Xb + x_lo_off, K, M, lda, 64, 128)) return 1;
    if (make_map(&blo, Wb + w_lo_off, K, N, ldb, 64, bn)) return 1;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ep.mode = pick_mode(ep, split, N);
  if (split) return launch_linear_split(a, alo, b, blo, M, N, K, ep, st);
  return wide ? launch_linear_fast<256>(a, b, M, N, K, ep, st) : launch_linear_fast<128>(a, b, M, N, K, ep, st);
}

// ---------------------------------------------------------------------------------------------------------------
// GEMM + LayerNorm in one kernel (fast mode, d_model = 256 rows owned by one CTA tile, path-level row counts)
// ---------------------------------------------------------------------------------------------------------------
extern "C" int dsvg_linear_ln_fusable(int M, int N, int n_planes) {
  static const bool off = [] { const char* e = getenv("DSVG_LN_FUSE"); return e && e[0] == '0'; }();
  return (!off && n_planes == 1 && N == 256 && M > 16384) ? 1 : 0;
}

static int ln_common_checks(const dsvg_bf16* X, size_t x_lo_off, int lda, const dsvg_bf16* W, size_t w_lo_off, int ldb, int M,
                            int N, int K) {
  DSVG_CHECK(X && W, "dsvg_linear_ln: null pointer");
  DSVG_CHECK(M > 0 && K > 0, "dsvg_linear_ln: bad shape");
  DSVG_CHECK(lda % 8 == 0 && ldb % 8 == 0, "dsvg_linear_ln: lda/ldb must be multiples of 8 elements");
  DSVG_CHECK(x_lo_off == 0 && w_lo_off == 0 && dsvg_linear_ln_fusable(M, N, 1),
             "dsvg_linear_ln: shape / mode not fusable (ask dsvg_linear_ln_fusable first)");
  return 0;
}

extern "C" int dsvg_linear_ln_fwd(const dsvg_bf16* X, size_t x_lo_off, int lda, const dsvg_bf16* W, size_t w_lo_off, int ldb,
                                  int M, int N, int K, const dsvg_epilogue* e, const float* gamma, const float* beta,
                                  dsvg_bf16* y, float* mean, float* rstd, void* stream) {
  if (ln_common_checks(X, x_lo_off, lda, W, w_lo_off, ldb, M, N, K)) return 1;
  DSVG_CHECK(e && gamma && beta && y && mean && rstd, "dsvg_linear_ln_fwd: null pointer");
  Epi ep{};
  if (fill_epi(ep, e, M, N)) return 1;
  DSVG_CHECK(ep.out_f32 && !ep.out_act && !ep.mask && !ep.relu && ep.scale_cols == 0 && !ep.acc_scale_dev && ep.vec == 1,
             "dsvg_linear_ln_fwd: the epilogue must be the residual-stream form (bias, dropout, row vector, residual -> fp32)");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  DSVG_CHECK(al16(gamma) && al16(beta) && al16(y), "dsvg_linear_ln_fwd: gamma / beta / y must be 16-byte aligned");
  ep.ln_gamma = gamma; ep.ln_beta = beta; ep.ln_out = reinterpret_cast<bf16*>(y); ep.ln_mean = mean; ep.ln_rstd = rstd;
  ep.mode = 8;
  CUtensorMap a, b;
  if (make_map(&a, reinterpret_cast<const bf16*>(X), K, M, lda, 64, 128)) return 1;
  if (make_map(&b, reinterpret_cast<const bf16*>(W), K, N, ldb, 64, 256)) return 1;
  return launch_linear_mode<256, 1, 8>(a, a, b, b, M, N, K, ep, static_cast<cudaStream_t>(stream));
}

extern "C" int dsvg_linear_ln_bwd(const dsvg_bf16* dY, size_t dy_lo_off, int lda, const dsvg_bf16* W, size_t w_lo_off, int ldb,
                                  int M, int N, int K, const float* x, const float* mean, const float* rstd,
                                  const float* gamma, const float* dx_in, float* dx_out, dsvg_bf16* dact, float drop_p,
                                  uint32_t drop_site, uint64_t seed, float* dgamma, float* dbeta, void* stream) {
  if (ln_common_checks(dY, dy_lo_off, lda, W, w_lo_off, ldb, M, N, K)) return 1;
  DSVG_CHECK(x && mean && rstd && gamma && (dx_out || dact), "dsvg_linear_ln_bwd: null pointer");
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  DSVG_CHECK(al16(x) && al16(gamma) && al16(dx_in) && al16(dx_out) && al16(dact),
             "dsvg_linear_ln_bwd: x / gamma / dx_in / dx_out / dact must be 16-byte aligned");
  Epi ep{};
  ep.rows_per_group = 1;
  ep.vec = 1;
  ep.drop = make_dropout(drop_p, drop_site, seed);
  ep.ln_x = x; ep.ln_mean = const_cast<float*>(mean); ep.ln_rstd = const_cast<float*>(rstd); ep.ln_gamma = gamma;
  ep.ln_dx_in = dx_in; ep.ln_dx_out = dx_out; ep.ln_out = reinterpret_cast<bf16*>(dact);
  ep.ln_dgamma = dgamma; ep.ln_dbeta = dbeta;
  ep.mode = 9;
  ep.dbg = nullptr;
  CUtensorMap a, b;
  if (make_map(&a, reinterpret_cast<const bf16*>(dY), K, M, lda, 64, 128)) return 1;
  if (make_map(&b, reinterpret_cast<const bf16*>(W), K, N, ldb, 64, 256)) return 1;
  return launch_linear_mode<256, 1, 9>(a, a, b, b, M, N, K, ep, static_cast<cudaStream_t>(stream));
}

extern "C" int dsvg_outer_group(int n, const dsvg_outer_problem* pr, int M, void* stream) {
  DSVG_CHECK(n >= 1 && n <= kMaxGroup && pr != nullptr && M > 0, "dsvg_outer_group: 1..%d problems", kMaxGroup);
  OuterGroup g{};
  g.n = n;
  g.M = M;
  int tiles_total = 0, tiles[kMaxGroup];
  for (int i = 0; i < n; ++i) {
    const dsvg_outer_problem& q = pr[i];
    DSVG_CHECK(q.A && q.B && q.C && q.P > 0 && q.Q > 0, "dsvg_outer_group: bad problem %d", i);
    DSVG_CHECK(q.lda % 8 == 0 && q.ldb % 8 == 0, "dsvg_outer_group: lda/ldb must be multiples of 8");
    if (make_map(&g.a[i], reinterpret_cast<const bf16*>(q.A), q.P, M, q.lda, 64, 64)) return 1;
    if (make_map(&g.b[i], reinterpret_cast<const bf16*>(q.B), q.Q, M, q.ldb, 64, 64)) return 1;
    g.C[i] = q.C; g.colsum[i] = q.colsum_out; g.alpha_dev[i] = q.alpha_dev; g.alpha[i] = q.alpha;
    g.P[i] = q.P; g.Q[i] = q.Q; g.ldc[i] = q.ldc;
    g.vec[i] = int(q.ldc % 4 == 0 && q.Q % 4 == 0 && (reinterpret_cast<uintptr_t>(q.C) & 15) == 0);
    tiles[i] = ceil_div(q.P, 256) * ceil_div(q.Q, 256);
    tiles_total += tiles[i];
  }
  const int total_mblk = ceil_div(M, 64);
  int splits = sm_count() / tiles_total;
  if (splits > total_mblk / 4) splits = total_mblk / 4;
  if (splits < 1) splits = 1;
  const int per = ceil_div(total_mblk, splits);
  splits = ceil_div(total_mblk, per);
  g.splits = splits;
  g.mblk_per_split = per;
  g.cta_begin[0] = 0;
  for (int i = 0; i < n; ++i) g.cta_begin[i + 1] = g.cta_begin[i] + tiles[i] * splits;
  for (int i = n; i < kMaxGroup; ++i) g.cta_begin[i + 1] = g.cta_begin[n];
  using Cfg = OuterCfg<256, 1, 256>;
  static bool configured[kMaxDevices] = {};
  if (first_use_on_device(configured)) {
    DSVG_CUDA(cudaFuncSetAttribute(outer_group_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  }
  DSVG_CUDA(launch_k(outer_group_kernel, dim3(g.cta_begin[n]), dim3(Cfg::kThreads), Cfg::kSmemBytes,
                     static_cast<cudaStream_t>(stream), g, g_outer_lbo ? g_outer_lbo : uint32_t(Cfg::kBoxBytes),
                     g_outer_sbo ? g_outer_sbo : 1024u));
  ++g_launches;
  return 0;
}

extern "C" int dsvg_outer(const dsvg_bf16* A, size_t a_lo_off, int lda, const dsvg_bf16* B, size_t b_lo_off, int ldb,
                          int M, int P, int Q, float alpha, const float* alpha_dev, float* C, int ldc, float* colsum_out,
                          void* stream) {
  DSVG_CHECK(A && B && C, "dsvg_outer: null pointer");
  DSVG_CHECK(M > 0 && P > 0 && Q > 0, "dsvg_outer: bad shape");
  DSVG_CHECK(lda % 8 == 0 && ldb % 8 == 0, "dsvg_outer: lda/ldb must be multiples of 8");
  DSVG_CHECK((a_lo_off == 0) == (b_lo_off == 0), "dsvg_outer: both operands must have the same number of planes");
  const bool wide = (Q > 128);
  CUtensorMap a, alo, b, blo;
  const bf16* Ab = reinterpret_cast<const bf16*>(A);
  const bf16* Bb = reinterpret_cast<const bf16*>(B);
  // dim0 = feature columns (contiguous), dim1 = M rows; the tensor extents clip (zero-fill) ragged edges
  if (make_map(&a, Ab, P, M, lda, 64, 64)) return 1;
  if (make_map(&b, Bb, Q, M, ldb, 64, 64)) return 1;
  alo = a;
  blo = b;
  const bool split = a_lo_off != 0;
  if (split) {
    if (make_map(&alo, Ab + a_lo_off, P, M, lda, 64, 64)) return 1;
    if (make_map(&blo, Bb + b_lo_off, Q, M, ldb, 64, 64)) return 1;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // tall 256 x 256 tiles for the path-level weight gradients (fast mode): a third less L2 -> SM traffic (see OuterCfg)
  static const bool tall_on = [] { const char* e = getenv("DSVG_OUTER_TALL"); return !(e && e[0] == '0'); }();
  // Measured per shape (tools/bench_outer.py, M = 131072 / 270336): 12 tall tiles 197 -> 155 us (2827 x 256) and 396 -> 291 us
  // (1536 x 512), 4 tiles 142 -> 115 us (512 x 512), 3 tiles equal (768 x 256: 65.5 us), 2 and 1 tiles slower (51 -> 54, 39 -> 45 us):
  // with few tiles the M range is split ~74-148 ways and the doubled reduction epilogue (256 KB of red.add per CTA) outweighs
  // the shorter main loop.
  if (tall_on && wide && !split && M >= 16384 && ceil_div(P, 256) * ceil_div(Q, 256) >= 4)
    return launch_outer<256, 1, 256>(a, alo, b, blo, M, P, Q, alpha, alpha_dev, C, ldc, colsum_out, st);
  if (wide) return split ? launch_outer<256, 2>(a, alo, b, blo, M, P, Q, alpha, alpha_dev, C, ldc, colsum_out, st)
                         : launch_outer<256, 1>(a, alo, b, blo, M, P, Q, alpha, alpha_dev, C, ldc, colsum_out, st);
  return split ? launch_outer<128, 2>(a, alo, b, blo, M, P, Q, alpha, alpha_dev, C, ldc, colsum_out, st)
               : launch_outer<128, 1>(a, alo, b, blo, M, P, Q, alpha, alpha_dev, C, ldc, colsum_out, st);
}
