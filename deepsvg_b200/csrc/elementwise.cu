// Small HBM-bound helpers around the GEMMs: fp32 -> split-bf16 casts (weights once per step, gradients of the tiny
// latent path), column sums for bias gradients, per-sequence sums for the broadcast `linear_global` branch, label
// embedding gather / scatter.
#include "../../include/dsvg_b200.h"
#include "common.cuh"

namespace dsvg {
extern unsigned long long g_launches;

// out[r, c] = in[r, c] * (mask[r, c] != 0 ? mask_scale : 0) * dropout ; columns C..ld_out-1 are zero-filled.
// Optional transposed copy outT[c, r] (ld_t >= R).  32x32 tiles through shared memory keep both writes coalesced.
__global__ void __launch_bounds__(256)
cast_act_kernel(const float* __restrict__ in, int ld_in, int R, int C, bf16* __restrict__ out, size_t out_lo, int ld_out,
                bf16* __restrict__ outT, size_t outT_lo, int ld_t, const bf16* __restrict__ mask, size_t mask_lo,
                int ld_mask, float mask_scale, Dropout drop) {
  pdl_launch_dependents();
  pdl_wait();
  drop_resolve(drop);
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 8 rows per pass
  for (int rr = ty; rr < 32; rr += 8) {
    const int r = r0 + rr, c = c0 + tx;
    float v = 0.f;
    if (r < R && c < C) {
      v = in[size_t(r) * ld_in + c];
      if (mask != nullptr) v = act_load(mask, mask_lo, size_t(r) * ld_mask + c) != 0.f ? v * mask_scale : 0.f;
      if (drop.p > 0.f) v *= dropout_mult(drop, (unsigned long long)r * C + c);
    }
    tile[rr][tx] = v;
    if (out != nullptr && r < R && c < ld_out) act_store(out, out_lo, size_t(r) * ld_out + c, v);
  }
  if (outT != nullptr) {
    __syncthreads();
    for (int cc = ty; cc < 32; cc += 8) {
      const int c = c0 + cc, r = r0 + tx;
      if (c < C && r < ld_t) act_store(outT, outT_lo, size_t(c) * ld_t + r, r < R ? tile[tx][cc] : 0.f);
    }
  }
}

// dst[c] += alpha * sum_r act[r, c]     (bias gradients).  Thread = column, blocks tile (columns x row chunks).
__global__ void __launch_bounds__(256)
colsum_kernel(const bf16* __restrict__ a, size_t lo, int ld, int M, int N, int rows_per_block,
              const float* __restrict__ alpha_dev, float* __restrict__ dst) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(M, r0 + rows_per_block);
  float s = 0.f;
  for (int r = r0; r < r1; ++r) s += act_load(a, lo, size_t(r) * ld + c);
  if (alpha_dev != nullptr) s *= *alpha_dev;
  atomicAdd(dst + c, s);
}

// out[q, c] = dropout( sum_{s < L} in[q*L + s, c] )  as act  (backward of a vector broadcast over a sequence)
__global__ void __launch_bounds__(256)
seg_sum_kernel(const float* __restrict__ in, int nseq, int L, int d, bf16* __restrict__ out, size_t out_lo,
               float* __restrict__ out_f32, Dropout drop) {
  pdl_launch_dependents();
  pdl_wait();
  drop_resolve(drop);
  const size_t n = size_t(nseq) * d;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x) {
    const size_t q = i / d, c = i % d;
    float s = 0.f;
    for (int t = 0; t < L; ++t) s += in[(q * L + t) * d + c];
    if (drop.p > 0.f) s *= dropout_mult(drop, i);
    if (out != nullptr) act_store(out, out_lo, i, s);
    if (out_f32 != nullptr) out_f32[i] = s;
  }
}

// label embedding: out[n] = table[label[n]] (act, and fp32 copy unused) ; backward: dtable[label[n]] += g[n]
// An id outside [0, n_rows) traps (the launch fails) -- nn.Embedding's device-side assert, not a silent read of
// someone else's memory.
__device__ __forceinline__ long long checked_row(const long long* idx, int i, int n_rows) {
  const long long r = idx[i];
  if (r < 0 || r >= n_rows) {
    printf("dsvg: label id %lld at position %d is outside the embedding table (%d rows)\n", r, i, n_rows);
    __trap();
  }
  return r;
}
__global__ void gather_rows_kernel(const float* __restrict__ table, const long long* __restrict__ idx, int n, int w,
                                   int n_rows, bf16* __restrict__ out, size_t out_lo) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * w; i += gridDim.x * blockDim.x)
    act_store(out, out_lo, i, table[size_t(checked_row(idx, i / w, n_rows)) * w + (i % w)]);
}
__global__ void scatter_rows_kernel(const float* __restrict__ g, const long long* __restrict__ idx, int n, int w,
                                    int n_rows, float* __restrict__ dtable) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n * w; i += gridDim.x * blockDim.x)
    atomicAdd(dtable + size_t(checked_row(idx, i / w, n_rows)) * w + (i % w), g[i]);
}

// y = a + b (fp32), used to merge gradient streams of the tiny latent path
__global__ void add_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ y, size_t n) {
  pdl_launch_dependents();
  pdl_wait();
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += size_t(gridDim.x) * blockDim.x)
    y[i] = a[i] + b[i];
}

}  // namespace dsvg
using namespace dsvg;

extern "C" int dsvg_cast_act(const float* in, int ld_in, int R, int C, dsvg_bf16* out, size_t out_lo_off, int ld_out,
                             dsvg_bf16* outT, size_t outT_lo_off, int ld_t, const dsvg_bf16* mask, size_t mask_lo_off,
                             int ld_mask, float mask_scale, float drop_p, uint32_t drop_site, uint64_t seed,
                             void* stream) {
  DSVG_CHECK(in && (out || outT) && R > 0 && C > 0, "dsvg_cast_act: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int cols = out ? (ld_out > C ? ld_out : C) : C;
  const int rows = outT ? (ld_t > R ? ld_t : R) : R;
  dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32));
  DSVG_CUDA(launch_k(cast_act_kernel, grid, dim3(256), 0, st, in, ld_in, R, C, reinterpret_cast<bf16*>(out), out_lo_off, ld_out,
                                        reinterpret_cast<bf16*>(outT), outT_lo_off, ld_t,
                                        reinterpret_cast<const bf16*>(mask), mask_lo_off, ld_mask, mask_scale,
                                        make_dropout(drop_p, drop_site, seed)));
  ++g_launches;
  return 0;
}

extern "C" int dsvg_colsum(const dsvg_bf16* a, size_t lo_off, int ld, int M, int N, const float* alpha_dev, float* dst,
                           void* stream) {
  DSVG_CHECK(a && dst && M > 0 && N > 0, "dsvg_colsum: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int colblocks = ceil_div(N, 256);
  int rowblocks = (148 * 4) / colblocks;
  if (rowblocks < 1) rowblocks = 1;
  int rpb = ceil_div(M, rowblocks);
  if (rpb < 64) rpb = 64;
  rowblocks = ceil_div(M, rpb);
  colsum_kernel<<<dim3(colblocks, rowblocks), 256, 0, st>>>(reinterpret_cast<const bf16*>(a), lo_off, ld, M, N, rpb,
                                                           alpha_dev, dst);
  ++g_launches;
  DSVG_LAUNCH_CHECK();
  return 0;
}

extern "C" int dsvg_seg_sum(const float* in, int nseq, int L, int d, dsvg_bf16* out, size_t out_lo_off, float* out_f32,
                            float drop_p, uint32_t drop_site, uint64_t seed, void* stream) {
  DSVG_CHECK(in && (out || out_f32) && nseq > 0 && L > 0, "dsvg_seg_sum: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int grid = ceil_div((long long)nseq * d, 256);
  if (grid > 148 * 8) grid = 148 * 8;
  DSVG_CUDA(launch_k(seg_sum_kernel, dim3(grid), dim3(256), 0, st, in, nseq, L, d, reinterpret_cast<bf16*>(out), out_lo_off, out_f32,
                                       make_dropout(drop_p, drop_site, seed)));
  ++g_launches;
  return 0;
}

extern "C" int dsvg_gather_rows(const float* table, const long long* idx, int n, int w, int n_rows, dsvg_bf16* out,
                                size_t out_lo_off, void* stream) {
  DSVG_CHECK(table && idx && out && n > 0 && w > 0 && n_rows > 0, "dsvg_gather_rows: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  gather_rows_kernel<<<ceil_div((long long)n * w, 256), 256, 0, st>>>(table, idx, n, w, n_rows, reinterpret_cast<bf16*>(out),
                                                                     out_lo_off);
  ++g_launches;
  DSVG_LAUNCH_CHECK();
  return 0;
}

extern "C" int dsvg_scatter_rows(const float* g, const long long* idx, int n, int w, int n_rows, float* dtable, void* stream) {
  DSVG_CHECK(g && idx && dtable && n > 0 && w > 0 && n_rows > 0, "dsvg_scatter_rows: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  scatter_rows_kernel<<<ceil_div((long long)n * w, 256), 256, 0, st>>>(g, idx, n, w, n_rows, dtable);
  ++g_launches;
  DSVG_LAUNCH_CHECK();
  return 0;
}

extern "C" int dsvg_add_f32(const float* a, const float* b, float* y, size_t n, void* stream) {
  DSVG_CHECK(a && b && y && n > 0, "dsvg_add_f32: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int grid = ceil_div((long long)n, 256);
  if (grid > 148 * 8) grid = 148 * 8;
  DSVG_CUDA(launch_k(add_f32_kernel, dim3(grid), dim3(256), 0, st, a, b, y, n));
  ++g_launches;
  return 0;
}
