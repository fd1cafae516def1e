// Input-side kernels: sequence bookkeeping from the command tensor, the command/argument/position embedding, and the
// constant (positional-only) decoder inputs -- all HBM / L2-bound gather work, coalesced and vectorised.
//
//   reference: masks            model/utils.py:7-66  (_get_key_padding_mask, _get_visibility_mask, _get_group_mask)
//              SVGEmbedding     model/model.py:46-57 (+ PositionalEncodingLUT positional_encoding.py:40-43)
//              ConstEmbedding   model/model.py:70-73
//
// The reference multiplies a (tokens x 704) concatenation of 11 argument embeddings by embed_fcn.weight.  Here the
// product is folded once per step into a table  T[k][v] = arg_embed[v] . W_k^T  (W_k = columns 64k..64k+63): a token's
// embedding is then a sum of 11 table rows.  Because 80 % of argument slots are PAD (-1 -> row 0), rows are stored as
// differences D[k][v] = T[k][v] - T[k][0] with the constant  base = bias + sum_k T[k][0]  added once, so only the
// non-PAD slots are gathered.  Everything is fp32 FMA (exact w.r.t. the fp32 reference up to summation order).
#include "../../include/dsvg_b200.h"
#include "common.cuh"

namespace dsvg {
extern unsigned long long g_launches;

constexpr int CMD_M = 0, CMD_EOS = 4;
__constant__ int c_nargs_of_cmd[7] = {2, 2, 6, 7, 0, 0, 0};  // row sums of CMD_ARGS_MASK (difflib/tensor.py:15-21)

// ---------------------------------------------------------------------------------------------------------
// per-sequence bookkeeping: one warp per sequence of L command ids
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
seq_prep_kernel(const float* __restrict__ commands, int nseq, int L, int* __restrict__ first_eos,
                uint8_t* __restrict__ visible, uint8_t* __restrict__ key_valid, uint8_t* __restrict__ grp,
                float* __restrict__ counts) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= nseq) return;
  const float* c = commands + size_t(q) * L;
  int fe = L, neos = 0, gbase = 0;
  for (int i0 = 0; i0 < L; i0 += 32) {
    const int i = i0 + lane;
    const int cmd = i < L ? int(c[i]) : -1;
    const unsigned eos = __ballot_sync(0xffffffffu, cmd == CMD_EOS);
    const unsigned mm = __ballot_sync(0xffffffffu, cmd == CMD_M);
    if (eos && fe == L) fe = i0 + __ffs(eos) - 1;
    neos += __popc(eos);
    if (i < L && grp != nullptr) grp[size_t(q) * L + i] = uint8_t(gbase + __popc(mm & (0xffffffffu >> (31 - lane))));
    gbase += __popc(mm);
  }
  const bool vis = neos < L - 1;  // model/utils.py:52
  float ccmd = 0.f, cargs = 0.f;
  for (int i0 = 0; i0 < L; i0 += 32) {
    const int i = i0 + lane;
    if (i < L) {
      if (key_valid != nullptr) key_valid[size_t(q) * L + i] = i < fe ? 1 : 0;  // model/utils.py:13,22
      if (i >= 1) {
        const int cmd = int(c[i]);
        cargs += float(c_nargs_of_cmd[cmd]);
        const bool ext = (i < fe) || (i >= 3 && i - 3 < fe);  // clean OR-shift-by-3 of the padding mask
        ccmd += (vis && ext) ? 1.f : 0.f;
      }
    }
  }
  ccmd = warp_sum(ccmd);
  cargs = warp_sum(cargs);
  if (lane == 0) {
    if (first_eos != nullptr) first_eos[q] = fe;
    if (visible != nullptr) visible[q] = vis ? 1 : 0;
    if (counts != nullptr) {
      atomicAdd(counts + 0, ccmd);
      atomicAdd(counts + 1, cargs);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// fold:  T[k*V + v][c] = sum_e arg_embed[v][e] * W[c][64k + e]
// ---------------------------------------------------------------------------------------------------------
// Block = (argument slot k, chunk of table rows); thread = output channel c, holding its 64 weights in registers.
__global__ void __launch_bounds__(512)
fold_kernel(const float* __restrict__ Ea, const float* __restrict__ W, float* __restrict__ T, int V, int n_args, int d,
            int rows_per_block) {
  __shared__ float ea[64];
  const int k = blockIdx.y, c = threadIdx.x;
  float w[64];
  const float4* wp = reinterpret_cast<const float4*>(W + size_t(c) * (64 * n_args) + 64 * k);
#pragma unroll
  for (int q = 0; q < 16; ++q) {
    float4 t = wp[q];
    w[4 * q] = t.x; w[4 * q + 1] = t.y; w[4 * q + 2] = t.z; w[4 * q + 3] = t.w;
  }
  const int v0 = blockIdx.x * rows_per_block, v1 = min(V, v0 + rows_per_block);
  for (int v = v0; v < v1; ++v) {
    __syncthreads();
    if (c < 64) ea[c] = Ea[size_t(v) * 64 + c];
    __syncthreads();
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < 64; ++e) s = fmaf(ea[e], w[e], s);
    T[(size_t(k) * V + v) * d + c] = s;
  }
}
// rows v >= 1 become differences to row 0; block (0,0) also writes base = bias + sum_k T[k][0]
__global__ void fold_sub_kernel(float* __restrict__ T, const float* __restrict__ bias, float* __restrict__ base, int V,
                                int n_args, int d) {
  const int k = blockIdx.y;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    const float t0 = T[size_t(k) * V * d + c];
    for (int v = 1 + blockIdx.x; v < V; v += gridDim.x) T[(size_t(k) * V + v) * d + c] -= t0;
    if (blockIdx.x == 0 && k == 0) {
      float s = bias[c];
      for (int kk = 0; kk < n_args; ++kk) s += T[size_t(kk) * V * d + c];
      base[c] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// embedding forward: one warp per token, lane owns channels {128 i + 4 lane .. +3}
// ---------------------------------------------------------------------------------------------------------
struct EmbedArgs {
  const float* commands;  // [T]
  const float* args;      // [T, n_args]
  const uint8_t* grp;     // [T] or nullptr
  const float* cmd_tab;   // [7, d]
  const float* D;         // [n_args*V, d] difference table
  const float* base;      // [d]
  const float* pos_tab;   // [L, d]
  const float* grp_tab;   // [G+2, d] or nullptr
  float* x;               // [T, d]
  int T, L, V, n_args;
  Dropout drop;
};

template <int NV>
__global__ void __launch_bounds__(256) embed_fwd_kernel(EmbedArgs a) {
  drop_resolve(a.drop);
  constexpr int d = NV * 128;
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= a.T) return;
  const int s = t % a.L;
  float mine = lane < a.n_args ? a.args[size_t(t) * a.n_args + lane] : (lane == 31 ? a.commands[t] : 0.f);
  const int cmd = int(__shfl_sync(0xffffffffu, mine, 31));
  float4 acc[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int c = 128 * i + 4 * lane;
    float4 b = *reinterpret_cast<const float4*>(a.base + c);
    float4 e = *reinterpret_cast<const float4*>(a.cmd_tab + size_t(cmd) * d + c);
    float4 p = *reinterpret_cast<const float4*>(a.pos_tab + size_t(s) * d + c);
    acc[i] = make_float4(b.x + e.x + p.x, b.y + e.y + p.y, b.z + e.z + p.z, b.w + e.w + p.w);
  }
  if (a.grp_tab != nullptr) {
    const int g = a.grp[t];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 e = *reinterpret_cast<const float4*>(a.grp_tab + size_t(g) * d + 128 * i + 4 * lane);
      acc[i].x += e.x; acc[i].y += e.y; acc[i].z += e.z; acc[i].w += e.w;
    }
  }
  for (int k = 0; k < a.n_args; ++k) {
    const int v = int(__shfl_sync(0xffffffffu, mine, k)) + 1;  // shift due to the -1 PAD value (model.py:50)
    if (v <= 0) continue;                                      // PAD: contributes T[k][0], already in base
    const float* row = a.D + (size_t(k) * a.V + v) * d;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float4 e = __ldg(reinterpret_cast<const float4*>(row + 128 * i + 4 * lane));
      acc[i].x += e.x; acc[i].y += e.y; acc[i].z += e.z; acc[i].w += e.w;
    }
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const size_t idx = size_t(t) * d + 128 * i + 4 * lane;
    float4 m = dropout_mult4(a.drop, idx);
    *reinterpret_cast<float4*>(a.x + idx) = make_float4(acc[i].x * m.x, acc[i].y * m.y, acc[i].z * m.z, acc[i].w * m.w);
  }
}

// ---------------------------------------------------------------------------------------------------------
// embedding backward.  Thread = channel (race-free accumulation); a block owns `spb` whole sequences.
//   dpos[s] += ..., dcmd[c] += ..., dgrp[g] += ... (block-local then one atomic per row/channel),
//   dD[k*V + v] += dx (global RED, non-PAD slots only).   dbase is recovered as sum_s dpos[s] by the caller.
// ---------------------------------------------------------------------------------------------------------
struct EmbedBwdArgs {
  const float* commands;
  const float* args;
  const uint8_t* grp;
  const float* dx;     // [T, d] gradient w.r.t. the (post-dropout) embedding output
  float* dcmd_tab;     // [7, d]
  float* dpos_tab;     // [L, d]
  float* dgrp_tab;     // [G+2, d] or nullptr
  float* dD;           // [n_args*V, d]
  int nseq, L, V, n_args, d, n_grp, spb;
  Dropout drop;
};

// The ids of the block's sequences are staged in shared memory once (commands, group ids, args + 1, and per token whether
// any argument slot is used: two thirds of the positions are EOS padding), and the dx rows are fetched eight positions at a
// time so that eight independent row loads per thread are in flight.  (The first version walked the tokens one by one
// with the id loads, the dx load and the reductions in one dependent chain: 380 us per launch at N = 512, latency-bound.)
__global__ void __launch_bounds__(512) embed_bwd_kernel(EmbedBwdArgs a) {
  drop_resolve(a.drop);
  extern __shared__ float sm[];  // [7 + n_grp][d] accumulators | int cmd[ntok] | int grp[ntok] | int used[ntok] | int arg[ntok][n_args]
  const int c = threadIdx.x;
  const int d = a.d, L = a.L, na = a.n_args;
  const int nsm = 7 + a.n_grp;
  const int q0 = blockIdx.x * a.spb;
  const int q1 = min(a.nseq, q0 + a.spb);
  const int nq = q1 - q0, ntok = nq * L, cap = a.spb * L;
  int* s_cmd = reinterpret_cast<int*>(sm + size_t(nsm) * d);
  int* s_grp = s_cmd + cap;
  int* s_used = s_grp + cap;
  int* s_arg = s_used + cap;
  for (int r = 0; r < nsm; ++r) sm[r * d + c] = 0.f;
  for (int i = c; i < ntok; i += blockDim.x) {
    const size_t t = size_t(q0) * L + i;
    s_cmd[i] = int(a.commands[t]);
    s_grp[i] = a.dgrp_tab != nullptr ? int(a.grp[t]) : 0;
    const float* ar = a.args + t * na;
    int used = 0;
    for (int k = 0; k < na; ++k) {
      const int vv = int(ar[k]) + 1;   // 0 = PAD (argument value -1)
      s_arg[i * na + k] = vv;
      used |= (vv > 0);
    }
    s_used[i] = used;
  }
  __syncthreads();
  constexpr int kPos = 8;
  for (int s0 = 0; s0 < L; s0 += kPos) {
    float accp[kPos];
#pragma unroll
    for (int j = 0; j < kPos; ++j) accp[j] = 0.f;
    for (int qi = 0; qi < nq; ++qi) {
      const size_t tbase = size_t(q0 + qi) * L + s0;
      float v[kPos];
#pragma unroll
      for (int j = 0; j < kPos; ++j) v[j] = (s0 + j < L) ? __ldg(a.dx + (tbase + j) * d + c) : 0.f;
#pragma unroll
      for (int j = 0; j < kPos; ++j) {
        if (s0 + j < L) {
          float x = v[j];
          if (a.drop.p > 0.f) x *= dropout_mult(a.drop, (tbase + j) * d + c);
          accp[j] += x;
          const int li = qi * L + s0 + j;
          sm[s_cmd[li] * d + c] += x;
          if (a.dgrp_tab != nullptr) sm[(7 + s_grp[li]) * d + c] += x;
          if (s_used[li]) {
            const int* ar = s_arg + li * na;
            for (int k = 0; k < na; ++k) {
              const int vv = ar[k];
              if (vv > 0) atomicAdd(a.dD + (size_t(k) * a.V + vv) * d + c, x);
            }
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < kPos; ++j)
      if (s0 + j < L) atomicAdd(a.dpos_tab + size_t(s0 + j) * d + c, accp[j]);
  }
  for (int r = 0; r < 7; ++r) atomicAdd(a.dcmd_tab + size_t(r) * d + c, sm[r * d + c]);
  if (a.dgrp_tab != nullptr)
    for (int r = 0; r < a.n_grp; ++r) atomicAdd(a.dgrp_tab + size_t(r) * d + c, sm[(7 + r) * d + c]);
}

// dbase[c] = sum_s dpos[s][c];  dbias = dbase;  dT[k][0] = dbase - sum_{v>=1} dD[k][v]   (in place in dD row 0)
__global__ void unfold_row0_kernel(float* __restrict__ dD, const float* __restrict__ dpos, float* __restrict__ dbias,
                                   int V, int L, int d) {
  const int k = blockIdx.x;
  for (int c = threadIdx.x; c < d; c += blockDim.x) {
    float tot = 0.f;
    for (int s = 0; s < L; ++s) tot += dpos[size_t(s) * d + c];
    float sub = 0.f;
    for (int v = 1; v < V; ++v) sub += dD[(size_t(k) * V + v) * d + c];
    dD[size_t(k) * V * d + c] = tot - sub;
    if (k == 0) atomicAdd(dbias + c, tot);
  }
}
// dEa[v][e] += sum_c dT[k][v][c] * W[c][64k+e]      (block = (v, k), thread = e)
__global__ void unfold_dEa_kernel(const float* __restrict__ dT, const float* __restrict__ W, float* __restrict__ dEa,
                                  int V, int n_args, int d) {
  const int v = blockIdx.x, k = blockIdx.y, e = threadIdx.x;  // 64 threads
  const float* row = dT + (size_t(k) * V + v) * d;
  const float* w = W + 64 * k + e;
  float s0 = 0.f, s1 = 0.f;
  for (int c = 0; c < d; c += 2) {
    s0 = fmaf(row[c], w[size_t(c) * (64 * n_args)], s0);
    s1 = fmaf(row[c + 1], w[size_t(c + 1) * (64 * n_args)], s1);
  }
  atomicAdd(dEa + size_t(v) * 64 + e, s0 + s1);
}
// dW[c][64k+e] += sum_v dT[k][v][c] * Ea[v][e]   (block = (chunk of v, k), thread = c with 64 register accumulators)
__global__ void __launch_bounds__(512)
unfold_dW_kernel(const float* __restrict__ dT, const float* __restrict__ Ea, float* __restrict__ dW, int V, int n_args,
                 int d, int rows_per_block) {
  __shared__ float ea[64];
  const int k = blockIdx.y, c = threadIdx.x;
  float acc[64];
#pragma unroll
  for (int e = 0; e < 64; ++e) acc[e] = 0.f;
  const int v0 = blockIdx.x * rows_per_block, v1 = min(V, v0 + rows_per_block);
  for (int v = v0; v < v1; ++v) {
    __syncthreads();
    if (c < 64) ea[c] = Ea[size_t(v) * 64 + c];
    __syncthreads();
    const float t = dT[(size_t(k) * V + v) * d + c];
#pragma unroll
    for (int e = 0; e < 64; ++e) acc[e] = fmaf(t, ea[e], acc[e]);
  }
  float* out = dW + size_t(c) * (64 * n_args) + 64 * k;
#pragma unroll
  for (int e = 0; e < 64; ++e) atomicAdd(out + e, acc[e]);
}

// ---------------------------------------------------------------------------------------------------------
// x[r] = dropout(add[r] + tab[r % L])  (E2 input: pooled path codes + group PE; decoder inputs: PE only, add = null)
// and the matching backward: dadd = mask * dx ; dtab[s] += sum over rows with r % L == s
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rows_embed_fwd_kernel(const float* __restrict__ add, const float* __restrict__ tab, float* __restrict__ x, int R,
                      int L, int d, Dropout drop) {
  drop_resolve(drop);
  const size_t n4 = size_t(R) * d / 4;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n4; i += size_t(gridDim.x) * blockDim.x) {
    const size_t idx = i * 4;
    const int r = int(idx / d), c = int(idx % d);
    float4 v = *reinterpret_cast<const float4*>(tab + size_t(r % L) * d + c);
    if (add != nullptr) {
      float4 u = *reinterpret_cast<const float4*>(add + idx);
      v.x += u.x; v.y += u.y; v.z += u.z; v.w += u.w;
    }
    float4 m = dropout_mult4(drop, idx);
    *reinterpret_cast<float4*>(x + idx) = make_float4(v.x * m.x, v.y * m.y, v.z * m.z, v.w * m.w);
  }
}
// thread = channel; block b owns rows {r : r / L in its sequence range}; dtab via one atomic per (s, c) per block
__global__ void __launch_bounds__(512)
rows_embed_bwd_kernel(const float* __restrict__ dx, float* __restrict__ dadd, float* __restrict__ dtab, int nseq, int L,
                      int d, int spb, Dropout drop) {
  drop_resolve(drop);
  const int c = threadIdx.x;
  const int q0 = blockIdx.x * spb, q1 = min(nseq, q0 + spb);
  for (int s = 0; s < L; ++s) {
    float acc = 0.f;
    for (int q = q0; q < q1; ++q) {
      const size_t idx = (size_t(q) * L + s) * d + c;
      float v = dx[idx];
      if (drop.p > 0.f) v *= dropout_mult(drop, idx);
      if (dadd != nullptr) dadd[idx] = v;
      acc += v;
    }
    atomicAdd(dtab + size_t(s) * d + c, acc);
  }
}

}  // namespace dsvg
using namespace dsvg;

extern "C" int dsvg_seq_prep(const float* commands, int nseq, int L, int* first_eos, uint8_t* visible,
                             uint8_t* key_valid, uint8_t* grp, float* counts, void* stream) {
  DSVG_CHECK(commands && nseq > 0 && L > 0, "dsvg_seq_prep: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  seq_prep_kernel<<<ceil_div(nseq, 8), 256, 0, st>>>(commands, nseq, L, first_eos, visible, key_valid, grp, counts);
  ++g_launches;
  DSVG_LAUNCH_CHECK();
  return 0;
}

extern "C" int dsvg_embed_fold(const float* arg_embed, const float* W, const float* bias, float* table, float* base,
                               int V, int n_args, int d, void* stream) {
  DSVG_CHECK(arg_embed && W && bias && table && base, "dsvg_embed_fold: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DSVG_CHECK(d <= 512 && d >= 64, "dsvg_embed_fold: d_model must be in [64, 512]");
  {
    const int rpb = 16;
    fold_kernel<<<dim3(ceil_div(V, rpb), n_args), d, 0, st>>>(arg_embed, W, table, V, n_args, d, rpb);
  }
  ++g_launches;
  DSVG_LAUNCH_CHECK();
  fold_sub_kernel<<<dim3(16, n_args), 256, 0, st>>>(table, bias, base, V, n_args, d);
  ++g_launches;
  DSVG_LAUNCH_CHECK();
  return 0;
}

extern "C" int dsvg_embed_fwd(const float* commands, const float* args, const uint8_t* grp, const float* cmd_tab,
                              const float* table, const float* base, const float* pos_tab, const float* grp_tab,
                              float* x, int T, int L, int V, int n_args, int d, float drop_p, uint32_t drop_site,
                              uint64_t seed, void* stream) {
  DSVG_CHECK(commands && args && cmd_tab && table && base && pos_tab && x && T > 0, "dsvg_embed_fwd: bad arguments");
  DSVG_CHECK(n_args <= 30, "dsvg_embed_fwd: n_args must be <= 30");
  DSVG_CHECK((grp_tab == nullptr) || (grp != nullptr), "dsvg_embed_fwd: group table without group indices");
  EmbedArgs a{commands, args, grp, cmd_tab, table, base, pos_tab, grp_tab, x, T, L, V, n_args,
              make_dropout(drop_p, drop_site, seed)};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int grid = ceil_div(T, 8);
  switch (d / 128) {
    case 1: embed_fwd_kernel<1><<<grid, 256, 0, st>>>(a); break;
    case 2: embed_fwd_kernel<2><<<grid, 256, 0, st>>>(a); break;
    case 4: embed_fwd_kernel<4><<<grid, 256, 0, st>>>(a); break;
    default: DSVG_CHECK(false, "dsvg_embed_fwd: d_model %d unsupported", d);
  }
  ++g_launches;
  DSVG_LAUNCH_CHECK();
  return 0;
}

extern "C" int dsvg_embed_bwd(const float* commands, const float* args, const uint8_t* grp, const float* dx,
                              const float* arg_embed, const float* W, float* d_cmd_tab, float* d_pos_tab,
                              float* d_grp_tab, float* d_arg_embed, float* d_W, float* d_bias, float* scratch_table,
                              int nseq, int L, int V, int n_args, int d, int n_grp, float drop_p, uint32_t drop_site,
                              uint64_t seed, void* stream) {
  DSVG_CHECK(commands && args && dx && arg_embed && W && d_cmd_tab && d_pos_tab && d_arg_embed && d_W && d_bias &&
                 scratch_table,
             "dsvg_embed_bwd: null pointer");
  DSVG_CHECK(d <= 512 && d % 32 == 0, "dsvg_embed_bwd: d_model must be <= 512");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  DSVG_CUDA(cudaMemsetAsync(scratch_table, 0, sizeof(float) * size_t(n_args) * V * d, st));
  EmbedBwdArgs a{};
  a.commands = commands; a.args = args; a.grp = grp; a.dx = dx;
  a.dcmd_tab = d_cmd_tab; a.dpos_tab = d_pos_tab; a.dgrp_tab = d_grp_tab; a.dD = scratch_table;
  a.nseq = nseq; a.L = L; a.V = V; a.n_args = n_args; a.d = d; a.n_grp = d_grp_tab ? n_grp : 0;
  a.spb = 4;
  a.drop = make_dropout(drop_p, drop_site, seed);
  const size_t smem = sizeof(float) * size_t(7 + a.n_grp) * d + sizeof(int) * size_t(a.spb) * L * (3 + n_args);
  DSVG_CHECK(smem <= 200 * 1024, "dsvg_embed_bwd: sequence too long for the id staging buffer (%zu bytes)", smem);
  if (smem > 48 * 1024) {     // long one-stage sequences: the group table alone is (max_total_len + 2) x d floats
    static bool configured[kMaxDevices] = {};
    if (first_use_on_device(configured))
      DSVG_CUDA(cudaFuncSetAttribute(embed_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  }
  embed_bwd_kernel<<<ceil_div(nseq, a.spb), d, smem, st>>>(a);
  ++g_launches;
  DSVG_LAUNCH_CHECK();
  unfold_row0_kernel<<<n_args, 256, 0, st>>>(scratch_table, d_pos_tab, d_bias, V, L, d);
  ++g_launches;
  DSVG_LAUNCH_CHECK();
  unfold_dEa_kernel<<<dim3(V, n_args), 64, 0, st>>>(scratch_table, W, d_arg_embed, V, n_args, d);
  ++g_launches;
  DSVG_LAUNCH_CHECK();
  unfold_dW_kernel<<<dim3(ceil_div(V, 32), n_args), d, 0, st>>>(scratch_table, arg_embed, d_W, V, n_args, d, 32);
  ++g_launches;
  DSVG_LAUNCH_CHECK();
  return 0;
}

extern "C" int dsvg_rows_embed_fwd(const float* add, const float* tab, float* x, int R, int L, int d, float drop_p,
                                   uint32_t drop_site, uint64_t seed, void* stream) {
  DSVG_CHECK(tab && x && R > 0 && L > 0 && d % 4 == 0, "dsvg_rows_embed_fwd: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t n4 = size_t(R) * d / 4;
  int grid = int((n4 + 255) / 256);
  if (grid > 148 * 16) grid = 148 * 16;
  rows_embed_fwd_kernel<<<grid, 256, 0, st>>>(add, tab, x, R, L, d, make_dropout(drop_p, drop_site, seed));
  ++g_launches;
  DSVG_LAUNCH_CHECK();
  return 0;
}

extern "C" int dsvg_rows_embed_bwd(const float* dx, float* dadd, float* dtab, int nseq, int L, int d, float drop_p,
                                   uint32_t drop_site, uint64_t seed, void* stream) {
  DSVG_CHECK(dx && dtab && nseq > 0 && L > 0 && d <= 512, "dsvg_rows_embed_bwd: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int spb = nseq >= 2048 ? 8 : 1;
  rows_embed_bwd_kernel<<<ceil_div(nseq, spb), d, 0, st>>>(dx, dadd, dtab, nseq, L, d, spb,
                                                          make_dropout(drop_p, drop_site, seed));
  ++g_launches;
  DSVG_LAUNCH_CHECK();
  return 0;
}
