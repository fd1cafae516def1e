// SVGLoss forward + the gradient of every loss term w.r.t. the logits, in one streaming pass per tensor.
//
//   reference: SVGLoss.forward, model/loss.py:19-65
//     loss_cmd  = mean over {positions selected by padding_mask*visibility} of CE_7 (command_logits, tgt_commands)
//     loss_args = mean over {arg slots selected by CMD_ARGS_MASK[tgt_commands]} of CE_257(args_logits, tgt_args + 1)
//     loss_visibility = mean CE_2(visibility_logits, visibility);  loss_kl = max(tol, -0.5 mean(1+ls-mu^2-e^ls))
//   The reference compacts logits with a boolean mask (a device->host sync); here the mask is a weight, nothing is
//   compacted, and the normalising counts come from dsvg_seq_prep (device memory, no sync).
//
// Gradients are written with UNIT upstream scale:  dlogits = w * (softmax - onehot) / count.  The per-term factors
// (loss weights x upstream autograd gradient) are applied by the consumers (alpha_dev of dsvg_linear / dsvg_outer).
#include "../../include/dsvg_b200.h"
#include "common.cuh"

namespace dsvg {
extern unsigned long long g_launches;

__constant__ uint8_t c_cmd_args_mask[7][11] = {
    {0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1}, {0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1}, {0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1},
    {1, 1, 1, 1, 1, 0, 0, 0, 0, 1, 1}, {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
    {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}};

// accumulators (float[8]): 0 = sum w*CE cmd, 1 = sum w*CE args, 2 = sum CE visibility, 3 = KL sum
// counts (float[2]):       0 = cmd positions, 1 = arg slots        (possibly all-reduced across ranks)

// ---------------------------------------------------------------------------------------------------------
// argument slots: one warp per target token, 11 slots x C classes
// ---------------------------------------------------------------------------------------------------------
template <int KR>   // classes per slot <= 32 * KR
__global__ void __launch_bounds__(256)
ce_args_kernel(const float* __restrict__ logits, int ld_logits, const float* __restrict__ commands,
               const float* __restrict__ args, const float* __restrict__ counts, bf16* __restrict__ dl, size_t dl_lo,
               int ld_dl, float* __restrict__ acc, int nseq, int L, int n_args, int C) {
  const int lane = threadIdx.x & 31;
  const int Ld = L - 1;
  const long long tok = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  float loss = 0.f;
  if (tok < (long long)nseq * Ld) {
    const int q = int(tok / Ld), sp = int(tok % Ld);
    const size_t src = size_t(q) * L + sp + 1;  // target position (SOS dropped: loss.py:49)
    const int cmd = int(commands[src]);
    const float inv_cnt = 1.f / counts[1];
    const float* lrow = logits + size_t(tok) * ld_logits;
    bf16* drow = dl + size_t(tok) * ld_dl;
    for (int k = 0; k < n_args; ++k) {
      const float* l = lrow + k * C;
      bf16* dk = drow + k * C;
      if (!c_cmd_args_mask[cmd][k]) {
        for (int j = lane; j < C; j += 32) {
          dk[j] = __float2bfloat16_rn(0.f);
          if (dl_lo) dk[j + dl_lo] = __float2bfloat16_rn(0.f);
        }
        continue;
      }
      const int tgt = int(args[src * n_args + k]) + 1;  // shift due to the -1 PAD value (loss.py:54)
      float v[KR];
      float m = -INFINITY;
#pragma unroll
      for (int i = 0; i < KR; ++i) {
        const int j = lane + 32 * i;
        v[i] = j < C ? l[j] : -INFINITY;
        m = fmaxf(m, v[i]);
      }
      m = warp_max(m);
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < KR; ++i) s += (lane + 32 * i < C) ? expf(v[i] - m) : 0.f;
      s = warp_sum(s);
      const float lse = m + logf(s);
      float lt = 0.f;  // logit of the target class: lane (tgt & 31), register (tgt >> 5); tgt is warp-uniform
#pragma unroll
      for (int i = 0; i < KR; ++i) {
        const float cand = __shfl_sync(0xffffffffu, v[i], tgt & 31);
        if (i == (tgt >> 5)) lt = cand;
      }
      if (lane == 0) loss += lse - lt;
#pragma unroll
      for (int i = 0; i < KR; ++i) {
        const int j = lane + 32 * i;
        if (j < C) {
          float g = (expf(v[i] - lse) - (j == tgt ? 1.f : 0.f)) * inv_cnt;
          act_store(dk, dl_lo, j, g);
        }
      }
    }
  }
  // block-level sum of the per-token losses -> one atomic
  __shared__ float red[8];
  if (lane == 0) red[threadIdx.x >> 5] = loss;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < (blockDim.x >> 5); ++w) t += red[w];
    if (t != 0.f) atomicAdd(acc + 1, t);
  }
}

// ---------------------------------------------------------------------------------------------------------
// commands: one thread per target token, <= 8 classes
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
ce_cmd_kernel(const float* __restrict__ logits, const float* __restrict__ commands, const int* __restrict__ first_eos,
              const uint8_t* __restrict__ visible, const float* __restrict__ counts, bf16* __restrict__ dl,
              size_t dl_lo, int ld_dl, float* __restrict__ acc, int nseq, int L, int C) {
  const int Ld = L - 1;
  const long long tok = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  float loss = 0.f;
  if (tok < (long long)nseq * Ld) {
    const int q = int(tok / Ld), i = int(tok % Ld) + 1;
    const int fe = first_eos[q];
    const bool w = visible[q] && ((i < fe) || (i >= 3 && i - 3 < fe));
    const float* l = logits + size_t(tok) * C;
    bf16* d = dl + size_t(tok) * ld_dl;
    if (w) {
      const int tgt = int(commands[size_t(q) * L + i]);
      float m = -INFINITY;
      for (int j = 0; j < C; ++j) m = fmaxf(m, l[j]);
      float s = 0.f;
      for (int j = 0; j < C; ++j) s += expf(l[j] - m);
      const float lse = m + logf(s);
      loss = lse - l[tgt];
      const float inv = 1.f / counts[0];
      for (int j = 0; j < C; ++j) act_store(d, dl_lo, j, (expf(l[j] - lse) - (j == tgt ? 1.f : 0.f)) * inv);
    } else {
      for (int j = 0; j < C; ++j) act_store(d, dl_lo, j, 0.f);
    }
    for (int j = C; j < ld_dl; ++j) act_store(d, dl_lo, j, 0.f);
  }
  loss = warp_sum(loss);
  if ((threadIdx.x & 31) == 0 && loss != 0.f) atomicAdd(acc + 0, loss);
}

// visibility: one thread per path, 2 classes, plain mean over n_total paths (loss.py:43)
__global__ void __launch_bounds__(256)
ce_vis_kernel(const float* __restrict__ logits, const uint8_t* __restrict__ visible, bf16* __restrict__ dl, size_t dl_lo,
              int ld_dl, float* __restrict__ acc, int nseq, float inv_total) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  float loss = 0.f;
  if (q < nseq) {
    const float a = logits[2 * q], b = logits[2 * q + 1];
    const float m = fmaxf(a, b);
    const float lse = m + logf(expf(a - m) + expf(b - m));
    const int t = visible[q] ? 1 : 0;
    loss = lse - (t ? b : a);
    bf16* d = dl + size_t(q) * ld_dl;
    act_store(d, dl_lo, 0, (expf(a - lse) - (t == 0 ? 1.f : 0.f)) * inv_total);
    act_store(d, dl_lo, 1, (expf(b - lse) - (t == 1 ? 1.f : 0.f)) * inv_total);
    for (int j = 2; j < ld_dl; ++j) act_store(d, dl_lo, j, 0.f);
  }
  loss = warp_sum(loss);
  if ((threadIdx.x & 31) == 0 && loss != 0.f) atomicAdd(acc + 2, loss);
}

// KL: sum over elements of (1 + ls - mu^2 - exp(ls))
__global__ void __launch_bounds__(256)
kl_sum_kernel(const float* __restrict__ mu, const float* __restrict__ ls, float* __restrict__ acc, int n) {
  float s = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    s += 1.f + ls[i] - mu[i] * mu[i] - expf(ls[i]);
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) atomicAdd(acc + 3, s);
}

// out[0..4] = loss, loss_cmd, loss_args, loss_visibility, loss_kl ; out[5] = 1 if the KL clamp is inactive else 0
__global__ void loss_finalize_kernel(const float* __restrict__ acc, const float* __restrict__ counts,
                                     float* __restrict__ out, float w_cmd, float w_args, float w_vis, float w_kl,
                                     float kl_tol, float inv_vis_total, float inv_kl_total, int has_vis, int has_kl) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const float lc = acc[0] / counts[0];
  const float la = acc[1] / counts[1];
  float total = w_cmd * lc + w_args * la;
  float lv = 0.f, lk = 0.f, kl_active = 0.f;
  if (has_vis) {
    lv = acc[2] * inv_vis_total;
    total += w_vis * lv;
  }
  if (has_kl) {
    const float raw = -0.5f * acc[3] * inv_kl_total;
    kl_active = raw > kl_tol ? 1.f : 0.f;  // clamp(min=tol) passes gradient only above the tolerance (loss.py:27)
    lk = fmaxf(raw, kl_tol);
    total += w_kl * lk;
  }
  out[0] = total; out[1] = lc; out[2] = la; out[3] = lv; out[4] = lk; out[5] = kl_active;
}

// latent: z = mu + exp(ls/2) * eps (model.py:182-185) and its backward incl. the KL term
__global__ void __launch_bounds__(256)
vae_fwd_kernel(const float* __restrict__ mu, const float* __restrict__ ls, const float* __restrict__ eps,
               float* __restrict__ z, int n) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    z[i] = mu[i] + expf(0.5f * ls[i]) * eps[i];
}
// dmu = dz + kl_scale * mu / n_total ;  dls = dz * eps * 0.5 * exp(ls/2) - kl_scale * 0.5 * (1 - exp(ls)) / n_total
// kl_scale = (*kl_coef_dev) * kl_active  (both device scalars; either may be null -> no KL contribution)
__global__ void __launch_bounds__(256)
vae_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ ls, const float* __restrict__ eps,
               const float* __restrict__ dz, const float* __restrict__ kl_coef_dev, const float* __restrict__ loss_out,
               float inv_total, float* __restrict__ dmu, float* __restrict__ dls, int n) {
  float ks = 0.f;
  if (kl_coef_dev != nullptr && loss_out != nullptr) ks = (*kl_coef_dev) * loss_out[5] * inv_total;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const float g = dz[i], e = expf(0.5f * ls[i]);
    dmu[i] = g + ks * mu[i];
    dls[i] = g * eps[i] * 0.5f * e - ks * 0.5f * (1.f - e * e);
  }
}

}  // namespace dsvg
using namespace dsvg;

extern "C" int dsvg_ce_args(const float* logits, int ld_logits, const float* commands, const float* args,
                            const float* counts, dsvg_bf16* dlogits, size_t dl_lo_off, int ld_dl, float* acc, int nseq,
                            int L, int n_args, int n_classes, void* stream) {
  DSVG_CHECK(logits && commands && args && counts && dlogits && acc, "dsvg_ce_args: null pointer");
  DSVG_CHECK(n_classes <= 512 && n_args <= 11, "dsvg_ce_args: at most 512 classes x 11 slots");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long toks = (long long)nseq * (L - 1);
  if (n_classes <= 288)
    ce_args_kernel<9><<<ceil_div(toks, 8), 256, 0, st>>>(logits, ld_logits, commands, args, counts,
                                                         reinterpret_cast<bf16*>(dlogits), dl_lo_off, ld_dl, acc, nseq, L,
                                                         n_args, n_classes);
  else   // relative-argument targets: 2 * args_dim classes (loss.py:15)
    ce_args_kernel<16><<<ceil_div(toks, 8), 256, 0, st>>>(logits, ld_logits, commands, args, counts,
                                                          reinterpret_cast<bf16*>(dlogits), dl_lo_off, ld_dl, acc, nseq, L,
                                                          n_args, n_classes);
  ++g_launches;
  DSVG_LAUNCH_CHECK();
  return 0;
}

extern "C" int dsvg_ce_cmd(const float* logits, const float* commands, const int* first_eos, const uint8_t* visible,
                           const float* counts, dsvg_bf16* dlogits, size_t dl_lo_off, int ld_dl, float* acc, int nseq,
                           int L, int n_classes, void* stream) {
  DSVG_CHECK(logits && commands && first_eos && visible && counts && dlogits && acc, "dsvg_ce_cmd: null pointer");
  DSVG_CHECK(n_classes <= ld_dl, "dsvg_ce_cmd: ld_dl too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long toks = (long long)nseq * (L - 1);
  ce_cmd_kernel<<<ceil_div(toks, 256), 256, 0, st>>>(logits, commands, first_eos, visible, counts,
                                                     reinterpret_cast<bf16*>(dlogits), dl_lo_off, ld_dl, acc, nseq, L,
                                                     n_classes);
  ++g_launches;
  DSVG_LAUNCH_CHECK();
  return 0;
}

extern "C" int dsvg_ce_vis(const float* logits, const uint8_t* visible, dsvg_bf16* dlogits, size_t dl_lo_off, int ld_dl,
                           float* acc, int nseq, float inv_total, void* stream) {
  DSVG_CHECK(logits && visible && dlogits && acc && ld_dl >= 2, "dsvg_ce_vis: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  ce_vis_kernel<<<ceil_div(nseq, 256), 256, 0, st>>>(logits, visible, reinterpret_cast<bf16*>(dlogits), dl_lo_off,
                                                     ld_dl, acc, nseq, inv_total);
  ++g_launches;
  DSVG_LAUNCH_CHECK();
  return 0;
}

extern "C" int dsvg_kl_sum(const float* mu, const float* logsigma, float* acc, int n, void* stream) {
  DSVG_CHECK(mu && logsigma && acc && n > 0, "dsvg_kl_sum: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int grid = ceil_div(n, 256);
  if (grid > 296) grid = 296;
  kl_sum_kernel<<<grid, 256, 0, st>>>(mu, logsigma, acc, n);
  ++g_launches;
  DSVG_LAUNCH_CHECK();
  return 0;
}

extern "C" int dsvg_loss_finalize(const float* acc, const float* counts, float* out, float w_cmd, float w_args,
                                  float w_vis, float w_kl, float kl_tolerance, float inv_vis_total, float inv_kl_total,
                                  int has_vis, int has_kl, void* stream) {
  DSVG_CHECK(acc && counts && out, "dsvg_loss_finalize: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  loss_finalize_kernel<<<1, 32, 0, st>>>(acc, counts, out, w_cmd, w_args, w_vis, w_kl, kl_tolerance, inv_vis_total,
                                         inv_kl_total, has_vis, has_kl);
  ++g_launches;
  DSVG_LAUNCH_CHECK();
  return 0;
}

extern "C" int dsvg_vae_fwd(const float* mu, const float* logsigma, const float* eps, float* z, int n, void* stream) {
  DSVG_CHECK(mu && logsigma && eps && z && n > 0, "dsvg_vae_fwd: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int grid = ceil_div(n, 256);
  if (grid > 296) grid = 296;
  vae_fwd_kernel<<<grid, 256, 0, st>>>(mu, logsigma, eps, z, n);
  ++g_launches;
  DSVG_LAUNCH_CHECK();
  return 0;
}

extern "C" int dsvg_vae_bwd(const float* mu, const float* logsigma, const float* eps, const float* dz,
                            const float* kl_coef_dev, const float* loss_out, float inv_total, float* dmu, float* dls,
                            int n, void* stream) {
  DSVG_CHECK(mu && logsigma && eps && dz && dmu && dls && n > 0, "dsvg_vae_bwd: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int grid = ceil_div(n, 256);
  if (grid > 296) grid = 296;
  vae_bwd_kernel<<<grid, 256, 0, st>>>(mu, logsigma, eps, dz, kl_coef_dev, loss_out, inv_total, dmu, dls, n);
  ++g_launches;
  DSVG_LAUNCH_CHECK();
  return 0;
}
