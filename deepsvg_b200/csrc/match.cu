// Hungarian self-matching of predicted path slots to target paths (cfg.self_match, HierarchicalSelfMatching).
//
//   reference: SVGTransformer.perfect_matching, model/model.py:311-350 -- cost[n, g, p] = 2 * masked-mean CE_args + masked-mean
//   CE_cmd + CE_visibility between target path g and predicted slot p (built there by repeating logits and targets G x Gp
//   times and calling F.cross_entropy: ~11 GB of temporaries at N = 512), then scipy.optimize.linear_sum_assignment per icon
//   on the host (a device->host sync and a Python loop), then torch.gather of the logits along the slot axis (:389-391).
//
//   here: (1) one pass over the logits for the log-sum-exp of every (slot, position, argument) group, (2) one warp per
//   (icon, target path, slot) gathers the target logits and reduces the three terms, (3) one thread per icon solves the
//   <= 16 x 16 assignment problem (shortest augmenting paths with potentials, fp64), (4) a group-granular copy kernel
//   applies the permutation (and its inverse in the backward pass).  No host round trip.
#include <cfloat>

#include "../../include/dsvg_b200.h"
#include "common.cuh"

namespace dsvg {
extern unsigned long long g_launches;

__constant__ uint8_t c_match_mask[7][11] = {
    {0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1}, {0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1}, {0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1},
    {1, 1, 1, 1, 1, 0, 0, 0, 0, 1, 1}, {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0},
    {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}};

// (1) log-sum-exp of the command logits and of each argument group of one predicted token; one warp per token
__global__ void __launch_bounds__(256)
match_lse_kernel(const float* __restrict__ cmd_logits, int n_cmd, const float* __restrict__ args_logits, int ld_args,
                 int n_args, int C, float* __restrict__ lse_c, float* __restrict__ lse_a, long long n_tok) {
  const int lane = threadIdx.x & 31;
  const long long tok = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (tok >= n_tok) return;
  {
    const float v = lane < n_cmd ? cmd_logits[tok * n_cmd + lane] : -INFINITY;
    const float m = warp_max(v);
    const float s = warp_sum(lane < n_cmd ? expf(v - m) : 0.f);
    if (lane == 0) lse_c[tok] = m + logf(s);
  }
  const float* row = args_logits + tok * ld_args;
  for (int k = 0; k < n_args; ++k) {
    const float* l = row + k * C;
    float m = -INFINITY;
    for (int j = lane; j < C; j += 32) m = fmaxf(m, l[j]);
    m = warp_max(m);
    float s = 0.f;
    for (int j = lane; j < C; j += 32) s += expf(l[j] - m);
    s = warp_sum(s);
    if (lane == 0) lse_a[tok * n_args + k] = m + logf(s);
  }
}

// (2) cost[n, g, p]; one warp per (n, g, p).  Targets are the SHIFTED sequences (commands[..., 1:]); visibility and the
// extended padding mask are taken on them exactly as perfect_matching does (model.py:314-315).
__global__ void __launch_bounds__(256)
match_cost_kernel(const float* __restrict__ cmd_logits, int n_cmd, const float* __restrict__ args_logits, int ld_args,
                  int n_args, int C, const float* __restrict__ vis_logits, const float* __restrict__ lse_c,
                  const float* __restrict__ lse_a, const float* __restrict__ commands, const float* __restrict__ args, int N,
                  int G, int Gp, int L, double* __restrict__ cost, uint8_t* __restrict__ vis_out) {
  const int lane = threadIdx.x & 31;
  const long long w = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= (long long)N * G * Gp) return;
  const int p = int(w % Gp), g = int((w / Gp) % G), n = int(w / ((long long)Gp * G));
  const int Ld = L - 1;
  const float* tc = commands + (size_t(n) * G + g) * L + 1;           // shifted targets
  const float* ta = args + ((size_t(n) * G + g) * L + 1) * n_args;
  // first EOS and EOS count of the shifted sequence
  int n_eos = 0, fe = Ld;
  for (int s = lane; s < Ld; s += 32) {
    if (int(tc[s]) == 4) {
      ++n_eos;
      fe = min(fe, s);
    }
  }
  n_eos = int(warp_sum(float(n_eos)));
  for (int o = 16; o > 0; o >>= 1) fe = min(fe, __shfl_xor_sync(0xffffffffu, fe, o));
  const bool visible = n_eos < Ld - 1;                                  // model/utils.py:45-56 on the shifted targets
  float num_c = 0.f, cnt_c = 0.f, num_a = 0.f, cnt_a = 0.f;
  for (int s = lane; s < Ld; s += 32) {
    const size_t tok = (size_t(n) * Gp + p) * Ld + s;
    const int c = int(tc[s]);
    const bool ext = visible && ((s < fe) || (s >= 3 && s < fe + 3));   // clean OR-shift-by-3 (SURVEY.md 8c hazard 1)
    if (ext) {
      num_c += lse_c[tok] - cmd_logits[tok * n_cmd + c];
      cnt_c += 1.f;
    }
    for (int k = 0; k < n_args; ++k) {
      if (c_match_mask[c][k]) {
        const int t = int(ta[size_t(s) * n_args + k]) + 1;
        num_a += lse_a[tok * n_args + k] - args_logits[tok * ld_args + k * C + t];
        cnt_a += 1.f;
      }
    }
  }
  num_c = warp_sum(num_c); cnt_c = warp_sum(cnt_c); num_a = warp_sum(num_a); cnt_a = warp_sum(cnt_a);
  if (lane == 0) {
    const float* vl = vis_logits + (size_t(n) * Gp + p) * 2;
    const float m = fmaxf(vl[0], vl[1]);
    const float lse = m + logf(expf(vl[0] - m) + expf(vl[1] - m));
    const float ce_v = lse - vl[visible ? 1 : 0];
    // 0 / 0 for invisible targets, as in the reference: those rows never reach the solver (costs[mask])
    cost[w] = 2.0 * double(num_a / cnt_a) + 1.0 * double(num_c / cnt_c) + 1.0 * double(ce_v);
    if (p == 0) vis_out[size_t(n) * G + g] = visible ? 1 : 0;
  }
}

// (3) rectangular assignment (rows = visible targets in order, columns = the Gp slots), one thread per icon.
// Shortest augmenting paths with dual potentials (the classical O(n^2 m) Hungarian formulation); n, m <= 16.
constexpr int kMaxSlots = 16;
__global__ void match_assign_kernel(const double* __restrict__ cost, const uint8_t* __restrict__ vis, int N, int G, int Gp,
                                    long long* __restrict__ assignment) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  int rows[kMaxSlots];
  int nr = 0;
  for (int g = 0; g < G; ++g)
    if (vis[size_t(n) * G + g]) rows[nr++] = g;
  const double* a = cost + size_t(n) * G * Gp;
  double u[kMaxSlots + 1], v[kMaxSlots + 1], minv[kMaxSlots + 1];
  int pcol[kMaxSlots + 1], way[kMaxSlots + 1];
  bool used[kMaxSlots + 1];
  for (int j = 0; j <= Gp; ++j) { v[j] = 0.0; pcol[j] = 0; }
  for (int i = 0; i <= nr; ++i) u[i] = 0.0;
  for (int i = 1; i <= nr; ++i) {
    pcol[0] = i;
    int j0 = 0;
    for (int j = 0; j <= Gp; ++j) { minv[j] = DBL_MAX; used[j] = false; }
    do {
      used[j0] = true;
      const int i0 = pcol[j0];
      double delta = DBL_MAX;
      int j1 = 0;
      for (int j = 1; j <= Gp; ++j) {
        if (!used[j]) {
          const double cur = a[size_t(rows[i0 - 1]) * Gp + (j - 1)] - u[i0] - v[j];
          if (cur < minv[j]) { minv[j] = cur; way[j] = j0; }
          if (minv[j] < delta) { delta = minv[j]; j1 = j; }
        }
      }
      for (int j = 0; j <= Gp; ++j) {
        if (used[j]) { u[pcol[j]] += delta; v[j] -= delta; }
        else minv[j] -= delta;
      }
      j0 = j1;
    } while (pcol[j0] != 0);
    do {
      const int j1 = way[j0];
      pcol[j0] = pcol[j1];
      j0 = j1;
    } while (j0 != 0);
  }
  // assignment list of the reference (model.py:342-346): slot of the i-th visible target, then the unused slots ascending
  long long* out = assignment + size_t(n) * Gp;
  bool taken[kMaxSlots];
  for (int j = 0; j < Gp; ++j) taken[j] = false;
  for (int j = 1; j <= Gp; ++j)
    if (pcol[j] != 0) { out[pcol[j] - 1] = j - 1; taken[j - 1] = true; }
  int k = nr;
  for (int j = 0; j < Gp; ++j)
    if (!taken[j]) out[k++] = j;
}

// (4) dst group (n, i) <- src group (n, asg[n, i])   (inverse: dst group (n, asg[n, i]) <- src group (n, i));
// a group is `group_bytes` contiguous bytes (multiple of 4); 16-byte vectors when everything is 16-byte aligned
__global__ void __launch_bounds__(256)
permute_groups_kernel(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, const long long* __restrict__ asg, int G,
                      size_t group_bytes, int inverse, int vec16) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t grp = blockIdx.x;                       // n * G + i
  const size_t n = grp / G, i = grp % G;
  const size_t other = n * G + size_t(asg[grp]);
  const uint8_t* s = src + (inverse ? grp : other) * group_bytes;
  uint8_t* d = dst + (inverse ? other : grp) * group_bytes;
  if (vec16) {
    const size_t nv = group_bytes / 16;
    for (size_t k = size_t(blockIdx.y) * blockDim.x + threadIdx.x; k < nv; k += size_t(gridDim.y) * blockDim.x)
      reinterpret_cast<uint4*>(d)[k] = reinterpret_cast<const uint4*>(s)[k];
  } else {
    const size_t nv = group_bytes / 4;
    for (size_t k = size_t(blockIdx.y) * blockDim.x + threadIdx.x; k < nv; k += size_t(gridDim.y) * blockDim.x)
      reinterpret_cast<uint32_t*>(d)[k] = reinterpret_cast<const uint32_t*>(s)[k];
  }
  (void)i;
}

}  // namespace dsvg
using namespace dsvg;

extern "C" int dsvg_match_assign(const float* cmd_logits, int n_cmd, const float* args_logits, int ld_args, int n_args,
                                 int n_classes, const float* vis_logits, const float* commands, const float* args, int N, int G,
                                 int Gp, int L, float* lse_cmd, float* lse_args, double* cost, unsigned char* visible,
                                 long long* assignment, void* stream) {
  DSVG_CHECK(cmd_logits && args_logits && vis_logits && commands && args && lse_cmd && lse_args && cost && visible && assignment,
             "dsvg_match_assign: null pointer");
  DSVG_CHECK(N > 0 && G > 0 && Gp > 0 && L > 1, "dsvg_match_assign: bad shape");
  DSVG_CHECK(G <= Gp && Gp <= kMaxSlots, "dsvg_match_assign: needs G <= num_groups_proposal <= %d", kMaxSlots);
  DSVG_CHECK(n_cmd <= 32 && n_args <= 11, "dsvg_match_assign: n_commands <= 32, n_args <= 11");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n_tok = (long long)N * Gp * (L - 1);
  match_lse_kernel<<<unsigned((n_tok + 7) / 8), 256, 0, st>>>(cmd_logits, n_cmd, args_logits, ld_args, n_args, n_classes, lse_cmd,
                                                               lse_args, n_tok);
  DSVG_LAUNCH_CHECK();
  const long long n_w = (long long)N * G * Gp;
  match_cost_kernel<<<unsigned((n_w + 7) / 8), 256, 0, st>>>(cmd_logits, n_cmd, args_logits, ld_args, n_args, n_classes, vis_logits,
                                                              lse_cmd, lse_args, commands, args, N, G, Gp, L, cost, visible);
  DSVG_LAUNCH_CHECK();
  match_assign_kernel<<<(N + 63) / 64, 64, 0, st>>>(cost, visible, N, G, Gp, assignment);
  DSVG_LAUNCH_CHECK();
  g_launches += 3;
  return 0;
}

extern "C" int dsvg_permute_groups(const void* src, void* dst, const long long* assignment, int N, int G, size_t group_bytes,
                                   int inverse, void* stream) {
  DSVG_CHECK(src && dst && assignment && N > 0 && G > 0 && group_bytes > 0 && group_bytes % 4 == 0 && src != dst,
             "dsvg_permute_groups: bad arguments");
  const int vec16 = (group_bytes % 16 == 0) && (reinterpret_cast<uintptr_t>(src) % 16 == 0) && (reinterpret_cast<uintptr_t>(dst) % 16 == 0);
  size_t per = group_bytes / (vec16 ? 16 : 4);
  unsigned gy = unsigned((per + 2047) / 2048);
  if (gy < 1) gy = 1;
  if (gy > 64) gy = 64;
  DSVG_CUDA(launch_k(permute_groups_kernel, dim3(unsigned(N) * unsigned(G), gy), dim3(256), 0, static_cast<cudaStream_t>(stream),
                     static_cast<const uint8_t*>(src), static_cast<uint8_t*>(dst), assignment, G, group_bytes, inverse, vec16));
  ++g_launches;
  return 0;
}
