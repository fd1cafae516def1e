// Input side of the hot path: the batch format and its two converters.
//
//   reference: SVGTensorDataset.get_data (svgtensor_dataset.py:164-205) pads an icon's per-path (len, 14) tensors to
//   MAX_NUM_GROUPS paths, wraps each in SOS ... EOS, pads with EOS to MAX_SEQ_LEN + 2 (SVGTensor.add_eos / add_sos / pad,
//   difflib/tensor.py:108-143) and stacks `cmds()` / `args()` (11 of the 14 columns: start_pos is dropped, tensor.py:23-46)
//   as float32 -- 12 floats = 48 bytes per position; the default collate then stacks icons.
//
//   here: the same batch as a PACKED host buffer -- command ids uint8 [N, G, S+2], arguments int16 [N, G, S+2, 11]
//   (-1 = PAD, 0..args_dim-1 otherwise; lossless) = 23 bytes per position -- assembled by one native call per batch
//   (dsvg_pack_icons, host code) and expanded on the GPU to the float tensors the forward kernels read
//   (dsvg_unpack_batch): 2.1x fewer H2D bytes and no per-icon Python tensor surgery.
#include <cmath>
#include <cstdint>

#include "../../include/dsvg_b200.h"
#include "common.cuh"

namespace dsvg {
extern unsigned long long g_launches;

constexpr int kRawCols = 14;                                   // difflib/tensor.py:23-32
static const int kArgCols[11] = {1, 2, 3, 4, 5, 8, 9, 10, 11, 12, 13};   // arg_keys order (tensor.py:41), start_pos skipped
constexpr uint8_t kEOS = 4, kSOS = 5;

__global__ void __launch_bounds__(256)
unpack_kernel(const uint8_t* __restrict__ cmd, const int16_t* __restrict__ args, float* __restrict__ cmd_f,
              float* __restrict__ args_f, size_t n_pos, int n_args) {
  pdl_launch_dependents();
  pdl_wait();
  const size_t n_el = n_pos * size_t(n_args);
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_el; i += stride) args_f[i] = float(args[i]);
  for (size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n_pos; i += stride) cmd_f[i] = float(cmd[i]);
}

}  // namespace dsvg
using namespace dsvg;

extern "C" int dsvg_pack_icons(const float* rows, const long long* group_offsets, int n_icons, int max_groups, int seq_len,
                               int grouped, unsigned char* cmd_out, short* args_out) {
  DSVG_CHECK(rows && group_offsets && cmd_out && args_out, "dsvg_pack_icons: null pointer");
  DSVG_CHECK(n_icons > 0 && max_groups > 0 && seq_len > 0, "dsvg_pack_icons: bad shape");
  const int L = seq_len + 2;
  const int G_out = grouped ? 1 : max_groups;
  for (int n = 0; n < n_icons; ++n) {
    for (int go = 0; go < G_out; ++go) {
      // rows of this output sequence: one path, or (grouped) the concatenation of all paths of the icon
      const long long r0 = group_offsets[size_t(n) * max_groups + (grouped ? 0 : go)];
      const long long r1 = group_offsets[size_t(n) * max_groups + (grouped ? max_groups : go + 1)];
      DSVG_CHECK(r1 >= r0, "dsvg_pack_icons: group offsets must be non-decreasing (icon %d)", n);
      const long long len = r1 - r0;
      DSVG_CHECK(len + 2 <= L, "dsvg_pack_icons: icon %d path %d has %lld commands, the window holds %d "
                 "(SVGTensor.pad would return an over-long tensor and torch.stack would fail)", n, go, len, seq_len);
      unsigned char* c = cmd_out + (size_t(n) * G_out + go) * L;
      short* a = args_out + (size_t(n) * G_out + go) * L * 11;
      for (int s = 0; s < L; ++s) {
        const bool body = s >= 1 && s <= len;
        if (!body) {
          c[s] = s == 0 ? kSOS : kEOS;                       // add_sos / add_eos / pad (pad token = EOS)
          for (int k = 0; k < 11; ++k) a[size_t(s) * 11 + k] = -1;   // PAD_VAL rows
          continue;
        }
        const float* src = rows + size_t(r0 + s - 1) * kRawCols;
        const float cv = src[0];
        DSVG_CHECK(cv >= 0.f && cv <= 6.f && cv == std::floor(cv), "dsvg_pack_icons: command id %g is not a token", cv);
        c[s] = static_cast<unsigned char>(cv);
        for (int k = 0; k < 11; ++k) {
          const float v = src[kArgCols[k]];
          DSVG_CHECK(v >= -1.f && v <= 32767.f && v == std::floor(v), "dsvg_pack_icons: argument %g is not an integer in "
                     "[-1, 32767] (the packed format holds numericalised SVGs)", v);
          a[size_t(s) * 11 + k] = static_cast<short>(v);
        }
      }
    }
  }
  return 0;
}

extern "C" int dsvg_unpack_batch(const unsigned char* cmd, const short* args, float* commands_f32, float* args_f32,
                                 size_t n_positions, int n_args, void* stream) {
  DSVG_CHECK(cmd && args && commands_f32 && args_f32 && n_positions > 0 && n_args > 0, "dsvg_unpack_batch: bad arguments");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  size_t blocks = (n_positions * n_args + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  DSVG_CUDA(launch_k(unpack_kernel, dim3(unsigned(blocks)), dim3(256), 0, st, cmd, reinterpret_cast<const int16_t*>(args),
                     commands_f32, args_f32, n_positions, n_args));
  ++g_launches;
  return 0;
}
