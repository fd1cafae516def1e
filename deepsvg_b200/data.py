"""Input side of the hot path: the packed batch format.

`pack_icons` replaces, for a whole batch, what `SVGTensorDataset.get_data` + the default collate do per icon in Python
(svgtensor_dataset.py:164-205; `SVGTensor.add_eos/add_sos/pad`, difflib/tensor.py:108-143): one native call writes command
ids as uint8 and arguments as int16 (23 bytes per position instead of 48) into pinned host memory; `PackedBatch.cuda()` is
one H2D copy per tensor and `unpack()` one CUDA kernel that yields the float32 `commands` / `args` tensors
`SVGTransformer.forward` takes.
"""
import ctypes as C

import torch

from . import _lib


class PackedBatch:
    """cmd: uint8 [N, G, L]; args: int16 [N, G, L, 11] (-1 = PAD); label: int64 [N] or None."""

    def __init__(self, cmd, args, label=None):
        self.cmd, self.args, self.label = cmd, args, label

    @property
    def nbytes(self):
        return self.cmd.numel() + 2 * self.args.numel() + (8 * self.label.numel() if self.label is not None else 0)

    def pin_memory(self):
        return PackedBatch(self.cmd.pin_memory(), self.args.pin_memory(),
                           self.label.pin_memory() if self.label is not None else None)

    def cuda(self, device=None, non_blocking=True, out=None):
        """Host -> device.  `out` (a device PackedBatch of the same shape) is reused when given."""
        if out is not None:
            out.cmd.copy_(self.cmd, non_blocking=non_blocking)
            out.args.copy_(self.args, non_blocking=non_blocking)
            if self.label is not None:
                out.label.copy_(self.label, non_blocking=non_blocking)
            return out
        dev = torch.device("cuda" if device is None else device)
        return PackedBatch(self.cmd.to(dev, non_blocking=non_blocking), self.args.to(dev, non_blocking=non_blocking),
                           self.label.to(dev, non_blocking=non_blocking) if self.label is not None else None)

    def unpack(self, out=None):
        """Device packed batch -> (commands float32 [N,G,L], args float32 [N,G,L,11]) in one kernel launch."""
        if not self.cmd.is_cuda:
            raise RuntimeError("deepsvg_b200.PackedBatch.unpack runs on the GPU: call .cuda() first (no CPU path)")
        if out is None:
            out = (torch.empty(self.cmd.shape, dtype=torch.float32, device=self.cmd.device),
                   torch.empty(self.args.shape, dtype=torch.float32, device=self.cmd.device))
        rc = _lib.load().dsvg_unpack_batch(self.cmd.data_ptr(), self.args.data_ptr(), out[0].data_ptr(), out[1].data_ptr(),
                                           self.cmd.numel(), self.args.shape[-1],
                                           torch.cuda.current_stream().cuda_stream)
        _lib.check(rc, "dsvg_unpack_batch")
        return out


def pack_icons(icons, max_num_groups, seq_len, grouped=False, labels=None, pin=True):
    """icons: list (batch) of lists (paths) of raw (len, 14) float tensors, exactly what `SVGTensorDataset._load_tensor`
    / `SVG.to_tensor(concat_groups=False)` hand to `get_data` (svgtensor_dataset.py:152-162).
    grouped=False: per-path tensors, seq_len = MAX_SEQ_LEN (fields `commands`, `args`);
    grouped=True : one concatenated sequence per icon, seq_len = MAX_TOTAL_LEN (fields `commands_grouped`, `args_grouped`)."""
    n = len(icons)
    offsets = [0]
    chunks = []
    for paths in icons:
        if len(paths) > max_num_groups:
            raise ValueError("an icon has %d paths, max_num_groups is %d" % (len(paths), max_num_groups))
        for g in range(max_num_groups):
            if g < len(paths) and paths[g].numel() > 0:
                t = paths[g].reshape(-1, 14)
                chunks.append(t)
                offsets.append(offsets[-1] + t.shape[0])
            else:
                offsets.append(offsets[-1])
    rows = torch.cat(chunks, 0).float().contiguous() if chunks else torch.zeros(1, 14)
    off = torch.tensor(offsets, dtype=torch.int64)
    G = 1 if grouped else max_num_groups
    cmd = torch.empty(n, G, seq_len + 2, dtype=torch.uint8)
    args = torch.empty(n, G, seq_len + 2, 11, dtype=torch.int16)
    if pin and torch.cuda.is_available():
        cmd, args = cmd.pin_memory(), args.pin_memory()
    rc = _lib.load().dsvg_pack_icons(rows.data_ptr(), off.data_ptr(), n, max_num_groups, seq_len, 1 if grouped else 0,
                                     cmd.data_ptr(), args.data_ptr())
    _lib.check(rc, "dsvg_pack_icons")
    lab = None
    if labels is not None:
        lab = torch.as_tensor(labels, dtype=torch.int64).reshape(n)
        if pin and torch.cuda.is_available():
            lab = lab.pin_memory()
    return PackedBatch(cmd, args, lab)


def pack_tensors(commands, args, label=None, pin=True):
    """Already-collated float batches (what a DataLoader over SVGTensorDataset yields) -> PackedBatch (host, lossless for
    numericalised SVGs: command ids 0..6, arguments -1..32767)."""
    cmd = commands.to(torch.uint8)
    a16 = args.to(torch.int16)
    if not (torch.equal(cmd.float(), commands.float()) and torch.equal(a16.float(), args.float())):
        raise ValueError("pack_tensors: values are not integral tokens / arguments")
    if pin and torch.cuda.is_available():
        cmd, a16 = cmd.pin_memory(), a16.pin_memory()
        label = label.pin_memory() if label is not None else None
    return PackedBatch(cmd.contiguous(), a16.contiguous(), label)
