"""`FusedAdamW`: torch.optim.AdamW semantics (+ optional global-norm gradient clipping) in two CUDA launches.

Replaces `optim.AdamW(model.parameters(), lr)` (deepsvg/config.py:64-65) and, when `max_grad_norm` is given, the
`clip_grad_norm_` call of deepsvg/train.py:100.  State lives in two flat fp32 buffers; a small device table of
(param, grad, exp_avg, exp_avg_sq, numel) rows is refreshed each step (gradient tensors change identity every backward).
"""
import math

import torch

from . import _lib


class FusedAdamW(torch.optim.Optimizer):
    RING = 4

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, max_grad_norm=None):
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, max_grad_norm=max_grad_norm)
        super().__init__(params, defaults)
        self._flat = {}

    def _group_state(self, gi, group):
        st = self._flat.get(gi)
        ps = [p for p in group["params"] if p.requires_grad]
        if st is not None and st["ids"] != [id(p) for p in ps]:
            # the set of trainable parameters changed (requires_grad toggled, add_param_group): the flat moment buffers
            # are laid out per parameter, so carry the existing moments over by identity and rebuild the table
            st = self._rebuild(st, ps)
            self._flat[gi] = st
        if st is None or st["device"] != ps[0].device:
            n = sum(p.numel() for p in ps)
            dev = ps[0].device
            st = dict(device=dev, step=0, m=torch.zeros(n, device=dev), v=torch.zeros(n, device=dev),
                      sq=torch.zeros(1, device=dev))
            off, rows = 0, []
            for p in ps:
                rows.append((off, p.numel()))
                off += p.numel()
            st["rows"] = rows
            st["ids"] = [id(p) for p in ps]
            # Ring of pinned pointer tables: the H2D copy of slot i is asynchronous, so the host may only rewrite slot
            # i after the copy enqueued from it has executed (event per slot).  The GPU-bound train loop lets the host
            # run several steps ahead; a single table would be overwritten under a queued copy.
            st["host"] = [torch.zeros(len(ps), 5, dtype=torch.int64).pin_memory() for _ in range(self.RING)]
            st["dev"] = [torch.zeros(len(ps), 5, dtype=torch.int64, device=dev) for _ in range(self.RING)]
            st["done"] = [None] * self.RING
            st["slot"] = 0
            st["chunks"] = min(64, max(1, math.ceil(max(p.numel() for p in ps) / 4096)))
            self._flat[gi] = st
        return st, ps

    def _rebuild(self, old, ps):
        dev = ps[0].device
        n = sum(p.numel() for p in ps)
        st = dict(device=dev, step=old["step"], m=torch.zeros(n, device=dev), v=torch.zeros(n, device=dev),
                  sq=torch.zeros(1, device=dev))
        where = {i: r for i, r in zip(old["ids"], old["rows"])}
        off, rows = 0, []
        for p in ps:
            rows.append((off, p.numel()))
            prev = where.get(id(p))
            if prev is not None and old["device"] == dev:
                st["m"][off:off + p.numel()].copy_(old["m"][prev[0]:prev[0] + prev[1]])
                st["v"][off:off + p.numel()].copy_(old["v"][prev[0]:prev[0] + prev[1]])
            off += p.numel()
        st["rows"], st["ids"] = rows, [id(p) for p in ps]
        st["host"] = [torch.zeros(len(ps), 5, dtype=torch.int64).pin_memory() for _ in range(self.RING)]
        st["dev"] = [torch.zeros(len(ps), 5, dtype=torch.int64, device=dev) for _ in range(self.RING)]
        st["done"] = [None] * self.RING
        st["slot"] = 0
        st["chunks"] = min(64, max(1, math.ceil(max(p.numel() for p in ps) / 4096)))
        return st

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.load()
        stream = torch.cuda.current_stream().cuda_stream
        for gi, group in enumerate(self.param_groups):
            st, ps = self._group_state(gi, group)
            if not ps[0].is_cuda:
                raise RuntimeError("deepsvg_b200.FusedAdamW has no CPU path")
            slot = st["slot"]
            st["slot"] = (slot + 1) % self.RING
            if st["done"][slot] is not None:
                st["done"][slot].synchronize()      # the copy that last read this pinned slot has executed
            host, table = st["host"][slot], st["dev"][slot]
            keep = []
            for i, (p, (off, n)) in enumerate(zip(ps, st["rows"])):
                g = p.grad
                if g is None:
                    host[i, 4] = 0
                    continue
                if not g.is_contiguous():
                    g = g.contiguous()
                    keep.append(g)
                host[i, 0], host[i, 1] = p.data_ptr(), g.data_ptr()
                host[i, 2], host[i, 3] = st["m"].data_ptr() + 4 * off, st["v"].data_ptr() + 4 * off
                host[i, 4] = n
            table.copy_(host, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record()
            st["done"][slot] = ev
            st["step"] += 1
            b1, b2 = group["betas"]
            t = st["step"]
            mx = group.get("max_grad_norm")
            sq = 0
            if mx is not None and mx > 0:
                st["sq"].zero_()
                _lib.check(lib.dsvg_grad_sqnorm(table.data_ptr(), len(ps), st["chunks"], st["sq"].data_ptr(), stream),
                           "dsvg_grad_sqnorm")
                sq = st["sq"].data_ptr()
            _lib.check(lib.dsvg_adamw_step(table.data_ptr(), len(ps), st["chunks"], group["lr"], b1, b2, group["eps"],
                                           group["weight_decay"], 1.0 - b1 ** t, 1.0 - b2 ** t,
                                           float(mx) if mx else 0.0, sq, stream), "dsvg_adamw_step")
            self._keep = keep   # contiguous gradient copies must outlive the asynchronous launch
            # The kernels wrote the parameters through raw pointers: tell autograd / version-keyed caches (the model's
            # bf16 weight-operand cache is keyed on `_version`) that every parameter changed.
            torch._C._increment_version([p for p in ps if p.grad is not None])
        return loss
