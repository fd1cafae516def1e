"""`SVGLoss`: drop-in for deepsvg.model.loss.SVGLoss (model/loss.py:9-65).

Same constructor and call (`loss_fn(output, labels, weights=...)` -> dict with "loss", "loss_cmd", "loss_args"
[, "loss_visibility", "loss_kl"], all 0-d tensors).  The cross-entropies run as streaming CUDA kernels that also
leave d(loss)/d(logits) behind for `SVGTransformer`'s backward (no boolean-mask compaction, no host sync, the
1.4 GB logits tensor is read exactly once).  The extended padding mask uses the clean OR-shift-by-3 semantics
(SURVEY.md 8c hazard 1).  Under data parallelism the normalising counts are all-reduced first so that the summed
per-rank gradients equal the single-process gradient of the global batch.
"""
import torch
import torch.nn as nn

from . import ops
from .model import CMD_ARGS_MASK, _r8
from .ops import Act


class _FusedLoss(torch.autograd.Function):
    """Loss terms from the logits produced by SVGTransformer; gradients travel through the LossHandle side channel."""

    @staticmethod
    def forward(ctx, mod, handle, weights, out_dict, token, *logits):
        ctx.set_materialize_grads(False)
        ctx.handle, ctx.weights = handle, weights
        vals = mod._run_kernels(handle, out_dict, weights)
        return tuple(vals[i] for i in range(5))

    @staticmethod
    def backward(ctx, g_loss, g_cmd, g_args, g_vis, g_kl):
        w, h = ctx.weights, ctx.handle
        dev = h.loss_out.device
        zero = torch.zeros((), device=dev)
        gl = g_loss if g_loss is not None else zero

        def eff(weight, g):
            e = gl * float(weight)
            return e + g if g is not None else e

        sc = torch.stack([eff(w.get("loss_args_weight", 0.0), g_args), eff(w.get("loss_cmd_weight", 0.0), g_cmd),
                          eff(w.get("loss_visibility_weight", 0.0), g_vis),
                          eff(w.get("loss_kl_weight", 0.0), g_kl)]).float().contiguous()
        bufs = getattr(h, "bufs", None)
        if bufs is not None:
            bufs["scales"].copy_(sc)       # the captured backward graphs read the static buffer
            sc = bufs["scales"]
        h.scales = sc
        n_logits = len(ctx.needs_input_grad) - 5
        return (None, None, None, None, torch.ones((), device=dev)) + (None,) * n_logits


class _StandaloneLoss(torch.autograd.Function):
    """Same kernels for logits that did not come with a LossHandle (e.g. gathered by DataParallel): the gradient is
    materialised as fp32 tensors for autograd."""

    @staticmethod
    def forward(ctx, mod, holder, weights, out_dict, *logits):
        ctx.set_materialize_grads(False)
        ctx.holder, ctx.weights = holder, weights
        vals = mod._run_kernels(holder, out_dict, weights)
        ctx.shapes = [t.shape for t in logits]
        return tuple(vals[i] for i in range(5))

    @staticmethod
    def backward(ctx, g_loss, g_cmd, g_args, g_vis, g_kl):
        w, h = ctx.weights, ctx.holder
        zero = torch.zeros((), device=h.loss_out.device)
        gl = g_loss if g_loss is not None else zero

        def eff(weight, g):
            e = gl * float(weight)
            return e + g if g is not None else e

        grads = [h.dl_cmd.float() * eff(w.get("loss_cmd_weight", 0.0), g_cmd),
                 h.dl_args.float() * eff(w.get("loss_args_weight", 0.0), g_args)]
        if h.dl_vis is not None:
            grads.append(h.dl_vis.float() * eff(w.get("loss_visibility_weight", 0.0), g_vis))
        if h.mu is not None:
            n = h.mu.numel() * h.world
            ek = eff(w.get("loss_kl_weight", 0.0), g_kl) * h.loss_out[5] / n
            grads.append(ek * h.mu)
            grads.append(-0.5 * ek * (1.0 - torch.exp(h.ls)))
        return (None, None, None, None) + tuple(g.reshape(s) for g, s in zip(grads, ctx.shapes))


class _Holder:
    pass


class SVGLoss(nn.Module):
    def __init__(self, cfg):
        super().__init__()
        self.cfg = cfg
        self.args_dim = 2 * cfg.args_dim if getattr(cfg, "rel_targets", False) else cfg.args_dim + 1   # loss.py:15
        self.register_buffer("cmd_args_mask", CMD_ARGS_MASK.clone())   # loss.py:17
        self.process_group = None

    # -------------------------------------------------------------------------------------------------
    def _run_kernels(self, h, out, weights):
        """Fills h.dl_* (unit-scale gradients) and returns the device vector
        [loss, loss_cmd, loss_args, loss_visibility, loss_kl, kl_active]."""
        cfg = self.cfg
        cl, al = out["command_logits"], out["args_logits"]
        tc, ta = out["tgt_commands"], out["tgt_args"]
        if not cl.is_cuda:
            raise RuntimeError("deepsvg_b200.SVGLoss has no CPU path")
        N, G, Ld, nc = cl.shape
        L = Ld + 1
        na, C = cfg.n_args, self.args_dim
        nseq, Md = N * G, N * G * Ld
        dev = cl.device
        if tc.shape[-1] != L:
            raise ValueError("targets must have %d positions" % L)
        tc = tc.detach().contiguous().float()
        ta = ta.detach().contiguous().float()
        cl2 = cl.detach().contiguous().view(Md, nc)
        al2 = al.detach().contiguous().view(Md, na * C)
        planes = getattr(h, "planes", 1)
        pg = getattr(h, "process_group", None) or self.process_group
        world = 1
        pre = getattr(h, "tgt_prep", None)
        if pre is not None and pre["src"] is out["tgt_commands"] and pg is not None:
            # SVGTransformer.forward already derived the target bookkeeping and started the global-count all-reduce
            import torch.distributed as dist
            world = dist.get_world_size(pg)
            first_eos, visible, counts = pre["first_eos"], pre["visible"], pre["counts"]
            pre["work"].wait()
        else:
            first_eos = torch.empty(nseq, dtype=torch.int32, device=dev)
            visible = torch.empty(nseq, dtype=torch.uint8, device=dev)
            counts = torch.zeros(2, device=dev)
            ops.seq_prep(tc, nseq, L, first_eos, visible, None, None, counts)
            if pg is not None:
                import torch.distributed as dist
                world = dist.get_world_size(pg)
                dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=pg)      # global masked counts
        h.world = world
        acc = torch.zeros(8, device=dev)
        vals = torch.zeros(8, device=dev)
        bufs = getattr(h, "bufs", None)
        h.dl_args = bufs["dl_args"] if bufs is not None else Act(Md, na * C, planes, dev, ld=_r8(na * C))
        h.dl_cmd = bufs["dl_cmd"] if bufs is not None else Act(Md, nc, planes, dev, ld=8)
        ops.ce_args(al2, na * C, tc, ta, counts, h.dl_args, acc, nseq, L, na, C)
        ops.ce_cmd(cl2, tc, first_eos, visible, counts, h.dl_cmd, acc, nseq, L, nc)
        two = cfg.decode_stages == 2
        h.dl_vis = None
        if two:
            vl = out["visibility_logits"].detach().contiguous().view(nseq, 2)
            h.dl_vis = bufs["dl_vis"] if bufs is not None else Act(nseq, 2, planes, dev, ld=8)
            ops.ce_vis(vl, visible, h.dl_vis, acc, nseq, 1.0 / (nseq * world))
        h.mu = h.ls = None
        has_kl = bool(cfg.use_vae)
        if has_kl:
            h.mu = out["mu"].detach().contiguous().view(-1)
            h.ls = out["logsigma"].detach().contiguous().view(-1)
            ops.kl_sum(h.mu, h.ls, acc, h.mu.numel())
        if pg is not None:
            import torch.distributed as dist
            dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=pg)
        ops.loss_finalize(acc, counts, vals, float(weights.get("loss_cmd_weight", 0.0)),
                          float(weights.get("loss_args_weight", 0.0)), float(weights.get("loss_visibility_weight", 0.0)),
                          float(weights.get("loss_kl_weight", 0.0)), float(weights.get("kl_tolerance", 0.0)),
                          1.0 / (nseq * world), 1.0 / (h.mu.numel() * world) if has_kl else 0.0, two, has_kl)
        if bufs is not None:
            bufs["loss_out"].copy_(vals)
            h.loss_out = bufs["loss_out"]
        else:
            h.loss_out = vals
        return vals

    # -------------------------------------------------------------------------------------------------
    def forward(self, output, labels=None, weights=None):
        """loss.py:19-65.  `labels` is unused (as in the reference); `weights` as produced by cfg.get_weights()."""
        cfg = self.cfg
        if weights is None:
            raise ValueError("weights dict required")
        two = cfg.decode_stages == 2
        handle = getattr(output["args_logits"], "_dsvg_handle", None)
        logits = [output["command_logits"], output["args_logits"]]
        if two:
            logits.append(output["visibility_logits"])
        if cfg.use_vae:
            logits += [output["mu"], output["logsigma"]]
        if handle is not None and not handle.used and torch.is_grad_enabled():
            handle.used = True
            vals = _FusedLoss.apply(self, handle, weights, output, handle.token, *logits)
        else:
            holder = _Holder()
            holder.planes = getattr(handle, "planes", 2) if handle is not None else 2
            vals = _StandaloneLoss.apply(self, holder, weights, output, *logits)
        res = {"loss": vals[0], "loss_cmd": vals[1], "loss_args": vals[2]}
        if two:
            res["loss_visibility"] = vals[3]
        if cfg.use_vae:
            res["loss_kl"] = vals[4]
        return res
