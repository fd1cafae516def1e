"""CPU oracle for the DeepSVG hot path -- TEST INFRASTRUCTURE ONLY.

A from-scratch, functional, batch-first ("token-major") restatement in plain PyTorch (CPU, fp32/fp64) of what the
reference computes in `SVGTransformer.forward` -> `SVGLoss.forward` -> `.backward()`.  It is the checker for the CUDA
path, never the product: only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs may import it.  The product (`deepsvg_b200`) has no CPU path at all.

Parity status: PINNED.  `tests/golden/make_golden.py` imports the real reference from /root/reference, runs it on
seeded weights/inputs and commits the outputs as fixtures; `tests/test_oracle_golden.py` checks this file against
them (logits, every loss term, every parameter gradient).  The reference has no tests or golden vectors of its
own, so executing it is the only possible pin.

Every function cites the reference lines it restates (paths relative to the reference repo root).
Deliberate deviation: the loss's "extended" padding mask uses the clean OR-shift-by-3 semantics instead of the
reference's aliased in-place add (model/utils.py:25-28), see SURVEY.md 8c hazard 1 -- the golden generator
patches the reference the same way ("de-aliased oracle").
"""
from __future__ import annotations

import math
from types import SimpleNamespace

import torch
import torch.nn.functional as F

# difflib/tensor.py:10 -- token vocabulary
CMD_M, CMD_L, CMD_C, CMD_A, CMD_EOS, CMD_SOS, CMD_Z = range(7)
# difflib/tensor.py:15-21 -- which of the 11 argument slots each command uses
CMD_ARGS_MASK = torch.tensor([[0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1],
                              [0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1],
                              [0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1],
                              [1, 1, 1, 1, 1, 0, 0, 0, 0, 1, 1],
                              [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
                              [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0],
                              [0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0]])


# --------------------------------------------------------------------------------------------------
# configuration (model/config.py:4-108) -- attribute names are the public contract
# --------------------------------------------------------------------------------------------------
def make_cfg(kind="hierarchical", **over):
    c = SimpleNamespace(
        args_dim=256, n_args=11, n_commands=7, dropout=0.1, model_type="transformer",
        encode_stages=1, decode_stages=1, use_resnet=True, use_vae=True, pred_mode="one_shot",
        rel_targets=False, label_condition=False, n_labels=100, dim_label=64, self_match=False,
        n_layers=4, n_layers_decode=4, n_heads=8, dim_feedforward=512, d_model=256, dim_z=256,
        max_num_groups=8, max_seq_len=30)
    if kind == "hierarchical":          # model/config.py:92-98
        c.encode_stages = c.decode_stages = 2
    elif kind != "one_stage":           # model/config.py:83-89
        raise ValueError(kind)
    for k, v in over.items():
        setattr(c, k, v)
    if "max_total_len" not in over:
        c.max_total_len = c.max_num_groups * c.max_seq_len
    if "num_groups_proposal" not in over:
        c.num_groups_proposal = c.max_num_groups
    return c


def param_shapes(cfg):
    """state_dict names/shapes of the reference module tree (model.py:16-309), parameters only."""
    d, dz, ff = cfg.d_model, cfg.dim_z, cfg.dim_feedforward
    out = {}

    def lin(name, o, i):
        out[name + ".weight"] = (o, i)
        out[name + ".bias"] = (o,)

    def ln(name):
        out[name + ".weight"] = (d,)
        out[name + ".bias"] = (d,)

    def layer(p, glob):
        out[p + ".self_attn.in_proj_weight"] = (3 * d, d)
        out[p + ".self_attn.in_proj_bias"] = (3 * d,)
        lin(p + ".self_attn.out_proj", d, d)
        if glob:
            lin(p + ".linear_global", d, dz)
        if cfg.label_condition:
            lin(p + ".linear_global2", d, cfg.dim_label)
        lin(p + ".linear1", ff, d)
        lin(p + ".linear2", d, ff)
        ln(p + ".norm1")
        ln(p + ".norm2")

    def stack(p, n, glob):
        for i in range(n):
            layer(f"{p}.layers.{i}", glob)
        ln(p + ".norm")

    two_e, two_d = cfg.encode_stages == 2, cfg.decode_stages == 2
    enc_len = cfg.max_seq_len if two_e else cfg.max_total_len
    out["encoder.embedding.command_embed.weight"] = (cfg.n_commands, d)
    out["encoder.embedding.arg_embed.weight"] = (cfg.args_dim + 1, 64)
    lin("encoder.embedding.embed_fcn", d, 64 * cfg.n_args)
    if not two_e:
        out["encoder.embedding.group_embed.weight"] = (cfg.max_num_groups + 2, d)
    out["encoder.embedding.pos_encoding.pos_embed.weight"] = (enc_len + 2, d)
    if cfg.label_condition:
        out["encoder.label_embedding.label_embedding.weight"] = (cfg.n_labels, cfg.dim_label)
    stack("encoder.encoder", cfg.n_layers, False)
    if two_e:
        if not getattr(cfg, "self_match", False):                      # model.py:114-115 (permutation-invariant E2 when matching)
            out["encoder.hierarchical_PE.pos_embed.weight"] = (cfg.max_num_groups, d)
        stack("encoder.hierarchical_encoder", cfg.n_layers, False)
    if cfg.use_resnet:
        for i in range(1, 5):
            lin(f"resnet.linear{i}.0", d, d)
    if cfg.use_vae:
        lin("vae.enc_mu_fcn", dz, d)
        lin("vae.enc_sigma_fcn", dz, d)
    else:
        lin("bottleneck.bottleneck", dz, d)
    if cfg.label_condition:
        out["decoder.label_embedding.label_embedding.weight"] = (cfg.n_labels, cfg.dim_label)
    if two_d:
        out["decoder.hierarchical_embedding.PE.pos_embed.weight"] = (cfg.num_groups_proposal, d)
        stack("decoder.hierarchical_decoder", cfg.n_layers_decode, True)
        lin("decoder.hierarchical_fcn.visibility_fcn", 2, d)
        lin("decoder.hierarchical_fcn.z_fcn", dz, d)
    dec_len = (cfg.max_seq_len if two_d else cfg.max_total_len) + 1
    if cfg.pred_mode == "autoregressive":                       # model.py:218-222: the decoder embeds its own (shifted) targets
        out["decoder.embedding.command_embed.weight"] = (cfg.n_commands, d)
        out["decoder.embedding.arg_embed.weight"] = (out_args_dim(cfg), 64)
        lin("decoder.embedding.embed_fcn", d, 64 * cfg.n_args)
        out["decoder.embedding.group_embed.weight"] = (cfg.max_total_len + 2, d)
        out["decoder.embedding.pos_encoding.pos_embed.weight"] = (cfg.max_total_len + 2, d)
    else:
        out["decoder.embedding.PE.pos_embed.weight"] = (dec_len, d)
    stack("decoder.decoder", cfg.n_layers_decode, True)
    lin("decoder.fcn.command_fcn", cfg.n_commands, d)
    lin("decoder.fcn.args_fcn", cfg.n_args * out_args_dim(cfg), d)
    return out


def out_args_dim(cfg):
    """Classes per argument slot (model.py:37,233; loss.py:15): relative targets span -(args_dim-1)..args_dim-1 (+PAD)."""
    return 2 * cfg.args_dim if cfg.rel_targets else cfg.args_dim + 1


def make_params(cfg, seed=0, dtype=torch.float32):
    """Deterministic test weights (NOT the reference initialiser): every tensor non-trivial so that bias / LayerNorm
    affine / embedding paths are all exercised.  Depends only on (name order, shape, seed) and the CPU generator."""
    g = torch.Generator().manual_seed(seed)
    params = {}
    for name, shape in param_shapes(cfg).items():
        t = torch.randn(*shape, generator=g, dtype=torch.float32)
        if name.endswith("norm.weight") or name.endswith("norm1.weight") or name.endswith("norm2.weight"):
            t = 1.0 + 0.1 * t
        elif name.endswith(".bias"):
            t = 0.05 * t
        elif len(shape) == 2 and ("embed" in name and "fcn" not in name or "embedding.weight" in name):
            t = t / math.sqrt(shape[1]) * 1.4
        else:
            t = t / math.sqrt(shape[-1])
        params[name] = t.to(dtype)
    return params


# --------------------------------------------------------------------------------------------------
# synthetic icons (SURVEY.md 8d value distribution)
# --------------------------------------------------------------------------------------------------
def synth_batch(cfg, n, seed=1234, dense=False, one_stage=None):
    """Returns float32 batch-first (commands [n,G,S+2], args [n,G,S+2,11]) like svgtensor_dataset.py:164-205 emits.
    Two-stage: G paths of <= max_seq_len commands, >= 1 visible path per icon.  One-stage: G = 1, a single sequence
    of <= max_total_len commands with 1..3 'm' sub-paths."""
    g = torch.Generator().manual_seed(seed)
    if one_stage is None:
        one_stage = cfg.encode_stages == 1
    G = 1 if one_stage else cfg.max_num_groups
    S = cfg.max_total_len if one_stage else cfg.max_seq_len
    L = S + 2
    cmd = torch.full((n, G, L), float(CMD_EOS))
    arg = torch.full((n, G, L, cfg.n_args), -1.0)
    mask = CMD_ARGS_MASK
    for i in range(n):
        nvis = G if dense else int(torch.randint(1, G + 1, (1,), generator=g))
        for p in range(G):
            cmd[i, p, 0] = CMD_SOS
            if p >= nvis:
                continue
            ln = S if dense else int(torch.randint(3, S + 1, (1,), generator=g))
            body = torch.randint(CMD_L, CMD_C + 1, (ln,), generator=g)
            body[0] = CMD_M
            if one_stage:
                for extra in range(int(torch.randint(0, 3, (1,), generator=g))):
                    body[int(torch.randint(1, ln, (1,), generator=g))] = CMD_M
            cmd[i, p, 1:1 + ln] = body.float()
            vals = torch.randint(0, cfg.args_dim, (ln, cfg.n_args), generator=g).float()
            m = mask[body].float()
            arg[i, p, 1:1 + ln] = vals * m - (1 - m)
    return cmd, arg


# --------------------------------------------------------------------------------------------------
# arithmetic
# --------------------------------------------------------------------------------------------------
class _MM:
    """matmul precision protocol: fp32 | bf16 (operands rounded) | bf16x3 (hi/lo split, 3 products)"""

    def __init__(self, mode):
        assert mode in ("fp32", "bf16", "bf16x3")
        self.mode = mode

    @staticmethod
    def _r(t):
        return t.to(torch.bfloat16).to(t.dtype)

    def mm(self, a, b):  # a @ b over the last/first dims (batched)
        if self.mode == "fp32":
            return a @ b
        ah, bh = self._r(a), self._r(b)
        if self.mode == "bf16":
            return ah @ bh
        al, bl = self._r(a - ah), self._r(b - bh)
        return ah @ bh + ah @ bl + al @ bh

    def linear(self, x, w, b=None):
        y = self.mm(x, w.t())
        return y if b is None else y + b


def _layer_norm(x, w, b):
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)  # improved_transformer.py:35-36 (nn.LayerNorm default eps)


def _drop(x, p):
    """train-mode dropout (torch's own RNG).  p = 0 (parity runs, eval mode) is the identity; p > 0 is used only by
    bench.py's CPU / stock-GPU baseline legs so that the timed arithmetic matches the reference's `model.train()` step."""
    return F.dropout(x, p, training=True) if p > 0 else x


def _self_attention(mm, x, p, pre, n_heads, key_pad, drop=0.0, causal=False):
    """functional.py:92-249 for query is key is value.  x [..., L, d]; key_pad bool [..., L] True = ignore key."""
    d = x.shape[-1]
    hd = d // n_heads
    qkv = mm.linear(x, p[pre + ".in_proj_weight"], p[pre + ".in_proj_bias"])           # :92
    q, k, v = qkv.split(d, dim=-1)
    q = q * (float(hd) ** -0.5)                                                           # :168 (after the bias)
    shp = x.shape[:-1] + (n_heads, hd)
    q, k, v = (t.reshape(shp).transpose(-2, -3) for t in (q, k, v))                       # [..., H, L, hd]
    s = mm.mm(q, k.transpose(-1, -2))                                                     # :228
    if causal:                                                                            # attn_mask = square_subsequent_mask
        Lq = s.shape[-1]                                                                  # (model/utils.py:69-72; functional.py:229)
        s = s.masked_fill(torch.triu(torch.ones(Lq, Lq, dtype=torch.bool, device=s.device), 1), float("-inf"))
    if key_pad is not None:
        s = s.masked_fill(key_pad[..., None, None, :], float("-inf"))                     # :235-240
    a = _drop(torch.softmax(s, dim=-1), drop)                                             # :243-244
    o = mm.mm(a, v).transpose(-2, -3).reshape(x.shape)                                    # :246-248
    return mm.linear(o, p[pre + ".out_proj.weight"], p[pre + ".out_proj.bias"])          # :249


def _layer(mm, x, p, pre, n_heads, key_pad, zglob, lab, drop=0.0, causal=False):
    """Pre-LN block: improved_transformer.py:42-54 (encoder) / :126-141 (decoder with linear_global).
    zglob / lab broadcast over the sequence axis (-2)."""
    h = x + _drop(_self_attention(mm, _layer_norm(x, p[pre + ".norm1.weight"], p[pre + ".norm1.bias"]), p,
                                  pre + ".self_attn", n_heads, key_pad, drop, causal), drop)
    if zglob is not None:
        h = h + _drop(mm.linear(zglob, p[pre + ".linear_global.weight"], p[pre + ".linear_global.bias"]), drop).unsqueeze(-2)
    if lab is not None:
        h = h + _drop(mm.linear(lab, p[pre + ".linear_global2.weight"], p[pre + ".linear_global2.bias"]), drop).unsqueeze(-2)
    f = _layer_norm(h, p[pre + ".norm2.weight"], p[pre + ".norm2.bias"])
    f = mm.linear(_drop(torch.relu(mm.linear(f, p[pre + ".linear1.weight"], p[pre + ".linear1.bias"])), drop),
                  p[pre + ".linear2.weight"], p[pre + ".linear2.bias"])
    return h + _drop(f, drop)


def _stack(mm, x, p, pre, n_layers, n_heads, key_pad=None, zglob=None, lab=None, drop=0.0, causal=False):
    """transformer.py:168-188 / :214-242: L layers then the final LayerNorm."""
    for i in range(n_layers):
        x = _layer(mm, x, p, f"{pre}.layers.{i}", n_heads, key_pad, zglob, lab, drop, causal)
    return _layer_norm(x, p[pre + ".norm.weight"], p[pre + ".norm.bias"])


# --------------------------------------------------------------------------------------------------
# masks (model/utils.py:7-66), batch-first, sequence on the last axis
# --------------------------------------------------------------------------------------------------
def key_padding(cmd):      # True from the first EOS on            (model/utils.py:7-17)
    return (cmd == CMD_EOS).cumsum(-1) > 0


def visibility(cmd):       # path has at least one real command    (model/utils.py:45-56)
    return (cmd == CMD_EOS).sum(-1) < cmd.shape[-1] - 1


def extended_padding(cmd):
    """model/utils.py:20-32 with extended=True, clean semantics: ext[i] = min(1, pad[i] + pad[i-3])."""
    pad = (~key_padding(cmd)).float()
    ext = pad.clone()
    ext[..., 3:] = torch.clamp(pad[..., 3:] + pad[..., :-3], max=1.0)
    return ext


def group_index(cmd):      # number of "m" so far                  (model/utils.py:35-42)
    return (cmd == CMD_M).cumsum(-1)


# --------------------------------------------------------------------------------------------------
# forward (model.py:352-412), eval mode (no dropout); VAE noise is an input
# --------------------------------------------------------------------------------------------------
def forward(params, cfg, commands, args, label=None, eps=None, matmul="fp32", z_in=None, train_dropout=False,
            commands_dec=None, args_dec=None):
    """commands [N,G,L] / args [N,G,L,11] float (encoder == decoder inputs, as model/config.py:47-60 wires them).
    Returns the reference's result dict (batch-first) plus 'z' [N, dz].
    train_dropout: draw the reference's train-mode dropout masks (cfg.dropout; 0.1 at the positional encodings,
    positional_encoding.py:26) -- baseline timing only, parity always runs eval-mode arithmetic."""
    mm = _MM(matmul)
    dr = float(cfg.dropout) if train_dropout else 0.0
    dr_pe = 0.1 if train_dropout else 0.0
    p = params
    dt = p["decoder.fcn.args_fcn.weight"].dtype
    N, G, L = commands.shape
    H = cfg.n_heads
    cmd = commands.long()
    two_e, two_d = cfg.encode_stages == 2, cfg.decode_stages == 2
    res = {}

    if z_in is None:
        # ---- SVGEmbedding (model.py:46-57) ----
        a_idx = (args + 1).long()
        emb = p["encoder.embedding.arg_embed.weight"][a_idx].reshape(N, G, L, -1)
        x = p["encoder.embedding.command_embed.weight"][cmd] + \
            mm.linear(emb, p["encoder.embedding.embed_fcn.weight"], p["encoder.embedding.embed_fcn.bias"])
        if not two_e:
            x = x + p["encoder.embedding.group_embed.weight"][group_index(cmd)]
        x = _drop(x + p["encoder.embedding.pos_encoding.pos_embed.weight"][:L], dr_pe)  # positional_encoding.py:40-43
        kp = key_padding(cmd)
        lab_e = p["encoder.label_embedding.label_embedding.weight"][label] if cfg.label_condition else None
        # ---- E1 (model.py:135) + masked mean over positions (:137) ----
        mem = _stack(mm, x, p, "encoder.encoder", cfg.n_layers, H, kp,
                     lab=None if lab_e is None else lab_e[:, None, :].expand(N, G, -1), drop=dr)
        w = (~kp).to(dt).unsqueeze(-1)
        z = (mem * w).sum(-2) / w.sum(-2)                                                 # [N,G,d]
        if two_e:
            vis = visibility(cmd)                                                         # [N,G]
            if not getattr(cfg, "self_match", False):
                z = _drop(z + p["encoder.hierarchical_PE.pos_embed.weight"][:G], dr_pe)    # model.py:157-158
            mem2 = _stack(mm, z, p, "encoder.hierarchical_encoder", cfg.n_layers, H, ~vis, lab=lab_e, drop=dr)   # :160
            wv = vis.to(dt).unsqueeze(-1)
            z = (mem2 * wv).sum(-2) / wv.sum(-2)                                           # :161  [N,d]
        else:
            z = z[:, 0]
        # ---- ResNet (basic_blocks.py:59-65) ----
        if cfg.use_resnet:
            for i in range(1, 5):
                z = z + torch.relu(mm.linear(z, p[f"resnet.linear{i}.0.weight"], p[f"resnet.linear{i}.0.bias"]))
        # ---- VAE (model.py:182-187) / Bottleneck (:196-197) ----
        if cfg.use_vae:
            mu = mm.linear(z, p["vae.enc_mu_fcn.weight"], p["vae.enc_mu_fcn.bias"])
            ls = mm.linear(z, p["vae.enc_sigma_fcn.weight"], p["vae.enc_sigma_fcn.bias"])
            if eps is None:
                eps = torch.zeros_like(mu)
            z = mu + torch.exp(ls / 2.0) * eps
            res["mu"], res["logsigma"] = mu.reshape(N, 1, 1, -1), ls.reshape(N, 1, 1, -1)
        else:
            z = mm.linear(z, p["bottleneck.bottleneck.weight"], p["bottleneck.bottleneck.bias"])
    else:
        z = z_in
    res["z"] = z

    # ---- Decoder (model.py:243-285) ----
    lab_d = p["decoder.label_embedding.label_embedding.weight"][label] if cfg.label_condition else None
    if two_d:
        Gp = cfg.num_groups_proposal
        src = _drop(p["decoder.hierarchical_embedding.PE.pos_embed.weight"][:Gp].unsqueeze(0).expand(N, Gp, -1), dr_pe)   # :251
        out = _stack(mm, src, p, "decoder.hierarchical_decoder", cfg.n_layers_decode, H, None, zglob=z, lab=lab_d, drop=dr)
        vis_logits = mm.linear(out, p["decoder.hierarchical_fcn.visibility_fcn.weight"],
                               p["decoder.hierarchical_fcn.visibility_fcn.bias"])          # basic_blocks.py:36
        zp = mm.linear(out, p["decoder.hierarchical_fcn.z_fcn.weight"], p["decoder.hierarchical_fcn.z_fcn.bias"])
        res["visibility_logits"] = vis_logits.reshape(N, Gp, 1, 2)
        zmem, Gd = zp, Gp                                                                   # [N,Gp,dz]
        lab_d1 = None if lab_d is None else lab_d[:, None, :].expand(N, Gp, -1)
    else:
        zmem, Gd = z[:, None, :], 1
        lab_d1 = None if lab_d is None else lab_d[:, None, :]
    Ld = (cfg.max_seq_len if two_d else cfg.max_total_len) + 1
    commands_dec = commands if commands_dec is None else commands_dec
    args_dec = args if args_dec is None else args_dec
    if cfg.pred_mode == "autoregressive":                                                  # model.py:262-272 (transformer)
        cd, ad = commands_dec[..., :-1].long(), args_dec[..., :-1, :]                      # teacher forcing: drop the last (:372)
        Ld = cd.shape[-1]
        emb = p["decoder.embedding.arg_embed.weight"][(ad + 1).long()].reshape(N, Gd, Ld, -1)
        src = p["decoder.embedding.command_embed.weight"][cd] + \
            mm.linear(emb, p["decoder.embedding.embed_fcn.weight"], p["decoder.embedding.embed_fcn.bias"]) + \
            p["decoder.embedding.group_embed.weight"][group_index(cd)]
        src = _drop(src + p["decoder.embedding.pos_encoding.pos_embed.weight"][:Ld], dr_pe)
        out = _stack(mm, src, p, "decoder.decoder", cfg.n_layers_decode, H, key_padding(cd), zglob=zmem, lab=lab_d1, drop=dr,
                     causal=True)                                                          # :269
    else:
        src = _drop(p["decoder.embedding.PE.pos_embed.weight"][:Ld].reshape(1, 1, Ld, -1).expand(N, Gd, Ld, -1), dr_pe)   # :278
        out = _stack(mm, src, p, "decoder.decoder", cfg.n_layers_decode, H, None, zglob=zmem, lab=lab_d1, drop=dr)        # :279
    res["command_logits"] = mm.linear(out, p["decoder.fcn.command_fcn.weight"], p["decoder.fcn.command_fcn.bias"])
    al = mm.linear(out, p["decoder.fcn.args_fcn.weight"], p["decoder.fcn.args_fcn.bias"])
    res["args_logits"] = al.reshape(N, Gd, Ld, cfg.n_args, out_args_dim(cfg))               # basic_blocks.py:21
    if getattr(cfg, "self_match", False) and two_d and z_in is None:                        # model.py:384-394
        asg = perfect_matching(res["command_logits"].detach(), res["args_logits"].detach(),
                               res["visibility_logits"].detach(), commands[..., 1:], args[..., 1:, :], cfg)
        res["assignment"] = asg
        for k in ("command_logits", "args_logits", "visibility_logits"):
            t = res[k]
            idx = asg.reshape(asg.shape + (1,) * (t.dim() - 2)).expand_as(t)
            res[k] = torch.gather(t, 1, idx)
    res["tgt_commands"], res["tgt_args"] = commands_dec, args_dec                          # model.py:404-405
    return res


# --------------------------------------------------------------------------------------------------
# Hungarian self-matching (model.py:311-350, cfg.self_match; model/config.py:101-108)
# --------------------------------------------------------------------------------------------------
def matching_costs(cl, al, vl, tgt_c, tgt_a):
    """cost[n, g, p] of explaining target path g with predicted slot p (model.py:313-337): 2 * masked-mean args CE +
    masked-mean command CE + visibility CE.  tgt_c / tgt_a are the SHIFTED targets (commands[..., 1:], args[..., 1:, :]), and
    -- as in the reference -- visibility and the extended padding mask are taken on those shifted sequences (so a path of a
    single command counts as invisible here, unlike in the loss).  Also returns the visibility mask [N, G]."""
    N, G, S, na = tgt_a.shape
    Gp = cl.shape[1]
    tc = tgt_c.long()
    vis = (tc == CMD_EOS).sum(-1) < S - 1                                                   # model/utils.py:45-56
    pad = extended_padding(tc).to(cl.dtype) * vis.unsqueeze(-1).to(cl.dtype)                # model.py:315
    lc_all = F.log_softmax(cl, -1)[:, None].expand(N, G, Gp, S, cl.shape[-1])
    la_all = F.log_softmax(al, -1)[:, None].expand(N, G, Gp, S, na, al.shape[-1])
    lv_all = F.log_softmax(vl.reshape(N, Gp, 2), -1)[:, None].expand(N, G, Gp, 2)
    ce_c = -lc_all.gather(-1, tc[:, :, None, :, None].expand(N, G, Gp, S, 1)).squeeze(-1)
    ce_a = -la_all.gather(-1, (tgt_a.long() + 1)[:, :, None, :, :, None].expand(N, G, Gp, S, na, 1)).squeeze(-1)
    ce_v = -lv_all.gather(-1, vis.long()[:, :, None, None].expand(N, G, Gp, 1)).squeeze(-1)
    mask = CMD_ARGS_MASK.to(tc.device)[tc].to(cl.dtype)                                     # [N,G,S,na]
    la = (ce_a * mask[:, :, None]).sum((-1, -2)) / mask.sum((-1, -2))[:, :, None]           # :334
    lc = (ce_c * pad[:, :, None]).sum(-1) / pad.sum(-1)[:, :, None]                         # :335
    return 2.0 * la + 1.0 * lc + 1.0 * ce_v, vis                                            # :337


def perfect_matching(cl, al, vl, tgt_c, tgt_a, cfg):
    """model.py:311-350.  Returns assignment [N, Gp] (long): output slot i of icon n takes predicted slot assignment[n, i].
    Rows of the cost matrix are the VISIBLE targets in order (compacted, costs[mask]); the unassigned slots follow in
    ascending order (`assign + list(full_set - set(assign))`).  The optimal assignment itself comes from the reference's own
    third-party solver, scipy.optimize.linear_sum_assignment (model.py:13,344)."""
    from scipy.optimize import linear_sum_assignment
    with torch.no_grad():
        cost, vis = matching_costs(cl, al, vl, tgt_c, tgt_a)
    Gp = cfg.num_groups_proposal
    rows = []
    for i in range(cost.shape[0]):
        _, assign = linear_sum_assignment(cost[i][vis[i]].cpu())
        assign = assign.tolist()
        rows.append(assign + sorted(set(range(Gp)) - set(assign)))
    return torch.tensor(rows, device=cl.device)


# --------------------------------------------------------------------------------------------------
# loss (model/loss.py:19-65)
# --------------------------------------------------------------------------------------------------
DEFAULT_WEIGHTS = {"kl_tolerance": 0.1, "loss_kl_weight": 1.0, "loss_cmd_weight": 1.0, "loss_args_weight": 2.0,
                   "loss_visibility_weight": 1.0}


def loss(out, cfg, weights=DEFAULT_WEIGHTS):
    res = {}
    total = 0.0
    if cfg.use_vae:                                                                        # loss.py:24-30
        mu, ls = out["mu"], out["logsigma"]
        kl = (-0.5 * torch.mean(1 + ls - mu.pow(2) - torch.exp(ls))).clamp(min=weights["kl_tolerance"])
        total = total + weights["loss_kl_weight"] * kl
        res["loss_kl"] = kl
    tc = out["tgt_commands"].long()
    ta = out["tgt_args"]
    vis = visibility(tc)                                                                   # :35
    wc = (extended_padding(tc) * vis.unsqueeze(-1).float())[..., 1:]                       # :36, :49
    if cfg.decode_stages == 2:                                                             # :41-46
        lv = F.cross_entropy(out["visibility_logits"].reshape(-1, 2), vis.reshape(-1).long())
        total = total + weights["loss_visibility_weight"] * lv
        res["loss_visibility"] = lv
    tc1, ta1 = tc[..., 1:], ta[..., 1:, :]
    wa = CMD_ARGS_MASK.to(tc1.device)[tc1].to(wc.dtype)                                    # :51
    cl, al = out["command_logits"], out["args_logits"]
    ce_c = F.cross_entropy(cl.reshape(-1, cl.shape[-1]), tc1.reshape(-1), reduction="none").reshape(tc1.shape)
    ce_a = F.cross_entropy(al.reshape(-1, al.shape[-1]), (ta1.long() + 1).reshape(-1),
                           reduction="none").reshape(ta1.shape)
    lc = (ce_c * wc).sum() / wc.sum()                                                      # :53 (mean over selected)
    la = (ce_a * wa).sum() / wa.sum()                                                      # :54
    total = total + weights["loss_cmd_weight"] * lc + weights["loss_args_weight"] * la    # :56-57
    res.update(loss=total, loss_cmd=lc, loss_args=la)
    return res


def train_step(params, cfg, commands, args, label=None, eps=None, weights=DEFAULT_WEIGHTS, matmul="fp32",
               train_dropout=False, commands_dec=None, args_dec=None):
    """forward + loss + backward (train.py:94-98; eval-mode arithmetic unless train_dropout).  Returns (out, losses, grads)."""
    leaves = {k: v.detach().clone().requires_grad_(True) for k, v in params.items()}
    out = forward(leaves, cfg, commands, args, label=label, eps=eps, matmul=matmul, train_dropout=train_dropout,
                  commands_dec=commands_dec, args_dec=args_dec)
    ls = loss(out, cfg, weights)
    ls["loss"].backward()
    grads = {k: (v.grad if v.grad is not None else torch.zeros_like(v)) for k, v in leaves.items()}
    return ({k: v.detach() for k, v in out.items()}, {k: v.detach() for k, v in ls.items()}, grads)
