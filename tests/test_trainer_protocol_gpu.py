"""The calls deepsvg/train.py makes, in its order, against deepsvg_b200's modules (the reference trainer itself cannot run on
the GPU box: /root/reference does not travel; and this container has no GPU).  Each step cites the trainer line it mirrors.
Covers: the single warm-up forward (train.py:67-72), nn.DataParallel wrap (:74), model(*model_args, params=...) (:94),
loss_fn(output, labels, weights=...) (:95), optimizer.zero_grad / backward / clip_grad_norm_ / step / schedulers (:92-106),
`.item()` on every loss entry (utils/stats.py:65-66), eval-mode visualisation call (:124-130 -> greedy_sample),
checkpoint save / load through `.module` (utils/train_utils.py:12-13,49,128,152)."""
import io

import pytest
import torch
import torch.nn as nn
from torch.utils.data import DataLoader, Dataset

from oracle import svg_oracle as O

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


class _Icons(Dataset):
    """Stands in for SVGTensorDataset: items are dicts keyed by cfg.model_args (svgtensor_dataset.py:164-205)."""

    def __init__(self, cfg, n, with_label):
        self.cmd, self.arg = O.synth_batch(cfg, n, seed=1)
        self.label = torch.randint(0, cfg.n_labels, (n,), generator=torch.Generator().manual_seed(2)) if with_label else None

    def __len__(self):
        return self.cmd.shape[0]

    def __getitem__(self, i):
        d = {"commands": self.cmd[i], "args": self.arg[i], "commands_grouped": self.cmd[i], "args_grouped": self.arg[i]}
        if self.label is not None:
            d["label"] = self.label[i]
        return d


@pytest.mark.parametrize("variant", ["hierarchical", "hierarchical_vae_label", "self_match"])
def test_reference_trainer_call_sequence(variant):
    from deepsvg_b200 import Hierarchical, HierarchicalSelfMatching, SVGLoss, SVGTransformer
    small = dict(d_model=128, n_heads=4, dim_feedforward=256, dim_z=64, n_layers=2, n_layers_decode=2, max_num_groups=4,
                 max_seq_len=10)
    if variant == "hierarchical":
        cfg = Hierarchical(use_vae=False, **small)
    elif variant == "self_match":
        cfg = HierarchicalSelfMatching(use_vae=False, **small)
    else:
        cfg = Hierarchical(use_vae=True, label_condition=True, n_labels=7, **small)
    model_args = cfg.get_model_args()                                  # model/config.py:47-60
    weights = {"kl_tolerance": 0.1, "loss_kl_weight": 1.0, "loss_cmd_weight": 1.0, "loss_args_weight": 2.0,
               "loss_visibility_weight": 1.0}                          # default_icons.py:66-73
    torch.manual_seed(0)
    model = SVGTransformer(cfg).to(DEV)                                # default_icons.py:59-60, train.py:36-37
    loss_fn = SVGLoss(cfg).to(DEV)                                     # default_icons.py:62-63
    optimizer = torch.optim.AdamW(model.parameters(), lr=2e-3)         # deepsvg/config.py:64-65
    sched = torch.optim.lr_scheduler.StepLR(optimizer, step_size=10, gamma=0.9)   # config.py:67-68
    ds = _Icons(O.make_cfg("hierarchical", n_labels=cfg.n_labels, **small), 48, cfg.label_condition)
    loader = DataLoader(ds, batch_size=8, shuffle=True, drop_last=True)
    data = next(iter(loader))
    model(*[data[a].to(DEV) for a in model_args], params={})           # train.py:67-72 (single warm-up forward)
    model = nn.DataParallel(model)                                     # train.py:74
    losses = []
    for epoch in range(4):
        for data in loader:
            model.train()                                              # train.py:86
            margs = [data[a].to(DEV) for a in model_args]              # :87
            labels = data["label"].to(DEV) if "label" in data else None   # :88
            optimizer.zero_grad()                                      # :92
            output = model(*margs, params={})                          # :94
            loss_dict = loss_fn(output, labels, weights=weights)       # :95
            loss_dict["loss"].backward()                               # :98
            nn.utils.clip_grad_norm_(model.parameters(), 1.0)          # :100
            optimizer.step()                                           # :102
            sched.step()                                               # :104
            vals = {k: v.item() for k, v in loss_dict.items()}         # stats.py:65-66
            assert all(v == v for v in vals.values())
            losses.append(vals["loss_cmd"] + vals["loss_args"])
    assert sum(losses[-4:]) < sum(losses[:4]) - 0.2, (losses[:4], losses[-4:])    # it trains through the wrapper
    # validation hook (train.py:121-130 -> cfg.visualize -> model.module.greedy_sample, default_icons.py:79-97)
    model.eval()
    with torch.no_grad():
        inner = model.module if isinstance(model, nn.DataParallel) else model     # train_utils.py:12-13
        if not cfg.label_condition:
            cy, ay = inner.greedy_sample(margs[0][:1], margs[1][:1], None, None)
            assert cy.shape[0] == 1 and ay.shape[-1] == 11
    # checkpoint round trip (train_utils.py:49,128,152: state_dict of .module, load_state_dict(strict=False))
    buf = io.BytesIO()
    torch.save({"model": inner.state_dict()}, buf)
    buf.seek(0)
    fresh = SVGTransformer(cfg).to(DEV).eval()
    missing, unexpected = fresh.load_state_dict(torch.load(buf)["model"], strict=False)
    assert not missing and not unexpected
    with torch.no_grad():
        kw = {"label": labels} if cfg.label_condition else {}
        if cfg.use_vae:                            # the VAE samples in eval mode too (model.py:182-187): inject the noise
            inner._eps_override = fresh._eps_override = torch.randn(margs[0].shape[0], cfg.dim_z, device=DEV)
        a = inner(margs[0], margs[1], margs[2], margs[3], **kw)["command_logits"]
        b = fresh(margs[0], margs[1], margs[2], margs[3], **kw)["command_logits"]
        # the trained model's cached bf16 weight operands must be casts of its CURRENT fp32 masters
        stale = [n for (n, d_), (_k, w, _wt) in inner._wcache.items()      # (entries of other devices belong to DP replicas)
                 if d_ == inner._param(n).device
                 and not torch.equal(w.t[0, :, :w.cols].float(), inner._param(n).detach().to(torch.bfloat16).float())]
        assert not stale, stale
    assert torch.equal(a, b)
