"""Development helper: per-parameter gradient error of the CUDA path vs the oracle (parity mode), sorted."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import svg_oracle as O
from tests.test_model_gpu import _build, _run, CASES

name = sys.argv[1] if len(sys.argv) > 1 else "hier"
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16x3"
kind, over, n = CASES[name]
cfg = O.make_cfg(kind, **over)
model, loss_fn, params = _build(cfg, prec)
cmd, arg = O.synth_batch(cfg, n, seed=21)
label = torch.randint(0, cfg.n_labels, (n,), generator=torch.Generator().manual_seed(2)) if cfg.label_condition else None
eps = torch.randn(n, cfg.dim_z, generator=torch.Generator().manual_seed(3)) if cfg.use_vae else None
out, ls, grads = _run(model, loss_fn, cmd, arg, label, eps)
ro, rl, rg = O.train_step(params, cfg, cmd, arg, label=label, eps=eps)
ro64, rl64, rg64 = O.train_step({k: v.double() for k, v in params.items()}, cfg, cmd.double(), arg.double(), label=label,
                                eps=None if eps is None else eps.double())
rows = []
for k, g in rg64.items():
    d = g.norm().item() + 1e-30
    rows.append((k, (grads[k].double() - g).norm().item() / d, (rg[k].double() - g).norm().item() / d))
rows.sort(key=lambda r: -r[1])
print("param | cuda-vs-fp64 | fp32oracle-vs-fp64")
for r in rows[:12]:
    print("%-60s %.2e %.2e" % r)
print("logits max err vs fp64:", (out["args_logits"].detach().cpu().double() - ro64["args_logits"]).abs().max().item(),
      " fp32 oracle:", (ro["args_logits"].double() - ro64["args_logits"]).abs().max().item())
print({k: (ls[k].item(), rl64[k].item()) for k in rl64})
