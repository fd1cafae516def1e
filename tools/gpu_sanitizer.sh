#!/bin/bash
# compute-sanitizer memcheck over the smoke step (eval, bf16x3) and a small train-mode step in fast mode (eager launches):
# every kernel family of the path runs at least once.  Output -> gpurun_out/<round>_sanitizer.txt
R=${1:-r2}
mkdir -p gpurun_out
OUT=gpurun_out/${R}_sanitizer.txt
echo "== compute-sanitizer --tool memcheck : __graft_entry__.smoke()" > $OUT
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -8 >> $OUT
echo "== compute-sanitizer --tool memcheck : one fast-mode train step (d_model 128, labels, VAE, dropout), eager" >> $OUT
DSVG_GRAPHS=0 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python - >> $OUT 2>&1 <<'PY'
import torch
from oracle import svg_oracle as O
from deepsvg_b200 import Hierarchical, SVGLoss, SVGTransformer, FusedAdamW
small = dict(d_model=128, n_heads=4, dim_feedforward=256, dim_z=64, n_layers=2, n_layers_decode=2, max_num_groups=4, max_seq_len=10)
cfg = Hierarchical(use_vae=True, label_condition=True, n_labels=7, **small)
m = SVGTransformer(cfg).cuda().train()
lf = SVGLoss(cfg).cuda()
opt = FusedAdamW(m.parameters(), lr=1e-3, max_grad_norm=1.0)
c, a = O.synth_batch(O.make_cfg("hierarchical", **small), 6, seed=1)
c, a = c.cuda(), a.cuda()
lab = torch.randint(0, 7, (6,), device="cuda")
W = {"kl_tolerance": 0.1, "loss_kl_weight": 1.0, "loss_cmd_weight": 1.0, "loss_args_weight": 2.0, "loss_visibility_weight": 1.0}
for _ in range(2):
    opt.zero_grad()
    ls = lf(m(c, a, c, a, lab, params={}), None, weights=W)
    ls["loss"].backward()
    opt.step()
torch.cuda.synchronize()
print("train steps ok, loss", ls["loss"].item())
PY
echo "== compute-sanitizer --tool memcheck : parity-mode (bf16x3) train steps: head_dim 32 / L 12 (attn_x3) and head_dim 64 / L 42 (attn_gx3), lean two-plane epilogues, dropout" >> $OUT
DSVG_GRAPHS=0 timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 python - >> $OUT 2>&1 <<'PY'
import torch
from oracle import svg_oracle as O
from deepsvg_b200 import Hierarchical, SVGLoss, SVGTransformer
W = {"kl_tolerance": 0.1, "loss_kl_weight": 1.0, "loss_cmd_weight": 1.0, "loss_args_weight": 2.0, "loss_visibility_weight": 1.0}
for heads, seq in ((4, 10), (2, 40)):
    small = dict(d_model=128, n_heads=heads, dim_feedforward=256, dim_z=64, n_layers=2, n_layers_decode=2, max_num_groups=4, max_seq_len=seq)
    cfg = Hierarchical(use_vae=True, **small)
    m = SVGTransformer(cfg, precision="bf16x3").cuda().train()
    lf = SVGLoss(cfg).cuda()
    c, a = O.synth_batch(O.make_cfg("hierarchical", **small), 6, seed=1)
    c, a = c.cuda(), a.cuda()
    for _ in range(2):
        m.zero_grad(set_to_none=True)
        ls = lf(m(c, a, c, a, params={}), None, weights=W)
        ls["loss"].backward()
    torch.cuda.synchronize()
    print("bf16x3 train steps ok (heads %d, seq %d), loss" % (heads, seq), ls["loss"].item())
PY
tail -16 $OUT
