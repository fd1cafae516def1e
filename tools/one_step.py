"""Development helper for ncu: 2 warm-up steps + 1 profiled train step of the bench workload."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import synth_icons, WEIGHTS
from deepsvg_b200 import Hierarchical, SVGLoss, SVGTransformer, _lib
B = int(sys.argv[1]) if len(sys.argv) > 1 else 512
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = "cuda:0"
cfg = Hierarchical(use_vae=False)
torch.manual_seed(0)
model = SVGTransformer(cfg, precision=os.environ.get("DSVG_PRECISION", "bf16")).to(dev).train()
loss_fn = SVGLoss(cfg).to(dev)
c, a = synth_icons(B)
c, a = c.to(dev), a.to(dev)
for i in range(steps):
    l0 = _lib.launch_count()
    model.zero_grad(set_to_none=True)
    out = model(c, a, c, a, params={})
    ls = loss_fn(out, None, weights=WEIGHTS)
    ls["loss"].backward()
    torch.cuda.synchronize()
    print("step", i, "launches", _lib.launch_count() - l0, "loss", ls["loss"].item(), flush=True)
