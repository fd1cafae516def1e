#!/bin/bash
mkdir -p gpurun_out
R=r2k
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --no-header -x -k "lean_epilogues or full_epilogue or linear_matches" 2>&1 | tail -8
timeout 120 python - <<'PY'
import os, sys, torch
sys.path.insert(0, os.getcwd())
from deepsvg_b200 import ops
dev = torch.device("cuda:0")
M = 131072
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
X = ops.Act(M, 256, 1, dev, zero=True); X.t.normal_()
W = ops.Act(512, 256, 1, dev, zero=True); W.t.normal_(std=0.06)
mk = ops.Act(M, 512, 1, dev, zero=True); mk.t.normal_(); mk.t[mk.t < 0] = 0
outs = [ops.Act(M, 512, 1, dev) for _ in range(2)]
i = [0]
def f():
    i[0] ^= 1
    ops.linear(X, W, M, 512, 256, mask=mk, mask_scale=1.1, out_act=outs[i[0]])
print("mask dgrad (mode 5) M=131072 N=512 K=256: %.1f us" % timeit(f))
ref = (X.float() @ W.float().t()) * 1.1 * (mk.float() != 0)
err = (outs[i[0]].float() - ref).abs().max().item() / ref.abs().max().item()
print("rel err vs fp32 torch: %.2e" % err)
PY
timeout 900 python -m pytest tests -m gpu -q --no-header 2>&1 | tail -5
timeout 600 python bench.py --no-cpu-baseline --no-parity-mode --no-ref-gpu 2>/dev/null | tail -1 | cut -c1-330
