"""CUDA-event microbenchmark of the weight-gradient kernel over the path-level shapes of the bench configs.
usage: DSVG_OUTER_TALL=0|1 python tools/bench_outer.py   (L2 is flushed by a 512 MB fill between launches)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepsvg_b200 import ops

dev = torch.device("cuda:0")
SHAPES = [(131072, 768, 256), (131072, 256, 256), (131072, 512, 256), (131072, 256, 512), (126976, 2827, 256),
          (270336, 1536, 512), (270336, 512, 512), (13312, 768, 256), (13312, 256, 256), (13312, 512, 256)]
flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)
for M, P, Q in SHAPES:
    A = ops.Act(M, P, 1, dev, ld=(P + 7) // 8 * 8, zero=True)
    A.t.normal_()
    B = ops.Act(M, Q, 1, dev, zero=True)
    B.t.normal_()
    Cw = torch.zeros(P, Q, device=dev)
    cs = torch.zeros(P, device=dev)
    ts = []
    for it in range(12):
        flush.fill_(it)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ops.outer(A, B, M, P, Q, Cw, colsum=cs)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts = sorted(ts[2:])
    med = ts[len(ts) // 2]
    print("outer M=%6d P=%4d Q=%3d  %7.1f us  %6.0f TFLOP/s  %5.2f TB/s algorithmic" % (
        M, P, Q, med, 2.0 * M * P * Q / med / 1e6, 2.0 * M * (P + Q) / med / 1e6), flush=True)
    del A, B
