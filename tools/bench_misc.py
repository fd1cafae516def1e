"""Development microbenchmarks (CUDA events): LayerNorm backward and the general tensor-core attention at the BASELINE shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepsvg_b200 import ops

dev = torch.device("cuda:0")


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for M, D in ((131072, 256), (270336, 512)):
    x, dxin = torch.randn(M, D, device=dev), torch.randn(M, D, device=dev)
    g = torch.ones(D, device=dev)
    mean, rstd = torch.zeros(M, device=dev), torch.ones(M, device=dev)
    dy = ops.Act(M, D, 1, dev, zero=True)
    dy.t.normal_()
    dx, dact = torch.empty(M, D, device=dev), ops.Act(M, D, 1, dev)
    dg, db = torch.zeros(D, device=dev), torch.zeros(D, device=dev)
    t = timeit(lambda: ops.ln_bwd(x, mean, rstd, g, M, D, dy=dy, dx_in=dxin, dx_out=dx, dact=dact, drop=(0.1, 3, 5), dgamma=dg,
                                  dbeta=db))
    print("ln_bwd  M=%d D=%d  %.1f us  %.2f TB/s" % (M, D, t, M * D * 16 / t / 1e6), flush=True)
    y = ops.Act(M, D, 1, dev)
    t = timeit(lambda: ops.ln_fwd(x, g, g, y, mean, rstd, M, D))
    print("ln_fwd  M=%d D=%d  %.1f us  %.2f TB/s" % (M, D, t, M * D * 6 / t / 1e6), flush=True)
    del x, dxin, dy, dx, dact, y

for nseq, L, H, hd in ((4096, 66, 8, 64), (4096, 65, 8, 64), (256, 52, 8, 32), (256, 16, 8, 64), (4096, 32, 8, 32)):
    d, M = H * hd, nseq * L
    qkv = ops.Act(M, 3 * d, 1, dev, zero=True)
    qkv.t.normal_(std=0.5)
    out, dout, dqkv = ops.Act(M, d, 1, dev), ops.Act(M, d, 1, dev, zero=True), ops.Act(M, 3 * d, 1, dev)
    dout.t.normal_()
    tf = timeit(lambda: ops.attn_fwd(qkv, None, out, nseq, L, H, hd, (0.1, 2, 9)))
    tb = timeit(lambda: ops.attn_bwd(qkv, None, dout, dqkv, nseq, L, H, hd, 0.125, (0.1, 2, 9)))
    print("attn  nseq=%d L=%d hd=%d  fwd %.1f us (%.2f TB/s)  bwd %.1f us (%.2f TB/s)" % (
        nseq, L, hd, tf, M * d * 8 / tf / 1e6, tb, M * d * 14 / tb / 1e6), flush=True)
    del qkv, out, dout, dqkv
