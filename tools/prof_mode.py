"""Development tool: launch one hot linear shape / epilogue mode a few times (for `ncu -k regex:linear_kernel`)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepsvg_b200 import ops

dev = torch.device("cuda:0")
which = sys.argv[1]
M = 131072


def go(M, N, K, **kw):
    X = ops.Act(M, K, 1, dev, zero=True)
    X.t.normal_()
    W = ops.Act(N, K, 1, dev, zero=True)
    W.t.normal_(std=K ** -0.5)
    for _ in range(4):
        ops.linear(X, W, M, N, K, **kw)
    torch.cuda.synchronize()


if which == "ffn1":
    go(M, 512, 256, bias=torch.zeros(512, device=dev), relu=True, drop=(0.1, 3, 7), out_act=ops.Act(M, 512, 1, dev))
elif which == "qkv":
    go(M, 768, 256, bias=torch.zeros(768, device=dev), scale_cols=256, scale=0.17, out_act=ops.Act(M, 768, 1, dev))
elif which == "proj":
    x = torch.zeros(M, 256, device=dev)
    go(M, 256, 256, bias=torch.zeros(256, device=dev), drop=(0.1, 4, 7), residual=x, out_f32=x)
elif which == "ffn2":
    x = torch.zeros(M, 256, device=dev)
    go(M, 256, 512, bias=torch.zeros(256, device=dev), drop=(0.1, 4, 7), residual=x, out_f32=x)
elif which == "mask":
    mk = ops.Act(M, 512, 1, dev, zero=True)
    mk.t.normal_()
    go(M, 512, 256, mask=mk, mask_scale=1.1, out_act=ops.Act(M, 512, 1, dev))
elif which == "dgrad":
    go(M, 256, 512, out_act=ops.Act(M, 256, 1, dev))
elif which == "small_outer":
    Ms = 4096
    A = ops.Act(Ms, 768, 1, dev, zero=True)
    A.t.normal_()
    B = ops.Act(Ms, 256, 1, dev, zero=True)
    B.t.normal_()
    Cw = torch.zeros(768, 256, device=dev)
    cs = torch.zeros(768, device=dev)
    for _ in range(4):
        ops.outer(A, B, Ms, 768, 256, Cw, colsum=cs)
    torch.cuda.synchronize()
elif which == "small_generic":
    g2 = torch.zeros(512, 256, device=dev)
    go(512, 256, 256, bias=torch.zeros(256, device=dev), drop=(0.1, 4, 7), rowvec=g2, rows_per_group=1,
       out_f32=torch.zeros(512, 256, device=dev))
elif which == "small_lean":
    x = torch.zeros(4096, 256, device=dev)
    go(4096, 256, 256, bias=torch.zeros(256, device=dev), drop=(0.1, 4, 7), residual=x, out_f32=x)
