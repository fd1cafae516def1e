"""Development tool (run under gpurun): per-tile clock64 stamps of CTA 0 of dsvg::linear_kernel for the hot shapes.
Slots: 0 producer tile start, 1 producer last-kb slot free, 2 mma loop top, 3 mma got TMEM stage, 4 mma first operands
landed, 5 mma last operands landed, 6 mma issued+committed, 8 epilogue loop top, 9 epilogue staging free (TMA-out modes),
10 epilogue accumulator ready, 11 epilogue TMEM drained, 12 epilogue all warps stored."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from deepsvg_b200 import _lib, ops

dev = torch.device("cuda:0")
lib = _lib.load()
lib.dsvg_debug_linear_trace.argtypes = [C.c_void_p]
lib.dsvg_debug_linear_trace.restype = None


def run(name, M, N, K, **kw):
    X = ops.Act(M, K, 1, dev, zero=True)
    X.t.normal_()
    W = ops.Act(N, K, 1, dev, zero=True)
    W.t.normal_(std=K ** -0.5)
    for _ in range(3):
        ops.linear(X, W, M, N, K, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.linear(X, W, M, N, K, **kw)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    buf = torch.zeros(256, dtype=torch.int64, device=dev)
    lib.dsvg_debug_linear_trace(C.c_void_p(buf.data_ptr()))
    ops.linear(X, W, M, N, K, **kw)
    torch.cuda.synchronize()
    lib.dsvg_debug_linear_trace(C.c_void_p(0))
    t = buf.cpu().view(16, 16)
    base = int(t[0, 2])
    print("== %s  M=%d N=%d K=%d : %.1f us/launch, %.0f TFLOP/s" % (name, M, N, K, us, 2.0 * M * N * K / us / 1e6))
    print("tile  prod0 prodL | mma_top mma_tmem mma_op0 mma_opL mma_done | epi_top epi_stg epi_acc epi_drn epi_sto   (clocks since CTA start)")
    for it in range(10):
        r = [int(t[it, s]) - base if int(t[it, s]) else -1 for s in (0, 1, 2, 3, 4, 5, 6, 8, 9, 10, 11, 12)]
        print("%4d %6d %6d | %6d %6d %6d %6d %6d | %6d %6d %6d %6d %6d" % tuple([it] + r))


M = 131072
bias512 = torch.zeros(512, device=dev)
bias256 = torch.zeros(256, device=dev)
bias768 = torch.zeros(768, device=dev)
out512 = ops.Act(M, 512, 1, dev)
out256 = ops.Act(M, 256, 1, dev)
out768 = ops.Act(M, 768, 1, dev)
res = torch.zeros(M, 256, device=dev)
outf = torch.zeros(M, 256, device=dev)
run("FFN1 (mode 3)", M, 512, 256, bias=bias512, relu=True, drop=(0.1, 3, 7), out_act=out512)
run("dgrad (mode 1) K=512", M, 256, 512, out_act=out256)
run("QKV (mode 2)", M, 768, 256, bias=bias768, scale_cols=256, scale=0.17, out_act=out768)
run("out-proj (mode 4)", M, 256, 256, bias=bias256, drop=(0.1, 4, 7), residual=res, out_f32=outf)
mk = ops.Act(M, 512, 1, dev, zero=True)
mk.t.normal_()
run("dgrad through mask (mode 5)", M, 512, 256, mask=mk, mask_scale=1.1, out_act=out512)
x2 = torch.zeros(M, 256, device=dev)
run("FFN2 (mode 4) K=512", M, 256, 512, bias=bias256, drop=(0.1, 4, 7), residual=x2, out_f32=x2)
