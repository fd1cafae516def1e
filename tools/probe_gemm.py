"""GPU probe for the tcgen05 contractions (development tool, run under gpurun).

Each case runs in its own subprocess with a timeout so that a trap / hang in one descriptor variant cannot take
the others (or the box) down.  Results: gpurun_out/probe_gemm.json
"""
import json
import os
import subprocess
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

CASES = {
    # name: (kind, M, N/P, K/Q, planes, extra)
    "lin_basic": ("linear", 512, 256, 256, 1, {}),
    "lin_ragged": ("linear", 300, 768, 256, 1, {}),
    "lin_n7": ("linear", 496, 7, 256, 1, {"f32": True}),
    "lin_head": ("linear", 992, 2827, 256, 1, {"f32": True}),
    "lin_k512": ("linear", 1024, 256, 512, 1, {"epi": True}),
    "lin_big": ("linear", 131072, 768, 256, 1, {"time": True}),
    "lin_x3": ("linear", 640, 512, 256, 2, {}),
    "out_basic": ("outer", 4096, 768, 256, 1, {}),
    "out_swapped": ("outer", 4096, 768, 256, 1, {"lbo": 1024, "sbo": 8192}),
    "out_ragged": ("outer", 1000, 300, 200, 1, {}),
    "out_small_q": ("outer", 2048, 2827, 64, 1, {}),
    "out_big": ("outer", 131072, 768, 256, 1, {"time": True}),
    "out_x3": ("outer", 4096, 256, 512, 2, {}),
}


def split_planes(x, planes, torch):
    hi = x.to(torch.bfloat16)
    if planes == 1:
        return hi.contiguous(), 0
    lo = (x - hi.float()).to(torch.bfloat16)
    buf = torch.stack([hi, lo]).contiguous()
    return buf, hi.numel()


def run_case(name):
    import ctypes as C
    import torch
    from deepsvg_b200 import _lib
    lib = _lib.load()
    kind, M, N, K, planes, extra = CASES[name]
    dev = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(1)
    st = torch.cuda.current_stream().cuda_stream
    res = {"case": name}
    if kind == "linear":
        X = torch.randn(M, K, generator=g).to(dev)
        W = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
        xb, xlo = split_planes(X, planes, torch)
        wb, wlo = split_planes(W, planes, torch)
        ep = _lib.Epilogue()
        bias = torch.randn(N, generator=g).to(dev)
        ep.bias = bias.data_ptr()
        out_f = torch.zeros(M, N, device=dev)
        out_a = torch.zeros(2 if planes == 2 else 1, M, N, device=dev, dtype=torch.bfloat16)
        ep.out_f32 = out_f.data_ptr()
        ep.out_f32_ld = N
        if not extra.get("f32"):
            ep.out_act = out_a.data_ptr()
            ep.out_act_ld = N
            ep.out_lo_off = M * N if planes == 2 else 0
        resid = None
        if extra.get("epi"):
            resid = torch.randn(M, N, generator=g).to(dev)
            ep.residual = resid.data_ptr()
            ep.res_ld = N
            ep.relu = 1
            ep.scale_cols = 64
            ep.scale = 0.5
        rc = lib.dsvg_linear(xb.data_ptr(), xlo, K, wb.data_ptr(), wlo, K, M, N, K, C.byref(ep), st)
        if rc != 0:
            res["error"] = lib.dsvg_last_error().decode()
            return res
        torch.cuda.synchronize()
        if planes == 1:
            ref = xb[0:M].float() @ wb.float().t() if xb.dim() == 2 else None
        else:
            ref = X.double() @ W.double().t()
            ref = ref.float()
        ref = ref + bias
        if extra.get("epi"):
            ref[:, :64] *= 0.5
            ref = torch.relu(ref) + resid
        err = (out_f - ref).abs().max().item()
        res["max_abs_err_f32"] = err
        res["ref_absmax"] = ref.abs().max().item()
        if not extra.get("f32"):
            got = out_a[0].float() + (out_a[1].float() if planes == 2 else 0)
            res["max_abs_err_act"] = (got - ref).abs().max().item()
        if extra.get("time"):
            for _ in range(3):
                lib.dsvg_linear(xb.data_ptr(), xlo, K, wb.data_ptr(), wlo, K, M, N, K, C.byref(ep), st)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                lib.dsvg_linear(xb.data_ptr(), xlo, K, wb.data_ptr(), wlo, K, M, N, K, C.byref(ep), st)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            res["ms"] = ms
            res["tflops"] = 2.0 * M * N * K / ms / 1e9
    else:
        P, Q = N, K
        if "lbo" in extra:
            lib.dsvg_debug_outer_desc(extra["lbo"], extra["sbo"])
        A = torch.randn(M, P, generator=g).to(dev)
        B = torch.randn(M, Q, generator=g).to(dev)
        lda = (P + 7) // 8 * 8
        ldb = (Q + 7) // 8 * 8
        Ap = torch.zeros(M, lda, device=dev)
        Ap[:, :P] = A
        Bp = torch.zeros(M, ldb, device=dev)
        Bp[:, :Q] = B
        ab, alo = split_planes(Ap, planes, torch)
        bb, blo = split_planes(Bp, planes, torch)
        Cout = torch.zeros(P, Q, device=dev)
        rc = lib.dsvg_outer(ab.data_ptr(), alo, lda, bb.data_ptr(), blo, ldb, M, P, Q, 1.0, Cout.data_ptr(), Q, st)
        if rc != 0:
            res["error"] = lib.dsvg_last_error().decode()
            return res
        torch.cuda.synchronize()
        if planes == 1:
            ref = (ab.float()[:, :P].double().t() @ bb.float()[:, :Q].double()).float()
        else:
            ref = (A.double().t() @ B.double()).float()
        res["max_abs_err_f32"] = (Cout - ref).abs().max().item()
        res["ref_absmax"] = ref.abs().max().item()
        if extra.get("time"):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(3):
                lib.dsvg_outer(ab.data_ptr(), alo, lda, bb.data_ptr(), blo, ldb, M, P, Q, 1.0, Cout.data_ptr(), Q, st)
            e0.record()
            for _ in range(10):
                lib.dsvg_outer(ab.data_ptr(), alo, lda, bb.data_ptr(), blo, ldb, M, P, Q, 1.0, Cout.data_ptr(), Q, st)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            res["ms"] = ms
            res["tflops"] = 2.0 * M * P * Q / ms / 1e9
    return res


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--case":
        try:
            r = run_case(sys.argv[2])
        except Exception as e:  # noqa
            r = {"case": sys.argv[2], "exception": repr(e)}
        print("RESULT " + json.dumps(r))
        return
    os.makedirs("gpurun_out", exist_ok=True)
    results = []
    names = sys.argv[1:] or list(CASES)
    for name in names:
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, __file__, "--case", name], capture_output=True, text=True, timeout=120)
            out = p.stdout + p.stderr
            line = [l for l in out.splitlines() if l.startswith("RESULT ")]
            r = json.loads(line[-1][7:]) if line else {"case": name, "crash": out[-1500:], "rc": p.returncode}
        except subprocess.TimeoutExpired:
            r = {"case": name, "timeout": True}
        r["wall_s"] = round(time.time() - t0, 1)
        print(json.dumps(r), flush=True)
        results.append(r)
        with open("gpurun_out/probe_gemm.json", "w") as f:
            json.dump(results, f, indent=1)


if __name__ == "__main__":
    main()
