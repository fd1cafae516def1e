#!/usr/bin/env python
"""bench.py -- icons/sec of one DeepSVG `hierarchical_ordered` train step (forward + SVGLoss + backward
[+ gradient all-reduce]) on N B200s.  Contract: see the task statement / DESIGN.md section "Measurement".

  python bench.py --gpus 1 --steps 20 --warmup 5                       # our arm (CUDA, tcgen05)
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference --steps 3 --warmup 1                # CPU arm: the oracle port on the host cores

One JSON line on stdout (rank 0).  `value` = device-resident inputs; `e2e` = through the public module API with host
buffers (pinned H2D of the step's inputs + D2H of the loss inside the timed region).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FWD_GFLOP_PER_ICON = 2.7052          # SURVEY.md 8d, hierarchical_ordered, padded shapes; train step = 3x forward
TRAIN_GFLOP_PER_ICON = 3 * FWD_GFLOP_PER_ICON
WEIGHTS = {"kl_tolerance": 0.1, "loss_kl_weight": 1.0, "loss_cmd_weight": 1.0, "loss_args_weight": 2.0,
           "loss_visibility_weight": 1.0}
MASK = np.array([[0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1], [0, 0, 0, 0, 0, 0, 0, 0, 0, 1, 1], [0, 0, 0, 0, 0, 1, 1, 1, 1, 1, 1],
                 [1, 1, 1, 1, 1, 0, 0, 0, 0, 1, 1], [0] * 11, [0] * 11, [0] * 11], dtype=np.float32)


def synth_icons(n, G=8, S=30, seed=1234):
    """Vectorised version of the SURVEY.md 8d generator: paths of U{3..S} commands (m then l/c), U{1..G} visible paths."""
    rng = np.random.default_rng(seed)
    L = S + 2
    cmd = np.full((n, G, L), 4.0, dtype=np.float32)
    cmd[:, :, 0] = 5.0
    nvis = rng.integers(1, G + 1, size=(n, 1))
    vis = np.arange(G)[None, :] < nvis
    ln = rng.integers(3, S + 1, size=(n, G))
    pos = np.arange(L)[None, None, :]
    body = rng.integers(1, 3, size=(n, G, L)).astype(np.float32)
    body[:, :, 1] = 0.0
    inside = (pos >= 1) & (pos <= ln[:, :, None]) & vis[:, :, None]
    cmd = np.where(inside, body, cmd)
    vals = rng.integers(0, 256, size=(n, G, L, 11)).astype(np.float32)
    m = MASK[cmd.astype(np.int64)]
    args = vals * m - (1 - m)
    return torch.from_numpy(cmd), torch.from_numpy(args.astype(np.float32))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1427.5)), d.get("bf16_tflops", 1654.1), \
            d.get("hbm_gbs", 6581.6), "measured"
    return 1400.0, 1590.0, 6650.0, "fallback"


class ClockSampler(threading.Thread):
    Q = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown," \
        "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu=0):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag = gpu, [], False

    def run(self):
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        mx = float(self.rows[0][2]) if self.rows and self.rows[0][2].replace(".", "").isdigit() else None
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(self.rows)}


# ---------------------------------------------------------------------------------------------------------
def cpu_port_rate(batch, steps, warmup, threads):
    """The oracle (CPU restatement of the reference path) timed on the host cores: icons/s of fwd+loss+bwd."""
    from oracle import svg_oracle as O
    torch.set_num_threads(threads)
    cfg = O.make_cfg("hierarchical", use_vae=False)
    params = O.make_params(cfg, seed=0)
    cmd, arg = synth_icons(batch, seed=99)
    best = None
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        O.train_step(params, cfg, cmd, arg)
        dt = time.perf_counter() - t0
        if i >= warmup:
            best = dt if best is None else min(best, dt)
    return batch / best, best


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    rate, dt = cpu_port_rate(a.cpu_batch, a.steps, a.warmup, threads)
    sample = "oracle port (eval-mode arithmetic, fp32 torch CPU), hierarchical_ordered, batch %d per step, best of %d" % (
        a.cpu_batch, a.steps)
    line = {"impl": "reference", "metric": "icons/sec train-step (fwd+loss+bwd) hierarchical_ordered", "value": rate,
            "unit": "icons/s", "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": dt * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": "hierarchical_ordered train step G=8 S=30 d_model=256 L=4 H=8 (BASELINE configs[1])",
                       "batch_per_step": a.cpu_batch},
            "cpu_baseline": {"value": rate, "unit": "icons/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": rate, "unit": "icons/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------
def run_ours(a):
    import torch.distributed as dist
    from deepsvg_b200 import Hierarchical, SVGLoss, SVGTransformer, _lib, ops
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py (our arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pg = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    cfg = Hierarchical(use_vae=False)                     # configs/deepsvg/hierarchical_ordered.py:4-9
    torch.manual_seed(1234)
    model = SVGTransformer(cfg, precision=a.precision, process_group=pg).to(dev)
    if world > 1:
        for p in model.parameters():                      # identical replicas
            dist.broadcast(p.data, 0)
    model.train()
    loss_fn = SVGLoss(cfg).to(dev)
    B = a.batch
    cmd_h, arg_h = synth_icons(B, seed=1234 + rank)
    cmd_h, arg_h = cmd_h.pin_memory(), arg_h.pin_memory()
    cmd_d, arg_d = cmd_h.to(dev), arg_h.to(dev)
    h2d = cmd_h.numel() * 4 + arg_h.numel() * 4
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    def step(c, a_):
        model.zero_grad(set_to_none=True)
        out = model(c, a_, c, a_, params={})
        ls = loss_fn(out, None, weights=WEIGHTS)
        ls["loss"].backward()
        return ls["loss"]

    def step_e2e():
        c = cmd_h.to(dev, non_blocking=True)
        a_ = arg_h.to(dev, non_blocking=True)
        l = step(c, a_)
        loss_host.copy_(l.detach(), non_blocking=True)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        barrier()
        return ms

    for _ in range(max(a.warmup, 3)):
        step(cmd_d, arg_d)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    l0 = _lib.launch_count()
    ms = timed(lambda: step(cmd_d, arg_d), a.steps)
    launches = (_lib.launch_count() - l0) / a.steps
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, a.steps)
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    final_loss = float(loss_host.item())

    # ---- per-family kernel timing (one extra, untimed step with events around every tensor-core launch) ----
    fam = {}
    if rank == 0:
        ops.PROFILE = []
        step(cmd_d, arg_d)
        torch.cuda.synchronize()
        for family, flops, e0, e1 in ops.PROFILE:
            f = fam.setdefault(family, [0, 0.0, 0.0])
            f[0] += 1
            f[1] += flops
            f[2] += e0.elapsed_time(e1)
        ops.PROFILE = None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    ips = world * B * a.steps / (ms / 1e3)
    ips_e2e = world * B * a.steps / (ms_e2e / 1e3)
    sus, burst, hbm, src = peaks()
    step_tflops = ips * TRAIN_GFLOP_PER_ICON / 1e3 / world
    lin = fam.get("linear", [1, 0.0, 1.0])
    lin_tflops = lin[1] / (lin[2] / 1e3) / 1e12 if lin[2] > 0 else 0.0
    cpu_threads = os.cpu_count() or 1
    cpu_rate, cpu_dt = (None, None)
    if world == 1 and not a.no_cpu_baseline:
        cpu_rate, cpu_dt = cpu_port_rate(a.cpu_batch, 2, 1, cpu_threads)
    line = {
        "metric": "icons/sec train-step (fwd+loss+bwd) hierarchical_ordered", "value": ips, "unit": "icons/s",
        "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3), "ms_per_step": ms / a.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if a.precision == "bf16" else "bf16x3(split-bf16, fp32-accurate)", "data": "synthetic",
        "config": {"workload": "hierarchical_ordered train step, G=8 S=30 n_args=11 d_model=256 L=4+4+4+4 H=8 ff=512, "
                               "dropout 0.1 (train mode), batch %d per GPU (BASELINE configs[%d])" % (B, 1 if world == 1 else 2),
                   "global_batch": B * world, "parallelism": "dp%d" % world,
                   "l2": "per-step working set (~10 GB of activations) >> 126 MB L2, no explicit flush needed",
                   "final_loss": final_loss},
        "e2e": {"value": ips_e2e, "unit": "icons/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / a.steps},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "kernel": "dsvg::linear_kernel (tcgen05 X.W^T, all forward + dgrad GEMMs)",
                     "achieved": lin_tflops, "peak": sus, "unit": "TFLOP/s", "frac": lin_tflops / sus if sus else None,
                     "peak_source": src + " (bf16_tflops_sustained)", "launches_per_step": lin[0],
                     "ms_per_step_in_kernel": lin[2], "traffic": None,
                     "step": {"achieved": step_tflops, "frac": step_tflops / sus, "gflop_per_icon": TRAIN_GFLOP_PER_ICON},
                     "families": {k: {"launches": v[0], "ms": v[2], "tflops": (v[1] / (v[2] / 1e3) / 1e12 if v[2] else 0)}
                                  for k, v in fam.items()}},
        "clocks": sampler.summary() if sampler else None,
    }
    if cpu_rate is not None:
        line["cpu_baseline"] = {"value": cpu_rate, "unit": "icons/s", "cores": cpu_threads, "kind": "port",
                                "sample": "oracle port (fp32 torch CPU, eval-mode arithmetic), batch %d, best of 2 steps "
                                          "(%.1f s each)" % (a.cpu_batch, cpu_dt)}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="icons per GPU per step")
    ap.add_argument("--precision", default=os.environ.get("DSVG_PRECISION", "bf16"))
    ap.add_argument("--cpu-batch", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
