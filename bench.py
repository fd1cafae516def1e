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


def ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel family, averaged over the launches of
    the committed `ncu --set full` capture (profiles/r*_ncu_linear.txt, written by tools/summarize_profiles.py)."""
    try:
        import glob
        files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_ncu_linear.txt")))
        tot, n = 0.0, 0
        unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        for line in open(files[-1]):
            if not line.startswith("{"):
                continue
            d = json.loads(line)
            b = 0.0
            for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                v, u = d[k].split()
                b += float(v) * unit[u]
            tot, n = tot + b, n + 1
        return (tot / n, "%s: mean over %d captured linear_kernel launches" % (os.path.basename(files[-1]), n)) if n else (None, None)
    except Exception:
        return None, None


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region, via in-process NVML (spawning nvidia-smi every 100 ms was
    measured to stall CUDA launches for 100-250 ms at a time on a multi-GPU box)."""

    def __init__(self, gpu=0):
        super().__init__(daemon=True)
        self.gpu, self.rows, self.stop_flag, self.recording = gpu, [], False, False
        self.h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(self._physical_index(gpu))
            # one-off firmware queries, synchronously and long before any timed region (they can take 100s of ms and
            # stall kernel submission while they run)
            self.mx = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.first = self._reasons()
        except Exception:
            self.nv, self.h = None, None

    @staticmethod
    def _physical_index(i):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        if vis:
            try:
                return int(vis.split(",")[i])
            except Exception:
                pass
        return i

    def _reasons(self):
        nv = self.nv
        try:
            f = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            return int(f(self.h))
        except Exception:
            return 0

    def run(self):
        """During the timed region only the SM clock is polled (a cached register read); the throttle-reason query is an
        RPC to the GPU's firmware that was measured to stall kernel submission on every GPU of the box for 100-250 ms, so
        it is issued once when the region starts being sampled and once when it stops."""
        if self.h is None:
            return
        nv = self.nv
        mx, first = self.mx, self.first
        while not self.stop_flag:
            try:
                clk = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                if self.recording:
                    self.rows.append((clk, mx, 0))
            except Exception:
                pass
            time.sleep(0.5)
        last = self._reasons()
        self.rows.append((self.rows[-1][0] if self.rows else 0, mx, first | last))

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "source": "nvml unavailable"}
        bits = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40}
        seen = sorted(k for k, b in bits.items() if any(r[2] & b for r in self.rows))
        return {"sm_mhz": float(np.median([r[0] for r in self.rows])), "sm_max_mhz": float(self.rows[0][1]),
                "reasons": seen, "samples": len(self.rows), "source": "nvml"}


# ---------------------------------------------------------------------------------------------------------
def host_threads():
    """Threads given to the CPU arm: all host cores up to 32 -- beyond that torch's intra-op parallelism over these small
    GEMMs (K = 256) only adds synchronisation (measured 38 s/step with 128 threads vs ~5 s with 32 on the same box)."""
    return max(1, min(os.cpu_count() or 1, 32))


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
    threads = host_threads()
    # bounded sample: probe one small step, then size the per-step batch so that W + K steps take about two minutes
    _, t_probe = cpu_port_rate(4, 1, 0, threads)
    budget_s = float(os.environ.get("DSVG_REF_BUDGET_S", "120"))   # wall-clock target for the W + K CPU steps
    per_icon = t_probe / 4.0
    batch = int(max(2, min(64, budget_s / (max(1, a.steps + a.warmup) * per_icon))))
    a.cpu_batch = batch
    rate, dt = cpu_port_rate(batch, a.steps, a.warmup, threads)
    sample = "oracle port (eval-mode arithmetic, fp32 torch CPU), hierarchical_ordered, batch %d per step, best of %d" % (
        batch, a.steps)
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

    cmd_in, arg_in = torch.empty_like(cmd_d), torch.empty_like(arg_d)   # device staging for the per-step H2D copies

    def step_e2e():
        cmd_in.copy_(cmd_h, non_blocking=True)       # pinned host -> device, every step
        arg_in.copy_(arg_h, non_blocking=True)
        l = step(cmd_in, arg_in)
        loss_host.copy_(l.detach(), non_blocking=True)   # device -> pinned host, every step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        import gc
        gc.collect()
        gc.disable()     # a generation-2 collection in the launching thread shows up as a 30-100 ms hole in the GPU queue
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        marks, host = [], []
        trace = bool(os.environ.get("DSVG_BENCH_TRACE"))
        st0 = torch.cuda.memory_stats() if trace else None
        e0.record()
        for _ in range(steps):
            h0 = time.perf_counter()
            fn()
            if trace:
                ev = torch.cuda.Event(enable_timing=True)
                ev.record()
                marks.append(ev)
                host.append((time.perf_counter() - h0) * 1e3)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if marks and rank == 0:
            ts = [e0.elapsed_time(m) for m in marks]
            st1 = torch.cuda.memory_stats()
            sys.stderr.write("per-step ms: " + " ".join("%.1f" % (b - a) for a, b in zip([0.0] + ts[:-1], ts)) + "\n")
            sys.stderr.write("host-side ms: " + " ".join("%.1f" % h for h in host) + "\n")
            sys.stderr.write("allocator: cudaMalloc +%d, retries +%d, segments %d\n" % (
                st1["num_device_alloc"] - st0["num_device_alloc"], st1["num_alloc_retries"] - st0["num_alloc_retries"],
                st1["segment.all.current"]))
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = t.item()
        barrier()
        gc.enable()
        return ms

    for _ in range(max(a.warmup, 3)):
        step(cmd_d, arg_d)
    sampler = ClockSampler(local) if (rank == 0 and not os.environ.get("DSVG_NO_CLOCKS")) else None
    if sampler:
        sampler.start()      # its one-off NVML firmware queries happen during the warm-up below, not in the timed region
    # Extended warm-up (untimed): the caching allocator needs a few more iterations to reach its steady-state pool, and
    # with NCCL peer mappings every late cudaMalloc costs 100-250 ms (measured as isolated spikes at N=2).  Continue
    # until five consecutive steps are within 10 % of the fastest seen AND trigger no new cudaMalloc (the default allocator
    # was measured to keep adding segments for ~80 steps: 238 cudaMallocs inside one 40-step timed region, each a
    # 20-350 ms hole); at most 120 extra steps, same count on all ranks.
    stable, best, extra = 0, None, 0
    n_malloc = torch.cuda.memory_stats()["num_device_alloc"]
    while extra < 120:
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        step(cmd_d, arg_d)
        step_e2e()
        t1.record()
        torch.cuda.synchronize()
        now_malloc = torch.cuda.memory_stats()["num_device_alloc"]
        dt = torch.tensor([t0.elapsed_time(t1), float(now_malloc - n_malloc)], device=dev)
        n_malloc = now_malloc
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        dt, grew = dt[0].item(), dt[1].item() > 0
        best = dt if best is None else min(best, dt)
        stable = stable + 1 if (dt <= 1.1 * best and not grew) else 0   # steady = fast AND no new device allocation
        extra += 1
        if stable >= 5 and extra >= 6:
            break
    if sampler:
        sampler.recording = True
    l0 = _lib.launch_count()
    ms = timed(lambda: step(cmd_d, arg_d), a.steps)
    launches = (_lib.launch_count() - l0) / a.steps
    for _ in range(5):
        step_e2e()
    ms_e2e = timed(step_e2e, a.steps)
    if sampler:
        sampler.stop_flag = True
        sampler.join(timeout=2)
    final_loss = float(loss_host.item())

    # ---- per-family kernel timing (one extra, untimed step with events around every tensor-core launch) ----
    fam = {}
    ops.PROFILE = [] if rank == 0 else None
    # Park the GPU behind a ~20 ms spin kernel first, so that the (slower, instrumented) host thread has enqueued the launches
    # and event records before the device reaches them: each event pair then brackets pure device time instead of the host's
    # launch latency (which dominated the 8-15 us group-level GEMMs).
    try:
        torch.cuda._sleep(40_000_000)
    except Exception:
        pass
    step(cmd_d, arg_d)              # every rank runs it: the step contains collectives
    torch.cuda.synchronize()
    shapes = {}
    if rank == 0:
        for family, flops, e0, e1, shape in ops.PROFILE:
            ms_k = e0.elapsed_time(e1)
            f = fam.setdefault(family, [0, 0.0, 0.0])
            f[0] += 1
            f[1] += flops
            f[2] += ms_k
            if shape is not None:
                g = shapes.setdefault((family,) + tuple(shape), [0, 0.0, 0.0])
                g[0] += 1
                g[1] += flops
                g[2] += ms_k
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
    cpu_threads = host_threads()
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
                   "final_loss": final_loss, "extra_untimed_warmup_steps": extra},
        "e2e": {"value": ips_e2e, "unit": "icons/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4,
                "ms_per_step": ms_e2e / a.steps},
        "gpu_launches": launches,
        "roofline": {"bound": "tensor", "kernel": "dsvg::linear_kernel (tcgen05 X.W^T, all forward + dgrad GEMMs)",
                     "achieved": lin_tflops, "peak": sus, "unit": "TFLOP/s", "frac": lin_tflops / sus if sus else None,
                     "peak_source": src + " (bf16_tflops_sustained)", "launches_per_step": lin[0],
                     "ms_per_step_in_kernel": lin[2], "traffic": ncu_traffic()[0], "traffic_source": ncu_traffic()[1],
                     "step": {"achieved": step_tflops, "frac": step_tflops / sus, "gflop_per_icon": TRAIN_GFLOP_PER_ICON},
                     "families": {k: {"launches": v[0], "ms": v[2], "tflops": (v[1] / (v[2] / 1e3) / 1e12 if v[2] else 0)}
                                  for k, v in fam.items()},
                     # the six most expensive GEMM shapes of the step, each with its own achieved rate (live CUDA events)
                     "top_shapes": [{"kernel": k[0], "MNK": list(k[1:]), "launches": v[0], "ms": round(v[2], 4),
                                     "tflops": round(v[1] / (v[2] / 1e3) / 1e12, 1),
                                     "frac_of_peak": round(v[1] / (v[2] / 1e3) / 1e12 / sus, 3)}
                                    for k, v in sorted(shapes.items(), key=lambda kv: -kv[1][2])[:6]]},
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
    import signal
    signal.alarm(int(os.environ.get("DSVG_BENCH_TIMEOUT", "900")))   # a hung collective must not hold the box
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
