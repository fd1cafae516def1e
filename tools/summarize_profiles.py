"""Turns gpurun_out/<round>_* (written by tools/gpu_artifacts.sh on the GPU box) into the tracked summaries under
profiles/: launch shares of one train step, key ncu metrics of the dominant kernels, the bench lines."""
import csv
import collections
import json
import os
import re
import subprocess
import sys

R = sys.argv[1] if len(sys.argv) > 1 else "r1"
OUT = "profiles"
os.makedirs(OUT, exist_ok=True)


def launches(path, last_n):
    rows = list(csv.reader(open(path, errors="ignore")))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
    data = []
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "")
        v = float(r[vi].replace(",", ""))
        v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
        data.append((name, v))
    data = data[-last_n:]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for n, v in data:
        agg[n][0] += 1
        agg[n][1] += v
    return data, agg


WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "smsp__inst_executed.sum"]


def ncu_raw(rep):
    """rep: path of the .ncu-rep; the GPU-side script leaves `<rep minus .ncu-rep>.raw.csv` (ncu --page raw --csv) next to it
    and deletes most reports (64 MiB copy-back limit), so the CSV is preferred."""
    pre = rep[:-len(".ncu-rep")] + ".raw.csv"
    if os.path.exists(pre):
        out = open(pre, errors="ignore").read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    if len(rows) < 3:
        return []
    H, U = rows[0], rows[1]
    res = []
    for r in rows[2:]:
        d = {"kernel": re.sub(r"\(.*", "", r[H.index("Kernel Name")])}
        for w in WANT:
            if w in H:
                d[w] = r[H.index(w)] + " " + U[H.index(w)]
        res.append(d)
    return res


MODES = {"qkv": (2, "QKV projection"), "ffn1": (3, "FFN first linear (ReLU + dropout)"),
         "proj": (4, "out-proj / FFN2 into the fp32 residual stream"), "lnfwd": (8, "mode 4 + fused LayerNorm"),
         "mask": (5, "dgrad through the ReLU/dropout mask"), "dgrad": (1, "plain dgrad"),
         "head_dgrad": (6, "args-head dgrad, fp32 accumulate"), "logits": (7, "2827-wide fp32 logits")}
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def to_bytes(s):
    v, u = s.split()
    return float(v) * UNIT[u]


def main():
    g = "gpurun_out/%s_" % R
    for src, dst in (("bench_n1", "bench_n1"), ("bench_reference", "bench_reference"), ("bench_fonts", "bench_fonts"),
                     ("bench_scaled", "bench_scaled"), ("bench_n2", "bench_n2")):
        if os.path.exists(g + src + ".json") and os.path.getsize(g + src + ".json") > 10:
            json.dump(json.load(open(g + src + ".json")), open("%s/%s_%s.json" % (OUT, R, dst), "w"), indent=1)
    if os.path.exists(g + "pytest_gpu.txt"):
        open("%s/%s_pytest_gpu.txt" % (OUT, R), "w").write(open(g + "pytest_gpu.txt").read())
    bench = json.load(open(g + "bench_n1.json"))
    n_per_step = int(round(bench["gpu_launches"])) + 25        # + torch fills / copies in the same step
    data, agg = launches(g + "launches.csv", n_per_step)
    tot = sum(a[1] for a in agg.values())
    with open("%s/%s_launch_shares.txt" % (OUT, R), "w") as f:
        f.write("# one train step (N=512, eager launches) under `ncu --metrics gpu__time_duration.sum --clock-control none`: last "
                "%d launches\n# (cold-cache, serialised: compare SHARES, not absolutes).  total %.0f us\n" % (len(data), tot))
        for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
            f.write("%9.0f us %5.1f%% %4d  %s\n" % (a[1], 100 * a[1] / tot, a[0], k))
    # per-role full-set captures of the GEMM family + launches per step of each role -> time-weighted family traffic
    per_mode = {}
    with open("%s/%s_ncu_linear.txt" % (OUT, R), "w") as f:
        f.write("# ncu --set full --clock-control none --import-source on: one launch of every lean epilogue role of "
                "dsvg::linear_kernel at the path-level shape (tools/prof_mode.py), selected raw metrics\n")
        for name, (mode, what) in MODES.items():
            rep = g + "mode_%s.ncu-rep" % name
            if not (os.path.exists(rep) or os.path.exists(rep[:-8] + ".raw.csv")):
                continue
            for d in ncu_raw(rep)[:1]:
                d["role"], d["mode"] = what, mode
                f.write(json.dumps(d) + "\n")
                n_l = sum(a[0] for k, a in agg.items() if "linear_kernel<256, 1, %d>" % mode in k)
                per_mode[name] = {"mode": mode, "launches_per_step": n_l,
                                  "dram_bytes": to_bytes(d["dram__bytes_read.sum"]) + to_bytes(d["dram__bytes_write.sum"]),
                                  "us": float(d["gpu__time_duration.sum"].split()[0])}
    json.dump(per_mode, open("%s/%s_linear_modes.json" % (OUT, R), "w"), indent=1)
    for name in ("outer", "attn", "ln_bwd"):
        rep = g + "mode_%s.ncu-rep" % name
        if not (os.path.exists(rep) or os.path.exists(rep[:-8] + ".raw.csv")):
            continue
        with open("%s/%s_ncu_%s.txt" % (OUT, R, name), "w") as f:
            f.write("# ncu --set full --clock-control none, selected raw metrics per captured launch (%s)\n" % rep)
            for d in ncu_raw(rep):
                f.write(json.dumps(d) + "\n")
    # parity mode (bf16x3): launch shares of one step + the captured roles
    if os.path.exists(g + "x3_launches.csv"):
        data, agg = launches(g + "x3_launches.csv", n_per_step)
        tot = sum(a[1] for a in agg.values())
        with open("%s/%s_x3_launch_shares.txt" % (OUT, R), "w") as f:
            f.write("# one PARITY-MODE (bf16x3) train step (N=512, eager launches) under `ncu --metrics gpu__time_duration.sum "
                    "--clock-control none`: last %d launches.  total %.0f us\n" % (len(data), tot))
            for k, a in sorted(agg.items(), key=lambda x: -x[1][1]):
                f.write("%9.0f us %5.1f%% %4d  %s\n" % (a[1], 100 * a[1] / tot, a[0], k))
    with open("%s/%s_x3_ncu.txt" % (OUT, R), "w") as f:
        f.write("# ncu --set full --clock-control none: parity-mode (two bf16 planes, three products) kernels at the path-level "
                "shape (DSVG_PLANES=2 tools/prof_mode.py)\n")
        for name in ("qkv", "ffn1", "proj", "attn"):
            rep = g + "x3mode_%s.ncu-rep" % name
            if os.path.exists(rep) or os.path.exists(rep[:-8] + ".raw.csv"):
                for d in ncu_raw(rep):
                    d["role"] = name
                    f.write(json.dumps(d) + "\n")
    print("wrote", sorted(x for x in os.listdir(OUT) if x.startswith(R)))


if __name__ == "__main__":
    main()
