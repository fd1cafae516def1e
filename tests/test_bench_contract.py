"""bench.py pieces that run without a GPU: the synthetic-icon generator and the reference arm's JSON line."""
import json
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402
from oracle import svg_oracle as O  # noqa: E402


def test_synthetic_icons_obey_the_dataset_format():
    """SURVEY.md 8d / svgtensor_dataset.py:164-205: SOS first, EOS padding, >= 1 visible path, masked arguments."""
    n, G, S = 64, 8, 30
    cmd, arg = bench.synth_icons(n, G, S, seed=5)
    assert cmd.shape == (n, G, S + 2) and arg.shape == (n, G, S + 2, 11)
    assert cmd.dtype == torch.float32 and arg.dtype == torch.float32
    c = cmd.long()
    assert (c[:, :, 0] == O.CMD_SOS).all()
    body = c[:, :, 1:]
    is_eos = body == O.CMD_EOS
    # once EOS starts it never stops (padding), and the last position is always EOS
    assert (is_eos[:, :, 1:] | ~is_eos[:, :, :-1]).all() and is_eos[:, :, -1].all()
    ln = (~is_eos).sum(-1)
    visible = ln > 0
    assert visible.any(1).all()                       # at least one non-empty path per icon
    assert ((ln[visible] >= 3) & (ln[visible] <= S)).all()
    assert (c[:, :, 1][visible] == O.CMD_M).all()     # paths start with a move
    m = O.CMD_ARGS_MASK[c].bool()
    assert ((arg >= 0) == m).all()                    # unused argument slots are -1, used ones are ids
    assert arg.max() <= 255 and (arg[m] == arg[m].round()).all()


def test_synthetic_icons_run_through_the_oracle():
    cfg = O.make_cfg("hierarchical", use_vae=False)
    cmd, arg = bench.synth_icons(2, seed=3)
    out = O.forward(O.make_params(cfg, seed=0), cfg, cmd, arg)
    assert all(torch.isfinite(v).all() for v in out.values() if torch.is_tensor(v))


def test_our_arm_refuses_to_run_without_a_gpu():
    if torch.cuda.is_available():
        return
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
    assert not any(line.startswith("{") for line in r.stdout.splitlines())   # no bench line is printed


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--cpu-batch", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better",
              "scaling", "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline"):
        assert k in line, k
    assert line["impl"] == "reference" and line["unit"] == "icons/s" and line["higher_is_better"] is True
    assert line["steps"] == 1 and line["n_gpus"] == 1 and line["vs_baseline"] is None
    assert "workload" in line["config"] and "model" not in line["config"]
    # both arms must print the SAME config object for the same command (the driver compares them)
    assert line["config"] == bench.bench_config("hier", bench.WORKLOADS["hier"]["batch"], 1)
    assert line["metric"] == bench.METRIC % "hierarchical_ordered"
    assert line["e2e"]["value"] == line["value"] and line["e2e"]["h2d_bytes_per_step"] == 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and cb["sample"]
    assert np.isfinite(line["value"]) and line["value"] > 0


def test_one_stage_synthetic_icons_and_workloads():
    """BASELINE configs[3]: G = 1 grouped tensors, 1..3 'm' sub-paths (group index stays inside group_embed), labels."""
    wl = bench.WORKLOADS["fonts"]
    c, a, lab = bench.workload_inputs(wl, 64, seed=3)
    assert c.shape == (64, 1, 52) and a.shape == (64, 1, 52, 11) and lab.shape == (64,) and lab.dtype == torch.int64
    assert int(lab.min()) >= 0 and int(lab.max()) < 52
    cl = c.long()
    n_m = (cl == O.CMD_M).sum(-1)
    assert int(n_m.min()) >= 1 and int(n_m.max()) <= 3
    assert (cl[:, :, 1] == O.CMD_M).all() and (cl[:, :, 0] == O.CMD_SOS).all()
    cfg = O.make_cfg(wl["kind"], **wl["over"])
    out = O.forward(O.make_params(cfg, seed=0), cfg, c[:2], a[:2], label=lab[:2])
    assert all(torch.isfinite(v).all() for v in out.values() if torch.is_tensor(v))
    c5, a5, l5 = bench.workload_inputs(bench.WORKLOADS["scaled"], 2, seed=1)
    assert c5.shape == (2, 16, 66) and a5.shape == (2, 16, 66, 11) and l5 is None
