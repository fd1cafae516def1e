#!/bin/bash
# 2-GPU pass: NCCL data-parallel correctness + weak scaling at N=2
mkdir -p gpurun_out
R=r2h
nvidia-smi -L
timeout 600 python -m pytest tests/test_trainer_protocol_gpu.py tests/test_pack.py -m gpu -q --no-header 2>&1 | tail -15
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q --no-header -k "nccl" 2>&1 | tail -15
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2> gpurun_out/${R}_bench_n2_stderr.log | tail -1 > gpurun_out/${R}_bench_n2.json
tail -5 gpurun_out/${R}_bench_n2_stderr.log
python - <<'PY'
import json
d=json.load(open("gpurun_out/r2h_bench_n2.json"))
print("N=2 value %.0f ms %.3f e2e %.0f" % (d["value"], d["ms_per_step"], d["e2e"]["value"]))
print("ddp_check", d.get("ddp_check"))
print("parity", d.get("parity_mode",{}).get("value"))
PY
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-ref-gpu --no-parity-mode 2>/dev/null | tail -1 | cut -c1-330
