#!/bin/bash
# parity-mode (bf16x3) attention on tensor cores: tests, launch shares and step time against the fp32 SIMT kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -x 2>&1 | tail -8 > gpurun_out/x3_pytest.txt
tail -4 gpurun_out/x3_pytest.txt
for mode in mma simt; do
  DSVG_ATTN=$mode DSVG_PRECISION=bf16x3 DSVG_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
     --log-file gpurun_out/x3_launches_$mode.csv python tools/one_step.py 512 2 hier > gpurun_out/x3_one_step_$mode.log 2>&1
  python tools/launch_summary.py gpurun_out/x3_launches_$mode.csv 409 > gpurun_out/x3_shares_$mode.txt
  head -12 gpurun_out/x3_shares_$mode.txt
done
timeout 600 python bench.py --precision bf16x3 --steps 10 --no-cpu-baseline --no-parity-mode --no-ref-gpu 2>gpurun_out/x3_bench.err | tail -1 > gpurun_out/x3_bench.json
cut -c1-400 gpurun_out/x3_bench.json
