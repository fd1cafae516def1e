#!/bin/bash
mkdir -p gpurun_out
R=r2f
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q --no-header -k "self_match or match_kernels or hierarch_logits" 2>&1 | tail -40 > gpurun_out/${R}_pytest_match.txt
tail -40 gpurun_out/${R}_pytest_match.txt
timeout 900 python -m pytest tests -m gpu -q --no-header 2>&1 | tail -30 > gpurun_out/${R}_pytest_gpu.txt
tail -8 gpurun_out/${R}_pytest_gpu.txt
for cfgname in fonts scaled; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches_${cfgname}.csv \
    python tools/one_step.py 0 2 ${cfgname} > gpurun_out/${R}_one_step_${cfgname}.log 2>&1
tail -2 gpurun_out/${R}_one_step_${cfgname}.log
done
python tools/launch_summary.py gpurun_out/${R}_launches_fonts.csv 266 | head -24
python tools/launch_summary.py gpurun_out/${R}_launches_scaled.csv 727 | head -24
