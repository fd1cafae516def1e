#!/bin/bash
# Runs on the GPU box (via gpurun): produces every measured artefact of a round under gpurun_out/.
#   bash tools/gpu_artifacts.sh r1
R=${1:-r1}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${R}_nvidia_smi.csv
timeout 500 python bench.py 2> gpurun_out/${R}_bench_stderr.log | tail -1 > gpurun_out/${R}_bench_n1.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 2>/dev/null | tail -1 > gpurun_out/${R}_bench_reference.json
# launch list of one full train step (cold-cache, serialised: compare SHARES)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches.csv \
    python tools/one_step.py 512 2 > gpurun_out/${R}_one_step.log 2>&1
# full-set captures of the dominant kernels (a few launches each)
timeout 400 ncu --set full --clock-control none --import-source on -k regex:linear_kernel -s 152 -c 12 \
    -o gpurun_out/${R}_prof_linear -f python tools/one_step.py 512 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:outer_kernel -s 20 -c 3 \
    -o gpurun_out/${R}_prof_outer -f python tools/one_step.py 512 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_mma -s 1 -c 2 \
    -o gpurun_out/${R}_prof_attn -f python tools/one_step.py 512 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:ln_bwd -s 4 -c 2 \
    -o gpurun_out/${R}_prof_ln_bwd -f python tools/one_step.py 512 1 > /dev/null 2>&1
ls -la gpurun_out/
