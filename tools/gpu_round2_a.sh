#!/bin/bash
# round-2 first GPU pass: full GPU test suite, then the three BASELINE workloads
mkdir -p gpurun_out
R=r2a
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${R}_nvidia_smi.csv
timeout 900 python -m pytest tests -m gpu -q -x --no-header -rf 2>&1 | tail -40 > gpurun_out/${R}_pytest_gpu.txt
tail -5 gpurun_out/${R}_pytest_gpu.txt
DSVG_BENCH_TRACE=1 timeout 600 python bench.py 2> gpurun_out/${R}_bench_stderr.log | tail -1 > gpurun_out/${R}_bench_hier.json
tail -4 gpurun_out/${R}_bench_stderr.log
timeout 400 python bench.py --config fonts --steps 10 2> gpurun_out/${R}_bench_fonts_stderr.log | tail -1 > gpurun_out/${R}_bench_fonts.json
timeout 600 python bench.py --config scaled --steps 5 --warmup 3 2> gpurun_out/${R}_bench_scaled_stderr.log | tail -1 > gpurun_out/${R}_bench_scaled.json
tail -3 gpurun_out/${R}_bench_fonts_stderr.log gpurun_out/${R}_bench_scaled_stderr.log
cut -c1-400 gpurun_out/${R}_bench_hier.json gpurun_out/${R}_bench_fonts.json gpurun_out/${R}_bench_scaled.json
