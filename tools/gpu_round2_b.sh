#!/bin/bash
# round-2 second GPU pass: CUDA-graph step
mkdir -p gpurun_out
R=r2b
timeout 600 python -m pytest tests/test_model_gpu.py -m gpu -q -x --no-header -k "graph or adamw or dropout_statistics or generic_autograd" 2>&1 | tail -30 > gpurun_out/${R}_pytest_graph.txt
tail -30 gpurun_out/${R}_pytest_graph.txt
DSVG_BENCH_TRACE=1 timeout 600 python bench.py --no-cpu-baseline 2> gpurun_out/${R}_bench_stderr.log | tail -1 > gpurun_out/${R}_bench_hier.json
tail -6 gpurun_out/${R}_bench_stderr.log
DSVG_BENCH_TRACE=1 timeout 400 python bench.py --config fonts --steps 10 --no-cpu-baseline 2> gpurun_out/${R}_bench_fonts_stderr.log | tail -1 > gpurun_out/${R}_bench_fonts.json
tail -6 gpurun_out/${R}_bench_fonts_stderr.log
timeout 600 python bench.py --config scaled --steps 5 --warmup 3 --no-cpu-baseline --no-parity-mode 2> gpurun_out/${R}_bench_scaled_stderr.log | tail -1 > gpurun_out/${R}_bench_scaled.json
tail -n 3 gpurun_out/${R}_bench_scaled_stderr.log
cut -c1-300 gpurun_out/${R}_bench_hier.json gpurun_out/${R}_bench_fonts.json gpurun_out/${R}_bench_scaled.json
timeout 900 python -m pytest tests -m gpu -q -x --no-header 2>&1 | tail -15 > gpurun_out/${R}_pytest_gpu.txt
tail -5 gpurun_out/${R}_pytest_gpu.txt
