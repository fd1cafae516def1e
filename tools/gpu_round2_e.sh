#!/bin/bash
mkdir -p gpurun_out
R=r2e
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --no-header -k "layernorm_fused" 2>&1 | tail -15
python tools/bench_lnfuse.py all 2>&1 | tee gpurun_out/${R}_lnfuse.txt
DSVG_LN9=2 python tools/bench_lnfuse.py bwd 2>&1 | tee -a gpurun_out/${R}_lnfuse.txt
timeout 600 python bench.py --no-cpu-baseline --no-parity-mode --no-ref-gpu --steps 10 2>/dev/null | tail -1 | cut -c1-400
