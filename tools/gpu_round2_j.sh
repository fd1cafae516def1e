#!/bin/bash
mkdir -p gpurun_out
R=r2j
timeout 600 python -m pytest tests/test_trainer_protocol_gpu.py tests/test_kernels_gpu.py -m gpu -q --no-header -k "trainer or attention or cast_transpose" 2>&1 | tail -8
python tools/bench_misc.py 2>&1 | grep attn | tee gpurun_out/${R}_misc.txt
timeout 600 python bench.py --config scaled --steps 5 --warmup 3 --no-cpu-baseline --no-parity-mode --no-ref-gpu 2>/dev/null | tail -1 | cut -c1-330
timeout 600 python bench.py --config fonts --steps 10 --no-cpu-baseline --no-parity-mode --no-ref-gpu 2>/dev/null | tail -1 | cut -c1-330
