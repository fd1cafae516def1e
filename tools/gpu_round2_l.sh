#!/bin/bash
timeout 300 python -m pytest tests/test_kernels_gpu.py -m gpu -q --no-header -x -k "lean_epilogues" 2>&1 | tail -3
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:linear_kernel python tools/prof_mode.py head_dgrad 2>&1 | grep -E "gpu__time_duration" | tail -3
timeout 600 python bench.py --no-cpu-baseline --no-parity-mode --no-ref-gpu 2>/dev/null | tail -1 | cut -c1-330
