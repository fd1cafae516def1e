#!/bin/bash
# round-2 third GPU pass: fused GEMM + LayerNorm epilogues
mkdir -p gpurun_out
R=r2c
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -x --no-header -k "layernorm_fused" 2>&1 | tail -40 > gpurun_out/${R}_pytest_lnfused.txt
tail -25 gpurun_out/${R}_pytest_lnfused.txt
timeout 900 python -m pytest tests -m gpu -q --no-header 2>&1 | tail -40 > gpurun_out/${R}_pytest_gpu.txt
tail -12 gpurun_out/${R}_pytest_gpu.txt
timeout 600 python bench.py --no-cpu-baseline --no-parity-mode --no-ref-gpu 2> gpurun_out/${R}_bench_stderr.log | tail -1 > gpurun_out/${R}_bench_hier.json
DSVG_LN_FUSE=0 timeout 600 python bench.py --no-cpu-baseline --no-parity-mode --no-ref-gpu 2> gpurun_out/${R}_bench_nofuse_stderr.log | tail -1 > gpurun_out/${R}_bench_hier_nofuse.json
cut -c1-330 gpurun_out/${R}_bench_hier.json gpurun_out/${R}_bench_hier_nofuse.json
tail -3 gpurun_out/${R}_bench_stderr.log
DSVG_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches.csv \
    python tools/one_step.py 512 2 > gpurun_out/${R}_one_step.log 2>&1
python tools/launch_summary.py gpurun_out/${R}_launches.csv 420 2>/dev/null | head -50 > gpurun_out/${R}_launch_shares.txt
head -30 gpurun_out/${R}_launch_shares.txt
