#!/bin/bash
mkdir -p gpurun_out
R=r2d
python tools/bench_lnfuse.py all 2>&1 | tee gpurun_out/${R}_lnfuse.txt
for v in 1 2; do echo "variant $v"; DSVG_LN9=$v python tools/bench_lnfuse.py bwd 2>&1 | tee -a gpurun_out/${R}_lnfuse.txt; done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_kernel -s 2 -c 1 -o gpurun_out/${R}_lnbwd512 -f python tools/bench_lnfuse.py lnbwd512 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:linear_kernel -s 2 -c 1 -o gpurun_out/${R}_lnfwd -f python tools/bench_lnfuse.py lnfwd > /dev/null 2>&1
ls -la gpurun_out | grep ${R}
timeout 300 python -m pytest tests/test_pack.py -m gpu -q --no-header 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline --no-parity-mode --no-ref-gpu --steps 10 2>/dev/null | tail -1 | cut -c1-900
