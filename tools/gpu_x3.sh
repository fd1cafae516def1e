#!/bin/bash
# parity mode (bf16x3): tests, launch shares of one hier step, step time of hier / fonts
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -x 2>&1 | tail -8 > gpurun_out/x3_pytest.txt
tail -4 gpurun_out/x3_pytest.txt
DSVG_PRECISION=bf16x3 DSVG_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv \
   --log-file gpurun_out/x3_launches.csv python tools/one_step.py 512 2 hier > gpurun_out/x3_one_step.log 2>&1
python tools/launch_summary.py gpurun_out/x3_launches.csv 409 > gpurun_out/x3_shares.txt
head -16 gpurun_out/x3_shares.txt
for cfg in hier fonts; do
  timeout 600 python bench.py --config $cfg --precision bf16x3 --steps 10 --warmup 3 --no-cpu-baseline --no-parity-mode --no-ref-gpu 2>gpurun_out/x3_bench_$cfg.err | tail -1 > gpurun_out/x3_bench_$cfg.json
  python -c "import json;b=json.load(open('gpurun_out/x3_bench_$cfg.json'));print('$cfg bf16x3',b['ms_per_step'],b['value'],b['run']['final_loss'],b['roofline']['hbm'])"
done
