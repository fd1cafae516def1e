#!/bin/bash
# parity-mode (bf16x3) attention on tensor cores: tests, then the parity-mode step of the three bench configs
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -x 2>&1 | tail -8 > gpurun_out/x3_pytest.txt
tail -4 gpurun_out/x3_pytest.txt
for cfg in hier fonts scaled; do
  steps=10; [ $cfg = scaled ] && steps=4
  timeout 600 python bench.py --config $cfg --precision bf16x3 --steps $steps --warmup 3 --no-cpu-baseline --no-parity-mode --no-ref-gpu 2>gpurun_out/x3_bench_$cfg.err | tail -1 > gpurun_out/x3_bench_$cfg.json
  python -c "import json;b=json.load(open('gpurun_out/x3_bench_$cfg.json'));print('$cfg bf16x3',b['ms_per_step'],b['value'],b['run']['final_loss'])"
done
