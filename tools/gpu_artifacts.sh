#!/bin/bash
# Runs on the GPU box (via gpurun): produces every measured artefact of a round under gpurun_out/.
#   bash tools/gpu_artifacts.sh r2
R=${1:-r2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > gpurun_out/${R}_nvidia_smi.csv
timeout 900 python -m pytest tests -m gpu -q --no-header 2>&1 | tail -12 > gpurun_out/${R}_pytest_gpu.txt
tail -3 gpurun_out/${R}_pytest_gpu.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${R}_smoke.txt 2>&1; tail -2 gpurun_out/${R}_smoke.txt
timeout 600 python bench.py 2> gpurun_out/${R}_bench_stderr.log | tail -1 > gpurun_out/${R}_bench_n1.json
timeout 400 python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tail -1 > gpurun_out/${R}_bench_reference.json
timeout 500 python bench.py --config fonts --steps 20 2>/dev/null | tail -1 > gpurun_out/${R}_bench_fonts.json
timeout 800 python bench.py --config scaled --steps 5 --warmup 3 2>/dev/null | tail -1 > gpurun_out/${R}_bench_scaled.json
cut -c1-260 gpurun_out/${R}_bench_n1.json gpurun_out/${R}_bench_fonts.json gpurun_out/${R}_bench_scaled.json gpurun_out/${R}_bench_reference.json
# launch list of one full train step, eager launches (cold-cache, serialised: compare SHARES)
DSVG_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_launches.csv \
    python tools/one_step.py 512 2 hier > gpurun_out/${R}_one_step.log 2>&1
# one full-set capture per GEMM role (lean epilogue mode) and per other hot kernel
for m in qkv ffn1 proj lnfwd mask dgrad head_dgrad logits; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:linear_kernel -s 2 -c 1 \
      -o gpurun_out/${R}_mode_${m} -f python tools/prof_mode.py ${m} > /dev/null 2>&1
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:outer_kernel -s 2 -c 1 -o gpurun_out/${R}_mode_outer -f python tools/prof_mode.py outer > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:attn_mma -s 2 -c 2 -o gpurun_out/${R}_mode_attn -f python tools/prof_mode.py attn > /dev/null 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:ln_bwd -s 1 -c 1 -o gpurun_out/${R}_mode_ln_bwd -f python tools/prof_mode.py ln_bwd > /dev/null 2>&1
# parity mode (bf16x3): launch list of one step, one capture per linear role at the path-level shape, the attention kernels
DSVG_PRECISION=bf16x3 DSVG_GRAPHS=0 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${R}_x3_launches.csv \
    python tools/one_step.py 512 2 hier > gpurun_out/${R}_x3_one_step.log 2>&1
for m in qkv ffn1 proj; do
  DSVG_PLANES=2 timeout 200 ncu --set full --clock-control none --import-source on -k regex:linear_kernel -s 2 -c 1 \
      -o gpurun_out/${R}_x3mode_${m} -f python tools/prof_mode.py ${m} > /dev/null 2>&1
done
DSVG_PLANES=2 timeout 200 ncu --set full --clock-control none --import-source on -k regex:attn_x3 -s 2 -c 2 -o gpurun_out/${R}_x3mode_attn -f python tools/prof_mode.py attn > /dev/null 2>&1
# gpurun copies back at most 64 MiB: keep the raw-metric CSV of every capture, and the full reports (with source) of three kernels
for rep in gpurun_out/${R}_*.ncu-rep; do
  ncu -i $rep --page raw --csv > ${rep%.ncu-rep}.raw.csv 2>/dev/null
  case $rep in
    *_mode_qkv.ncu-rep|*_mode_lnfwd.ncu-rep|*_x3mode_qkv.ncu-rep) ;;
    *) rm -f $rep ;;
  esac
done
ls gpurun_out | grep ${R}_ | wc -l
du -sh gpurun_out
