#!/bin/bash
mkdir -p gpurun_out
R=r2i
echo "== default (double-buffered gmma fwd, hd32 grid 32/SM)"; python tools/bench_misc.py 2>&1 | grep attn | tee gpurun_out/${R}_misc.txt
echo "== DSVG_GMMA_DB=0"; DSVG_GMMA_DB=0 python tools/bench_misc.py 2>&1 | grep attn | tee -a gpurun_out/${R}_misc.txt
echo "== DSVG_ATTN_GRID=occ"; DSVG_ATTN_GRID=occ python tools/bench_misc.py 2>&1 | grep attn | tee -a gpurun_out/${R}_misc.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:attn_gmma -s 2 -c 2 -o gpurun_out/${R}_gattn -f python tools/prof_mode.py gattn > /dev/null 2>&1
ls -la gpurun_out | grep ${R}
