#!/bin/bash
mkdir -p gpurun_out
for t in 0 1; do echo "== DSVG_OUTER_TALL=$t"; DSVG_OUTER_TALL=$t timeout 300 python tools/bench_outer.py; done 2>&1 | tee gpurun_out/tall_shapes.txt
