#!/bin/bash
mkdir -p gpurun_out
for m in qkv proj ffn1; do
  DSVG_PLANES=2 timeout 200 ncu --set full --clock-control none --import-source on -k regex:linear_kernel -s 2 -c 1 \
      -o gpurun_out/x3lin_${m} -f python tools/prof_mode.py ${m} > /dev/null 2>&1
done
ls -la gpurun_out/x3lin_*
