#!/bin/bash
# 2-GPU pass (gpurun --gpus 2): the trainer call sequence under a real 2-device nn.DataParallel, the NCCL data-parallel
# gradient test, and the N=2 bench line.  usage: bash tools/gpu_2gpu.sh r2
R=${1:-r2}
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m pytest tests/test_trainer_protocol_gpu.py -m gpu -q --no-header 2>&1 | tail -25
timeout 900 python -m pytest tests/test_model_gpu.py -m gpu -q --no-header -k "nccl" 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2> gpurun_out/${R}_bench_n2_stderr.log | tail -1 > gpurun_out/${R}_bench_n2.json
cut -c1-300 gpurun_out/${R}_bench_n2.json
