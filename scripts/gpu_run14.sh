#!/bin/bash
set -u
mkdir -p gpurun_out
exec > >(tee -a gpurun_out/run14.log) 2>&1
nvidia-smi -L
timeout 300 python -c "import torch; torch.zeros(1).cuda(); print('torch warm', torch.cuda.device_count())"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --bwd > gpurun_out/bench_c2_2gpu.json 2> gpurun_out/bench_c2_2gpu.err; cat gpurun_out/bench_c2_2gpu.json; tail -3 gpurun_out/bench_c2_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 | tail -1 | cut -c1-300
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-cpu-baseline --no-e2e | cut -c1-400
