#!/bin/bash
# round 2: the full bench line on 4 GPUs under torchrun
set -u
O=gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 3 --warmup 3 > $O/r02_bench_4gpu.json 2> $O/r02_bench_4gpu.err
tail -c 300 $O/r02_bench_4gpu.json
