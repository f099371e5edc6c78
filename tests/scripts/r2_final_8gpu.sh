#!/bin/bash
# round 2: the full bench line (headline + secondary cfg 1 / 3 / 4 / 5) on 8 GPUs under torchrun
set -u
O=gpurun_out
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 3 --warmup 3 > $O/r02_bench_8gpu.json 2> $O/r02_bench_8gpu.err
tail -c 400 $O/r02_bench_8gpu.json
tail -3 $O/r02_bench_8gpu.err
