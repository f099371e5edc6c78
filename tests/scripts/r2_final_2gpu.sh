#!/bin/bash
# round 2: two-GPU checks (NCCL gradient parity test, the full bench line incl. the cfg-5 all-reduce under torchrun)
set -u
O=gpurun_out
L=$O/r2_final_2gpu.log
: > $L
nvidia-smi -L >> $L 2>&1
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k nccl 2>&1 | tail -4 >> $L
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $O/r02_bench_2gpu.json 2> $O/r02_bench_2gpu.err
tail -c 600 $O/r02_bench_2gpu.json >> $L
tail -5 $O/r02_bench_2gpu.err >> $L
tail -3 $L
