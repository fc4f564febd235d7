#!/bin/bash
# multi-GPU pass: NCCL check of the sharded operations, then the bench at N GPUs (N = $1)
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tools/multi_gpu_check.py > gpurun_out/multi_gpu_check_n$N.txt 2>&1; echo "check exit $?"; tail -3 gpurun_out/multi_gpu_check_n$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench exit $?"; tail -3 gpurun_out/bench_n$N.err; cat gpurun_out/bench_n$N.json
