#!/bin/bash
# multi-GPU pass, sampler only: N GPUs ($1), optional emulation of more ranks' store volume ($2 = peer repeat)
N=${1:-2}; REP=${2:-1}
mkdir -p gpurun_out
ALABI_B200_PEER_REPEAT=$REP timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_mcmc_n${N}_r$REP.json 2> gpurun_out/bench_mcmc_n${N}_r$REP.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_mcmc_n${N}_r$REP.json')); m=d['mcmc']
print({k:m[k] for k in ('value','e2e','kernel_only','gathered_chain_ok','value_with_nccl_allgather_after_the_run')})
PY
