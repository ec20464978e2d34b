#!/bin/bash
# round-2 GPU run Y (N GPUs): the NCCL tests, then ONE step chain sharded over the ranks (--chain-shard)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests -m gpu -q -k "nccl" > gpurun_out/y_pytest_nccl_n$N.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/y_pytest_nccl_n$N.txt; tail -5 gpurun_out/y_pytest_nccl_n$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --chain-steps 64 --chain-shard > gpurun_out/y_chain_shard_n$N.json 2> gpurun_out/y_chain_shard_n$N.err; echo "rc=$?"; tail -c 800 gpurun_out/y_chain_shard_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/y_chain_shard_n$N.json')); print('sharded chain', d['n_gpus'], d['value'], d['sharded_step']['matches_unsharded_step'], d['gpu_launches_per_step']); print(d['phase_ms_rank0'])"
