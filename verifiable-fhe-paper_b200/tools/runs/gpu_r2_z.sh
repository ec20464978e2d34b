#!/bin/bash
# round-2 GPU run Z (N GPUs): per-phase host times of the step, unsharded on one GPU and sharded over N
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python bench.py --chain-steps 64 > gpurun_out/z_chain_n1.json 2> gpurun_out/z_chain_n1.err; echo "rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --chain-steps 64 --chain-shard > gpurun_out/z_chain_shard_n$N.json 2> gpurun_out/z_chain_shard_n$N.err; echo "rc=$?"; tail -c 600 gpurun_out/z_chain_shard_n$N.err
python -c "
import json
a=json.load(open('gpurun_out/z_chain_n1.json')); b=json.load(open('gpurun_out/z_chain_shard_n$N.json'))
print('n1', a['value'], a['phase_ms_rank0'])
print('n$N', b['value'], b['sharded_step']['matches_unsharded_step'], b['phase_ms_rank0'])"
