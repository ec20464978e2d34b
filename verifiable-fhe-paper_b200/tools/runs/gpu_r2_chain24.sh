#!/bin/bash
# round-2 (4 GPUs): configs[4] stand-in at G = 2 and 4 (G = 1 and 8 are in gpu_r2_final.sh / gpu_r2_m8.sh)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for G in 2 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 2952$G bench.py --gpus $G --chain-steps 64 > gpurun_out/chain_g$G.json 2> gpurun_out/chain_g$G.err; echo "chain $G rc=$?"
  python -c "
import json; d=json.load(open('gpurun_out/chain_g$G.json')); print(d['n_gpus'], d['value'], d['steps_per_s_all_gpus'])"
done
