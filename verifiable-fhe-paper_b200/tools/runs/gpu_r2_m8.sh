#!/bin/bash
# round-2 GPU run M (8 GPUs): NCCL sharded-commit test, 8-rank bench line, 8 chains (configs[4]), 4-rank line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/m8_smi.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "nccl" -rs > gpurun_out/m8_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/m8_pytest.txt; tail -4 gpurun_out/m8_pytest.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/m8_bench.json 2> gpurun_out/m8_bench.err; echo "bench8 rc=$?"; tail -c 300 gpurun_out/m8_bench.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/m4_bench.json 2> gpurun_out/m4_bench.err; echo "bench4 rc=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --chain-steps 64 > gpurun_out/m8_chain.json 2> gpurun_out/m8_chain.err; echo "chain8 rc=$?"; cat gpurun_out/m8_chain.json | cut -c1-900
python - <<'PY'
import json
for f in ("m8_bench", "m4_bench"):
    d = json.load(open("gpurun_out/%s.json" % f))
    print(f, "value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "eager", d["e2e_eager"]["value"])
    print("shard", {k: d["shard_commit"].get(k) for k in ("ms_per_step", "single_gpu_ms", "strong_scaling_efficiency", "matches_single_gpu")}, "checks", d["self_checks"])
PY
