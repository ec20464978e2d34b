#!/bin/bash
# round-2 GPU run M (2 GPUs): NCCL sharded-commit test, 2-rank bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/m2_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "nccl or commit_multi" -rs > gpurun_out/m2_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/m2_pytest.txt; tail -6 gpurun_out/m2_pytest.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/m2_bench.json 2> gpurun_out/m2_bench.err; echo "bench rc=$?"; tail -c 500 gpurun_out/m2_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/m2_bench.json"))
print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"], "eager", d["e2e_eager"]["value"])
print("shard", d["shard_commit"]); print("checks", d["self_checks"])
PY
