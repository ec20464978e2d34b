#!/bin/bash
# round-2 GPU run X (1 GPU): sharded resident batches (vpbs_ctx_set_shard) parity tests, whole suite,
# the unsharded chain after the refactoring
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/x_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/x_pytest.txt; tail -15 gpurun_out/x_pytest.txt
timeout 600 python bench.py --chain-steps 64 > gpurun_out/x_chain.json 2> gpurun_out/x_chain.err; echo "chain rc=$?"; tail -c 400 gpurun_out/x_chain.err
python -c "
import json; d=json.load(open('gpurun_out/x_chain.json')); print('chain', d['value'], d['gpu_launches_per_step'])"
