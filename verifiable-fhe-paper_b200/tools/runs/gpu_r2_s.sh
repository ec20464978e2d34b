#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "eval_ext2 or openings or proof_of_work or fri" > gpurun_out/s_pytest.txt 2>&1; tail -3 gpurun_out/s_pytest.txt
timeout 600 python bench.py --chain-steps 64 > gpurun_out/s_chain64.json 2> gpurun_out/s_chain64.err; echo "chain rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/s_chain64.json')); print(d['value'], d['gpu_launches_per_step'])"
timeout 300 python bench.py --fri-commit-phase > gpurun_out/s_fri.json 2> gpurun_out/s_fri.err; python -c "
import json; d=json.load(open('gpurun_out/s_fri.json')); print(d['value'], d['matches_oracle'], d['pow_witness_valid'])"
