#!/bin/bash
# round-2 GPU run U: register-only short strided passes — parity suite, 2^13 chain, quick timings
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/u_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/u_pytest.txt; tail -4 gpurun_out/u_pytest.txt
timeout 600 python bench.py --chain-steps 200 --chain-log-n 13 --chain-eager > gpurun_out/u_chain_n8.json 2> gpurun_out/u_chain_n8.err; echo "chain n8 rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/u_chain_n8.json')); print(d['value'], d['eager_commits_ms_per_step'], d['gpu_launches_per_step'])"
timeout 300 python bench.py --fri-commit-phase > gpurun_out/u_fri.json 2> gpurun_out/u_fri.err; python -c "
import json; d=json.load(open('gpurun_out/u_fri.json')); print('fri', d['value'], d['matches_oracle'])"
python verifiable-fhe-paper_b200/tools/quick_commit_timing.py 2>&1 | tail -2
