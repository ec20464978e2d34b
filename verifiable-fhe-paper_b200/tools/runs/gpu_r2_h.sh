#!/bin/bash
# round-2 GPU run H: quick A/B of the two commit orders + NTT bench + FRI phase
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T=verifiable-fhe-paper_b200/tools
$T/ntt_bench > gpurun_out/h_nb.txt 2>&1; cat gpurun_out/h_nb.txt
python $T/quick_commit_timing.py > gpurun_out/h_quick.txt 2>&1; cat gpurun_out/h_quick.txt
timeout 300 python bench.py --fri-commit-phase > gpurun_out/h_fri.json 2> gpurun_out/h_fri.err; echo "fri rc=$?"; tail -c 300 gpurun_out/h_fri.err; python -c "
import json; d=json.load(open('gpurun_out/h_fri.json')); print(d['value'], d['per_layer_host_calls_ms'], d['matches_oracle'], d['gpu_launches'])"
