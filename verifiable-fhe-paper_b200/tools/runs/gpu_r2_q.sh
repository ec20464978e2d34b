#!/bin/bash
# round-2 GPU run Q (2 GPUs): vpbs_commit_multi with single upload + peer copies
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "commit_multi or scattered or nccl" > gpurun_out/q_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/q_pytest.txt; tail -5 gpurun_out/q_pytest.txt
for G in 1 2; do timeout 600 python bench.py --multi-commit $G > gpurun_out/q_multi_g$G.json 2> gpurun_out/q_multi_g$G.err; echo "multi $G rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/q_multi_g$G.json')); print(d['n_gpus'], d['ms_per_step'], d['matches_single_gpu_commit'], d['e2e']['slowest_device_phase_ms'])"; done
