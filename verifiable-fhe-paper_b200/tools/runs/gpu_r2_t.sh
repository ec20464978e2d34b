#!/bin/bash
# round-2 GPU run T: launch list of the N=8-class (2^13 rows) chained step
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/t_chain13_launches.csv python bench.py --chain-steps 2 --chain-log-n 13 > gpurun_out/t_ncu.log 2>&1; echo "rc=$?"
