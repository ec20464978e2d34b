#!/bin/bash
# round-2 GPU run R: launch list of the chained N=1024 step (where do the 14.3 ms go)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r_chain_launches.csv python bench.py --chain-steps 2 > gpurun_out/r_ncu_chain.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/r_ncu_chain.log | cut -c1-300
