#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T=verifiable-fhe-paper_b200/tools
python $T/quick_commit_timing.py 2>&1 | head -2
ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --csv --log-file gpurun_out/o_traffic_warm.csv python $T/commit_workload.py 3 > gpurun_out/o_traffic.log 2>&1; tail -1 gpurun_out/o_traffic.log
