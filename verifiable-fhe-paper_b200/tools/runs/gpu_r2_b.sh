#!/bin/bash
# round-2 GPU run B: ncu --set full with source counters for hash_leaves and the NTT passes
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T=verifiable-fhe-paper_b200/tools
ncu --set full --import-source on --clock-control none -k regex:hash_leaves -s 2 -c 1 -f -o gpurun_out/r2_hash_leaves_v12 $T/pb_new > gpurun_out/b_ncu_hash.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:r16p -s 10 -c 2 -f -o gpurun_out/r2_ntt_v12 $T/nb_new > gpurun_out/b_ncu_ntt.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -3 gpurun_out/b_ncu_hash.log gpurun_out/b_ncu_ntt.log
