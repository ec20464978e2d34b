#!/bin/bash
# round-2 GPU run F: zs debug, ncu --set full of hash_leaves and the NTT passes, tool timings
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T=verifiable-fhe-paper_b200/tools
python $T/dbg/zs_debug.py > gpurun_out/f_zs.txt 2>&1; cat gpurun_out/f_zs.txt | tail -15
$T/selftest > gpurun_out/f_selftest.txt 2>&1; tail -3 gpurun_out/f_selftest.txt
$T/poseidon_bench > gpurun_out/f_pb.txt 2>&1; cat gpurun_out/f_pb.txt
$T/ntt_bench > gpurun_out/f_nb.txt 2>&1; cat gpurun_out/f_nb.txt
ncu --set full --import-source on --clock-control none -k regex:hash_leaves -s 2 -c 1 -f -o gpurun_out/r2_hash_leaves_v12 $T/poseidon_bench > gpurun_out/f_ncu_hash.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:r16p -s 10 -c 2 -f -o gpurun_out/r2_ntt_v12 $T/ntt_bench > gpurun_out/f_ncu_ntt.log 2>&1
ls -la gpurun_out/*.ncu-rep
