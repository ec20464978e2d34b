#!/bin/bash
# round-2 GPU run K: prove_openings front half + 128-bit leaf loads — parity suite, kernel timings
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T=verifiable-fhe-paper_b200/tools
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/k_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/k_pytest.txt
tail -15 gpurun_out/k_pytest.txt
(cd $T; ./poseidon_bench; ./poseidon_bench 135; ./poseidon_bench 20) > gpurun_out/k_pb.txt 2>&1; cat gpurun_out/k_pb.txt
python $T/quick_commit_timing.py > gpurun_out/k_quick.txt 2>&1; cat gpurun_out/k_quick.txt
