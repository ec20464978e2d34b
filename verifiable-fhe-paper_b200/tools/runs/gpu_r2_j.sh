#!/bin/bash
# round-2 GPU run J: TMA tile loads in the 256-point NTT passes — parity suite, timings, sanitizer
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T=verifiable-fhe-paper_b200/tools
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/j_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/j_pytest.txt
tail -8 gpurun_out/j_pytest.txt
python $T/quick_commit_timing.py > gpurun_out/j_quick.txt 2>&1; cat gpurun_out/j_quick.txt
timeout 900 compute-sanitizer --tool memcheck python $T/sanitizer_workload.py > gpurun_out/j_sanitizer.txt 2>&1; tail -4 gpurun_out/j_sanitizer.txt
