#!/bin/bash
# round-2 GPU run A: parity of the new multiply + per-block hashing, A/B kernel timings
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T=verifiable-fhe-paper_b200/tools
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/a_smi.txt
$T/selftest > gpurun_out/a_selftest.txt 2>&1
for b in pb_old pb_new nb_old nb_new; do echo "== $b" >> gpurun_out/a_kernels.txt; $T/$b >> gpurun_out/a_kernels.txt 2>&1; done
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/a_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/a_pytest.txt
python $T/quick_commit_timing.py > gpurun_out/a_quick.txt 2>&1
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/a_bench.json 2> gpurun_out/a_bench.err
tail -3 gpurun_out/a_pytest.txt; cat gpurun_out/a_kernels.txt gpurun_out/a_quick.txt; tail -c 1500 gpurun_out/a_bench.err
