#!/bin/bash
# round-2 GPU run E (re-entry): full GPU parity suite, bench line, FRI commit phase, launch list
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/e_smi.txt
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/e_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/e_pytest.txt
tail -25 gpurun_out/e_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/e_bench.err
timeout 300 python bench.py --fri-commit-phase > gpurun_out/e_fri.json 2> gpurun_out/e_fri.err; echo "fri rc=$?"; tail -c 600 gpurun_out/e_fri.err; cat gpurun_out/e_fri.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/e_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --e2e-steps 1 > gpurun_out/e_ncu_bench.log 2>&1; echo "ncu rc=$?"
