#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/d_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/d_pytest.txt
tail -25 gpurun_out/d_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/d_bench.json 2> gpurun_out/d_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/d_bench.err
timeout 300 python bench.py --fri-commit-phase > gpurun_out/d_fri.json 2> gpurun_out/d_fri.err; echo "fri rc=$?"; tail -c 600 gpurun_out/d_fri.err; cat gpurun_out/d_fri.json
