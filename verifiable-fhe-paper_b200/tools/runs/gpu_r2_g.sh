#!/bin/bash
# round-2 GPU run G: block-by-block commit order — parity suite + bench
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/g_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/g_pytest.txt
tail -12 gpurun_out/g_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/g_bench.json 2> gpurun_out/g_bench.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/g_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/g_bench.json"))
print("ms_per_step", d["ms_per_step"], "phase", d["phase_ms"], "whole", d["roofline_whole_commit"]["int_frac"], "frac", d["roofline"]["frac"])
print("e2e", d["e2e"]["ms_per_step"], "eager", d["e2e_eager"]["ms_per_step"], "standin", d["step_standin"]["kernels_ms"], d["step_standin"]["resident_pipeline_ms"])
print("checks", d["self_checks"])
PY
