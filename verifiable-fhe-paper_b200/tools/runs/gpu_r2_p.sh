#!/bin/bash
# round-2 GPU run P: column-chunked leaf hashing on the resident host path — parity suite, bench, chain
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/p_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/p_pytest.txt
tail -8 gpurun_out/p_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/p_bench.json 2> gpurun_out/p_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/p_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/p_bench.json"))
print("ms_per_step", d["ms_per_step"], "whole", d["roofline_whole_commit"]["int_frac"], "frac", d["roofline"]["frac"])
print("e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"], "eager", d["e2e_eager"]["ms_per_step"], "standin", d["step_standin"]["kernels_ms"], d["step_standin"]["resident_pipeline_ms"])
print("checks", d["self_checks"])
PY
timeout 600 python bench.py --chain-steps 64 > gpurun_out/p_chain64.json 2> gpurun_out/p_chain64.err; echo "chain rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/p_chain64.json')); print(d['value'], d['gpu_launches_per_step'])"
