#!/bin/bash
# round-2 GPU run I: integer-biased full rounds — selftest, parity suite, bench, ncu of hash_leaves
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T=verifiable-fhe-paper_b200/tools
$T/selftest > gpurun_out/i_selftest.txt 2>&1; tail -3 gpurun_out/i_selftest.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/i_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/i_pytest.txt
tail -5 gpurun_out/i_pytest.txt
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/i_bench.json 2> gpurun_out/i_bench.err; echo "bench rc=$?"
tail -c 800 gpurun_out/i_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/i_bench.json"))
print("ms_per_step", d["ms_per_step"], "phase", d["phase_ms"], "whole", d["roofline_whole_commit"]["int_frac"], "frac", d["roofline"]["frac"])
print("e2e", d["e2e"]["ms_per_step"], "eager", d["e2e_eager"]["ms_per_step"], "standin", d["step_standin"]["kernels_ms"], d["step_standin"]["resident_pipeline_ms"])
print("cpu", d["cpu_baseline"]["value"], "checks", d["self_checks"])
PY
ncu --set full --import-source on --clock-control none -k regex:hash_leaves -s 2 -c 1 -f -o gpurun_out/r2_hash_leaves_v13 $T/poseidon_bench > gpurun_out/i_ncu_hash.log 2>&1
