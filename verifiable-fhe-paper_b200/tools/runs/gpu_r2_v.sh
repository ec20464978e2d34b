#!/bin/bash
# round-2 GPU run V: 3-product squaring — selftest, parity suite, bench, ncu of hash_leaves
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T=verifiable-fhe-paper_b200/tools
$T/selftest > gpurun_out/v_selftest.txt 2>&1; tail -3 gpurun_out/v_selftest.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/v_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/v_pytest.txt; tail -4 gpurun_out/v_pytest.txt
timeout 900 python bench.py > gpurun_out/v_bench.json 2> gpurun_out/v_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/v_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/v_bench.json"))
print("ms_per_step", d["ms_per_step"], "value", d["value"], "phase", d["phase_ms"])
print("frac", d["roofline"]["frac"], "whole", d["roofline_whole_commit"]["int_frac"])
print("e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"], "standin", d["step_standin"]["kernels_ms"], d["step_standin"]["resident_pipeline_ms"])
print("checks", d["self_checks"])
PY
timeout 600 python bench.py --chain-steps 128 > gpurun_out/v_chain.json 2> gpurun_out/v_chain.err; python -c "
import json; d=json.load(open('gpurun_out/v_chain.json')); print('chain', d['value'])"
ncu --set full --import-source on --clock-control none -k regex:hash_leaves -s 2 -c 1 -f -o gpurun_out/r2_hash_leaves_v14 $T/poseidon_bench > gpurun_out/v_ncu_hash.log 2>&1
