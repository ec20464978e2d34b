#!/bin/bash
# round-2 final evidence (1 GPU): parity suite, bench line of record + CPU arm, chains, FRI phase, launch list
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/final_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/final_pytest.txt; tail -4 gpurun_out/final_pytest.txt
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/final_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_ref.json 2> gpurun_out/final_bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --chain-steps 730 > gpurun_out/final_chain730.json 2> gpurun_out/final_chain730.err; echo "chain730 rc=$?"; tail -c 300 gpurun_out/final_chain730.err
timeout 600 python bench.py --chain-steps 64 --chain-eager > gpurun_out/final_chain64.json 2> gpurun_out/final_chain64.err; echo "chain64 rc=$?"
timeout 600 python bench.py --chain-steps 200 --chain-log-n 13 --chain-eager > gpurun_out/final_chain_n8.json 2> gpurun_out/final_chain_n8.err; echo "chain n8 rc=$?"
timeout 300 python bench.py --fri-commit-phase > gpurun_out/final_fri.json 2> gpurun_out/final_fri.err; echo "fri rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/final_ncu_bench.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json
d = json.load(open("gpurun_out/final_bench.json"))
print("ms_per_step", d["ms_per_step"], "value", d["value"], "phase", d["phase_ms"])
print("frac", d["roofline"]["frac"], "whole", d["roofline_whole_commit"]["int_frac"], "hbm", d["roofline_hbm"]["frac"])
print("e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"], "eager", d["e2e_eager"]["ms_per_step"], "pageable", d["e2e_pageable"]["ms_per_step"])
print("standin", d["step_standin"]["kernels_ms"], d["step_standin"]["resident_pipeline_ms"], d["step_standin"].get("constants_sigmas_commit_ms"))
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["us_per_permutation_per_core"], "checks", d["self_checks"])
for f in ("final_chain730", "final_chain64", "final_chain_n8"):
    c = json.load(open("gpurun_out/%s.json" % f)); print(f, c["value"], c.get("eager_commits_ms_per_step"), c["gpu_launches_per_step"], c["full_pbs_730_steps_s"])
f = json.load(open("gpurun_out/final_fri.json")); print("fri", f["value"], f["per_layer_host_calls_ms"], f["matches_oracle"])
r = json.load(open("gpurun_out/final_bench_ref.json")); print("ref", r["value"], r["ms_per_step"])
PY
