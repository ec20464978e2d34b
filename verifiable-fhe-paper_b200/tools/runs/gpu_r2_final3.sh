#!/bin/bash
# round-2 FINAL evidence (1 GPU): parity suite, bench line of record + CPU arm, chains
# (with per-phase times), FRI phase, launch list, sanitizer over the new host / shard / sponge paths
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/f3_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/f3_pytest.txt; tail -4 gpurun_out/f3_pytest.txt
timeout 900 python bench.py > gpurun_out/f3_bench.json 2> gpurun_out/f3_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/f3_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f3_bench_ref.json 2> gpurun_out/f3_bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py --chain-steps 730 > gpurun_out/f3_chain730.json 2> gpurun_out/f3_chain730.err; echo "chain730 rc=$?"; tail -c 300 gpurun_out/f3_chain730.err
timeout 600 python bench.py --chain-steps 200 --chain-log-n 13 > gpurun_out/f3_chain_n8.json 2> gpurun_out/f3_chain_n8.err; echo "chain n8 rc=$?"
timeout 300 python bench.py --fri-commit-phase > gpurun_out/f3_fri.json 2> gpurun_out/f3_fri.err; echo "fri rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/f3_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/f3_ncu_bench.log 2>&1; echo "ncu rc=$?"
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python verifiable-fhe-paper_b200/tools/sanitizer_workload.py > gpurun_out/f3_sanitizer.txt 2>&1; echo "sanitizer rc=$?"; tail -4 gpurun_out/f3_sanitizer.txt
python - <<'PY'
import json
d = json.load(open("gpurun_out/f3_bench.json"))
print("ms_per_step", d["ms_per_step"], "value", d["value"], "phase", d["phase_ms"])
print("frac", d["roofline"]["frac"], "whole", d["roofline_whole_commit"]["int_frac"], "hbm", d["roofline_hbm"]["frac"])
print("e2e", d["e2e"]["ms_per_step"], d["e2e"]["value"], "eager", d["e2e_eager"]["ms_per_step"], "pageable", d["e2e_pageable"]["ms_per_step"], d["e2e_pageable"]["inputs_only"]["ms_per_step"], d["e2e_pageable"]["inputs_only"]["ms_per_step_driver_staging"])
print("standin", d["step_standin"]["kernels_ms"], d["step_standin"]["resident_pipeline_ms"], d["step_standin"].get("constants_sigmas_commit_ms"))
print("cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["us_per_permutation_per_core"], "checks", d["self_checks"])
for f in ("f3_chain730", "f3_chain_n8"):
    c = json.load(open("gpurun_out/%s.json" % f)); print(f, c["value"], c["gpu_launches_per_step"], c["full_pbs_730_steps_s"], c["phase_ms_rank0"])
f = json.load(open("gpurun_out/f3_fri.json")); print("fri", f["value"], f["matches_oracle"])
r = json.load(open("gpurun_out/f3_bench_ref.json")); print("ref", r["value"], r["ms_per_step"])
PY
