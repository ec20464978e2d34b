#!/bin/bash
# round-2 GPU run N: evidence for the TMA NTT passes — ncu --set full, warm-cache DRAM traffic of a
# whole commit, launch list of the bench command, full bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T=verifiable-fhe-paper_b200/tools
ncu --set full --import-source on --clock-control none -k regex:r16t -s 12 -c 2 -f -o gpurun_out/r2_ntt_tma python $T/commit_workload.py 2 > gpurun_out/n_ncu_ntt.log 2>&1; tail -2 gpurun_out/n_ncu_ntt.log
ncu --cache-control none --clock-control none --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --csv --log-file gpurun_out/n_traffic_warm.csv python $T/commit_workload.py 3 > gpurun_out/n_traffic.log 2>&1; tail -1 gpurun_out/n_traffic.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/n_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/n_ncu_bench.log 2>&1; echo "ncu launches rc=$?"
timeout 600 python bench.py > gpurun_out/n_bench.json 2> gpurun_out/n_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/n_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/n_bench_ref.json 2> gpurun_out/n_bench_ref.err; echo "ref rc=$?"; cat gpurun_out/n_bench_ref.json | cut -c1-600
python - <<'PY'
import json
d = json.load(open("gpurun_out/n_bench.json"))
print("ms_per_step", d["ms_per_step"], "phase", d["phase_ms"], "whole", d["roofline_whole_commit"]["int_frac"], "frac", d["roofline"]["frac"], "hbm", d["roofline_hbm"]["frac"])
print("e2e", d["e2e"]["ms_per_step"], "eager", d["e2e_eager"]["ms_per_step"], "standin", d["step_standin"]["kernels_ms"], d["step_standin"]["resident_pipeline_ms"], d["step_standin"].get("constants_sigmas_commit_ms"))
print("cpu", d["cpu_baseline"]["value"], "checks", d["self_checks"])
PY
