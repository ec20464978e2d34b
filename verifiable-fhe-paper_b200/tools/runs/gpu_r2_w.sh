#!/bin/bash
# round-2 GPU run W: pinned staging ring for pageable host columns (parity suite, bench with the
# e2e_pageable.inputs_only leg); NTT butterflies with carry materialisation on the FMA pipe (A/B)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T=verifiable-fhe-paper_b200/tools
for v in base add mul addmul base addmul; do echo "== $v"; $T/nb_$v; done > gpurun_out/w_nb.txt 2>&1; cat gpurun_out/w_nb.txt | tail -40
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/w_pytest.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/w_pytest.txt; tail -6 gpurun_out/w_pytest.txt
timeout 900 python bench.py > gpurun_out/w_bench.json 2> gpurun_out/w_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/w_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/w_bench.json"))
print("ms_per_step", d["ms_per_step"], "value", d["value"], "phase", d["phase_ms"])
print("frac", d["roofline"]["frac"], "whole", d["roofline_whole_commit"]["int_frac"])
print("e2e", d["e2e"]["ms_per_step"], "eager", d["e2e_eager"]["ms_per_step"], "pageable", d["e2e_pageable"]["ms_per_step"], d["e2e_pageable"]["inputs_only"])
print("checks", d["self_checks"])
PY
