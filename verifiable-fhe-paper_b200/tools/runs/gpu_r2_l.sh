#!/bin/bash
# round-2 GPU run L: chained step stand-ins (configs 0 / 2-3), step-shape parity, bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "step_shapes or prove_openings" > gpurun_out/l_pytest.txt 2>&1; tail -3 gpurun_out/l_pytest.txt
timeout 600 python bench.py --chain-steps 64 --chain-eager > gpurun_out/l_chain64.json 2> gpurun_out/l_chain64.err; echo "chain64 rc=$?"; tail -c 600 gpurun_out/l_chain64.err; cat gpurun_out/l_chain64.json
timeout 600 python bench.py --chain-steps 730 > gpurun_out/l_chain730.json 2> gpurun_out/l_chain730.err; echo "chain730 rc=$?"; cat gpurun_out/l_chain730.json
timeout 600 python bench.py --chain-steps 200 --chain-log-n 13 --chain-eager > gpurun_out/l_chain_n8.json 2> gpurun_out/l_chain_n8.err; echo "chain n8 rc=$?"; tail -c 300 gpurun_out/l_chain_n8.err; cat gpurun_out/l_chain_n8.json
