#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
T=verifiable-fhe-paper_b200/tools
$T/pb_dmma > gpurun_out/c_dmma.txt 2>&1
for b in pb_t256 pb_t64; do echo "== $b"; $T/$b; done > gpurun_out/c_hash_variants.txt 2>&1
cat gpurun_out/c_dmma.txt gpurun_out/c_hash_variants.txt
