#!/bin/bash
# Programmatic dependent launch, second call: the solve kernel's map / hint loads ahead of the wait, edge types of the captured
# sweep reported (stats graph_edges / graph_pdl_edges), main build and early-load variant, GPU suite with the option on.
set -u
out=gpurun_out
mkdir -p $out
t0=$(date +%s)
timeout 150 python bench/ab_pdl.py 4 > $out/pdl2_ab.jsonl 2> $out/pdl2_ab.err; echo "ab rc=$? t=$(( $(date +%s) - t0 ))"; cut -c1-400 $out/pdl2_ab.jsonl; tail -3 $out/pdl2_ab.err
BSDE_B200_LIB=$PWD/bsde-tensor-networks_b200/libbsde_b200_variant.so timeout 150 python bench/ab_pdl.py 4 > $out/pdl2_ab_early.jsonl 2> $out/pdl2_ab_early.err; echo "ab early rc=$? t=$(( $(date +%s) - t0 ))"; cut -c1-400 $out/pdl2_ab_early.jsonl; tail -3 $out/pdl2_ab_early.err
BSDE_B200_PDL=1 timeout 300 python -m pytest tests -m gpu -q > $out/pdl2_pytest.txt 2>&1; echo "pytest pdl=1 rc=$? t=$(( $(date +%s) - t0 ))"; tail -3 $out/pdl2_pytest.txt
