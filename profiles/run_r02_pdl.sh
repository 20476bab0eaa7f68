#!/bin/bash
# Round 2, programmatic dependent launch of the micro-step kernels: same-process A/B (option off/on, cores compared bit for
# bit), the early-load variant build, then the GPU suite and the SALSA / implicit driver with the option forced on.
set -u
out=gpurun_out
mkdir -p $out
t0=$(date +%s)
timeout 150 python bench/ab_pdl.py 4 > $out/pdl_ab.jsonl 2> $out/pdl_ab.err; echo "ab rc=$? t=$(( $(date +%s) - t0 ))"; cat $out/pdl_ab.jsonl; tail -3 $out/pdl_ab.err
BSDE_B200_LIB=$PWD/bsde-tensor-networks_b200/libbsde_b200_variant.so timeout 150 python bench/ab_pdl.py 4 > $out/pdl_ab_early.jsonl 2> $out/pdl_ab_early.err; echo "ab early rc=$? t=$(( $(date +%s) - t0 ))"; cat $out/pdl_ab_early.jsonl; tail -3 $out/pdl_ab_early.err
BSDE_B200_PDL=1 timeout 300 python -m pytest tests -m gpu -q > $out/pdl_pytest.txt 2>&1; echo "pytest pdl=1 rc=$? t=$(( $(date +%s) - t0 ))"; tail -3 $out/pdl_pytest.txt
