#!/bin/bash
# Round 2, build with programmatic dependent launch (default on, single GPU), 16-warp reduction blocks and the batched moment
# staging of the solve kernel: option A/B in one process, GPU suite, bench line, smoke.
set -u
out=gpurun_out
mkdir -p $out
t0=$(date +%s)
timeout 120 python bench/ab_opts.py --fits 4 --rounds 2 "pdl=0" "pdl=1" > $out/last2_ab.jsonl 2> $out/last2_ab.err; echo "ab rc=$? t=$(( $(date +%s) - t0 ))"; cat $out/last2_ab.jsonl; tail -3 $out/last2_ab.err
timeout 300 python -m pytest tests -m gpu -q > $out/last2_pytest.txt 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - t0 ))"; tail -3 $out/last2_pytest.txt
timeout 170 python bench.py > $out/last2_bench_1gpu.json 2> $out/last2_bench_1gpu.err; echo "bench rc=$? t=$(( $(date +%s) - t0 ))"
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $out/last2_smoke.txt 2>&1; echo "smoke rc=$? t=$(( $(date +%s) - t0 ))"; tail -1 $out/last2_smoke.txt
