#!/bin/bash
# Round 2, last GPU call: GPU suite, bench line and smoke of the final build (no profiler), then one ncu launch list of the
# TT evaluation / gradient driver (k_tt_eval after its rewrite).  Every step has its own time limit.
set -u
out=gpurun_out
mkdir -p $out
t0=$(date +%s)
timeout 360 python -m pytest tests -m gpu -q --durations=15 > $out/last_pytest.txt 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - t0 ))"; tail -3 $out/last_pytest.txt
timeout 170 python bench.py > $out/last_bench_1gpu.json 2> $out/last_bench_1gpu.err; echo "bench rc=$? t=$(( $(date +%s) - t0 ))"
timeout 60 python -c "import __graft_entry__ as g; g.smoke()" > $out/last_smoke.txt 2>&1; echo "smoke rc=$? t=$(( $(date +%s) - t0 ))"; tail -1 $out/last_smoke.txt
timeout 60 python profiles/time_eval_grad.py > $out/last_eval_grad.txt 2>&1; echo "eval_grad rc=$? t=$(( $(date +%s) - t0 ))"; cat $out/last_eval_grad.txt
timeout 90 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 700 --csv --log-file $out/last_launches_eval_grad.csv python profiles/time_eval_grad.py > $out/last_ncu.log 2>&1; echo "ncu rc=$? t=$(( $(date +%s) - t0 ))"
