#!/bin/bash
# Round 2, final build: GPU suite, smoke, bench line and reference-arm line of the same box (no profiler).
set -u
out=gpurun_out
mkdir -p $out
timeout 600 python -m pytest tests -m gpu -q > $out/final_pytest.txt 2>&1; echo "pytest rc=$?"; tail -2 $out/final_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" > $out/final_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 $out/final_smoke.txt
python bench.py --impl reference > $out/final_bench_ref.json 2> $out/final_bench_ref.err; echo "bench ref rc=$?"
python bench.py > $out/final_bench_1gpu.json 2> $out/final_bench_1gpu.err; echo "bench rc=$?"
