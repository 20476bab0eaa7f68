#!/bin/bash
# Round 2, session 3, call A: truncated solve of cores with 65..128 unknowns (odd-even register-resident Jacobi with 16 lanes per
# column pair against the round-robin tournament with 32), GPU suite, the reference's own benchmark grid, bench lines.
set -u
out=gpurun_out
mkdir -p $out
( cd bench/micro && ./vh_solve vh_corpus.bin ) > $out/s3a_vh_solve.txt 2>&1; echo "vh_solve rc=$?"; tail -2 $out/s3a_vh_solve.txt
timeout 900 python -m pytest tests -m gpu -x -q > $out/s3a_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 $out/s3a_pytest.txt
python bench/sweep.py --vary d --values 10,20,50,100 --samples 2000 --rank 5 --basis 5 --sweeps 20 > $out/s3a_sweep_d.jsonl 2>&1; echo "sweep d rc=$?"
python bench/sweep.py --vary N --values 1000,2000,5000,10000,20000 --assets 10 --rank 5 --basis 5 --sweeps 20 > $out/s3a_sweep_N.jsonl 2>&1; echo "sweep N rc=$?"
python bench/sweep.py --vary r --values 2,3,4,5 --assets 10 --samples 2000 --basis 5 --sweeps 20 > $out/s3a_sweep_r.jsonl 2>&1; echo "sweep r rc=$?"
python bench.py > $out/s3a_bench_1gpu.json 2> $out/s3a_bench_1gpu.err; echo "bench rc=$?"
python bench.py --impl reference > $out/s3a_bench_ref.json 2> $out/s3a_bench_ref.err; echo "bench ref rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > $out/s3a_smoke.txt 2>&1; echo "smoke rc=$?"; tail -1 $out/s3a_smoke.txt
