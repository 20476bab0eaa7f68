#!/bin/bash
# Round 2, closing build (programmatic dependent launch, 16-warp reduction blocks, batched moment staging): GPU suite, bench
# line, then under ncu the launch list of two bench fits and full captures of the reduction and the solve kernel.
set -u
out=gpurun_out
mkdir -p $out
t0=$(date +%s)
timeout 300 python -m pytest tests -m gpu -q > $out/closing_pytest.txt 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - t0 ))"; tail -3 $out/closing_pytest.txt
timeout 170 python bench.py > $out/closing_bench_1gpu.json 2> $out/closing_bench_1gpu.err; echo "bench rc=$? t=$(( $(date +%s) - t0 ))"; tail -2 $out/closing_bench_1gpu.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 13000 --csv --log-file $out/closing_launches.csv \
    python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extras > $out/closing_ncu_list.log 2>&1
rc=$?; echo "launch list rc=$rc t=$(( $(date +%s) - t0 ))"
if [ $rc -ne 0 ]; then
  BSDE_B200_PDL=0 timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 13000 --csv --log-file $out/closing_launches.csv \
      python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extras > $out/closing_ncu_list.log 2>&1
  echo "launch list (pdl off) rc=$? t=$(( $(date +%s) - t0 ))"
fi
gzip -f $out/closing_launches.csv
cap() {   # name, kernel regex, launches to skip
    timeout 70 ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c 1 -o /tmp/$1 -f \
        python profiles/run_fit.py 2 > $out/closing_ncu_$1.log 2>&1
    echo "$1 capture rc=$? t=$(( $(date +%s) - t0 ))"
    ncu -i /tmp/$1.ncu-rep --page raw --csv > $out/closing_$1_raw.csv 2>/dev/null
    rm -f /tmp/$1.ncu-rep
}
cap reduce 'k_moment_reduce' 20
cap solve 'k_micro_solve_fast' 20
ls -la $out | grep closing_
