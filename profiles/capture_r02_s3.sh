#!/bin/bash
# Round 2, session 3 (k_assemble_r4 with V^T y in one tile; cores up to 512 unknowns): bench line first (no profiler), then
# the launch list of the same bench command and one `ncu --set full` capture of the assembly, solve and reduce kernels.
set -u
out=gpurun_out
mkdir -p $out
python bench.py > $out/s3_bench_1gpu.json 2> $out/s3_bench_1gpu.err; echo "bench rc=$?"
python bench.py --impl reference > $out/s3_bench_ref.json 2> $out/s3_bench_ref.err; echo "bench ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file $out/s3_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $out/s3_ncu_list.log 2>&1
echo "launch list rc=$?"
cap() {   # name, driver + args, kernel regex, launches to skip, launches to take
    ncu --set full --clock-control none --import-source on -k regex:"$3" -s $4 -c $5 -o /tmp/$1 -f \
        python $2 > $out/s3_ncu_$1.log 2>&1
    echo "$1 capture rc=$?"
    ncu -i /tmp/$1.ncu-rep --page raw --csv > $out/s3_$1_raw.csv 2>/dev/null
    ncu -i /tmp/$1.ncu-rep --page source --csv 2>/dev/null | gzip > $out/s3_$1_source.csv.gz
    rm -f /tmp/$1.ncu-rep
}
cap asm "profiles/run_fit.py 2" 'k_assemble_r4' 8 1
cap solve "profiles/run_fit.py 2" 'k_micro_solve_fast' 6 1
cap reduce "profiles/run_fit.py 2" 'k_moment_reduce' 6 1
ls -la $out | grep s3_
