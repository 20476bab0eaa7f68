#!/bin/bash
# Runs on the GPU box (gpurun): launch list of the timed bench fit + one `ncu --set full` capture per hot kernel.
# The .ncu-rep files embed the whole library image (~30 MB each), so they are reduced to CSV pages on the box and
# only those travel back:  gpurun_out/launches.csv, gpurun_out/<name>_raw.csv, gpurun_out/<name>_source.csv.gz
set -u
out=gpurun_out
mkdir -p $out
ncu --metrics gpu__time_duration.sum --clock-control none -s 24070 -c 1700 --csv --log-file $out/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline > $out/ncu_list.log 2>&1
echo "launch list rc=$?"
cap() {   # name, kernel regex, launches to skip, launches to take
    ncu --set full --clock-control none --import-source on -k regex:"$2" -s $3 -c $4 -o /tmp/$1 -f \
        python profiles/run_fit.py 2 > $out/ncu_$1.log 2>&1
    echo "$1 capture rc=$?"
    ncu -i /tmp/$1.ncu-rep --page raw --csv > $out/$1_raw.csv 2>/dev/null
    ncu -i /tmp/$1.ncu-rep --page source --csv 2>/dev/null | gzip > $out/$1_source.csv.gz
    rm -f /tmp/$1.ncu-rep
}
cap asm 'k_assemble_r4' 8 1
cap solve 'k_micro_solve_fast' 6 1
cap reduce 'k_moment_reduce' 6 1
ls -la $out
