#!/bin/bash
# Round 2, runs on the GPU box (gpurun): launch list of the bench command + one `ncu --set full` capture per kernel.
# The .ncu-rep files embed the library image (~30 MB each), so they are reduced to CSV pages on the box and only those
# travel back: gpurun_out/r02_launches.csv, gpurun_out/r02_<name>_raw.csv, gpurun_out/r02_<name>_source.csv.gz
set -u
out=gpurun_out
mkdir -p $out
# every launch of: 3 warm-up fits, 1 timed fit (one CUDA graph per sweep: ncu lists the kernel nodes), 1 eager profiling fit
ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file $out/r02_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $out/r02_ncu_list.log 2>&1
echo "launch list rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file $out/r02_launches_c4.csv \
    python bench/c4_implicit.py --steps 1 --K 1 --sweeps 4 > $out/r02_ncu_list_c4.log 2>&1
echo "C4 launch list rc=$?"
cap() {   # name, driver + args, kernel regex, launches to skip, launches to take
    ncu --set full --clock-control none --import-source on -k regex:"$3" -s $4 -c $5 -o /tmp/$1 -f \
        python $2 > $out/r02_ncu_$1.log 2>&1
    echo "$1 capture rc=$?"
    ncu -i /tmp/$1.ncu-rep --page raw --csv > $out/r02_$1_raw.csv 2>/dev/null
    ncu -i /tmp/$1.ncu-rep --page source --csv 2>/dev/null | gzip > $out/r02_$1_source.csv.gz
    rm -f /tmp/$1.ncu-rep
}
cap asm "profiles/run_fit.py 2" 'k_assemble_r4' 8 1
cap solve "profiles/run_fit.py 2" 'k_micro_solve_fast' 6 1
cap solve_trunc "profiles/run_fit_converging.py 6" 'k_micro_solve_fast' 70 1
cap reduce "profiles/run_fit.py 2" 'k_moment_reduce' 6 1
cap hbm "profiles/run_hbm.py" 'k_paths|k_terminal|k_target|k_tt_eval|k_grad_step|k_interface_right' 0 8
cap iface_left "profiles/run_fit.py 1 x fuse_interface=0" 'k_interface_left' 3 1
ls -la $out | grep r02_
