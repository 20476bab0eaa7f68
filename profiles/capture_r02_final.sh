#!/bin/bash
# Round 2, final build, runs on the GPU box (gpurun): launch lists of the bench fit and of the converging fit, `ncu --set full`
# captures of the truncated solve and of the assembly kernels of the drop-in convention (Legendre in-kernel, explicit values).
# Reports are reduced to CSV pages on the box; only those travel back (gpurun_out/r02f_*).
set -u
out=gpurun_out
mkdir -p $out
python profiles/run_fit_legendre.py legendre 2 > $out/r02f_legendre_time.txt 2>&1
python profiles/run_fit_legendre.py phi 2 > $out/r02f_phi_time.txt 2>&1
tail -1 $out/r02f_legendre_time.txt $out/r02f_phi_time.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file $out/r02f_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > $out/r02f_ncu_list.log 2>&1
echo "launch list rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file $out/r02f_launches_conv.csv \
    python profiles/run_fit_converging.py 6 > $out/r02f_ncu_list_conv.log 2>&1
echo "converging launch list rc=$?"
cap() {   # name, driver + args, kernel regex, launches to skip, launches to take
    ncu --set full --clock-control none --import-source on -k regex:"$3" -s $4 -c $5 -o /tmp/$1 -f \
        python $2 > $out/r02f_ncu_$1.log 2>&1
    echo "$1 capture rc=$?"
    ncu -i /tmp/$1.ncu-rep --page raw --csv > $out/r02f_$1_raw.csv 2>/dev/null
    ncu -i /tmp/$1.ncu-rep --page source --csv 2>/dev/null | gzip > $out/r02f_$1_source.csv.gz
    rm -f /tmp/$1.ncu-rep
}
cap solve_trunc "profiles/run_fit_converging.py 6" 'k_micro_solve_fast' 70 1
cap asm_legendre "profiles/run_fit_legendre.py legendre 1" 'k_assemble' 20 1
cap asm_phi "profiles/run_fit_legendre.py phi 1" 'k_assemble' 20 1
ls -la $out | grep r02f_
