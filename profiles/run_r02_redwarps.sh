#!/bin/bash
# Reduction kernel with 8 / 16 / 32 warps per block (builds libbsde_b200_rw8 / _rw16 / default) x 3 or 6 row groups, C3 fit.
set -u
out=gpurun_out
mkdir -p $out
t0=$(date +%s)
: > $out/redwarps_ab.jsonl
for lib in libbsde_b200.so libbsde_b200_rw16.so libbsde_b200_rw8.so; do
  BSDE_B200_LIB=$PWD/bsde-tensor-networks_b200/$lib timeout 120 python bench/ab_opts.py --fits 4 --rounds 2 "red_groups=3" "red_groups=6" "red_groups=2" >> $out/redwarps_ab.jsonl 2> $out/redwarps_ab.err; echo "$lib rc=$? t=$(( $(date +%s) - t0 ))"
done
cat $out/redwarps_ab.jsonl; tail -3 $out/redwarps_ab.err
