#!/bin/bash
# 2-GPU bench line of the closing build (16-warp reduction blocks; the sharded sweeps keep plain launches): replica check included.
set -u
out=gpurun_out
mkdir -p $out
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $out/last2_bench_2gpu.json 2> $out/last2_bench_2gpu.err; echo "bench 2gpu rc=$?"
tail -c 1500 $out/last2_bench_2gpu.json; tail -3 $out/last2_bench_2gpu.err
