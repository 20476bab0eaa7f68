for v in "" "_variant"; do
  BSDE_B200_LIB=$PWD/bsde-tensor-networks_b200/libbsde_b200$v.so python bench.py --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/ph$v.json 2> gpurun_out/ph$v.err
  python - "gpurun_out/ph$v.json" <<PY
import json,sys
j=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[1], round(j["value"]/1e6,2), "M", j["roofline"]["class_us_per_launch"], {k:round(v) for k,v in j["roofline"]["solve_kernel_phase_cycles_per_launch"].items()})
PY
done
