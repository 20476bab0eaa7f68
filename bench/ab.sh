#!/bin/bash
# Same-box A/B of library builds and run-time switches (run inside ONE gpurun call: fit times differ by +-1 % between boxes).
#   bash bench/ab.sh "" "|red_groups=1" "_variant" "_variant|gate_tier2=0"
# Each argument is "<library suffix>|<option=value,...>": libbsde_b200<suffix>.so (built with
# `python bsde-tensor-networks_b200/build.py -DNAME=VALUE` -> suffix _variant) and bsde_ctx_set_option switches
# (INTEGRATION.md section 4).  Prints samples*sweeps/s, ms per fit and the per-class times of bench.py for each.
for spec in "$@"; do
  v="${spec%%|*}"; opts="${spec#*|}"; [ "$opts" == "$spec" ] && opts=""
  args=""; IFS=',' read -ra O <<< "$opts"; for o in "${O[@]}"; do [ -n "$o" ] && args="$args --option $o"; done
  lib=$PWD/bsde-tensor-networks_b200/libbsde_b200$v.so
  tag=$(echo "$spec" | tr -c 'A-Za-z0-9_=' '_')
  BSDE_B200_LIB=$lib python bench.py --steps 4 --warmup 2 --no-e2e --no-cpu-baseline $args > gpurun_out/bv$tag.json 2> gpurun_out/bv$tag.err
  python - "$spec" "gpurun_out/bv$tag" <<PY
import json,sys
v,f=sys.argv[1],sys.argv[2]
try:
    j=json.loads(open(f+".json").read().strip().splitlines()[-1])
    print("variant", repr(v), round(j["value"]/1e6,2), "M", round(j["ms_per_step"],1), "ms", {k:(round(x,1) if x else x) for k,x in j["roofline"]["class_us_per_launch"].items()}, j["solves_per_fit"]["cholesky"])
except Exception as e:
    print("variant", repr(v), "failed", e); print(open(f+".err").read()[-800:])
PY
done
