#!/bin/bash
# A/B of environment switches on one box: tools/ab_env.sh <tag> "ENV1=a ENV2=b" "ENV1=c" ...  (each config is run twice, interleaved)
tag=$1; shift
i=0
for rep in a b; do
  j=0
  for cfg in "$@"; do
    env $cfg python bench.py --mode forward --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/${tag}_${j}_$rep.json 2> gpurun_out/${tag}_${j}_$rep.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/${tag}_${j}_$rep.json").read().strip().splitlines()[-1]); st=d["stages"]
print("$rep [$cfg]", d["ms_per_step"], d["e2e"]["ms_per_step"], d["clocks"]["sm_mhz"], {k.replace("_gemm",""):v["avg_us"] for k,v in st.items() if k in ("qkv_gemm","outproj_gemm","fc1_gemm","fc2_gemm","attention")})
PY
    j=$((j+1))
  done
done
