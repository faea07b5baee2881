#!/bin/bash
# A/B of the tower's LayerNorm schedules on one box, back to back: tools/ab_fold.sh <tag> [extra env assignments for the fold arm]
tag=$1; shift
for i in a b; do for f in 1 0; do
  if [ $f = 1 ]; then envs="$*"; else envs=""; fi
  env HVLM_LN_FOLD=$f $envs python bench.py --mode forward --steps 20 --warmup 5 --no-cpu-baseline --no-eager-baseline > gpurun_out/${tag}_fold${f}_$i.json 2> gpurun_out/${tag}_fold${f}_$i.err
done; done
python - <<PY
import json
for n in ("fold1_a","fold0_a","fold1_b","fold0_b"):
    try:
        d=json.loads(open("gpurun_out/${tag}_%s.json"%n).read().strip().splitlines()[-1])
        st=d["stages"]
        print(n, d["ms_per_step"], d["value"], d["e2e"]["ms_per_step"], d["clocks"]["sm_mhz"], {k:(v["avg_us"],v["launches_per_step"]) for k,v in st.items() if k in ("layernorm","qkv_gemm","outproj_gemm","fc1_gemm","fc2_gemm","attention")})
    except Exception as e: print(n, "ERR", e)
PY
