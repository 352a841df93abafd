#!/bin/bash
# A/B of an environment switch on the headline bench: tools/ab_env.sh CER_STRIP   (runs =1 then =0)
VAR=${1:?variable name}
mkdir -p gpurun_out
for s in 1 0; do
  env $VAR=$s timeout 300 python bench.py --no-sub-records --no-library-bar --no-cpu-baseline > gpurun_out/ab_${VAR}_$s.json 2> gpurun_out/ab_${VAR}_$s.err
  echo "bench rc=$?"
  python - "$VAR" "$s" <<'PY'
import json, sys
var, s = sys.argv[1], sys.argv[2]
d = json.load(open(f"gpurun_out/ab_{var}_{s}.json"))
print(f"{var}={s}", round(d["value"]), round(d["ms_per_step"], 3), "long", round(d["long_run"]["value"]), "e2e", round(d["e2e"]["value"]), "ir50 ms", round(d["ir50"]["ms"], 3))
for r in d["ir50_layers"][:3]:
    print("   ", r["class"], r["variant"], r["launches"], r["ms"], round(r["tflops"]))
PY
done
