#!/bin/bash
# A/B of one environment switch on one box: bash tools/ab_env.sh VAR  (runs VAR=0 / default, twice each)
V=$1
for r in 0 1 0 1; do
  if [ $r = 0 ]; then export $V=0; else unset $V; fi
  timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-bar --no-sub-records 2>/dev/null > gpurun_out/ab_${V}_$r.json
  python - <<PY
import json
j=[json.loads(l) for l in open("gpurun_out/ab_${V}_$r.json") if l.startswith("{")][0]
print("$V", "off" if $r == 0 else "default", "value", round(j["value"]), "long", round(j["long_run"]["value"]), "e2e", round(j["e2e"]["value"]))
print("   ", " | ".join("%s%s:%.3f" % (L["class"].split()[0], L["class"].split()[1], L["ms"]) for L in j["ir50_layers"]))
PY
done
