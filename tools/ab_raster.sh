for r in 0 1 0 1; do
CER_RASTER=$r timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-library-bar --no-sub-records 2>/dev/null > gpurun_out/ab_raster_$r.json
python - <<PY
import json
j=[json.loads(l) for l in open("gpurun_out/ab_raster_$r.json") if l.startswith("{")][0]
print("raster", $r, "value", round(j["value"]), "long", round(j["long_run"]["value"]), "e2e", round(j["e2e"]["value"]), "ir50_ms", round(j["ir50"]["ms"],3))
for L in j["ir50_layers"]:
    if "128" in str(L.get("layer","")) or "raster" in str(L.get("kernel","")) or "bres" in str(L.get("kernel","")):
        print("   ", {k:(round(v,3) if isinstance(v,float) else v) for k,v in L.items()})
PY
done
