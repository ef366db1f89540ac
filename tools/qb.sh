#!/bin/bash
# quick bench summary: tools/qb.sh <workload> [extra bench args]
w=$1; shift
python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/qb_$w.json 2> gpurun_out/qb_$w.err || tail -5 gpurun_out/qb_$w.err
python - <<PY
import json
r=json.load(open("gpurun_out/qb_$w.json")); f=r["roofline"]
print("$w", "ms/step %.4f" % r["ms_per_step"], "k1 %.4f k2 %.4f k3 %.4f" % (f["k1_ms"], f["k2_ms"], f["k3_ms"]), "frac %.3f" % f["frac"], "e2e_ms %.4f" % r["e2e"]["ms_per_step"], "value %.4g" % r["value"])
PY
