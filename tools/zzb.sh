# quick look at the 4096-pulse zz batch (C4): per-stage times in both modes
python bench.py --workload zz_batch --no-configs --no-cpu-baseline > gpurun_out/zzb.json 2> gpurun_out/zzb.err || tail -5 gpurun_out/zzb.err
python - <<PY
import json
r=json.load(open("gpurun_out/zzb.json"))
rf=r["roofline"]
print("frechet ms/step %.3f frac %.3f k1 %.3f k2 %.3f k3 %.3f exec_frac %s" % (r["ms_per_step"], rf["frac"], rf["k1_ms"], rf["k2_ms"], rf["k3_ms"], rf["k1_executed_frac"]))
t=r["taylor3"]; rf=t["roofline"]
print("taylor3 ms/step %.3f frac %.3f k1 %.3f k2 %.3f k3 %.3f" % (t["ms_per_step"], rf["frac"], rf["k1_ms"], rf["k2_ms"], rf["k3_ms"]))
PY
