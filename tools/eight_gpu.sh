N=${1:-8}
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline"
$T > gpurun_out/r01e_${N}gpu_weak_bus.json 2> gpurun_out/${N}g.err
$T --workload zz_batch > gpurun_out/r01e_${N}gpu_weak_zz_batch.json 2>> gpurun_out/${N}g.err
$T --shard time > gpurun_out/r01e_${N}gpu_time_bus.json 2>> gpurun_out/${N}g.err
$T --shard time --workload synth16x100000 > gpurun_out/r01e_${N}gpu_time_synth16x100000.json 2>> gpurun_out/${N}g.err
for f in gpurun_out/r01e_${N}gpu_*.json; do python - $f <<'PY'
import json,sys
try:
    r=json.loads(open(sys.argv[1]).read().strip().split("\n")[-1]); print(sys.argv[1], r["n_gpus"], r["scaling"], "ms %.4f value %.4g" % (r["ms_per_step"], r["value"]))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
tail -3 gpurun_out/${N}g.err
