N=${1:-8}
for w in synth64x100000 synth128x100000 synth256x100000; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 --no-cpu-baseline --shard time --workload $w > gpurun_out/r01f_${N}gpu_time_$w.json 2> gpurun_out/r01f_${N}gpu_time_$w.err || tail -5 gpurun_out/r01f_${N}gpu_time_$w.err
  python - gpurun_out/r01f_${N}gpu_time_$w.json <<'PY'
import json,sys
try:
    r=json.loads(open(sys.argv[1]).read().strip().split("\n")[-1]); print(sys.argv[1], r["n_gpus"], r["scaling"], "ms %.2f value %.4g" % (r["ms_per_step"], r["value"]))
except Exception as e: print(sys.argv[1], "ERR", e)
PY
done
