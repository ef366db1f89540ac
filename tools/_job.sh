set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02b_pytest.log
for wl in synth64x2000 synth128x512 synth256x296 cavity20 cavity40; do
  python tools/run_once.py $wl frechet 3 > gpurun_out/r02b_plain_$wl.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02b_launches_$wl.csv python tools/run_once.py $wl frechet 2 > gpurun_out/r02b_ncu_$wl.log 2>&1
done
( time python bench.py > gpurun_out/r02b_bench.json 2> gpurun_out/r02b_bench.err ) 2> gpurun_out/r02b_bench.time
( time python bench.py --impl reference > gpurun_out/r02b_bench_ref.json 2> gpurun_out/r02b_bench_ref.err ) 2> gpurun_out/r02b_bench_ref.time
