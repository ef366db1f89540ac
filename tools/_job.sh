set -x
( time timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r02n_gputests.log 2>&1 ) 2> gpurun_out/r02n_gputests.time
tail -3 gpurun_out/r02n_gputests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > gpurun_out/r02n_smoke.log 2>&1; tail -2 gpurun_out/r02n_smoke.log
( time python bench.py > gpurun_out/r02n_bench.json 2> gpurun_out/r02n_bench.err ) 2> gpurun_out/r02n_bench.time
( time python bench.py --impl reference > gpurun_out/r02n_bench_ref.json 2> gpurun_out/r02n_bench_ref.err ) 2> gpurun_out/r02n_bench_ref.time
tail -c 600 gpurun_out/r02n_bench.json; cat gpurun_out/r02n_bench.time
