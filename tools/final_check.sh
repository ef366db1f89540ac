python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 300 gpurun_out/final_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_ref.json 2>> gpurun_out/final_bench.err; head -c 200 gpurun_out/final_ref.json; echo
