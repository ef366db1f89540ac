# general-path quick look: parity tests of the path + one evaluation per workload
python -m pytest tests -m gpu -x -q -k "general_path or fullsize or cavity or synthetic or sharding" 2>&1 | tail -4
for wl in cavity20 cavity40 synth32x4000 synth64x2000 synth128x512 synth256x296; do python tools/run_once.py $wl frechet 3 2>&1 | tail -1; done
