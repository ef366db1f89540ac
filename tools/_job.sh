set -x
( time python bench.py > gpurun_out/r02j_bench.json 2> gpurun_out/r02j_bench.err ) 2> gpurun_out/r02j_bench.time
( time python bench.py --impl reference > gpurun_out/r02j_bench_ref.json 2> gpurun_out/r02j_bench_ref.err ) 2> gpurun_out/r02j_bench_ref.time
# launch lists (ncu serialises and runs cold: shares, not absolutes) -- only after the plain runs above exited
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02j_launches_bus.csv python tools/prof_run.py bus 10000 3 0 > gpurun_out/r02j_ncu_l_bus.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02j_launches_zz_batch.csv python tools/prof_run.py zz_batch 4096 3 0 > gpurun_out/r02j_ncu_l_zzb.log 2>&1
# full captures of the last evaluation's kernels; only the raw metric page comes home (the reports are too big)
ncu --set full --clock-control none -s 6 -c 3 -o /tmp/r02j_full_bus -f python tools/prof_run.py bus 10000 3 0 > gpurun_out/r02j_ncu_f_bus.log 2>&1
ncu -i /tmp/r02j_full_bus.ncu-rep --page raw --csv > gpurun_out/r02j_full_bus_raw.csv
ncu --set full --clock-control none -s 6 -c 3 -o /tmp/r02j_full_zzb -f python tools/prof_run.py zz_batch 4096 3 0 > gpurun_out/r02j_ncu_f_zzb.log 2>&1
ncu -i /tmp/r02j_full_zzb.ncu-rep --page raw --csv > gpurun_out/r02j_full_zzb_raw.csv
ncu --set full --clock-control none -k regex:g_gemm2 -s 40 -c 1 -o /tmp/r02j_full_gemm128 -f python tools/run_once.py synth128x512 frechet 1 > gpurun_out/r02j_ncu_f_g128.log 2>&1
ncu -i /tmp/r02j_full_gemm128.ncu-rep --page raw --csv > gpurun_out/r02j_full_gemm128_raw.csv
du -sh gpurun_out
