#!/bin/bash
# Round-2 profile capture (run under gpurun on one B200).  Every ncu run follows a plain run of the same command.
mkdir -p gpurun_out
timeout 150 python scratch/kbench.py 5 6 3 > gpurun_out/kb_plain_r2.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"estep_kernel|mstats_kernel" -c 4 -o gpurun_out/prof_stream_r2 -f python scratch/kbench.py 5 6 3 > gpurun_out/ncu_stream_r2.log 2>&1
timeout 100 python scratch/sbench.py 30 2000 40 2 > gpurun_out/sb_plain_r2f.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"search_persistent_kernel" -c 1 -o gpurun_out/prof_search_r2 -f python scratch/sbench.py 30 2000 40 1 > gpurun_out/ncu_search_r2.log 2>&1
CMD="python bench.py --config 5 --starts 6 --steps 1 --warmup 0 --no-cpu --no-e2e"
timeout 200 $CMD > gpurun_out/bench_for_launchlist_r2.json 2> gpurun_out/bench_for_launchlist_r2.err && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_r2.csv $CMD > gpurun_out/ncu_launches_r2.log 2>&1
wc -l gpurun_out/launches_r2.csv; ls -la gpurun_out/*_r2.ncu-rep
