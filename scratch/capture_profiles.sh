#!/bin/bash
# Profile capture for profiles/ (run under gpurun on one B200).  Every ncu run follows a plain run of the same command.
mkdir -p gpurun_out
python scratch/kbench.py 5 4 3 > gpurun_out/kb_plain_r1b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"estep_kernel|mstats_kernel" -c 4 -o gpurun_out/prof_stream_r1b -f python scratch/kbench.py 5 4 3 > gpurun_out/ncu_a.log 2>&1
python scratch/sbench.py 30 2000 40 1 > gpurun_out/sb_plain_r1b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"score_dict_kernel|finish_gen_kernel" -s 60 -c 4 -o gpurun_out/prof_search_r1b -f python scratch/sbench.py 30 2000 40 1 > gpurun_out/ncu_b.log 2>&1
python bench.py --config 5 --starts 4 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/bench_for_launchlist_r1b.json 2> gpurun_out/bench_for_launchlist_r1b.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/launches_r1b.csv python bench.py --config 5 --starts 4 --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_c.log 2>&1
wc -l gpurun_out/launches_r1b.csv; ls -la gpurun_out/*r1b*
