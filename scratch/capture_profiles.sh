#!/bin/bash
# Profile capture for profiles/ (run under gpurun on one B200).  Every ncu run follows a plain run of the same command
# and carries its own timeout.  The launch list runs with MNEMCU_CHAINS=1: ncu serialises kernels, and the two search
# chains wait on each other's events (a full-application capture with two chains stalled for the whole time limit).
mkdir -p gpurun_out
timeout 150 python scratch/kbench.py 5 4 3 > gpurun_out/kb_plain_r1c.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"estep_kernel|mstats_kernel" -c 4 -o gpurun_out/prof_stream_r1c -f python scratch/kbench.py 5 4 3 > gpurun_out/ncu_a.log 2>&1
export MNEMCU_CHAINS=1
CMD="python bench.py --config 5 --starts 2 --steps 1 --warmup 0 --no-cpu --no-e2e"
timeout 200 $CMD > gpurun_out/bench_for_launchlist_r1c.json 2> gpurun_out/bench_for_launchlist_r1c.err && \
timeout 700 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_r1c.csv $CMD > gpurun_out/ncu_c.log 2>&1
wc -l gpurun_out/launches_r1c.csv; ls -la gpurun_out/*r1c*
