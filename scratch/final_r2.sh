#!/bin/bash
# Final round-2 verification + profile refresh (one B200 under gpurun).  Every ncu run follows a plain run of the same command.
mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_final_r2.log 2>&1; tail -2 gpurun_out/pytest_final_r2.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/smoke_final_r2.log 2>&1; tail -1 gpurun_out/smoke_final_r2.log
timeout 300 python bench.py > gpurun_out/bench_final_r2.json 2> gpurun_out/bench_final_r2.err; tail -c 300 gpurun_out/bench_final_r2.json
timeout 100 python scratch/sbench.py 30 2000 40 2 > gpurun_out/sb_plain_r2g.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"search_persistent_kernel" -c 1 -o gpurun_out/prof_search_r2g -f python scratch/sbench.py 30 2000 40 1 > gpurun_out/ncu_search_r2g.log 2>&1
CMD="python bench.py --config 5 --starts 6 --steps 1 --warmup 0 --no-cpu --no-e2e"
timeout 200 $CMD > gpurun_out/bench_for_launchlist_r2g.json 2> gpurun_out/bench_for_launchlist_r2g.err && \
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_r2g.csv $CMD > gpurun_out/ncu_launches_r2g.log 2>&1
wc -l gpurun_out/launches_r2g.csv; ls -la gpurun_out/prof_search_r2g.ncu-rep; tail -2 gpurun_out/sb_plain_r2g.log
