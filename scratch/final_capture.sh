timeout 400 python bench.py > gpurun_out/bench_default6.json 2> gpurun_out/bench_default6.err
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_ref6.json 2> gpurun_out/bench_ref6.err
timeout 150 python scratch/kbench.py 5 4 3 > gpurun_out/kb_plain_r1d.log 2>&1 && timeout 300 ncu --set full --clock-control none --import-source on -k regex:"estep_kernel|mstats_kernel" -c 4 -o gpurun_out/prof_stream_r1d -f python scratch/kbench.py 5 4 3 > gpurun_out/ncu_a.log 2>&1
python -c "
import json
d=json.load(open('gpurun_out/bench_ref6.json')); print(d['value'], d['ms_per_step'], d['cpu_baseline']['cores']); print(d['cpu_baseline']['sample'])
d=json.load(open('gpurun_out/bench_default6.json')); print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['breakdown_last_step_rank0']); print(d['cpu_baseline'])"
