#!/bin/bash
# First GPU pass: smoke, parity tests, bench, launch list, one full ncu capture of the apply kernel.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -25 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
for w in 8 16 32; do
  echo "== bench group_width=$w"; RG_GROUP_WIDTH=$w timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_w$w.json 2> gpurun_out/bench_w$w.err; echo "exit $?"
  python -c "import json;d=json.load(open('gpurun_out/bench_w$w.json'));print(d['ms_per_step'],d['config']['apply_ms_per_step'],d['roofline']['frac'])"
done
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1"
echo "== ncu"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:apply_columns -s 3 -c 2 -f -o gpurun_out/prof_apply $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
ls -la gpurun_out
