#!/bin/bash
# Standard GPU pass: smoke, full gpu test-suite (+ the parity files on every apply variant), bench, reference arm, launch list
# and one full ncu capture of the apply kernel (kept as CSV pages: the .ncu-rep embeds the 40 MB module).
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4
for av in 1 2 3 4; do echo "== pytest parity, apply_variant=$av"; RG_APPLY_VARIANT_TEST=$av timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_known_answers.py -m gpu -q -p no:cacheprovider 2>&1 | tail -2; done
echo "== bench"; timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/bench.json'))
print({k: d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}); print(d['roofline']); print(d['e2e']); print(d['e2e_products_only']); print(d['clocks']); print(d.get('cpu_baseline')); print(d['config']['pack_ms_per_step'], d['config']['apply_ms_per_step'])"; tail -3 gpurun_out/bench.err
echo "== bench cfg1"; timeout 300 python bench.py --workload cfg1 --steps 200 --no-cpu-baseline --e2e-steps 4 > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_cfg1.json')); print('cfg1', d['ms_per_step'], d['config']['apply_ms_per_step'], d['roofline']['frac'])"
echo "== bench reference arm"; timeout 1200 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "exit $?"; cut -c1-1500 gpurun_out/bench_ref.json
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "ncu list exit $?"
bash tools/gpu_ncu_apply.sh ${1:-final_apply}
