#!/bin/bash
# Round-2 baseline: bare-copy ceiling of the box, bench of the committed kernels, one full ncu capture of the apply kernel.
set -u
mkdir -p gpurun_out
nvidia-smi -L; nproc; grep -E "MemTotal|Hugepagesize|HugePages_Total|AnonHugePages" /proc/meminfo; cat /sys/kernel/mm/transparent_hugepage/enabled
echo "== pcie ceiling (one copy per transfer)"; timeout 200 python tools/pcie_ceiling.py --out gpurun_out/pcie_ceiling.json
echo "== pcie ceiling (7 pieces per transfer)"; timeout 200 python tools/pcie_ceiling.py --pieces 7 --out gpurun_out/pcie_ceiling_7pieces.json
echo "== bench"; timeout 600 python bench.py --steps 100 --warmup 10 > gpurun_out/r2_base_bench.json 2> gpurun_out/r2_base_bench.err; echo "bench exit $?"; python -c "
import json; d=json.load(open('gpurun_out/r2_base_bench.json'))
print({k: d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}); print(d['roofline']); print(d['e2e']); print(d['clocks']); print(d['config']['pack_ms_per_step'], d['config']['apply_ms_per_step'])"; tail -3 gpurun_out/r2_base_bench.err
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:apply_columns -s 3 -c 1 -f -o gpurun_out/r2_base_apply $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"; ls -la gpurun_out/*.ncu-rep
