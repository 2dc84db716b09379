#!/bin/bash
# Column-pair kernel against the column-group kernel per field count (cfg3 table, first k fields): where "auto" should switch.
#   usage: gpu_duo_fields.sh [library variant]
set -u
mkdir -p gpurun_out
if [ -n "${1:-}" ]; then export RADAR_GRID_B200_LIB=radar-processor_b200/lib/libradargrid_b200_$1.so; fi
for k in 1 2 3 4 5; do
  for duo in 0 2; do
    out=gpurun_out/duof_${k}_${duo}.json
    RG_BENCH_FIELDS=$k RADAR_GRID_B200_DUO=$duo timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 > $out 2> ${out%.json}.err
    python -c "import json;d=json.load(open('$out'));print('F=$k duo=$duo', 'step %.4f ms apply %.4f ms pack %.4f ms same=%s'%(d['ms_per_step'],d['config']['apply_ms_per_step'],d['config']['pack_ms_per_step'],d['config']['device_vs_host_path_identical']))" || tail -3 ${out%.json}.err
  done
done
