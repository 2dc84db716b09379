#!/bin/bash
set -u
mkdir -p gpurun_out
L=radar-processor_b200/lib
echo "== pytest default"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -3
run() { # name lib apply_variant width extra
  out=gpurun_out/n_$1.json
  RADAR_GRID_B200_LIB=$2 RG_APPLY_VARIANT=$3 RG_GROUP_WIDTH=${4:-0} timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 $5 > $out 2> ${out%.json}.err
  python -c "import json;d=json.load(open('$out'));print('$1', 'step %.3f ms apply %.3f ms pack %.3f ms frac %.3f'%(d['ms_per_step'],d['config']['apply_ms_per_step'],d['config']['pack_ms_per_step'],d['roofline']['frac']))" || tail -3 ${out%.json}.err
}
for v in u v w x; do run group_$v $L/libradargrid_b200_$v.so 0 0 ""; done
for v in u v; do run cfg1_$v $L/libradargrid_b200_$v.so 0 0 "--workload cfg1"; done
