#!/bin/bash
set -u
mkdir -p gpurun_out
L=radar-processor_b200/lib
for av in 0 3; do echo "== pytest default apply_variant=$av"; RG_APPLY_VARIANT_TEST=$av timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -3; done
run() { # name lib apply_variant width extra
  out=gpurun_out/m_$1.json
  RADAR_GRID_B200_LIB=$2 RG_APPLY_VARIANT=$3 RG_GROUP_WIDTH=${4:-0} timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 $5 > $out 2> ${out%.json}.err
  python -c "import json;d=json.load(open('$out'));print('$1', 'step %.3f ms apply %.3f ms pack %.3f ms frac %.3f'%(d['ms_per_step'],d['config']['apply_ms_per_step'],d['config']['pack_ms_per_step'],d['roofline']['frac']))" || tail -3 ${out%.json}.err
}
for v in s t; do run group_$v $L/libradargrid_b200_$v.so 0 0 ""; done
run cfg1_s $L/libradargrid_b200_s.so 0 0 "--workload cfg1"
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1"
export RADAR_GRID_B200_LIB=$L/libradargrid_b200_s.so
timeout 300 $CMD > gpurun_out/plain_group_s.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:apply_columns -s 3 -c 1 -f -o gpurun_out/prof_group_s $CMD > gpurun_out/ncu_group_s.log 2>&1
echo "ncu exit $?"
