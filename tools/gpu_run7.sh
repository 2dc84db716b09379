#!/bin/bash
set -u
mkdir -p gpurun_out
L=radar-processor_b200/lib
echo "== pytest W (group, tex)"; RADAR_GRID_B200_LIB=$L/libradargrid_b200_W.so RG_APPLY_VARIANT_TEST=1 timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -3
echo "== pytest W (sell, tex)"; RADAR_GRID_B200_LIB=$L/libradargrid_b200_W.so timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -3
run() { # name lib apply_variant width
  out=gpurun_out/g_$1.json
  RADAR_GRID_B200_LIB=$2 RG_APPLY_VARIANT=$3 RG_GROUP_WIDTH=${4:-0} timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 > $out 2> ${out%.json}.err
  python -c "import json;d=json.load(open('$out'));print('$1', 'step %.3f ms apply %.3f ms pack %.3f ms frac %.3f'%(d['ms_per_step'],d['config']['apply_ms_per_step'],d['config']['pack_ms_per_step'],d['roofline']['frac']))" || tail -3 ${out%.json}.err
}
for v in V W X Y; do run group_$v $L/libradargrid_b200_$v.so 1; done
run group_W_w16 $L/libradargrid_b200_W.so 1 16
for v in V W; do run sell_$v $L/libradargrid_b200_$v.so 0; done
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1"
export RADAR_GRID_B200_LIB=$L/libradargrid_b200_W.so
RG_APPLY_VARIANT=1 timeout 300 $CMD > gpurun_out/plain_groupW.log 2>&1 && \
RG_APPLY_VARIANT=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:apply_columns -s 3 -c 1 -f -o gpurun_out/prof_groupW $CMD > gpurun_out/ncu_groupW.log 2>&1
echo "ncu group exit $?"
