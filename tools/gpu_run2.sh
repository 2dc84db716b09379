#!/bin/bash
# A/B of the apply-kernel variants (record layout / accumulate / row order) on cfg3.
set -u
mkdir -p gpurun_out
L=radar-processor_b200/lib
for v in v0 v1 v2 v3; do
  echo "== pytest $v"; RADAR_GRID_B200_LIB=$L/libradargrid_b200_$v.so timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -4
done
for v in v0 v1 v2 v3; do for srt in 0 1; do for w in 8 16; do
  out=gpurun_out/ab_${v}_s${srt}_w${w}.json
  RADAR_GRID_B200_LIB=$L/libradargrid_b200_$v.so RG_SORT_ROWS=$srt RG_GROUP_WIDTH=$w timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 > $out 2> ${out%.json}.err
  python -c "import json;d=json.load(open('$out'));print('$v sort=$srt W=$w', 'step %.3f ms apply %.3f ms pack %.3f ms frac %.3f build %.1f ms'%(d['ms_per_step'],d['config']['apply_ms_per_step'],d['config']['pack_ms_per_step'],d['roofline']['frac'],d['config']['geometry_build_ms_device']))" || tail -3 ${out%.json}.err
done; done; done
# one full profile of the most promising candidates
for v in v2 v3; do
  CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1"
  RADAR_GRID_B200_LIB=$L/libradargrid_b200_$v.so timeout 300 $CMD > gpurun_out/plain_$v.log 2>&1 && \
  RADAR_GRID_B200_LIB=$L/libradargrid_b200_$v.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:apply_columns -s 3 -c 1 -f -o gpurun_out/prof_$v $CMD > gpurun_out/ncu_$v.log 2>&1
  echo "ncu $v exit $?"
done
