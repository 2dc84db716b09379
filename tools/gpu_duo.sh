#!/bin/bash
# Column-pair kernel pass: its tests, then cfg3 / cfg1 bench lines with option duo = 0 / 1 / 2 (RADAR_GRID_B200_DUO) for every
# library variant named on the command line ("default" = the shipped one).   usage: gpu_duo.sh [-t] default h4u4 ...
set -u
mkdir -p gpurun_out
L=radar-processor_b200/lib
if [ "${1:-}" = "-t" ]; then shift; echo "== pytest duo"; timeout 900 python -m pytest tests/test_gpu_duo.py -q -p no:cacheprovider -x 2>&1 | tail -15; fi
run() { # name lib duo extra
  out=gpurun_out/duo_$1.json
  RADAR_GRID_B200_DUO=$3 RADAR_GRID_B200_LIB=$2 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 $4 > $out 2> ${out%.json}.err
  python -c "import json;d=json.load(open('$out'));print('$1', 'step %.4f ms apply %.4f ms pack %.4f ms frac %.3f same=%s'%(d['ms_per_step'],d['config']['apply_ms_per_step'],d['config']['pack_ms_per_step'],d['roofline']['frac'],d['config']['device_vs_host_path_identical']))" || tail -3 ${out%.json}.err
}
lib() { if [ "$1" = default ]; then echo $L/libradargrid_b200.so; else echo $L/libradargrid_b200_$1.so; fi; }
first=1
for v in "$@"; do
  if [ $first = 1 ]; then run cfg3_${v}_duo0 $(lib $v) 0 ""; fi
  run cfg3_${v}_duo1 $(lib $v) 1 ""
  first=0
done
for v in "$@"; do
  if [ "$v" = default ]; then run cfg1_${v}_duo0 $(lib $v) 0 "--workload cfg1"; fi
  run cfg1_${v}_duo2 $(lib $v) 2 "--workload cfg1"
done
