#!/bin/bash
# A/B of library variants built with radar-processor_b200/build.py --variant X ...   usage: gpu_ab.sh a b c ...
set -u
mkdir -p gpurun_out
L=radar-processor_b200/lib
echo "== pytest default"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -3
run() { # name lib extra
  out=gpurun_out/ab_$1.json
  RADAR_GRID_B200_LIB=$2 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 $3 > $out 2> ${out%.json}.err
  python -c "import json;d=json.load(open('$out'));print('$1', 'step %.3f ms apply %.3f ms pack %.3f ms frac %.3f'%(d['ms_per_step'],d['config']['apply_ms_per_step'],d['config']['pack_ms_per_step'],d['roofline']['frac']))" || tail -3 ${out%.json}.err
}
for v in "$@"; do run cfg3_$v $L/libradargrid_b200_$v.so ""; done
for v in "$@"; do run cfg1_$v $L/libradargrid_b200_$v.so "--workload cfg1"; done
