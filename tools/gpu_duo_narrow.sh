#!/bin/bash
# One- and two-field passes of the column-pair kernel: library variants against each other on cfg1, the cfg3 table (first k
# fields) and cfg5 slab by slab.   usage: gpu_duo_narrow.sh default n8 ...
set -u
mkdir -p gpurun_out
L=radar-processor_b200/lib
lib() { if [ "$1" = default ]; then echo $L/libradargrid_b200.so; else echo $L/libradargrid_b200_$1.so; fi; }
for v in "$@"; do
  export RADAR_GRID_B200_LIB=$(lib $v)
  out=gpurun_out/narrow_cfg1_$v.json
  timeout 300 python bench.py --workload cfg1 --steps 200 --warmup 5 --no-cpu-baseline --e2e-steps 2 > $out 2> ${out%.json}.err
  python -c "import json;d=json.load(open('$out'));print('$v cfg1', 'step %.4f ms apply %.4f ms frac %.3f'%(d['ms_per_step'],d['config']['apply_ms_per_step'],d['roofline']['frac']))" || tail -3 ${out%.json}.err
  for k in 1 2; do
    out=gpurun_out/narrow_cfg3_${k}_$v.json
    RG_BENCH_FIELDS=$k timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 > $out 2> ${out%.json}.err
    python -c "import json;d=json.load(open('$out'));print('$v cfg3 F=$k', 'step %.4f ms apply %.4f ms'%(d['ms_per_step'],d['config']['apply_ms_per_step']))" || tail -3 ${out%.json}.err
  done
  timeout 600 python examples/zslab_colmax.py --spec cfg5 --slabs 8 2>gpurun_out/narrow_cfg5_$v.err | tee gpurun_out/narrow_cfg5_$v.json | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$v cfg5 apply_ms', d['apply_ms_per_rank'], [s['apply_ms'] for s in d['per_slab_rank0']])"
done
