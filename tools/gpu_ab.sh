#!/bin/bash
# A/B of library variants built with radar-processor_b200/build.py --variant X ...   usage: gpu_ab.sh [-t "pytest args"] a b c ...
# ("default" = the shipped library).  Prints step / apply / pack ms and the roofline fraction for cfg3 and cfg1.
set -u
mkdir -p gpurun_out
L=radar-processor_b200/lib
PYT="tests -m gpu"
if [ "${1:-}" = "-t" ]; then PYT="$2"; shift 2; fi
# -q: everything but the three-minute cfg5 brute-force test
if [ "${1:-}" = "-q" ]; then PYT="tests -m gpu --deselect tests/test_gpu_configs.py::test_cfg5_slab_rows_against_bruteforce_and_colmax_against_the_oracle"; shift; fi
if [ -n "$PYT" ]; then echo "== pytest $PYT"; timeout 1500 python -m pytest $PYT -q -p no:cacheprovider -x 2>&1 | tail -15; fi
run() { # name lib extra
  out=gpurun_out/ab_$1.json
  RADAR_GRID_B200_LIB=$2 timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 $3 > $out 2> ${out%.json}.err
  python -c "import json;d=json.load(open('$out'));print('$1', 'step %.4f ms apply %.4f ms pack %.4f ms frac %.3f same=%s'%(d['ms_per_step'],d['config']['apply_ms_per_step'],d['config']['pack_ms_per_step'],d['roofline']['frac'],d['config']['device_vs_host_path_identical']))" || tail -3 ${out%.json}.err
}
lib() { if [ "$1" = default ]; then echo $L/libradargrid_b200.so; else echo $L/libradargrid_b200_$1.so; fi; }
for v in "$@"; do run cfg3_$v $(lib $v) ""; done
for v in "$@"; do run cfg1_$v $(lib $v) "--workload cfg1"; done
