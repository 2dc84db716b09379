#!/bin/bash
# A/B of library variants on cfg1, single-field cfg3 and cfg3: preload of the next level's pairs (pre3), PDL between the
# heavy-row kernel and the column kernel (default vs nopdl).
set -u
mkdir -p gpurun_out
L=radar-processor_b200/lib
python -c "
import sys; sys.path.insert(0,'.')
from oracle import build_ref; print('oracle/_ref available on this box:', build_ref.available())"
echo "== parity files on the default library (PDL)"; timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_known_answers.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
run() { out=gpurun_out/ab_$1.json
  RADAR_GRID_B200_LIB=$2 timeout 300 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --e2e-steps 2 $3 > $out 2> ${out%.json}.err
  python -c "import json;d=json.load(open('$out'));print('$1', 'step %.4f ms apply %.4f ms pack %.4f ms frac %.3f same=%s'%(d['ms_per_step'],d['config']['apply_ms_per_step'],d['config']['pack_ms_per_step'],d['roofline']['frac'],d['config']['device_vs_host_path_identical']))" || tail -3 ${out%.json}.err; }
for r in 1 2; do
run cfg3_default $L/libradargrid_b200.so ""
run cfg3_nopdl $L/libradargrid_b200_nopdl.so ""
run cfg1_default $L/libradargrid_b200.so "--workload cfg1"
run cfg1_nopdl $L/libradargrid_b200_nopdl.so "--workload cfg1"
run cfg1_pre3 $L/libradargrid_b200_pre3.so "--workload cfg1"
done
RG_BENCH_FIELDS=1 run cfg3f1_default $L/libradargrid_b200.so ""
RG_BENCH_FIELDS=1 run cfg3f1_pre3 $L/libradargrid_b200_pre3.so ""
echo "== parity files on pre3"; RADAR_GRID_B200_LIB=$L/libradargrid_b200_pre3.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -2
