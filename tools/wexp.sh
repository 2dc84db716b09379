#!/bin/bash
# warp-slice copy x group width x state-in-smem: parity tests on (variant 4, W=4), then timings
RG_GROUP_WIDTH=4 RG_APPLY_VARIANT=4 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_known_answers.py tests/test_gpu_fullsize.py -m gpu -q -p no:cacheprovider 2>&1 | tail -3
for L in "" _b; do for F in 5 4; do for V in 0 4; do for W in 4 8; do
RADAR_GRID_B200_LIB=radar-processor_b200/lib/libradargrid_b200$L.so RG_BENCH_FIELDS=$F RG_GROUP_WIDTH=$W RG_APPLY_VARIANT=$V timeout 300 python bench.py --workload cfg3 --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 > gpurun_out/w.json 2> gpurun_out/w.err
python -c "import json;d=json.load(open('gpurun_out/w.json'));print('cfg3 lib=$L F=$F variant=$V W=$W', 'apply %.3f ms frac %.3f'%(d['config']['apply_ms_per_step'],d['roofline']['frac']))" || tail -3 gpurun_out/w.err
done; done; done; done
