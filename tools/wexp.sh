#!/bin/bash
# warp-slice copy experiment: parity tests on variant 4, then cfg3 / cfg1 timings, default vs variant 4
RG_APPLY_VARIANT=4 RG_APPLY_VARIANT_TEST=4 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_known_answers.py tests/test_gpu_fullsize.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
for wl in cfg3 cfg1; do for V in 0 4; do
RG_APPLY_VARIANT=$V timeout 300 python bench.py --workload $wl --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 > gpurun_out/w.json 2> gpurun_out/w.err
python -c "import json;d=json.load(open('gpurun_out/w.json'));print('$wl variant=$V', 'apply %.3f ms frac %.3f'%(d['config']['apply_ms_per_step'],d['roofline']['frac']))" || tail -3 gpurun_out/w.err
done; done
