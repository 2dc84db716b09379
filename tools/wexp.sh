#!/bin/bash
# CTA launch order experiment: full gpu tests, then cfg3 timings with / without the heaviest-first order
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -3
for T in 1 0 1 0; do
RG_TILE_ORDER=$T timeout 300 python bench.py --workload cfg3 --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 > gpurun_out/w.json 2> gpurun_out/w.err
python -c "import json;d=json.load(open('gpurun_out/w.json'));print('cfg3 tile_order=$T', 'step %.3f apply %.3f ms frac %.3f'%(d['ms_per_step'],d['config']['apply_ms_per_step'],d['roofline']['frac']))" || tail -3 gpurun_out/w.err
done
