#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest"; timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -6
for av in 1 0; do
  out=gpurun_out/c_av${av}.json
  RG_APPLY_VARIANT=$av timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu-baseline --e2e-steps 2 > $out 2> ${out%.json}.err
  python -c "import json;d=json.load(open('$out'));print('apply_variant=$av', 'step %.3f ms apply %.3f ms pack %.3f ms frac %.3f build %.1f ms'%(d['ms_per_step'],d['config']['apply_ms_per_step'],d['config']['pack_ms_per_step'],d['roofline']['frac'],d['config']['geometry_build_ms_device']))" || tail -3 ${out%.json}.err
done
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1"
timeout 300 $CMD > gpurun_out/plain_sell.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:apply_sell -s 3 -c 1 -f -o gpurun_out/prof_sell $CMD > gpurun_out/ncu_sell.log 2>&1
echo "ncu exit $?"
