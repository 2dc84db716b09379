#!/bin/bash
# One full ncu capture of the apply kernel (after the same command ran clean without ncu).  The .ncu-rep embeds the whole
# 40 MB module, so only its raw-metric and per-instruction CSV pages are kept (tools/ncu_summary.py / ncu_lines.py read those).
#   usage: gpu_ncu_apply.sh <name> [library variant]
set -u
mkdir -p gpurun_out
NAME=${1:-apply}
if [ -n "${2:-}" ]; then export RADAR_GRID_B200_LIB=radar-processor_b200/lib/libradargrid_b200_$2.so; fi
CMD="python bench.py --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:apply_(columns|duo)" -s 3 -c 1 -f -o /tmp/$NAME $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
ncu -i /tmp/$NAME.ncu-rep --page raw --csv > gpurun_out/$NAME.raw.csv 2>/dev/null
ncu -i /tmp/$NAME.ncu-rep --page source --csv > gpurun_out/$NAME.source.csv 2>/dev/null
ls -la /tmp/$NAME.ncu-rep gpurun_out/$NAME.*.csv
