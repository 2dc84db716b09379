#!/bin/bash
# ncu capture of the one-field apply kernel at cfg1 (CSV pages only).
set -u
mkdir -p gpurun_out
NAME=${1:-r02_cfg1_apply}
CMD="python bench.py --workload cfg1 --steps 4 --warmup 3 --no-cpu-baseline --e2e-steps 1"
timeout 300 $CMD > gpurun_out/plain.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:apply_(columns|duo)" -s 3 -c 1 -f -o /tmp/$NAME $CMD > gpurun_out/ncu_full.log 2>&1
echo "ncu full exit $?"
ncu -i /tmp/$NAME.ncu-rep --page raw --csv > gpurun_out/$NAME.raw.csv 2>/dev/null
ncu -i /tmp/$NAME.ncu-rep --page source --csv > gpurun_out/$NAME.source.csv 2>/dev/null
ls -la gpurun_out/$NAME.*.csv
