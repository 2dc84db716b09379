#!/bin/bash
# BASELINE.json configs[3]: 256-volume time series, COLMAX + CAPPI per volume, on 1/2/4/8 ranks of one box.
set -u
mkdir -p gpurun_out
: > gpurun_out/timeseries_cfg4.jsonl
for n in 1 2 4 8; do
  if [ "$n" = 1 ]; then
    timeout 200 python examples/timeseries_batch.py 2> gpurun_out/timeseries_n$n.err | grep '^{' >> gpurun_out/timeseries_cfg4.jsonl
  else
    timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2957$n examples/timeseries_batch.py 2> gpurun_out/timeseries_n$n.err | grep '^{' >> gpurun_out/timeseries_cfg4.jsonl
  fi
  echo "n=$n exit $?"; tail -2 gpurun_out/timeseries_n$n.err | cut -c1-300
done
python -c "
import json
for l in open('gpurun_out/timeseries_cfg4.jsonl'):
    d=json.loads(l); print(d['world'], round(d['volumes_per_s'],1), '%.3g'%d['voxels_per_s'], round(d['ms_per_volume_per_rank'],3), d['pipelined_equals_synchronous'], d['gpu_launches_rank0'], d['table_build_s_wall'])"
