#!/bin/bash
# One 8-GPU box: BASELINE.json configs[4] (cfg5 z-slabs, one per GPU, NCCL max-reduce + the CAPPI two-party sum), then
# the volume-batch bench at 8 and 4 ranks.
set -u
mkdir -p gpurun_out
nvidia-smi -L | wc -l
echo "== cfg5 z-slab COLMAX + CAPPI 4100 m, 8 ranks"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 examples/zslab_colmax.py --spec cfg5 --slabs 8 --cappi 4100 2>gpurun_out/zslab_cfg5_n8.err | grep '^{' > gpurun_out/zslab_cfg5_n8.json
python -c "import json; d=json.load(open('gpurun_out/zslab_cfg5_n8.json')); print({k:d[k] for k in d if k!='per_slab'}); print(d['per_slab'])"; tail -2 gpurun_out/zslab_cfg5_n8.err
for n in 8 4; do
echo "== bench $n GPUs"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2955$n bench.py --gpus $n --steps 100 --warmup 5 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err; echo "exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n$n.json'))
print({k: d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','scaling')}); print(d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['clocks'])"; tail -2 gpurun_out/bench_n$n.err
done
