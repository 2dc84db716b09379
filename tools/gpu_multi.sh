#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L
echo "== z-slab COLMAX with NCCL all-reduce, 2 ranks, cfg3 (checked against the unsharded build)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 examples/zslab_colmax.py --spec cfg3 --slabs 4 --check 2>&1 | grep '^{' | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in d if k!='per_slab'})"
echo "== z-slab COLMAX, 2 ranks, cfg5"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 examples/zslab_colmax.py --spec cfg5 --slabs 16 2>&1 | grep '^{' > gpurun_out/zslab_cfg5_n2.json; python -c "import json; d=json.load(open('gpurun_out/zslab_cfg5_n2.json')); print({k:d[k] for k in d if k!='per_slab'}, len(d['per_slab']))"
echo "== bench 2 GPUs"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json'))
print({k: d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','scaling')}); print(d['roofline']['frac'], d['e2e']['value'], d['e2e']['ms_per_step'], d['clocks'])"; tail -3 gpurun_out/bench_n2.err
