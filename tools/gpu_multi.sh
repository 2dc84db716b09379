#!/bin/bash
# N ranks of one box (default 2): z-slab products over NCCL checked against the unsharded build, then the bench line (with its
# z-slab record), then the bare-copy ceiling with N concurrent processes.    usage: gpu_multi.sh [N]
set -u
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
F='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(json.dumps({k:d[k] for k in d if k not in ("per_slab_rank0",)}))'
echo "== z-slab COLMAX + CAPPI + PPI over NCCL, $N ranks, cfg3 in $N balanced slabs (checked against the unsharded build)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 examples/zslab_colmax.py --spec cfg3 --slabs $N --cappi 4750 --ppi 1.0 --check --repeat 10 2>gpurun_out/zslab_cfg3_n$N.err | grep '^{' | tee gpurun_out/zslab_cfg3_n$N.json | python -c "$F"; tail -2 gpurun_out/zslab_cfg3_n$N.err | cut -c1-300
if [ "$N" -ge 8 ]; then
echo "== BASELINE configs[4]: cfg5 (80 x 2001 x 2001), one balanced z-slab per rank, COLMAX + CAPPI 4100 m + PPI 1.0 deg over NCCL"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 examples/zslab_colmax.py --spec cfg5 --slabs $N --cappi 4100 --ppi 1.0 --repeat 10 2>gpurun_out/zslab_cfg5_n$N.err | grep '^{' | tee gpurun_out/zslab_cfg5_n$N.json | python -c "$F"; tail -2 gpurun_out/zslab_cfg5_n$N.err | cut -c1-300
fi
echo "== bench $N GPUs"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n$N.json'))
print({k: d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','scaling')}); print(d['roofline']['frac'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'products-only', d['e2e_products_only']['value'], d['e2e_products_only']['ms_per_step'], d['clocks']); print(d['zslab'])"; tail -3 gpurun_out/bench_n$N.err | cut -c1-300
if [ -z "${SKIP_PCIE:-}" ]; then echo "== pcie ceiling, up to $N concurrent processes"; timeout 300 python tools/pcie_ceiling.py --out gpurun_out/pcie_ceiling_n$N.json; fi
