#!/bin/bash
# 2-rank NCCL run of the z-slab COLMAX + CAPPI example (cfg3, checked against the unsharded build), then 1 GPU for comparison.
set -u
mkdir -p gpurun_out
nvidia-smi -L
F='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(json.dumps({k:d[k] for k in d if k!="per_slab"}))'
echo "== 1 rank, cfg3, 4 slabs, CAPPI 4750 m (levels 9|10 straddle slabs 0|1)"
timeout 300 python examples/zslab_colmax.py --spec cfg3 --slabs 4 --cappi 4750 --check 2>&1 | python -c "$F" | tee gpurun_out/zslab_cappi_cfg3_n1.json
echo "== 2 ranks (NCCL), same"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 examples/zslab_colmax.py --spec cfg3 --slabs 4 --cappi 4750 --check 2>&1 | python -c "$F" | tee gpurun_out/zslab_cappi_cfg3_n2.json
echo "== 2 ranks, 2 slabs, CAPPI 9900 m (levels 19|20: one on each rank)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29532 examples/zslab_colmax.py --spec cfg3 --slabs 2 --cappi 9900 --check 2>&1 | python -c "$F" | tee gpurun_out/zslab_cappi_cfg3_n2_2slabs.json
