#!/bin/bash
# One GPU: the z-slab example checked against the unsharded build (cfg3), then BASELINE configs[4] (cfg5) slab by slab,
# with slabs of equal level count (round-1 layout, for comparison) and balanced by pair count.
set -u
mkdir -p gpurun_out
F='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(json.dumps({k:d[k] for k in d if k not in ("per_slab_rank0",)})); [print("   ", s) for s in d["per_slab_rank0"]]'
echo "== adapter + new tests"; timeout 600 python -m pytest tests/test_adapter.py tests/test_gpu_configs.py -m gpu -q -p no:cacheprovider -x --deselect tests/test_gpu_configs.py::test_cfg5_slab_rows_against_bruteforce_and_colmax_against_the_oracle 2>&1 | tail -3
echo "== cfg3, 4 balanced slabs, COLMAX + CAPPI + PPI, checked"; timeout 300 python examples/zslab_colmax.py --spec cfg3 --slabs 4 --cappi 4750 --ppi 1.0 --check 2>gpurun_out/zslab_cfg3.err | tee gpurun_out/zslab_cfg3_1gpu.json | python -c "$F"; tail -2 gpurun_out/zslab_cfg3.err
echo "== cfg5, 8 slabs of equal level count (split where > 2^32 pairs)"; timeout 900 python examples/zslab_colmax.py --spec cfg5 --slabs 8 --even 2>gpurun_out/zslab_cfg5_even.err | tee gpurun_out/zslab_cfg5_even_1gpu.json | python -c "$F"; tail -2 gpurun_out/zslab_cfg5_even.err
echo "== cfg5, 8 slabs balanced by pair count"; timeout 900 python examples/zslab_colmax.py --spec cfg5 --slabs 8 2>gpurun_out/zslab_cfg5_bal.err | tee gpurun_out/zslab_cfg5_balanced_1gpu.json | python -c "$F"; tail -2 gpurun_out/zslab_cfg5_bal.err
