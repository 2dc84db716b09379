#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== device-buffer tests"; timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k "device_buffers or zslab or pipeline" 2>&1 | tail -2
echo "== small, 3 slabs, checked against the unsharded build"; timeout 300 python examples/zslab_colmax.py --spec small --slabs 3 --check 2>&1 | tail -1 | cut -c1-300
echo "== cfg3, 4 slabs, checked"; timeout 300 python examples/zslab_colmax.py --spec cfg3 --slabs 4 --check 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print({k:d[k] for k in d if k!='per_slab'})"
echo "== cfg5, 8 slabs on one GPU in sequence"; timeout 1200 python examples/zslab_colmax.py --spec cfg5 --slabs 8 > gpurun_out/zslab_cfg5.json 2> gpurun_out/zslab_cfg5.err; echo "exit $?"; cut -c1-2500 gpurun_out/zslab_cfg5.json; tail -3 gpurun_out/zslab_cfg5.err
