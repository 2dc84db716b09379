#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== small, 3 slabs, checked against the unsharded build"; timeout 300 python examples/zslab_colmax.py --spec small --slabs 3 --check 2>&1 | tail -2
echo "== cfg3, 4 slabs, checked"; timeout 300 python examples/zslab_colmax.py --spec cfg3 --slabs 4 --check 2>&1 | tail -2 | cut -c1-900
echo "== cfg5, 8 slabs on one GPU in sequence"; timeout 900 python examples/zslab_colmax.py --spec cfg5 --slabs 8 > gpurun_out/zslab_cfg5.json 2> gpurun_out/zslab_cfg5.err; echo "exit $?"; cut -c1-1500 gpurun_out/zslab_cfg5.json; tail -3 gpurun_out/zslab_cfg5.err
