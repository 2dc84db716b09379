#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L
echo "== zslab test on the A/B kernel"; RG_APPLY_VARIANT_TEST=2 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider -k zslab 2>&1 | tail -2
echo "== bench 2 GPUs"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 5 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "exit $?"
python -c "
import json; d=json.load(open('gpurun_out/bench_n2.json'))
print({k: d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches','scaling')}); print(d['roofline']['frac'], d['e2e']['value'], d['clocks'])"; tail -5 gpurun_out/bench_n2.err
echo "== reference arm under torchrun"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/bench_ref_n2.json 2> gpurun_out/bench_ref_n2.err; echo "exit $?"; cut -c1-300 gpurun_out/bench_ref_n2.json
echo "== bench 1 GPU"
timeout 900 python bench.py --steps 50 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['roofline']['traffic'])"
