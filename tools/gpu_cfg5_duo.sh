F='import sys,json
for l in sys.stdin:
    if l.startswith("{"):
        d=json.loads(l); print(json.dumps({k:d[k] for k in d if k not in ("per_slab_rank0",)})); [print("   ", s) for s in d["per_slab_rank0"]]'
for duo in 0 2; do echo "== cfg5 balanced duo=$duo"; RADAR_GRID_B200_DUO=$duo timeout 900 python examples/zslab_colmax.py --spec cfg5 --slabs 8 2>gpurun_out/zslab_cfg5_duo$duo.err | tee gpurun_out/zslab_cfg5_duo${duo}_1gpu.json | python -c "$F"; tail -2 gpurun_out/zslab_cfg5_duo$duo.err; done
