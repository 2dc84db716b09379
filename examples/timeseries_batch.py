#!/usr/bin/env python
"""
Time-series batch (BASELINE.json configs[3]): 256 cfg1-shaped volumes of one scan strategy, COLMAX + CAPPI 4000 m
per volume, sharded across the GPUs of one box with no inter-GPU communication.

Every rank builds its own replica of the neighbour table (once), takes the volume ids `shard_volumes` gives it and
feeds them through a `VolumePipeline`: pinned host fields in, the two product planes out, the 3-D grid never written
to HBM (`want_grid=False`).  This is what the reference's examples/batch_processing.py does with a
ThreadPoolExecutor around its CPU path.

    python examples/timeseries_batch.py [--volumes 256] [--distinct 16]
    torchrun --nproc-per-node 8 examples/timeseries_batch.py

Synthetic volumes cost 0.45 s each to generate on the host, so every rank generates `--distinct` different ones
(seeds = its first volume ids) before the clock starts and cycles through them; every volume is copied to the
device in full each time.  Timed region: wall clock around the rank's whole share, bracketed by barriers, max over
ranks.  Prints one JSON line.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "radar-processor_b200")):
    sys.path.insert(0, p)

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spec", default="cfg1")
    ap.add_argument("--volumes", type=int, default=256)
    ap.add_argument("--distinct", type=int, default=16)
    ap.add_argument("--streams", type=int, default=3)
    ap.add_argument("--altitude", type=float, default=4000.0)
    ap.add_argument("--repeat", type=int, default=1, help="run the whole series this many times inside the timed region")
    args = ap.parse_args()

    import torch
    import radar_grid_b200 as rg
    from radar_grid_b200 import _native as N, distributed as D, synthetic as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    affinity = D.bind_host_to_gpu(local) if world > 1 else {"bound": False, "why": "single rank"}
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = N.Context(local)
    spec = S.SPECS[args.spec]
    nz, ny, nx = spec.grid_shape
    G, F = spec.n_gates, len(spec.fields)
    gates = S.gate_coordinates(spec)
    t0 = time.perf_counter()
    geom = rg.DeviceGeometry.build(*gates, spec.grid_shape, spec.grid_limits, min_radius=spec.min_radius,
                                   beam_factor=spec.beam_factor, weighting=spec.weighting, toa=spec.toa, ctx=ctx)
    t_build = time.perf_counter() - t0

    mine = D.shard_volumes(args.volumes, world, rank)
    n_distinct = max(1, min(args.distinct, len(mine)))
    host = []                                                 # pinned inputs: F fields of G gates, NaN = masked
    for vid in mine[:n_distinct]:
        fields = S.make_fields(spec, seed=vid, gates=gates)
        bufs = []
        for name in spec.fields:
            b = rg.pinned_empty((G,), np.float32)
            b[:] = np.ma.getdata(fields[name])
            b[np.ma.getmaskarray(fields[name])] = np.nan
            bufs.append(b)
        host.append(bufs)
    products = [rg.ColumnMax(), rg.CAPPI(args.altitude)]
    outs = [[rg.pinned_empty((F, ny, nx), np.float32) for _ in products] for _ in mine]
    jobs = [dict(fields=host[i % n_distinct], mask_invalid=True, want_grid=False, products=products,
                 out_products=outs[i]) for i in range(len(mine))]

    pipe = rg.VolumePipeline(geom, n_streams=args.streams)
    pipe.map(jobs[:2 * args.streams])                         # warm-up: staging buffers, slice copy of the table
    launches0 = pipe.kernel_launches()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(max(1, args.repeat)):
        pipe.map(jobs)
    torch.cuda.synchronize()
    elapsed = time.perf_counter() - t0
    launches = pipe.kernel_launches() - launches0
    if world > 1:
        t = torch.tensor([elapsed], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        elapsed = float(t.item())

    # the pipelined result of a volume is bit-identical to a synchronous single call on the same volume
    k = min(len(mine), n_distinct) - 1
    one = rg.grid_fields(geom, host[k], mask_invalid=True, want_grid=False, products=products, ctx=ctx)["products"]
    same = all(np.array_equal(a, b, equal_nan=True) for a, b in zip(one, outs[k]))
    cyc = k + n_distinct
    if cyc < len(mine):                                       # ... and so is its later repetition in the cycle
        same = same and all(np.array_equal(a, b, equal_nan=True) for a, b in zip(outs[k], outs[cyc]))

    if rank == 0:
        print(json.dumps({
            "workload": f"{args.volumes} x {spec.name}-shaped volumes ({F} field(s), {G} gates) -> COLMAX + CAPPI "
                        f"{args.altitude:g} m on {nz}x{ny}x{nx}, products only",
            "world": world, "volumes": args.volumes, "volumes_per_rank": len(mine), "distinct_volumes_per_rank": n_distinct,
            "streams": args.streams, "seconds": elapsed, "repeat": max(1, args.repeat),
            "volumes_per_s": args.volumes * max(1, args.repeat) / elapsed,
            "voxels_per_s": args.volumes * max(1, args.repeat) * F * nz * ny * nx / elapsed,
            "ms_per_volume_per_rank": 1e3 * elapsed / max(len(mine) * max(1, args.repeat), 1),
            "h2d_bytes_per_volume": 4 * F * G, "d2h_bytes_per_volume": 4 * F * ny * nx * len(products),
            "table_build_s_wall": round(t_build, 3), "pairs": geom.n_pairs, "gpu_launches_rank0": launches,
            "pipelined_equals_synchronous": bool(same), "host_affinity_rank0": affinity,
            "timing": "wall clock around the rank's share incl. H2D/D2H, barrier before, max over ranks"}))
    pipe.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
