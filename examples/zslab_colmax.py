#!/usr/bin/env python
"""
Large-domain products by z-slab sharding (BASELINE.json configs[4]: 80 x 2001 x 2001 grid at 250 m).

The whole neighbour table of this grid has ~2.8e10 pairs (226 GB) — more than one GPU holds and more than the
reference's int32 indptr can index — but voxel rows are z-major, so a z-slab is a contiguous row range.  The levels
are cut into slabs of (nearly) equal PAIR count (a level census, ``DeviceGeometry.level_pairs``: the lowest levels
hold several times the pairs of the highest), every rank builds and holds only its slab(s), and ONE fused pass per
slab writes the slab's TERMS of every requested product (COLMAX with -inf for "no data", CAPPI / PPI as the sums over
the levels the slab owns: ``partial=True`` requests) — no 3-D grid, no torch arithmetic around the collective.  One
all-reduce(MAX) and one all-reduce(SUM) over NCCL finish the planes (``distributed.zslab_finish``).

    python examples/zslab_colmax.py [--spec cfg5] [--slabs 11]                      # one GPU, slabs in sequence
    torchrun --nproc-per-node 8 examples/zslab_colmax.py --slabs 8                  # one slab per GPU

Prints one JSON line: per-slab build / apply times of rank 0, per-rank totals, collective and wall times, checksums.
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

MAX_PAIRS = 0xFFFFFFFF - 1          # a slab table is indexed with 32 bits


def balanced_slabs(census, n_slabs, D):
    """Contiguous slabs of equal pair count; more slabs than asked for if one would exceed the 32-bit pair index."""
    n = n_slabs
    while True:
        ranges = [r for r in D.zslab_ranges(len(census), n, weights=census) if r[1] > r[0]]
        if all(census[a:b].sum() * 1.02 < MAX_PAIRS or b - a == 1 for a, b in ranges):
            return ranges
        n += 1


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--spec", default="cfg5")
    ap.add_argument("--slabs", type=int, default=8)
    ap.add_argument("--even", action="store_true", help="slabs of equal level count instead of equal pair count")
    ap.add_argument("--cappi", type=float, default=None, metavar="ALT_M", help="also a CAPPI at this altitude")
    ap.add_argument("--ppi", type=float, default=None, metavar="ELEV_DEG", help="also a PPI at this elevation")
    ap.add_argument("--repeat", type=int, default=3, help="timed repetitions of the gridding pass over the resident tables")
    ap.add_argument("--check", action="store_true", help="also build the whole grid in one piece and compare (small specs)")
    args = ap.parse_args()

    import torch
    import radar_grid_b200 as rg
    from radar_grid_b200 import _native as N, distributed as D, synthetic as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = N.Context(local)
    if args.check:
        # bit-identity with the unsharded products needs the same summation order everywhere: the group width is
        # otherwise chosen per table from its mean row length
        ctx.set_option("group_width", 8)
    spec = S.SPECS[args.spec]
    nz, ny, nx = spec.grid_shape
    gates = S.gate_coordinates(spec)
    fields = S.make_fields(spec, seed=0, gates=gates)
    name = spec.fields[0]
    data = np.ma.getdata(fields[name]).copy()
    data[np.ma.getmaskarray(fields[name])] = np.nan
    dgates = [torch.from_numpy(g).cuda() for g in gates]
    dfield = torch.from_numpy(data).cuda()
    kw = dict(min_radius=spec.min_radius, beam_factor=spec.beam_factor, toa=spec.toa)

    t0 = time.perf_counter()
    census = rg.DeviceGeometry.level_pairs(*dgates, spec.grid_shape, spec.grid_limits, column_stride=4, ctx=ctx, **kw)
    census_s = time.perf_counter() - t0
    slabs = D.zslab_ranges(nz, args.slabs) if args.even else balanced_slabs(census, args.slabs, D)
    if args.even:                                             # halve slabs that would overflow the 32-bit pair index
        fixed = []
        for a, b in slabs:
            while census[a:b].sum() * 1.02 >= MAX_PAIRS and b - a > 1:
                m = (a + b) // 2
                fixed.append((a, m))
                a = m
            fixed.append((a, b))
        slabs = fixed
    mine = [slabs[i] for i in D.shard_volumes(len(slabs), world, rank)]
    products = [rg.ColumnMax()] + ([] if args.cappi is None else [rg.CAPPI(args.cappi)]) + ([] if args.ppi is None else [rg.PPI(args.ppi)])

    # one slab per rank stays resident (the multi-GPU case); a rank walking several slabs builds, grids and frees them in turn
    resident = len(mine) <= 1
    log, acc, geoms = [], None, []
    t_wall0 = time.perf_counter()
    for z0, z1 in mine:
        t0 = time.perf_counter()
        geom = rg.DeviceGeometry.build(*dgates, spec.grid_shape, spec.grid_limits, weighting=spec.weighting, z_range=(z0, z1), ctx=ctx, **kw)
        t_build = time.perf_counter() - t0
        rg.grid_fields(geom, [dfield], mask_invalid=True, want_grid=False, products=products[:1], ctx=ctx)   # first touch: lazily built copies
        ctx.synchronize()
        t0 = time.perf_counter()
        terms = D.zslab_terms(geom, [dfield], products, mask_invalid=True, ctx=ctx)
        ctx.synchronize()
        t_apply = time.perf_counter() - t0
        acc = terms if acc is None else D.zslab_merge(acc, terms, products)
        info = geom.info
        log.append({"z": [z0, z1], "pairs": info["n_pairs"], "build_ms_device": round(info["build_ms"], 1), "build_s_wall": round(t_build, 3),
                    "apply_ms": round(t_apply * 1e3, 3), "table_GB": round(info["n_pairs"] * 8 / 1e9, 2),
                    "pair_stream_TBs": round(info["n_pairs"] * 8 / max(t_apply, 1e-9) / 1e12, 2)})
        if resident:
            geoms.append(geom)
        else:
            geom.close()
    if acc is None:                                           # a rank without a slab contributes the neutral terms
        empty = rg.DeviceGeometry.build(*dgates, spec.grid_shape, spec.grid_limits, weighting=spec.weighting, z_range=(nz, nz), ctx=ctx, **kw)
        acc = D.zslab_terms(empty, [dfield], products, mask_invalid=True, ctx=ctx)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    planes = D.zslab_finish(acc, products)
    torch.cuda.synchronize()
    t_coll = time.perf_counter() - t0
    t_wall = time.perf_counter() - t_wall0

    # steady state (tables resident, as for every later volume of the scan strategy): gridding pass + collectives
    steady = None
    if resident and args.repeat > 0:
        walls, colls = [], []
        for _ in range(args.repeat):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            tm = {}
            t0 = time.perf_counter()
            if geoms:
                D.zslab_products(geoms[0], [dfield], products, mask_invalid=True, ctx=ctx, timings=tm)
            else:
                D.zslab_finish(D.zslab_terms(empty, [dfield], products, mask_invalid=True, ctx=ctx), products)
            torch.cuda.synchronize()
            walls.append((time.perf_counter() - t0) * 1e3)
            colls.append(tm.get("allreduce_ms", 0.0))
        steady = {"wall_ms": float(np.median(walls)), "allreduce_ms": float(np.median(colls))}

    colmax = planes[0]
    out = {"spec": spec.name, "grid": list(spec.grid_shape), "slabs": len(slabs), "slab_levels": [list(s) for s in slabs],
           "balanced_by_pairs": not args.even, "census_s": round(census_s, 3), "world": world,
           "valid_pixels": int((~torch.isnan(colmax)).sum().item()),
           "colmax_sum": float(torch.nan_to_num(colmax).double().sum().item()), "per_slab_rank0": log,
           "first_volume_wall_s_rank0": round(t_wall, 3), "finish_collectives_ms_rank0": round(t_coll * 1e3, 3)}
    per_rank = torch.tensor([sum(s["build_s_wall"] for s in log), sum(s["apply_ms"] for s in log), float(sum(s["pairs"] for s in log)),
                             steady["wall_ms"] if steady else 0.0, steady["allreduce_ms"] if steady else 0.0], dtype=torch.float64, device="cuda")
    if world > 1:
        allr = [torch.zeros_like(per_rank) for _ in range(world)]
        dist.all_gather(allr, per_rank)
        allr = torch.stack(allr).cpu().numpy()
    else:
        allr = per_rank.cpu().numpy()[None]
    out.update(total_pairs=int(allr[:, 2].sum()), build_s_per_rank=[round(float(v), 3) for v in allr[:, 0]],
               apply_ms_per_rank=[round(float(v), 3) for v in allr[:, 1]], pairs_per_rank=[int(v) for v in allr[:, 2]])
    if steady:
        out["steady_state"] = {"wall_ms_max_over_ranks": round(float(allr[:, 3].max()), 3),
                               "allreduce_ms_max_over_ranks": round(float(allr[:, 4].max()), 3),
                               "sequential_one_gpu_apply_ms": round(float(allr[:, 1].sum()), 3),
                               "speedup_vs_sequential_apply": round(float(allr[:, 1].sum() / max(allr[:, 3].max(), 1e-9)), 2)}
    k = 1
    if args.cappi is not None:
        out.update(cappi_altitude=args.cappi, cappi_valid_pixels=int((~torch.isnan(planes[k])).sum().item()),
                   cappi_sum=float(torch.nan_to_num(planes[k]).double().sum().item()))
        k += 1
    if args.ppi is not None:
        out.update(ppi_elevation=args.ppi, ppi_dtype=str(planes[k].dtype), ppi_valid_pixels=int((~torch.isnan(planes[k])).sum().item()),
                   ppi_sum=float(torch.nan_to_num(planes[k]).double().sum().item()))
    if args.check:
        for g in geoms:
            g.close()
        whole = rg.DeviceGeometry.build(*dgates, spec.grid_shape, spec.grid_limits, weighting=spec.weighting, ctx=ctx, **kw)
        refs = rg.grid_fields(whole, [dfield], mask_invalid=True, want_grid=False, products=products, ctx=ctx)["products"]
        same = lambda a, b: bool(a.dtype == b.dtype and torch.equal(torch.nan_to_num(a, nan=-1e30), torch.nan_to_num(b, nan=-1e30)))
        out["identical_to_unsharded"] = [same(a, b) for a, b in zip(planes, refs)]
        out["census_exact"] = bool(np.array_equal(
            rg.DeviceGeometry.level_pairs(*dgates, spec.grid_shape, spec.grid_limits, column_stride=1, ctx=ctx, **kw),
            np.diff(whole.export_csr()[0].astype(np.int64)[::ny * nx])))
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
