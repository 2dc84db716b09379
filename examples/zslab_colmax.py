#!/usr/bin/env python
"""
Large-domain COLMAX by z-slab sharding (BASELINE.json configs[4]: 80 x 2001 x 2001 grid at 250 m).

The whole neighbour table of this grid has ~2.8e10 pairs (226 GB) — more than one GPU holds and more than the
reference's int32 indptr can index — but voxel rows are z-major, so a z-slab is a contiguous row range: every
rank builds and holds only its slab(s), grids them products-only (no 3-D grid is ever written), and the partial
COLMAX planes are combined with ONE all-reduce(max) (NaN = "no data" travels as -inf).

    python examples/zslab_colmax.py [--spec cfg5] [--slabs 8]                       # one GPU, slabs in sequence
    torchrun --nproc-per-node 8 examples/zslab_colmax.py --slabs 8                  # one slab per GPU + NCCL max

Prints one JSON line with build/apply times per slab and the COLMAX checksum.
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
    ap.add_argument("--spec", default="cfg5")
    ap.add_argument("--slabs", type=int, default=8)
    ap.add_argument("--cappi", type=float, default=None, metavar="ALT_M",
                    help="also a CAPPI at this altitude: level picks out of the same pass, one all-reduce(sum)")
    ap.add_argument("--ppi", type=float, default=None, metavar="ELEV_DEG",
                    help="also a PPI at this elevation: partial blends from each slab's 3-D grid, one all-reduce(sum)")
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
    spec = S.SPECS[args.spec]
    nz, ny, nx = spec.grid_shape
    gates = S.gate_coordinates(spec)
    fields = S.make_fields(spec, seed=0, gates=gates)
    name = spec.fields[0]
    data = np.ma.getdata(fields[name]).copy()
    data[np.ma.getmaskarray(fields[name])] = np.nan
    dgates = [torch.from_numpy(g).cuda() for g in gates]
    dfield = torch.from_numpy(data).cuda()

    slabs = D.zslab_ranges(nz, args.slabs)
    mine = [slabs[i] for i in D.shard_volumes(len(slabs), world, rank)]
    partial = torch.full((1, ny, nx), float("nan"), device="cuda")
    cappi_req = None if args.cappi is None else rg.CAPPI(args.cappi)
    cappi_acc, cappi_owned = None, []
    ppi_req = None if args.ppi is None else rg.PPI(args.ppi)
    ppi_plan = None if ppi_req is None else D.ppi_zslab_plan(ppi_req, spec.grid_shape, spec.grid_limits)
    ppi_acc = None
    log = []
    todo = list(mine)
    while todo:
        z0, z1 = todo.pop(0)
        t0 = time.perf_counter()
        try:
            geom = rg.DeviceGeometry.build(*dgates, spec.grid_shape, spec.grid_limits, min_radius=spec.min_radius,
                                           beam_factor=spec.beam_factor, weighting=spec.weighting, toa=spec.toa,
                                           z_range=(z0, z1), ctx=ctx)
        except NotImplementedError:
            # a slab table is indexed with 32 bits: more than 2^32-1 pairs -> halve the slab (the lowest levels are the densest)
            if z1 - z0 < 2:
                raise
            zm = (z0 + z1) // 2
            todo[:0] = [(z0, zm), (zm, z1)]
            continue
        t_build = time.perf_counter() - t0
        t0 = time.perf_counter()
        levels = []
        if cappi_req is not None:
            plan = D.cappi_zslab_terms(cappi_req, spec.grid_shape, spec.grid_limits, (z0, z1))
            levels = [] if plan is None else [z for z, _ in plan[0]]
        res = rg.grid_fields(geom, [dfield], mask_invalid=True, want_grid=ppi_req is not None,
                             products=[rg.ColumnMax()] + [rg.LevelPick(z) for z in levels], ctx=ctx)
        ctx.synchronize()
        if ppi_req is not None:                               # the beam's level pair differs per pixel: from the slab grid
            part = D.ppi_zslab_partial(ppi_plan, (z0, z1), torch.stack(res["grids"]))
            ppi_acc = part if ppi_acc is None else ppi_acc + part
        if levels:
            part = D.cappi_zslab_partial(cappi_req, spec.grid_shape, spec.grid_limits, (z0, z1),
                                         lambda lv: res["products"][1:], partial)
            cappi_acc = part if cappi_acc is None else cappi_acc + part
            cappi_owned += levels
        t_apply = time.perf_counter() - t0
        plane = res["products"][0]
        partial = torch.fmax(partial, plane)                 # fmax ignores NaN, as np.nanmax does
        info = geom.info
        log.append({"z": [z0, z1], "pairs": info["n_pairs"], "build_ms_device": round(info["build_ms"], 1),
                    "build_s_wall": round(t_build, 3), "apply_ms": round(t_apply * 1e3, 2),
                    "table_GB": round(info["device_bytes"] / 1e9, 2)})
        geom.close()
    colmax = D.allreduce_nanmax(partial)                      # one NCCL all-reduce(max) across the ranks
    cappi = None
    if cappi_req is not None:
        plan = D.cappi_zslab_terms(cappi_req, spec.grid_shape, spec.grid_limits, (0, nz))
        if plan is None:
            cappi = torch.full_like(partial, float("nan"))
        else:
            if cappi_acc is None:                             # this rank owns neither level: the neutral element
                cappi_acc = torch.full_like(partial, -0.0, dtype=torch.float32 if plan[1] == np.float32 else torch.float64)
            if world > 1:
                dist.all_reduce(cappi_acc, op=dist.ReduceOp.SUM)      # the two-party sum of SURVEY 8e
            cappi = cappi_acc.to(torch.float32)
    ppi = None
    if ppi_req is not None:
        if ppi_acc is None:
            ppi_acc = D.ppi_zslab_partial(ppi_plan, (0, 0), torch.empty((1, 0, ny, nx), device="cuda"))
        if world > 1:
            dist.all_reduce(ppi_acc, op=dist.ReduceOp.SUM)
        ppi = ppi_acc
    torch.cuda.synchronize()

    out = {"spec": spec.name, "grid": list(spec.grid_shape), "slabs": len(log) if world == 1 else len(slabs), "world": world,
           "total_pairs": sum(s["pairs"] for s in log), "valid_pixels": int((~torch.isnan(colmax)).sum().item()),
           "colmax_sum": float(torch.nan_to_num(colmax).double().sum().item()), "per_slab": log}
    if cappi is not None:
        out.update(cappi_altitude=args.cappi, cappi_levels_owned_here=cappi_owned,
                   cappi_valid_pixels=int((~torch.isnan(cappi)).sum().item()),
                   cappi_sum=float(torch.nan_to_num(cappi).double().sum().item()))
    if ppi is not None:
        out.update(ppi_elevation=args.ppi, ppi_dtype=str(ppi.dtype), ppi_valid_pixels=int((~torch.isnan(ppi)).sum().item()),
                   ppi_sum=float(torch.nan_to_num(ppi).double().sum().item()))
    if args.check:
        whole = rg.DeviceGeometry.build(*dgates, spec.grid_shape, spec.grid_limits, min_radius=spec.min_radius,
                                        beam_factor=spec.beam_factor, weighting=spec.weighting, toa=spec.toa, ctx=ctx)
        prods = [rg.ColumnMax()] + ([] if cappi_req is None else [cappi_req]) + ([] if ppi_req is None else [ppi_req])
        refs = rg.grid_fields(whole, [dfield], mask_invalid=True, want_grid=False, products=prods, ctx=ctx)["products"]
        ref = refs[0]
        if cappi is not None:
            out["cappi_identical_to_unsharded"] = bool(torch.equal(torch.nan_to_num(refs[1], nan=-1e30),
                                                                   torch.nan_to_num(cappi, nan=-1e30)))
        if ppi is not None:
            out["ppi_identical_to_unsharded"] = bool(refs[-1].dtype == ppi.dtype and torch.equal(
                torch.nan_to_num(refs[-1], nan=-1e30), torch.nan_to_num(ppi, nan=-1e30)))
        out["identical_to_unsharded"] = bool(torch.equal(torch.nan_to_num(ref, nan=-1e30), torch.nan_to_num(colmax, nan=-1e30)))
    if rank == 0:
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
