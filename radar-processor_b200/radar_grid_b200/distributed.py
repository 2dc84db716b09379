"""
Multi-GPU sharding of the gridding path — one process per GPU, torch.distributed for the plumbing.

The path shards two ways, neither of which the reference has (it only has a
multiprocessing.Pool over z-levels inside the table build, compute.py:203-222):

  volume batch   volumes of one scan strategy are independent given the neighbour table: every rank holds a
                 replica of the table (built locally on its GPU) and grids the volume ids `shard_volumes`
                 assigns to it.  No data-path collective; results stay on the rank or are gathered to rank 0.
  z-slab         voxel rows are z-major, so a z-slab is a contiguous CSR row range: every rank builds and
                 holds only its slab (`zslab_ranges`), grids it, and the partial COLMAX planes are combined by
                 ONE all-reduce(max) over NCCL/NVLink (`allreduce_nanmax`; NaN = "no data" is carried as -inf
                 because max(NaN, x) is unspecified in NCCL).  3-D grids need no exchange: slabs concatenate.
                 A CAPPI blends two adjacent levels that may sit in different slabs: every rank contributes
                 weight x level for the levels it owns and ONE all-reduce(sum) finishes the blend
                 (`cappi_zslab`; bit-identical to the unsharded CAPPI, see there).  A PPI follows the beam, so
                 its level pair differs per pixel: every rank gathers, from its slab's 3-D grid, the levels it
                 owns pixel by pixel and the same all-reduce(sum) finishes the blend (`ppi_zslab`).

The collective helpers work on CPU tensors with the gloo backend too, which is how the host-side logic is
tested without GPUs.
"""

from __future__ import annotations

from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np


def shard_volumes(n_volumes: int, world_size: int, rank: int) -> List[int]:
    """Round-robin assignment of volume ids to ranks (balanced to within one volume)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank outside [0, world_size)")
    return list(range(rank, n_volumes, world_size))


def zslab_ranges(nz: int, world_size: int, weights: Optional[Sequence[float]] = None) -> List[Tuple[int, int]]:
    """
    Contiguous [z0, z1) level ranges, one per rank.

    Without ``weights``: sizes differing by at most one level.  With ``weights`` (one number per level, normally the
    pair counts from ``DeviceGeometry.level_pairs``: table bytes and gridding time of a slab are proportional to its
    pairs, and the low levels of a radar grid hold several times the pairs of the high ones): the contiguous
    partition that minimises the heaviest slab (exact, by dynamic programming; nz and world_size are tiny).
    Ranks beyond the number of levels get empty slabs (z0 == z1).
    """
    if world_size < 1 or nz < 0:
        raise ValueError("bad arguments")
    if weights is None:
        base, extra = divmod(nz, world_size)
        out, z = [], 0
        for r in range(world_size):
            n = base + (1 if r < extra else 0)
            out.append((z, z + n))
            z += n
        return out
    w = [float(x) for x in weights]
    if len(w) != nz or any(x < 0 for x in w):
        raise ValueError("weights must hold one non-negative number per level")
    pre = [0.0]
    for x in w:
        pre.append(pre[-1] + x)
    k = min(world_size, max(nz, 1))
    inf = float("inf")
    # best[j][i]: minimal heaviest slab when the first i levels go to j slabs, every slab at least one level
    best = [[inf] * (nz + 1) for _ in range(k + 1)]
    cut = [[0] * (nz + 1) for _ in range(k + 1)]
    best[0][0] = 0.0
    for j in range(1, k + 1):
        for i in range(j, nz + 1):
            for m in range(j - 1, i):
                cand = max(best[j - 1][m], pre[i] - pre[m])
                if cand < best[j][i]:
                    best[j][i], cut[j][i] = cand, m
    bounds, i = [nz], nz
    for j in range(k, 0, -1):
        i = cut[j][i] if nz > 0 else 0
        bounds.append(i)
    bounds.reverse()
    out = [(bounds[r], bounds[r + 1]) for r in range(k)] if nz > 0 else []
    out += [(nz, nz)] * (world_size - len(out))
    return out


def _dist():
    import torch.distributed as dist
    return dist


def allreduce_nanmax(plane, group=None, minimum: bool = False, encoded: bool = False):
    """
    In-place combine of partial column-max (or -min) planes across ranks with NumPy's nanmax semantics:
    a pixel is NaN only if it is NaN on every rank.  `plane` is a torch tensor (CUDA with nccl, CPU with gloo).

    ``encoded``: the plane comes from a ``ColumnMax(partial=True)`` request, i.e. the fused epilogue already wrote
    -inf (+inf) where the slab has no data, and only the decode after the collective is left.  The sentinel doubles as
    a value: a genuine -inf (+inf) column extremum comes back as NaN -- reflectivity-like fields never hold one.
    """
    import torch
    dist = _dist()
    sentinel = float("inf") if minimum else float("-inf")
    if not encoded:
        plane.masked_fill_(torch.isnan(plane), sentinel)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(plane, op=dist.ReduceOp.MIN if minimum else dist.ReduceOp.MAX, group=group)
    plane.masked_fill_(plane == sentinel, float("nan"))
    return plane


def allreduce_nanmean(total, count, group=None):
    """Column mean across slabs: sums and valid counts are added, then divided (NaN where the count is 0)."""
    dist = _dist()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(count, op=dist.ReduceOp.SUM, group=group)
    return total / count


def gather_to_rank0(local: Dict[int, np.ndarray], group=None) -> Optional[Dict[int, np.ndarray]]:
    """Collect {volume id: result} dictionaries on rank 0 (host-side gather of small 2-D products)."""
    dist = _dist()
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return dict(local)
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    bucket = [None] * world if rank == 0 else None
    dist.gather_object(local, bucket, dst=0, group=group)
    if rank != 0:
        return None
    merged: Dict[int, np.ndarray] = {}
    for part in bucket:
        merged.update(part)
    return merged


def grid_volume_batch(volume_ids: Sequence[int], load_volume: Callable[[int], Sequence], grid_one: Callable) -> Dict[int, object]:
    """Grid this rank's share of a time series.  `grid_one(fields)` is normally a closure over
    ``engine.grid_fields`` and the rank's table replica; it is injectable so the sharding logic can be tested on CPU."""
    return {vid: grid_one(load_volume(vid)) for vid in volume_ids}


def colmax_zslab(grid_slab: Callable[[], "object"], group=None):
    """z-slab COLMAX: `grid_slab()` returns this rank's partial COLMAX planes (torch tensor, NaN = no data),
    e.g. from ``grid_fields(slab_geometry, ..., products=[ColumnMax()], want_grid=False)``; the result is the
    global COLMAX on every rank."""
    return allreduce_nanmax(grid_slab(), group=group)


def cappi_zslab_terms(request, grid_shape, grid_limits, z_range):
    """
    Split a CAPPI request (engine.CAPPI; reference products.py:317-415) over z-slabs.

    Returns None when the altitude is outside the grid (all-NaN plane, products.py:370-372), else
    ``(terms, dtype)``: ``terms`` = [(global level, weight)] for the levels inside ``z_range`` = [z0, z1) — none, one
    or two entries — and ``dtype`` the arithmetic type of the blend (float32, or float64 when the grid limits are
    NumPy scalars; see engine.CAPPI.resolve).  A level pick (nearest / exact level / clamped) is one term of weight 1.
    """
    from . import _native as N
    pr, _ = request.resolve(grid_shape, grid_limits, True)
    if pr is None:
        return None
    if pr.mode == N.RG_BLEND_PICK:
        parts, dtype = [(int(pr.z_lo), 1.0)], np.float32
    else:
        parts = [(int(pr.z_lo), float(pr.w_lo)), (int(pr.z_hi), float(pr.w_hi))]
        dtype = np.float32 if pr.mode == N.RG_BLEND_F32 else np.float64
    z0, z1 = int(z_range[0]), int(z_range[1])
    return [(z, w) for z, w in parts if z0 <= z < z1], dtype


def cappi_zslab_partial(request, grid_shape, grid_limits, z_range, level_planes: Callable[[List[int]], Sequence], like):
    """This rank's contribution to a z-slab CAPPI, before the all-reduce: ``sum(weight * plane)`` over the levels
    it owns, in the blend's arithmetic type, starting from -0.0 (the neutral element of IEEE addition, signed zeros
    included).  None when the altitude is outside the grid.  Arguments as in `cappi_zslab`."""
    import torch
    plan = cappi_zslab_terms(request, grid_shape, grid_limits, z_range)
    if plan is None:
        return None
    terms, dtype = plan
    tdt = torch.float32 if dtype == np.float32 else torch.float64
    acc = torch.full_like(like, -0.0, dtype=tdt)
    if terms:
        planes = level_planes([z for z, _ in terms])
        if len(planes) != len(terms):
            raise ValueError("level_planes returned a different number of planes than levels asked for")
        for (z, w), plane in zip(terms, planes):
            plane = plane.to(tdt)
            acc += plane if w == 1.0 else plane * w
    return acc


def cappi_zslab(request, grid_shape, grid_limits, z_range, level_planes: Callable[[List[int]], Sequence], like,
                group=None):
    """
    CAPPI of a grid that is split into z-slabs across ranks; the result is the global plane on every rank.

    ``level_planes(levels)`` returns this rank's planes (torch tensors shaped like ``like``, float32, NaN = no data)
    for the GLOBAL level indices asked for — normally a closure over
    ``grid_fields(slab_geometry, ..., products=[LevelPick(z) for z in levels], want_grid=False)``, so the levels
    come out of the same fused pass as the slab's COLMAX; it is not called on ranks that own neither level.
    ``like`` is a tensor giving shape and device of a plane (F, ny, nx).

    Every rank forms its partial blend (`cappi_zslab_partial`) and one all-reduce(sum) adds the at most two
    non-neutral contributions: ``w_lo*g[lo] + w_hi*g[hi]`` with one rounding per operation, exactly the reference's
    expression (products.py:411), NaN in either level giving NaN.  Weight-1 picks skip the multiplication.
    """
    import torch
    dist = _dist()
    acc = cappi_zslab_partial(request, grid_shape, grid_limits, z_range, level_planes, like)
    if acc is None:
        return torch.full_like(like, float("nan"), dtype=torch.float32)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=group)
    return acc.to(torch.float32)


def ppi_zslab_plan(request, grid_shape, grid_limits):
    """
    Per-pixel level pair and weights of a constant-elevation PPI (engine.PPI) — the host-side half of
    reference products.py:227-309, in the reference's own dtypes: float32 pixel coordinates and horizontal distance,
    float64 beam height (radar_altitude 0: heights are relative to the radar), float64 fractional level.

    linear : {"lo", "hi"} int64 (ny, nx) clipped to [0, nz-1], {"w_lo", "w_hi"} float64, "nan" bool (beam below /
             above the grid, products.py:306-309)
    nearest: {"lo"} = round(z_frac) clipped, "nan" = index outside [0, nz) (products.py:263-272)
    """
    from .products import compute_beam_height, compute_beam_height_flat
    if request.interpolation not in ("linear", "nearest"):
        raise ValueError(f"Unknown interpolation method: {request.interpolation}")
    nz, ny, nx = grid_shape
    z_min, z_max = grid_limits[0]
    yc = np.linspace(grid_limits[1][0], grid_limits[1][1], ny, dtype="float32")
    xc = np.linspace(grid_limits[2][0], grid_limits[2][1], nx, dtype="float32")
    yy, xx = np.meshgrid(yc, xc, indexing="ij")
    hdist = np.sqrt(xx ** 2 + yy ** 2)
    tz = (compute_beam_height(hdist, request.elevation_angle, 0.0, request.ke) if request.earth_curvature
          else compute_beam_height_flat(hdist, request.elevation_angle, 0.0))
    z_step = (z_max - z_min) / (nz - 1) if nz > 1 else 1.0
    z_frac = (tz - z_min) / z_step
    if request.interpolation == "nearest":
        zi = np.round(z_frac).astype(np.int64)
        return {"mode": "nearest", "lo": np.clip(zi, 0, nz - 1), "nan": ~((zi >= 0) & (zi < nz))}
    lo = np.floor(z_frac).astype(np.int64)
    w_hi = z_frac - lo
    return {"mode": "linear", "lo": np.clip(lo, 0, nz - 1), "hi": np.clip(lo + 1, 0, nz - 1),
            "w_lo": 1.0 - w_hi, "w_hi": w_hi, "nan": (tz < z_min) | (tz > z_max)}


def ppi_zslab_partial(plan, z_range, slab_grids):
    """This rank's contribution to a z-slab PPI: ``slab_grids`` is a torch tensor (F, z1-z0, ny, nx) float32 (the
    slab's 3-D grids, NaN = no data); per pixel the owned levels of the pair are gathered and weighted, levels of
    other slabs contribute -0.0 (neutral for IEEE addition), pixels the beam leaves the grid at are NaN on every
    rank.  float64 (F, ny, nx) for 'linear' — the reference returns float64 there — float32 for 'nearest'."""
    import torch
    z0, z1 = int(z_range[0]), int(z_range[1])
    dev = slab_grids.device
    F, nzs = slab_grids.shape[0], slab_grids.shape[1]
    if nzs != z1 - z0:
        raise ValueError("slab_grids does not have z1 - z0 levels")

    def pick(levels):
        lv = torch.from_numpy(levels).to(dev)
        own = (lv >= z0) & (lv < z1)
        idx = (lv - z0).clamp_(0, max(nzs - 1, 0))
        vals = torch.gather(slab_grids, 1, idx.expand(F, 1, *idx.shape)).squeeze(1) if nzs > 0 else \
            torch.zeros((F,) + tuple(lv.shape), device=dev)
        return own, vals

    nan = torch.from_numpy(plan["nan"]).to(dev)
    if plan["mode"] == "nearest":
        own, v = pick(plan["lo"])
        out = torch.where(own, v, torch.full_like(v, -0.0))
    else:
        own_lo, v_lo = pick(plan["lo"])
        own_hi, v_hi = pick(plan["hi"])
        w_lo, w_hi = torch.from_numpy(plan["w_lo"]).to(dev), torch.from_numpy(plan["w_hi"]).to(dev)
        neutral = torch.full(v_lo.shape, -0.0, dtype=torch.float64, device=dev)
        out = torch.where(own_lo, w_lo * v_lo.double(), neutral) + torch.where(own_hi, w_hi * v_hi.double(), neutral)
    out.masked_fill_(nan, float("nan"))
    return out


def ppi_zslab(request, grid_shape, grid_limits, z_range, slab_grids, group=None):
    """Constant-elevation PPI of a grid split into z-slabs: `ppi_zslab_partial` on every rank, then ONE
    all-reduce(sum) of the (F, ny, nx) planes — per pixel ``w_lo*g[lo] + w_hi*g[hi]`` with the reference's roundings
    (products.py:294-304), whichever ranks hold the two levels.  The global plane is returned on every rank."""
    dist = _dist()
    out = ppi_zslab_partial(ppi_zslab_plan(request, grid_shape, grid_limits), z_range, slab_grids)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


def _partial_requests(products):
    import dataclasses
    from .engine import CAPPI, PPI, ColumnMax, ColumnMean, ColumnMin, LevelPick
    reqs = []
    for p in products:
        if isinstance(p, ColumnMean) or not isinstance(p, (ColumnMax, ColumnMin, CAPPI, PPI, LevelPick)):
            raise ValueError(f"{type(p).__name__} has no z-slab form (use allreduce_nanmean on sums and counts)")
        reqs.append(dataclasses.replace(p, partial=True))
    return reqs


def _reduce_kind(p) -> str:
    from .engine import ColumnMax, ColumnMin
    return "min" if isinstance(p, ColumnMin) else "max" if isinstance(p, ColumnMax) else "sum"


def zslab_terms(slab_geom, fields, products, **grid_kwargs) -> List:
    """This slab's TERMS of the requested 2-D products (torch tensors, one (F, ny, nx) plane set per request) from one
    fused pass with ``partial=True`` requests and no 3-D grid: see `zslab_products`."""
    import torch
    from .engine import grid_fields
    res = grid_fields(slab_geom, fields, products=_partial_requests(products), want_grid=False, **grid_kwargs)
    return [pl if hasattr(pl, "is_cuda") else torch.from_numpy(pl) for pl in res["products"]]


def zslab_merge(acc: List, terms: List, products) -> List:
    """Merge another slab's terms into ``acc`` in place (what the all-reduce does between ranks), for a rank that walks
    several slabs one after the other."""
    import torch
    for p, a, t in zip(products, acc, terms):
        kind = _reduce_kind(p)
        if kind == "max":
            torch.maximum(a, t, out=a)
        elif kind == "min":
            torch.minimum(a, t, out=a)
        else:
            a.add_(t)
    return acc


def zslab_finish(terms: List, products, group=None) -> List:
    """all-reduce(MAX | MIN | SUM) of the terms -- planes of one reduction kind and dtype travel in ONE buffer -- then
    decode: the -inf / +inf "no data" sentinel of COLMAX / COLMIN becomes NaN, a float64 CAPPI blend is rounded to
    float32 (reference products.py:412).  In place where the dtype allows; the finished planes are returned."""
    import torch
    from .engine import CAPPI, ColumnMax, ColumnMin
    dist = _dist()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        ops = {"max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN, "sum": dist.ReduceOp.SUM}
        buckets: Dict[Tuple[str, object], List[int]] = {}
        for i, (p, pl) in enumerate(zip(products, terms)):
            buckets.setdefault((_reduce_kind(p), pl.dtype), []).append(i)
        for (kind, _), idxs in buckets.items():
            if len(idxs) == 1:
                dist.all_reduce(terms[idxs[0]], op=ops[kind], group=group)
                continue
            flat = torch.cat([terms[i].reshape(-1) for i in idxs])
            dist.all_reduce(flat, op=ops[kind], group=group)
            off = 0
            for i in idxs:
                n = terms[i].numel()
                terms[i].copy_(flat[off:off + n].view_as(terms[i]))
                off += n
    out = []
    for p, pl in zip(products, terms):
        if isinstance(p, ColumnMin):
            pl.masked_fill_(pl == float("inf"), float("nan"))
        elif isinstance(p, ColumnMax):
            pl.masked_fill_(pl == float("-inf"), float("nan"))
        elif isinstance(p, CAPPI) and pl.dtype == torch.float64:
            pl = pl.to(torch.float32)
        out.append(pl)
    return out


def zslab_products(slab_geom, fields, products, group=None, timings: Optional[dict] = None, **grid_kwargs) -> List:
    """
    2-D products of a grid that is split into z-slabs across ranks, from ONE fused pass per rank and ONE collective
    per reduction kind; the finished planes are returned on every rank, bit-identical to the unsharded products.

    Every request (``ColumnMax``, ``ColumnMin``, ``CAPPI``, ``PPI``, ``LevelPick``) is issued with ``partial=True``: the
    epilogue of ``grid_fields(slab_geom, ..., want_grid=False)`` writes this slab's TERM of the product -- running max
    with -inf for "no data here"; for the level blends ``w_lo*g[lo] + w_hi*g[hi]`` (reference products.py:294-304, 411)
    the products of the levels the slab owns and -0.0, the neutral element of IEEE addition, for the others -- so no
    slab 3-D grid is materialised and no torch arithmetic surrounds the collective: all-reduce(MAX | MIN | SUM), decode
    the sentinel, round a float64 CAPPI blend to float32.  ``fields`` are torch CUDA tensors (NCCL) or NumPy arrays (the
    planes then go through torch CPU tensors: gloo, the CPU tests).  ``timings`` (optional dict) receives the wall times
    ``apply_ms`` and ``allreduce_ms``.
    """
    import time
    import torch
    t0 = time.perf_counter()
    terms = zslab_terms(slab_geom, fields, products, **grid_kwargs)
    cuda = bool(terms) and terms[0].is_cuda
    if cuda:
        torch.cuda.synchronize()
    t1 = time.perf_counter()
    out = zslab_finish(terms, products, group=group)
    if cuda:
        torch.cuda.synchronize()
    if timings is not None:
        timings["apply_ms"] = (t1 - t0) * 1e3
        timings["allreduce_ms"] = (time.perf_counter() - t1) * 1e3
    return out


def parse_cpulist(text: str) -> List[int]:
    """'0-3,8,10-11' (the sysfs cpulist format) -> [0, 1, 2, 3, 8, 10, 11]."""
    cpus: List[int] = []
    for part in text.strip().split(","):
        if not part:
            continue
        a, _, b = part.partition("-")
        cpus.extend(range(int(a), int(b or a) + 1))
    return cpus


def bind_host_to_gpu(device_index: int, sysfs_root: str = "/sys/bus/pci/devices") -> Dict[str, object]:
    """
    Pin this process to the CPUs of the NUMA node its GPU hangs off, BEFORE pinned host buffers are allocated: with
    one process per GPU the end-to-end path moves ~300 MB per volume between pinned host memory and the device, and
    first-touch places those pages on the node of the allocating thread — the wrong socket for half the GPUs of an
    8-GPU box unless the process is bound.  Best effort: returns what it did ({"bound": False, "why": ...} when the
    topology cannot be read), never raises.
    """
    import os
    try:
        import torch
        pr = torch.cuda.get_device_properties(device_index)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(os.path.join(sysfs_root, bdf, "local_cpulist")) as fh:
            local = parse_cpulist(fh.read())
        try:
            with open(os.path.join(sysfs_root, bdf, "numa_node")) as fh:
                node = int(fh.read().strip())
        except OSError:
            node = -1
        allowed = sorted(set(local) & set(os.sched_getaffinity(0)))
        if not allowed:
            return {"bound": False, "why": "no local CPU in this process's affinity mask", "pci": bdf, "numa_node": node}
        if len(allowed) == len(os.sched_getaffinity(0)):
            return {"bound": False, "why": "single NUMA domain (every allowed CPU is local)", "pci": bdf, "numa_node": node}
        os.sched_setaffinity(0, allowed)
        return {"bound": True, "pci": bdf, "numa_node": node, "cpus": len(allowed)}
    except Exception as e:  # noqa: BLE001 — topology files differ between boxes; binding is an optimisation only
        return {"bound": False, "why": f"{type(e).__name__}: {e}"}
