"""
Host-side logic of the multi-GPU paths on CPU: world_size-2 gloo process groups.  The per-rank compute is
the oracle here (tests may use it); on GPUs it is grid_fields on the rank's table replica / z-slab.
"""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import golden_case, load_golden
from radar_grid_b200 import distributed as D


def test_shard_volumes_partition():
    for n, w in ((256, 8), (10, 4), (3, 8), (0, 2)):
        shards = [D.shard_volumes(n, w, r) for r in range(w)]
        flat = sorted(v for s in shards for v in s)
        assert flat == list(range(n))
        assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1
    with pytest.raises(ValueError):
        D.shard_volumes(4, 2, 2)


def test_zslab_ranges_cover_contiguously():
    for nz, w in ((80, 8), (40, 3), (5, 8), (20, 1)):
        r = D.zslab_ranges(nz, w)
        assert r[0][0] == 0 and r[-1][1] == nz and len(r) == w
        assert all(a[1] == b[0] for a, b in zip(r, r[1:]))
        sizes = [b - a for a, b in r]
        assert max(sizes) - min(sizes) <= 1


def test_zslab_ranges_balanced_by_weights():
    """With per-level weights the partition minimises the heaviest slab (checked against brute force on small cases),
    stays contiguous, and hands empty slabs to ranks beyond the number of levels."""
    import itertools
    rng = np.random.default_rng(3)
    for nz, w in ((7, 3), (9, 4), (6, 6), (5, 2)):
        weights = rng.integers(0, 50, size=nz).astype(float)
        r = D.zslab_ranges(nz, w, weights=weights)
        assert r[0][0] == 0 and r[-1][1] == nz and len(r) == w and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        got = max(weights[a:b].sum() for a, b in r)
        best = min(max(weights[a:b].sum() for a, b in zip((0,) + c, c + (nz,)))
                   for c in itertools.combinations(range(1, nz), w - 1)) if w > 1 else weights.sum()
        assert got == best, (nz, w, got, best)
    r = D.zslab_ranges(3, 5, weights=[4.0, 1.0, 1.0])
    assert r == [(0, 1), (1, 2), (2, 3), (3, 3), (3, 3)]
    # a radar-like profile: the low levels hold most of the pairs
    prof = [100, 90, 70, 50, 30, 20, 10, 5, 2, 1]
    r = D.zslab_ranges(10, 4, weights=prof)
    loads = [sum(prof[a:b]) for a, b in r]
    even = [sum(prof[a:b]) for a, b in D.zslab_ranges(10, 4)]
    assert max(loads) == 118 < max(even) == 260
    with pytest.raises(ValueError):
        D.zslab_ranges(4, 2, weights=[1.0, 2.0])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import sys
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from oracle import radar_grid_oracle as O
        spec, radar, gates, fields, g = golden_case("tiny")
        nz, ny, nx = spec.grid_shape
        # --- z-slab COLMAX: partial nanmax of this rank's levels, one all-reduce(max)
        z0, z1 = D.zslab_ranges(nz, world)[rank]
        lo, hi = g["indptr"][z0 * ny * nx], g["indptr"][z1 * ny * nx]
        slab = O.apply_geometry(g["indptr"][z0 * ny * nx:z1 * ny * nx + 1] - lo, g["gate_indices"][lo:hi],
                                g["weights"][lo:hi], (z1 - z0, ny, nx), fields["DBZH"])
        assert np.array_equal(slab, g["grid_DBZH"][z0:z1], equal_nan=True)
        partial = torch.from_numpy(O.column_reduce("max", slab).copy())
        full = D.colmax_zslab(lambda: partial).numpy()
        assert np.array_equal(full, g["colmax"], equal_nan=True), "z-slab COLMAX != reference COLMAX"
        pmin = torch.from_numpy(O.column_reduce("min", slab).copy())
        assert np.array_equal(D.allreduce_nanmax(pmin, minimum=True).numpy(), g["colmin"], equal_nan=True)
        # --- z-slab CAPPI: levels in different slabs, one all-reduce(sum); bit-identical to the unsharded CAPPI
        from radar_grid_b200 import CAPPI
        grid = g["grid_DBZH"]
        like = torch.empty((1, ny, nx))
        z_top = spec.grid_limits[0][1]
        step = z_top / (nz - 1)
        boundary = D.zslab_ranges(nz, world)[0][1]                      # first level of rank 1
        lims64 = tuple(tuple(np.float64(v) for v in ax) for ax in spec.grid_limits)      # float64 blend
        cases = [(CAPPI((boundary - 0.6) * step), spec.grid_limits),    # lo on rank 0, hi on rank 1
                 (CAPPI((boundary - 0.6) * step), lims64),
                 (CAPPI(0.3 * step), spec.grid_limits),                 # both levels on rank 0
                 (CAPPI((nz - 1.5) * step), spec.grid_limits),          # both on the last rank
                 (CAPPI(boundary * step), spec.grid_limits),            # exact level
                 (CAPPI((boundary - 0.4) * step, "nearest"), spec.grid_limits),
                 (CAPPI(z_top + 1.0), spec.grid_limits)]                # outside: all NaN
        for req, lims in cases:
            asked = []

            def level_planes(levels):
                asked.extend(levels)
                assert all(z0 <= z < z1 for z in levels), "asked for a level outside the slab"
                return [torch.from_numpy(grid[z][None].copy()) for z in levels]

            got = D.cappi_zslab(req, spec.grid_shape, lims, (z0, z1), level_planes, like).numpy()[0]
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                want = O.cappi(grid, spec.grid_shape, lims, req.altitude, req.interpolation)
            assert got.dtype == np.float32 and np.array_equal(got, np.asarray(want, dtype=np.float32), equal_nan=True), \
                f"z-slab CAPPI {req} != unsharded CAPPI"
            assert np.array_equal(np.signbit(got), np.signbit(np.asarray(want, dtype=np.float32)))
        # --- zslab_finish: the terms a fused partial=True pass would write (emulated from the reference grid), batched
        # all-reduce(MAX) + all-reduce(SUM) over gloo, decode; equals the reference products
        from radar_grid_b200 import ColumnMax as CMax, LevelPick
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            cm = np.nanmax(grid[z0:z1], axis=0) if z1 > z0 else np.full((ny, nx), np.nan, np.float32)
        t_max = torch.from_numpy(np.where(np.isnan(cm), -np.inf, cm).astype(np.float32)[None].copy())
        lvl = boundary                                                   # owned by rank 1 only
        t_pick = torch.from_numpy((grid[lvl] if z0 <= lvl < z1 else np.full((ny, nx), -0.0, np.float32))[None].copy())
        lvl2 = 0
        t_pick2 = torch.from_numpy((grid[lvl2] if z0 <= lvl2 < z1 else np.full((ny, nx), -0.0, np.float32))[None].copy())
        reqs = [CMax(), LevelPick(lvl), LevelPick(lvl2)]
        fin = D.zslab_finish([t_max, t_pick, t_pick2], reqs)
        assert np.array_equal(fin[0].numpy()[0], g["colmax"], equal_nan=True)
        assert np.array_equal(fin[1].numpy()[0], grid[lvl], equal_nan=True) and np.array_equal(fin[2].numpy()[0], grid[lvl2], equal_nan=True)
        # --- z-slab PPI: per-pixel level pairs, gathered from the slab's grid, one all-reduce(sum)
        from radar_grid_b200 import PPI
        slab_t = torch.from_numpy(grid[z0:z1][None].copy())
        for el in (0.5, 6.0, 20.0):
            for interp in ("linear", "nearest"):
                for curved in (True, False):
                    req = PPI(el, interp, curved)
                    got = D.ppi_zslab(req, spec.grid_shape, spec.grid_limits, (z0, z1), slab_t).numpy()[0]
                    want = O.ppi(grid, spec.grid_shape, spec.grid_limits, el, interp, curved)
                    assert got.dtype == want.dtype and np.array_equal(got, want, equal_nan=True), f"z-slab PPI {req}"
                    assert np.array_equal(np.signbit(got), np.signbit(want))
        # --- volume batch: each rank grids its share, rank 0 gathers
        ids = D.shard_volumes(5, world, rank)
        local = D.grid_volume_batch(ids, lambda vid: vid, lambda v: np.full((2, 2), float(v)))
        merged = D.gather_to_rank0(local)
        if rank == 0:
            assert sorted(merged) == list(range(5)) and all(merged[v][0, 0] == v for v in merged)
        else:
            assert merged is None
        open(os.path.join(out_dir, f"ok{rank}"), "w").write("ok")
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_allreduce_nanmax_single_process_semantics():
    p = torch.tensor([[1.0, float("nan")], [float("nan"), -2.0]])
    out = D.allreduce_nanmax(p.clone())
    assert torch.equal(torch.isnan(out), torch.isnan(p)) and out[0, 0] == 1.0 and out[1, 1] == -2.0


def test_cpulist_parsing_and_best_effort_binding(tmp_path):
    assert D.parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert D.parse_cpulist("5") == [5] and D.parse_cpulist("") == []
    before = os.sched_getaffinity(0)
    r = D.bind_host_to_gpu(0, sysfs_root=str(tmp_path))        # no GPU / no topology here: reports, never raises
    assert r["bound"] is False and "why" in r and os.sched_getaffinity(0) == before
