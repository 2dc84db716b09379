"""
GPU tests of the column-pair ("duo") apply kernel (radar-processor_b200/csrc/rg_duo.cu): the multi-field pass over the
merged rows of two adjacent columns.  Reference arithmetic: src/radar_grid/interpolate.py:59-104 (apply_geometry),
products: src/radar_grid/products.py:168-580.

Option "duo": 0 = never, 1 = auto (two or more fields, tables whose merge saves enough), 2 = whenever the table allows it
(rows sorted by gate id).  The tests force 2 so that the tiny fixtures take the kernel, and compare with the golden grids
of the real reference (fast-path bar: 1e-4 absolute / 1e-5 relative, identical NaN mask), with the column-group kernel
(duo = 0) and with the stand-alone products of the kernel's own grid (bit for bit).
"""
import warnings

import numpy as np
import pytest

import radar_grid_b200 as rg
from radar_grid_b200 import engine
from conftest import assert_same, golden_case
from oracle import radar_grid_oracle as O

pytestmark = pytest.mark.gpu
ATOL, RTOL = 1e-4, 1e-5


def assert_close_same_mask(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, what
    np.testing.assert_array_equal(np.isnan(a), np.isnan(b), err_msg=f"{what}: NaN mask differs")
    np.testing.assert_allclose(a, b, rtol=RTOL, atol=ATOL, equal_nan=True, err_msg=what)


def build(spec, gates, weighting="barnes2", alt=0, **kw):
    toa = spec.toa if alt == 0 else 4000.0
    return rg.DeviceGeometry.build(*gates, spec.grid_shape, spec.grid_limits, radar_altitude=float(alt),
                                   min_radius=spec.min_radius, beam_factor=spec.beam_factor, weighting=weighting,
                                   toa=toa, **kw)


@pytest.mark.parametrize("spec_name,weighting,alt", [("tiny", "barnes2", 0), ("tiny", "cressman", 0), ("tiny", "nearest", 0),
                                                     ("tiny", "barnes2", 350), ("small", "barnes2", 0)])
def test_duo_grids_match_the_reference_goldens(spec_name, weighting, alt):
    spec, radar, gates, fields, g = golden_case(spec_name, weighting, alt)
    dev = build(spec, gates, weighting, alt)
    names = list(fields)
    data = [np.ma.getdata(fields[n]) for n in names]
    masks = [np.ma.getmaskarray(fields[n]) for n in names]
    try:
        dev.ctx.set_option("duo", 2)
        for nf in range(1, len(names) + 1):
            res = rg.grid_fields(dev, data[:nf], masks=masks[:nf])
            for n, grid in zip(names[:nf], res["grids"]):
                assert grid.dtype == np.float32
                assert_close_same_mask(grid, g[f"grid_{n}"], f"duo F={nf} grid_{n}")
        assert dev.duo_slots > 0, "the column-pair copy was not built: the pass fell back to the column-group kernel"
    finally:
        dev.ctx.set_option("duo", 1)


def test_duo_fused_products_equal_the_products_of_its_grid_and_the_column_group_kernel():
    """Every product kind in the fused epilogue of the duo kernel: bit-identical to the stand-alone products of the grid
    the same kernel wrote (one ColumnState class), with and without the 3-D grid; the grids agree with the column-group
    kernel (duo = 0) within the fast-path tolerance (the association order of the row sums differs)."""
    spec, radar, gates, fields, g = golden_case("small")
    dev = build(spec, gates)
    names = list(fields)
    data = [np.ma.getdata(fields[n]) for n in names]
    masks = [np.ma.getmaskarray(fields[n]) for n in names]
    simple = [rg.ColumnMax(), rg.CAPPI(1234.5)]
    full = [rg.ColumnMax(), rg.ColumnMin(), rg.ColumnMean(), rg.CAPPI(1234.5), rg.CAPPI(3000.0, interpolation="nearest"),
            rg.PPI(1.3), rg.PPI(2.3, interpolation="nearest")]
    try:
        for reqs in (simple, full):
            dev.ctx.set_option("duo", 0)
            old = rg.grid_fields(dev, data, masks=masks, products=reqs)
            dev.ctx.set_option("duo", 2)
            new = rg.grid_fields(dev, data, masks=masks, products=reqs)
            only = rg.grid_fields(dev, data, masks=masks, products=reqs, want_grid=False)
            for f, n in enumerate(names):
                assert_close_same_mask(new["grids"][f], old["grids"][f], f"grid {n}")
                assert_close_same_mask(new["grids"][f], g[f"grid_{n}"], f"golden grid {n}")
            alone = engine.run_products(new["grids"], spec.grid_shape, spec.grid_limits, reqs, ctx=dev.ctx)
            for i, (a, b) in enumerate(zip(new["products"], only["products"])):
                assert_same(a, b, f"product {i}: with vs without the 3-D grid")
            for i, (a, b) in enumerate(zip(new["products"], alone)):
                assert_same(a, b, f"product {i}: fused vs stand-alone")
            # COLMAX of the fused pass is the nanmax of the grid it wrote
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", RuntimeWarning)
                for f in range(len(names)):
                    assert_same(new["products"][0][f], np.nanmax(new["grids"][f], axis=0), "fused COLMAX")
    finally:
        dev.ctx.set_option("duo", 1)


def test_duo_unmasked_nonfinite_values_propagate_only_into_rows_that_hold_the_gate():
    """An UNMASKED NaN / inf propagates into every voxel whose row holds the gate and nowhere else (interpolate.py:78-82).
    The merged entries carry a zero weight for the column that does not hold the gate: the pack kernel flags such volumes
    and the kernel takes its predicated path.  Against the oracle, and the same volume again without the specials (the
    flag must not stick)."""
    spec, radar, gates, fields, g = golden_case("tiny")
    dev = build(spec, gates)
    indptr, idx, w = dev.export_csr()
    names = list(fields)[:3]
    rng = np.random.default_rng(5)
    data, masks = [], []
    for n in names:
        v = np.ma.getdata(fields[n]).copy()
        m = np.ma.getmaskarray(fields[n]).copy()
        hit = rng.choice(v.size, size=40, replace=False)
        v[hit[:20]] = np.nan
        v[hit[20:30]] = np.inf
        v[hit[30:]] = -np.inf
        m[hit] = False
        data.append(v)
        masks.append(m)
    want = [O.apply_geometry(indptr.astype(np.int64), idx, w, spec.grid_shape, np.ma.MaskedArray(v, mask=m)) for v, m in zip(data, masks)]
    try:
        dev.ctx.set_option("duo", 2)
        res = rg.grid_fields(dev, data, masks=masks, mask_invalid=False)
        for f, n in enumerate(names):
            got = res["grids"][f]
            np.testing.assert_array_equal(np.isnan(got), np.isnan(want[f]), err_msg=f"{n}: NaN mask")
            np.testing.assert_array_equal(np.isinf(got), np.isinf(want[f]), err_msg=f"{n}: inf mask")
            fin = np.isfinite(want[f])
            np.testing.assert_allclose(got[fin], want[f][fin], rtol=RTOL, atol=ATOL)
        assert np.isnan(want[0]).sum() > np.isnan(g[f"grid_{names[0]}"]).sum()      # the specials did reach some voxels
        clean = rg.grid_fields(dev, [np.ma.getdata(fields[n]) for n in names], masks=[np.ma.getmaskarray(fields[n]) for n in names])
        for f, n in enumerate(names):
            assert_close_same_mask(clean["grids"][f], g[f"grid_{n}"], f"clean volume after a flagged one: {n}")
    finally:
        dev.ctx.set_option("duo", 1)


def test_duo_heavy_rows_odd_shapes_and_every_field_count():
    """A hand-made table (rows sorted by gate id) with rows of 600, 5 000 and 9 000 pairs next to ordinary and empty rows
    on a grid whose ny is odd and whose nx is not a multiple of 8: every field count against the oracle."""
    rng = np.random.default_rng(43)
    shape, limits = (3, 3, 9), ((0.0, 2000.0), (-500.0, 500.0), (-2000.0, 2000.0))
    n_gates = 12000
    lens = rng.integers(0, 40, size=int(np.prod(shape)))
    lens[[4, 5, 13]] = (5000, 600, 9000)
    lens[22] = 513
    lens[[0, 30, 50]] = 0
    indptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    rows = []
    base = np.sort(rng.choice(n_gates, size=60, replace=False))         # neighbouring rows share gates, as real ones do
    for n in lens:
        if n <= 40:
            rows.append(np.sort(rng.choice(base, size=n, replace=False)))
        else:
            rows.append(np.sort(rng.choice(n_gates, size=n, replace=False)))
    idx = np.concatenate(rows).astype(np.int32)
    w = rng.uniform(0.02, 1.0, size=idx.shape[0]).astype(np.float32)
    dev = rg.DeviceGeometry.from_csr(indptr, idx, w, shape, limits, n_gates)
    fields = []
    for f in range(8):
        v = rng.normal(10.0 * f, 5.0, size=n_gates).astype(np.float32)
        v[rng.random(n_gates) < (0.5 if f == 0 else 0.05)] = np.nan
        fields.append(np.ma.masked_invalid(v))
    want = [O.apply_geometry(indptr, idx, w, shape, f) for f in fields]
    try:
        dev.ctx.set_option("duo", 2)
        for nf in range(1, 9):
            res = rg.grid_fields(dev, [np.ma.getdata(f) for f in fields[:nf]], masks=[np.ma.getmaskarray(f) for f in fields[:nf]],
                                 products=[rg.ColumnMax(), rg.ColumnMean()])
            for f in range(nf):
                assert_close_same_mask(res["grids"][f], want[f], f"F={nf} field {f}")
                with warnings.catch_warnings():
                    warnings.simplefilter("ignore", RuntimeWarning)
                    assert_same(res["products"][0][f], np.nanmax(res["grids"][f], axis=0), "fused COLMAX")
        assert dev.duo_slots > 0
    finally:
        dev.ctx.set_option("duo", 1)


def test_duo_is_not_taken_for_tables_in_another_row_order():
    """The reference's own table keeps the KD-tree order of every row: no merge by gate id is possible, the pass must fall
    back to the column-group kernel and still be right."""
    spec, radar, gates, fields, g = golden_case("tiny")
    dev = rg.DeviceGeometry.from_csr(g["indptr"], g["gate_indices"], g["weights"], spec.grid_shape, spec.grid_limits,
                                     n_gates=len(gates[0]))
    names = list(fields)
    try:
        dev.ctx.set_option("duo", 2)
        res = rg.grid_fields(dev, [np.ma.getdata(fields[n]) for n in names], masks=[np.ma.getmaskarray(fields[n]) for n in names])
        for n, grid in zip(names, res["grids"]):
            assert_close_same_mask(grid, g[f"grid_{n}"], f"grid_{n}")
        assert dev.duo_slots < 0
    finally:
        dev.ctx.set_option("duo", 1)


def test_duo_zslab_rows_and_partial_level_ranges():
    """A z-slab build feeds the duo kernel the same rows as the full build (bit-identical grids), and a products-only
    request that walks a sub-range of levels gives the same planes as the full walk."""
    spec, radar, gates, fields, g = golden_case("small")
    full = build(spec, gates)
    nz = spec.grid_shape[0]
    z0, z1 = nz // 3, nz - 1
    slab = build(spec, gates, z_range=(z0, z1))
    names = list(fields)[:4]
    data = [np.ma.getdata(fields[n]) for n in names]
    masks = [np.ma.getmaskarray(fields[n]) for n in names]
    try:
        for d in (full, slab):
            d.ctx.set_option("duo", 2)
        a = rg.grid_fields(full, data, masks=masks, products=[rg.ColumnMax(z_min_idx=z0, z_max_idx=z1 - 1)])
        b = rg.grid_fields(slab, data, masks=masks)
        only = rg.grid_fields(full, data, masks=masks, products=[rg.ColumnMax(z_min_idx=z0, z_max_idx=z1 - 1)], want_grid=False)
        for f in range(len(names)):
            assert_same(a["grids"][f][z0:z1], b["grids"][f], f"slab rows, field {f}")
            assert_same(a["products"][0][f], only["products"][0][f], "COLMAX over a level sub-range")
    finally:
        for d in (full, slab):
            d.ctx.set_option("duo", 1)
