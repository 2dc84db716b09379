"""
GPU parity tests proper: the CUDA path (through the C ABI) against the committed reference fixtures
(tests/golden, produced by the real reference) and against the CPU oracle on the same seeded inputs.

Bars (north_star): neighbour sets bit-exact; weights <= 1 float32 ulp (float64 exp differs between libm
and CUDA by <= 1 ulp before the float32 rounding); grids/products bit-exact in reference-order mode on the
reference's own table; fast path within 1e-4 absolute / 1e-5 relative with identical NaN mask.
"""
import numpy as np
import pytest

import radar_grid_b200 as rg
from conftest import assert_same, canonical, golden_case, ulp_diff_f32
from oracle import radar_grid_oracle as O

pytestmark = pytest.mark.gpu

CASES = [("tiny", "barnes2", 0), ("tiny", "cressman", 0), ("tiny", "nearest", 0),
         ("tiny", "barnes2", 350), ("small", "barnes2", 0)]
ATOL, RTOL = 1e-4, 1e-5


def assert_close_same_mask(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, what
    np.testing.assert_array_equal(np.isnan(a), np.isnan(b), err_msg=f"{what}: NaN mask differs")
    np.testing.assert_allclose(a, b, rtol=RTOL, atol=ATOL, equal_nan=True, err_msg=what)


def build(spec, gates, weighting, alt, **kw):
    toa = spec.toa if alt == 0 else 4000.0
    return rg.DeviceGeometry.build(*gates, spec.grid_shape, spec.grid_limits, radar_altitude=float(alt),
                                   min_radius=spec.min_radius, beam_factor=spec.beam_factor, weighting=weighting,
                                   toa=toa, **kw)


@pytest.mark.parametrize("spec_name,weighting,alt", CASES)
def test_neighbour_table_sets_and_weights(spec_name, weighting, alt):
    spec, radar, gates, fields, g = golden_case(spec_name, weighting, alt)
    dev = build(spec, gates, weighting, alt)
    indptr, idx, w = dev.export_csr()
    assert_same(indptr.astype(np.int32), g["indptr"], "row lengths (indptr)")
    ci, cw = canonical(indptr, idx, w)
    ri, rw = canonical(g["indptr"], g["gate_indices"], g["weights"])
    assert_same(ci, ri, "neighbour sets")
    ulps = ulp_diff_f32(cw, rw)
    assert ulps.max(initial=0) <= 1, f"weights differ by {ulps.max()} ulp"
    assert (ulps > 0).mean() < 1e-4 if ulps.size else True
    info = dev.info
    assert info["n_pairs"] == len(g["gate_indices"])
    assert info["max_row_len"] == int(np.diff(g["indptr"]).max())
    assert info["n_empty_rows"] == int((np.diff(g["indptr"]) == 0).sum())


def test_build_is_deterministic_and_rows_sorted_within_cells():
    spec, radar, gates, fields, g = golden_case("tiny")
    a = build(spec, gates, "barnes2", 0).export_csr()
    b = build(spec, gates, "barnes2", 0).export_csr()
    for x, y in zip(a, b):
        assert_same(x, y, "two builds differ")


def test_zslab_build_is_a_row_slice():
    spec, radar, gates, fields, g = golden_case("tiny")
    nz, ny, nx = spec.grid_shape
    full = build(spec, gates, "barnes2", 0).export_csr()
    slab = build(spec, gates, "barnes2", 0, z_range=(2, 5)).export_csr()
    lo, hi = full[0][2 * ny * nx], full[0][5 * ny * nx]
    ci, cw = canonical(slab[0], slab[1], slab[2])
    fi, fw = canonical(full[0][2 * ny * nx:5 * ny * nx + 1] - lo, full[1][lo:hi], full[2][lo:hi])
    assert_same(ci, fi)
    assert_same(cw, fw)


@pytest.mark.parametrize("spec_name,weighting,alt", CASES)
def test_reference_order_apply_is_bit_exact_on_reference_table(spec_name, weighting, alt):
    spec, radar, gates, fields, g = golden_case(spec_name, weighting, alt)
    dev = rg.DeviceGeometry.from_csr(g["indptr"], g["gate_indices"], g["weights"], spec.grid_shape,
                                     spec.grid_limits, n_gates=len(gates[0]))
    names = list(fields)
    data = [np.ma.getdata(fields[n]) for n in names]
    masks = [np.ma.getmaskarray(fields[n]) for n in names]
    res = rg.grid_fields(dev, data, masks=masks, reference_order=True)
    for n, grid in zip(names, res["grids"]):
        assert_same(grid, g[f"grid_{n}"], f"grid_{n}")
    # GateFilter mask OR-ed on the host, as the reference does
    masks_qc = [m | g["rhohv_excluded"] for m in masks]
    res = rg.grid_fields(dev, data, masks=masks_qc, reference_order=True)
    for n, grid in zip(names, res["grids"]):
        assert_same(grid, g[f"gridqc_{n}"], f"gridqc_{n}")
    res = rg.grid_fields(dev, data[:1], masks=masks[:1], fill_value=-9999.0, reference_order=True)
    assert_same(res["grids"][0], g["grid_fill_DBZH"])


@pytest.mark.parametrize("spec_name,weighting,alt", CASES)
@pytest.mark.parametrize("own_table", [False, True])
def test_fast_apply_within_tolerance(spec_name, weighting, alt, own_table):
    spec, radar, gates, fields, g = golden_case(spec_name, weighting, alt)
    if own_table:
        dev = build(spec, gates, weighting, alt)
    else:
        dev = rg.DeviceGeometry.from_csr(g["indptr"], g["gate_indices"], g["weights"], spec.grid_shape,
                                         spec.grid_limits, n_gates=len(gates[0]))
    names = list(fields)
    data = [np.ma.getdata(fields[n]) for n in names]
    masks = [np.ma.getmaskarray(fields[n]) for n in names]
    res = rg.grid_fields(dev, data, masks=masks)
    for n, grid in zip(names, res["grids"]):
        assert grid.dtype == np.float32
        assert_close_same_mask(grid, g[f"grid_{n}"], f"grid_{n}")
    # every group width gives the same answer up to rounding
    for width in (4, 8, 16, 32):
        dev.ctx.set_option("group_width", width)
        r1 = rg.grid_fields(dev, data[:1], masks=masks[:1])
        assert_close_same_mask(r1["grids"][0], g["grid_DBZH"], f"group_width={width}")
        r5 = rg.grid_fields(dev, data, masks=masks)
        assert_close_same_mask(r5["grids"][-1], g[f"grid_{names[-1]}"], f"group_width={width} 5 fields")
    dev.ctx.set_option("group_width", 0)


@pytest.mark.parametrize("spec_name", ["tiny", "small"])
def test_slice_copy_and_csr_copy_sum_in_the_same_order(spec_name):
    """The warp-slice copy only changes where a lane's pairs are read from, not which pairs it takes or in which order:
    at equal group width the two table layouts must agree bit for bit, grids and fused products alike."""
    spec, radar, gates, fields, g = golden_case(spec_name)
    dev = build(spec, gates, "barnes2", 0)
    names = list(fields)
    data = [np.ma.getdata(fields[n]) for n in names]
    masks = [np.ma.getmaskarray(fields[n]) for n in names]
    reqs = [rg.ColumnMax(), rg.CAPPI(1234.5)]
    try:
        for width in (4, 8, 16, 32):
            dev.ctx.set_option("group_width", width)
            for nf in (1, 2, len(names)):
                out = {}
                for variant in (1, 4):
                    dev.ctx.set_option("apply_variant", variant)
                    out[variant] = rg.grid_fields(dev, data[:nf], masks=masks[:nf], products=reqs)
                for f in range(nf):
                    assert_same(out[1]["grids"][f], out[4]["grids"][f], f"W={width} F={nf} grid {f}")
                for a, b in zip(out[1]["products"], out[4]["products"]):
                    assert_same(a, b, f"W={width} F={nf} products")
    finally:
        import os
        dev.ctx.set_option("group_width", 0)
        dev.ctx.set_option("apply_variant", int(os.environ.get("RG_APPLY_VARIANT_TEST") or 0))   # what the session runs on


def test_fused_qc_rule_equals_host_mask_and_masked_invalid_on_device():
    spec, radar, gates, fields, g = golden_case("small")
    dev = build(spec, gates, "barnes2", 0)
    names = list(fields)
    raw = [np.ma.getdata(fields[n]).copy() for n in names]
    for n, r in zip(names, raw):            # put the NaNs back: the device derives the masks itself
        r[np.ma.getmaskarray(fields[n])] = np.nan
    rhohv = raw[names.index("RHOHV")]
    rules = [rg.RangeRule(rhohv, lo=0.8), rg.RangeRule(rhohv, hi=1.0)]
    res = rg.grid_fields(dev, raw, mask_invalid=True, rules=rules, products=[rg.ColumnMax()])
    for n, grid in zip(names, res["grids"]):
        assert_close_same_mask(grid, g[f"gridqc_{n}"], f"gridqc_{n}")
    assert_close_same_mask(res["products"][0][0], g["colmax_qc"], "fused colmax with QC")


@pytest.mark.parametrize("spec_name,weighting,alt", CASES)
def test_products_bit_exact_on_reference_grid(spec_name, weighting, alt):
    spec, radar, gates, fields, g = golden_case(spec_name, weighting, alt)
    grid = g["grid_DBZH"]
    geom = rg.GridGeometry(spec.grid_shape, spec.grid_limits, g["indptr"], g["gate_indices"], g["weights"], spec.toa)
    zmax = spec.grid_limits[0][1]
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        assert_same(rg.column_max(grid), g["colmax"], "colmax")
        assert_same(rg.column_max(grid, z_min_idx=1, z_max_idx=3), g["colmax_idx_1_3"])
        assert_same(rg.column_max(grid, z_min_alt=1500.0, z_max_alt=0.6 * zmax, geometry=geom), g["colmax_alt"])
        assert_same(rg.column_min(grid), g["colmin"], "colmin")
        assert_same(rg.column_mean(grid), g["colmean"], "colmean")
        assert_same(rg.column_max(g["grid_fill_DBZH"]), g["colmax_fill"], "colmax with numeric fill")
    for alt_m in (4000.0, 0.0, zmax, 1234.5, zmax / (spec.grid_shape[0] - 1) * 2):
        tag = f"{alt_m:.1f}"
        assert_same(rg.constant_altitude_ppi(grid, geom, alt_m), g[f"cappi_lin_{tag}"], f"cappi_lin {tag}")
        assert_same(rg.constant_altitude_ppi(grid, geom, alt_m, "nearest"), g[f"cappi_near_{tag}"], f"cappi_near {tag}")
    assert_same(rg.constant_altitude_ppi(grid, geom, zmax + 1.0), g["cappi_oob"])
    geom64 = rg.GridGeometry(spec.grid_shape, tuple(tuple(np.float64(v) for v in l) for l in spec.grid_limits),
                             g["indptr"], g["gate_indices"], g["weights"], spec.toa)
    assert_same(rg.constant_altitude_ppi(grid, geom64, 1234.5), g["cappi_lin64_1234.5"], "float64-weight CAPPI")
    for elev in (0.5, 2.3, 6.9, 25.0):
        tag = f"{elev:.1f}"
        assert_same(rg.constant_elevation_ppi(grid, geom, elev), g[f"ppi_lin_{tag}"], f"ppi_lin {tag}")
        assert_same(rg.constant_elevation_ppi(grid, geom, elev, interpolation="nearest"), g[f"ppi_near_{tag}"], f"ppi_near {tag}")
        assert_same(rg.constant_elevation_ppi(grid, geom, elev, earth_curvature=False), g[f"ppi_flat_{tag}"], f"ppi_flat {tag}")


@pytest.mark.parametrize("reference_order", [False, True])
def test_fused_epilogue_equals_products_of_the_grid(reference_order):
    """COLMAX / CAPPI / PPI from the epilogue are bit-identical to the stand-alone products of the same grid,
    and a products-only call (no 3-D output) gives the same planes."""
    spec, radar, gates, fields, g = golden_case("small")
    dev = build(spec, gates, "barnes2", 0)
    names = list(fields)
    data = [np.ma.getdata(fields[n]) for n in names]
    masks = [np.ma.getmaskarray(fields[n]) for n in names]
    reqs = [rg.ColumnMax(), rg.ColumnMin(z_min_idx=2, z_max_idx=7), rg.ColumnMean(), rg.CAPPI(4000.0),
            rg.CAPPI(1234.5), rg.PPI(0.5), rg.PPI(2.3, "nearest")]
    both = rg.grid_fields(dev, data, masks=masks, products=reqs, reference_order=reference_order)
    only = rg.grid_fields(dev, data, masks=masks, products=reqs, want_grid=False, reference_order=reference_order)
    assert all(x is None for x in only["grids"])
    alone = rg.run_products(both["grids"], spec.grid_shape, spec.grid_limits, reqs)
    for r, a, b, c in zip(reqs, both["products"], only["products"], alone):
        assert_same(a, c, f"fused vs stand-alone: {r}")
        assert_same(a, b, f"with vs without 3-D output: {r}")
    # and against the oracle's products of the oracle's grid (tolerance: the grid itself is within tolerance)
    og = O.apply_geometry(g["indptr"], g["gate_indices"], g["weights"], spec.grid_shape, fields["DBZH"])
    assert_close_same_mask(both["products"][0][0], O.column_reduce("max", og), "colmax vs oracle")
    assert_close_same_mask(both["products"][3][0], O.cappi(og, spec.grid_shape, spec.grid_limits, 4000.0), "cappi vs oracle")
    assert_close_same_mask(both["products"][5][0], O.ppi(og, spec.grid_shape, spec.grid_limits, 0.5), "ppi vs oracle")


def test_linearity_and_constant_field_properties():
    """Size-independent properties: a constant field grids to the constant wherever defined; the operator is linear."""
    spec, radar, gates, fields, g = golden_case("small")
    dev = build(spec, gates, "barnes2", 0)
    n = len(gates[0])
    rng = np.random.default_rng(3)
    a = rng.normal(size=n).astype(np.float32)
    b = rng.normal(size=n).astype(np.float32)
    const = np.full(n, 7.25, dtype=np.float32)
    res = rg.grid_fields(dev, [a, b, (2 * a + 3 * b).astype(np.float32), const])["grids"]
    defined = ~np.isnan(res[3])
    np.testing.assert_array_equal(defined, np.diff(g["indptr"]).reshape(spec.grid_shape) > 0)
    np.testing.assert_allclose(res[3][defined], 7.25, rtol=1e-6)
    np.testing.assert_allclose(res[2][defined], (2 * res[0] + 3 * res[1])[defined], rtol=1e-4, atol=1e-4)


def test_zslab_shards_concatenate_and_colmax_reduces():
    """z-slab sharding: slabs built and gridded independently give the same rows bit for bit, and the
    nan-aware max of the partial COLMAX planes is the COLMAX of the whole grid."""
    import torch
    from radar_grid_b200 import distributed as D
    spec, radar, gates, fields, g = golden_case("small")
    names = list(fields)
    data = [np.ma.getdata(fields[n]) for n in names]
    masks = [np.ma.getmaskarray(fields[n]) for n in names]
    whole = build(spec, gates, "barnes2", 0)
    # bit-for-bit equality needs the same summation order, i.e. the same group width, for the slabs and the whole grid
    # (left to itself the library picks it from each table's mean row length)
    whole.ctx.set_option("group_width", 8)
    try:
        _zslab_checks(spec, gates, names, data, masks, whole, D, torch)
    finally:
        whole.ctx.set_option("group_width", 0)


def _zslab_checks(spec, gates, names, data, masks, whole, D, torch):
    full = rg.grid_fields(whole, data, masks=masks, products=[rg.ColumnMax(), rg.ColumnMin()])
    parts, cmax, cmin = [], None, None
    for z0, z1 in D.zslab_ranges(spec.grid_shape[0], 3):
        slab = build(spec, gates, "barnes2", 0, z_range=(z0, z1))
        r = rg.grid_fields(slab, data, masks=masks, products=[rg.ColumnMax(), rg.ColumnMin()])
        assert r["grids"][0].shape == (z1 - z0, spec.grid_shape[1], spec.grid_shape[2])
        parts.append(r["grids"])
        pm, pn = torch.from_numpy(r["products"][0].copy()), torch.from_numpy(r["products"][1].copy())
        # what all_reduce(MAX/MIN) does across ranks, with the -inf/+inf encoding of "no data"
        enc = lambda t, s: torch.where(torch.isnan(t), torch.full_like(t, s), t)
        cmax = enc(pm, -np.inf) if cmax is None else torch.maximum(cmax, enc(pm, -np.inf))
        cmin = enc(pn, np.inf) if cmin is None else torch.minimum(cmin, enc(pn, np.inf))
    for f in range(len(names)):
        assert_same(np.concatenate([p[f] for p in parts], axis=0), full["grids"][f], f"slab concat field {f}")
    dec = lambda t, s: torch.where(t == s, torch.full_like(t, float("nan")), t).numpy()
    assert_same(dec(cmax, -np.inf), full["products"][0], "z-slab COLMAX")
    assert_same(dec(cmin, np.inf), full["products"][1], "z-slab COLMIN")
    one = D.allreduce_nanmax(torch.from_numpy(full["products"][0].copy()))      # world size 1: identity incl. NaNs
    assert_same(one.numpy(), full["products"][0])


def test_zslab_cappi_two_party_sum_is_bit_identical():
    """A CAPPI whose two levels sit in different z-slabs: every slab contributes weight x level for the
    levels it owns, out of the same fused pass as its COLMAX (LevelPick), and the sum of the contributions — what
    all_reduce(SUM) does across ranks — is the unsharded fused CAPPI bit for bit, for the float32 and float64 blends,
    level picks and the out-of-range NaN plane."""
    import torch
    from radar_grid_b200 import distributed as D
    spec, radar, gates, fields, g = golden_case("small")
    names = list(fields)
    data = [np.ma.getdata(fields[n]) for n in names]
    masks = [np.ma.getmaskarray(fields[n]) for n in names]
    nz, ny, nx = spec.grid_shape
    z_top = spec.grid_limits[0][1]
    step = z_top / (nz - 1)
    ranges = D.zslab_ranges(nz, 3)
    b = ranges[0][1]
    lims64 = tuple(tuple(np.float64(v) for v in ax) for ax in spec.grid_limits)
    whole = build(spec, gates, "barnes2", 0)
    whole.ctx.set_option("group_width", 8)           # same summation order for slabs and whole grid, as above
    try:
        slabs = [build(spec, gates, "barnes2", 0, z_range=zr) for zr in ranges]
        like = torch.empty((len(names), ny, nx))
        for alt, interp, lims in [((b - 0.6) * step, "linear", spec.grid_limits), ((b - 0.6) * step, "linear", lims64),
                                  ((ranges[1][1] - 0.25) * step, "linear", spec.grid_limits),
                                  (0.3 * step, "linear", spec.grid_limits), (b * step, "linear", spec.grid_limits),
                                  ((b - 0.4) * step, "nearest", spec.grid_limits), (z_top + 5.0, "linear", spec.grid_limits)]:
            req = rg.CAPPI(alt, interp)
            whole.grid_limits = lims
            want = rg.grid_fields(whole, data, masks=masks, want_grid=False, products=[req])["products"][0]
            total, owners = None, 0
            for zr, slab in zip(ranges, slabs):
                def level_planes(levels, slab=slab):
                    r = rg.grid_fields(slab, data, masks=masks, want_grid=False,
                                       products=[rg.ColumnMax()] + [rg.LevelPick(z) for z in levels])
                    return [torch.from_numpy(p.copy()) for p in r["products"][1:]]
                terms = D.cappi_zslab_terms(req, spec.grid_shape, lims, zr)
                owners += 0 if terms is None else len(terms[0])
                part = D.cappi_zslab_partial(req, spec.grid_shape, lims, zr, level_planes, like)
                if part is not None:
                    total = part if total is None else total + part          # what all_reduce(SUM) does
            if total is None:                                                # outside the grid on every rank
                total = D.cappi_zslab(req, spec.grid_shape, lims, ranges[0], None, like)
            got = total.to(torch.float32).numpy()
            assert_same(got, want, f"z-slab CAPPI {alt} {interp}")
            assert np.array_equal(np.signbit(got), np.signbit(want))
            if alt <= z_top:
                assert owners == (1 if (interp == "nearest" or alt == b * step) else 2)
    finally:
        whole.grid_limits = spec.grid_limits
        whole.ctx.set_option("group_width", 0)


def test_zslab_ppi_partial_blends_sum_to_the_fused_ppi():
    """PPI over z-slabs: per pixel every slab weights the levels of the beam's pair that it owns, gathered
    from its own 3-D grid on the device; the sum of the partial planes (all_reduce(SUM) across ranks) is the fused PPI
    of the unsharded grid bit for bit — float64 for 'linear', float32 for 'nearest', NaN where the beam leaves the grid."""
    import torch
    from radar_grid_b200 import distributed as D
    spec, radar, gates, fields, g = golden_case("small")
    names = list(fields)
    data = [torch.from_numpy(np.ma.getdata(fields[n]).copy()).cuda() for n in names]
    masks = [torch.from_numpy(np.ma.getmaskarray(fields[n]).copy()).cuda() for n in names]
    ranges = D.zslab_ranges(spec.grid_shape[0], 3)
    whole = build(spec, gates, "barnes2", 0)
    whole.ctx.set_option("group_width", 8)           # same summation order for slabs and whole grid, as above
    try:
        slab_grids = []
        for zr in ranges:
            slab = build(spec, gates, "barnes2", 0, z_range=zr)
            slab_grids.append(torch.stack(rg.grid_fields(slab, data, masks=masks)["grids"]))
        for el, interp, curved in [(0.5, "linear", True), (4.0, "linear", True), (4.0, "nearest", True),
                                   (12.0, "linear", False), (30.0, "nearest", False)]:
            req = rg.PPI(el, interp, curved)
            want = rg.grid_fields(whole, data, masks=masks, want_grid=False, products=[req])["products"][0]
            plan = D.ppi_zslab_plan(req, spec.grid_shape, spec.grid_limits)
            total = None
            for zr, sg in zip(ranges, slab_grids):
                part = D.ppi_zslab_partial(plan, zr, sg)
                assert part.is_cuda
                total = part if total is None else total + part
            assert_same(total.cpu().numpy(), want.cpu().numpy(), f"z-slab PPI {el} {interp}")
            assert int((~torch.isnan(want)).sum()) > 0
    finally:
        whole.ctx.set_option("group_width", 0)


def test_device_buffers_match_host_buffers():
    """torch CUDA tensors in / out (zero-copy, RG_DEVICE) give bit-identical results to NumPy buffers (RG_HOST)."""
    torch = pytest.importorskip("torch")
    if not torch.cuda.is_available():
        pytest.skip("torch sees no CUDA device")
    spec, radar, gates, fields, g = golden_case("small")
    dev = build(spec, gates, "barnes2", 0)
    names = list(fields)
    raw = [np.ma.getdata(fields[n]).copy() for n in names]
    for n, r in zip(names, raw):
        r[np.ma.getmaskarray(fields[n])] = np.nan
    reqs = [rg.ColumnMax(), rg.CAPPI(1234.5), rg.PPI(2.3)]
    host = rg.grid_fields(dev, raw, mask_invalid=True, products=reqs)
    dten = [torch.from_numpy(r).cuda() for r in raw]
    torch.cuda.synchronize()
    devr = rg.grid_fields(dev, dten, mask_invalid=True, products=reqs)
    dev.ctx.synchronize()
    for a, b in zip(host["grids"], devr["grids"]):
        assert b.is_cuda and b.dtype == torch.float32
        assert_same(a, b.cpu().numpy())
    for a, b in zip(host["products"], devr["products"]):
        assert_same(a, b.cpu().numpy())


def test_errors_mirror_the_reference():
    spec, radar, gates, fields, g = golden_case("tiny")
    with pytest.raises(ValueError):
        rg.compute_grid_geometry(*gates, spec.grid_shape, spec.grid_limits, temp_dir="/nonexistent/dir")     # compute.py:173
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        with pytest.raises(ValueError):
            rg.compute_grid_geometry(*gates, spec.grid_shape, spec.grid_limits, temp_dir=tmp, weighting="gauss")   # :176
        geom = rg.compute_grid_geometry(*gates, spec.grid_shape, spec.grid_limits, temp_dir=tmp, min_radius=spec.min_radius,
                                        beam_factor=spec.beam_factor, radar_altitude=0.0)
    assert geom.radar_altitude == 0.0 and geom.n_pairs() == len(g["gate_indices"])
    assert_same(geom.indptr.astype(np.int32), g["indptr"])
    with pytest.raises(ValueError):
        rg.apply_geometry(geom, fields["DBZH"], additional_filters="nope")                      # interpolate.py:56
    grid = rg.apply_geometry(geom, fields["DBZH"])
    with pytest.raises(ValueError):
        rg.constant_altitude_ppi(grid, geom, 1000.0, interpolation="cubic")                     # products.py:415
    with pytest.raises(ValueError):
        rg.constant_elevation_ppi(grid, geom, 1.0, interpolation="cubic")                       # products.py:312
    with pytest.raises(ValueError):
        rg.column_max(grid, z_min_alt=1000.0)                                                  # products.py:467
    bad = gates[0].copy()
    bad[5] = np.nan
    with pytest.raises(ValueError):
        rg.DeviceGeometry.build(bad, gates[1], gates[2], spec.grid_shape, spec.grid_limits)     # cKDTree refuses non-finite data
    # save / load round trip of a GPU-built geometry, then apply through the loaded (host CSR) twin
    with tempfile.TemporaryDirectory() as tmp:
        path = tmp + "/g.npz"
        rg.save_geometry(geom, path)
        again = rg.load_geometry(path)
    assert_same(rg.apply_geometry(again, fields["DBZH"]), grid, "apply through a reloaded geometry")


def test_volume_pipeline_equals_sequential_calls():
    """A time series through VolumePipeline (3 streams, copies overlapped) gives exactly the per-volume results."""
    from radar_grid_b200 import synthetic as S
    spec, radar, gates, fields, g = golden_case("small")
    dev = build(spec, gates, "barnes2", 0)
    nz, ny, nx = spec.grid_shape
    jobs, expect = [], []
    for seed in range(7):
        vol = S.make_fields(spec, seed=seed, gates=gates)
        raw = []
        for n in spec.fields:
            a = rg.pinned_empty((spec.n_gates,), np.float32)
            a[:] = np.ma.getdata(vol[n])
            a[np.ma.getmaskarray(vol[n])] = np.nan
            raw.append(a)
        jobs.append({"fields": raw, "mask_invalid": True, "products": [rg.ColumnMax(), rg.CAPPI(4000.0)],
                     "out_grids": [rg.pinned_empty((nz, ny, nx), np.float32) for _ in spec.fields]})
        expect.append(rg.grid_fields(dev, raw, mask_invalid=True, products=[rg.ColumnMax(), rg.CAPPI(4000.0)]))
    pipe = rg.VolumePipeline(dev, n_streams=3)
    import os
    if os.environ.get("RG_APPLY_VARIANT_TEST"):          # keep the A/B kernel selection consistent with `expect`
        pipe.map(jobs[:1])                               # (the A/B kernel builds its table copy lazily: do it once, serially)
        for c in pipe.ctxs:
            c.set_option("apply_variant", int(os.environ["RG_APPLY_VARIANT_TEST"]))
        pipe.ctxs[0].synchronize()
        rg.grid_fields(dev, ctx=pipe.ctxs[0], **jobs[0])
    got = pipe.map(jobs)
    pipe.close()
    for e, r in zip(expect, got):
        for a, b in zip(e["grids"], r["grids"]):
            assert_same(a, np.asarray(b))
        for a, b in zip(e["products"], r["products"]):
            assert_same(a, b)
    # volumes differ from one another (the pipeline did not hand back one buffer seven times)
    assert not np.array_equal(got[0]["products"][0], got[1]["products"][0], equal_nan=True)


def test_heavy_rows_go_through_the_chunk_kernel():
    """Rows longer than 512 pairs (the voxels next to the radar) are reduced by heavy_rows_kernel, in chunks of 4096
    pairs, and picked up by the column kernel: a hand-made table with rows of 600, 5 000 and 9 000 pairs next to
    ordinary and empty rows, every field count and both table layouts, against the oracle."""
    rng = np.random.default_rng(42)
    shape, limits = (3, 2, 9), ((0.0, 2000.0), (-500.0, 500.0), (-2000.0, 2000.0))
    n_gates = 12000
    lens = rng.integers(0, 40, size=int(np.prod(shape)))
    lens[[4, 5, 13]] = (5000, 600, 9000)            # same slice / same column (rows 4 and 13 are not; 4 and 22 would be)
    lens[22] = 513
    lens[[0, 30]] = 0
    indptr = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    idx = np.concatenate([np.sort(rng.choice(n_gates, size=n, replace=False)) for n in lens]).astype(np.int32)
    w = rng.uniform(0.02, 1.0, size=idx.shape[0]).astype(np.float32)
    dev = rg.DeviceGeometry.from_csr(indptr, idx, w, shape, limits, n_gates)
    fields = []
    for f in range(8):
        v = rng.normal(10.0 * f, 5.0, size=n_gates).astype(np.float32)
        v[rng.random(n_gates) < (0.5 if f == 0 else 0.05)] = np.nan
        fields.append(np.ma.masked_invalid(v))
    want = [O.apply_geometry(indptr, idx, w, shape, f) for f in fields]
    assert not np.isnan(want[0].ravel()[[4, 5, 13, 22]]).any()
    import os, warnings
    try:
        for variant in (1, 4):
            dev.ctx.set_option("apply_variant", variant)
            for nf in (1, 2, 3, 4, 5, 6, 7, 8):
                res = rg.grid_fields(dev, [np.ma.getdata(f) for f in fields[:nf]], masks=[np.ma.getmaskarray(f) for f in fields[:nf]],
                                     products=[rg.ColumnMax()])
                for f in range(nf):
                    assert_close_same_mask(res["grids"][f], want[f], f"variant {variant} F={nf} field {f}")
                    with warnings.catch_warnings():
                        warnings.simplefilter("ignore", RuntimeWarning)
                        assert_same(res["products"][0][f], np.nanmax(res["grids"][f], axis=0), "fused COLMAX")
                exact = rg.grid_fields(dev, [np.ma.getdata(f) for f in fields[:nf]], masks=[np.ma.getmaskarray(f) for f in fields[:nf]],
                                       reference_order=True)
                for f in range(nf):
                    assert_same(exact["grids"][f], want[f], f"reference order F={nf} field {f}")
    finally:
        dev.ctx.set_option("apply_variant", int(os.environ.get("RG_APPLY_VARIANT_TEST") or 0))


def test_empty_zslab_still_writes_its_product_planes():
    """z_begin == z_end (more ranks than levels): no rows, but COLMAX / CAPPI / COLMEAN planes of a fused request must be
    the no-data planes, not whatever the output buffer held."""
    spec, radar, gates, fields, g = golden_case("tiny")
    dev = build(spec, gates, "barnes2", 0, z_range=(3, 3))
    assert dev.n_rows == 0
    name = list(fields)[0]
    _, ny, nx = spec.grid_shape
    outs = [np.full((1, ny, nx), 123.0, dtype=np.float32) for _ in range(3)]
    res = rg.grid_fields(dev, [np.ma.getdata(fields[name])], masks=[np.ma.getmaskarray(fields[name])], want_grid=False,
                         products=[rg.ColumnMax(), rg.LevelPick(2), rg.ColumnMean()], out_products=outs)
    for p in res["products"]:
        assert np.isnan(p).all()


def test_zslab_fused_terms_of_every_product_kind_merge_to_the_unsharded_products():
    """distributed.zslab_terms / zslab_merge / zslab_finish: ONE fused pass per slab with partial=True requests (no slab
    3-D grid, no torch arithmetic before the collective) -- COLMAX / COLMIN with -inf / +inf for "no data", CAPPI
    (float32 and float64 blends, level pick), PPI (linear -> float64, nearest) and LevelPick as sums over the owned
    levels with -0.0 elsewhere -- merged the way all-reduce(MAX | MIN | SUM) merges ranks, is bit-identical to the
    unsharded fused products, for balanced (uneven) slabs and with an empty slab in the list."""
    import torch
    from radar_grid_b200 import distributed as D
    spec, radar, gates, fields, g = golden_case("small")
    names = list(fields)
    data = [np.ma.getdata(fields[n]) for n in names]
    masks = [np.ma.getmaskarray(fields[n]) for n in names]
    nz = spec.grid_shape[0]
    step = spec.grid_limits[0][1] / (nz - 1)
    lims64 = tuple(tuple(np.float64(v) for v in ax) for ax in spec.grid_limits)
    pairs = rg.DeviceGeometry.level_pairs(*gates, spec.grid_shape, spec.grid_limits, min_radius=spec.min_radius,
                                          beam_factor=spec.beam_factor, toa=spec.toa, column_stride=1)
    whole = build(spec, gates, "barnes2", 0)
    indptr = whole.export_csr()[0].astype(np.int64)
    ny, nx = spec.grid_shape[1:]
    np.testing.assert_array_equal(pairs, np.diff(indptr[::ny * nx]))         # exact census at stride 1
    approx = rg.DeviceGeometry.level_pairs(*gates, spec.grid_shape, spec.grid_limits, min_radius=spec.min_radius,
                                           beam_factor=spec.beam_factor, toa=spec.toa, column_stride=3)
    assert abs(approx.sum() - pairs.sum()) < 0.1 * pairs.sum()
    ranges = D.zslab_ranges(nz, 4, weights=pairs)
    assert ranges[0][0] == 0 and ranges[-1][1] == nz and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
    ranges = ranges[:2] + [(ranges[1][1], ranges[1][1])] + ranges[2:]         # plus an empty slab
    whole.ctx.set_option("group_width", 8)            # same summation order for slabs and whole grid
    try:
        for lims in (spec.grid_limits, lims64):
            reqs = [rg.ColumnMax(), rg.ColumnMin(z_min_idx=1), rg.CAPPI((ranges[0][1] - 0.6) * step), rg.CAPPI(2.0 * step),
                    rg.CAPPI(1234.5, "nearest"), rg.PPI(2.3)]
            reqs2 = [rg.ColumnMax(z_max_idx=nz - 2), rg.PPI(6.9, "nearest"), rg.PPI(0.5, earth_curvature=False), rg.LevelPick(nz - 1),
                     rg.CAPPI(spec.grid_limits[0][1] + 10.0)]
            reqs3 = [rg.ColumnMax(), rg.CAPPI((ranges[0][1] - 0.6) * step)]     # the operational pair: its own kernel variant
            for rq in (reqs, reqs2, reqs3):
                whole.grid_limits = lims
                want = rg.grid_fields(whole, data, masks=masks, want_grid=False, products=rq)["products"]
                acc = None
                for zr in ranges:
                    slab = build(spec, gates, "barnes2", 0, z_range=zr)
                    slab.grid_limits = lims
                    terms = D.zslab_terms(slab, data, masks=masks, products=rq)
                    acc = terms if acc is None else D.zslab_merge(acc, terms, rq)
                got = D.zslab_finish(acc, rq)
                for p, a, b in zip(rq, got, want):
                    assert_same(a.numpy(), b, f"z-slab {type(p).__name__} {p}")
                    assert np.array_equal(np.signbit(a.numpy()), np.signbit(b))
    finally:
        whole.grid_limits = spec.grid_limits
        whole.ctx.set_option("group_width", 0)


def test_caller_buffers_are_validated_not_reinterpreted():
    """out_grids / out_products / device fields and masks reach the library as raw pointers: anything that is not
    exactly the contiguous array of the right dtype, size and side must raise instead of being read as something else."""
    import torch
    spec, radar, gates, fields, g = golden_case("tiny")
    dev = build(spec, gates, "barnes2", 0)
    name = list(fields)[0]
    data, mask = np.ma.getdata(fields[name]), np.ma.getmaskarray(fields[name])
    nz, ny, nx = spec.grid_shape
    ok = rg.grid_fields(dev, [data], masks=[mask], out_grids=[np.empty((nz, ny, nx), np.float32)],
                        products=[rg.PPI(2.3)], out_products=[np.empty((1, ny, nx), np.float64)])
    assert ok["products"][0].dtype == np.float64
    for bad in (np.empty((nz, ny, nx), np.float64), np.empty((nz, ny, 2 * nx), np.float32)[:, :, ::2], np.empty((nz, ny, nx - 1), np.float32),
                torch.empty((nz, ny, nx), dtype=torch.float32, device="cuda")):
        with pytest.raises(ValueError):
            rg.grid_fields(dev, [data], masks=[mask], out_grids=[bad])
    with pytest.raises(ValueError):                                          # PPI 'linear' planes are float64
        rg.grid_fields(dev, [data], masks=[mask], products=[rg.PPI(2.3)], out_products=[np.empty((1, ny, nx), np.float32)])
    d_data = torch.from_numpy(data).cuda()
    with pytest.raises(ValueError):
        rg.grid_fields(dev, [d_data.double()])
    with pytest.raises(ValueError):
        rg.grid_fields(dev, [d_data], masks=[torch.from_numpy(mask).cuda()[:-1]])
    with pytest.raises(ValueError):
        rg.grid_fields(dev, [d_data], out_grids=[np.empty((nz, ny, nx), np.float32)])
    out = rg.grid_fields(dev, [d_data], masks=[torch.from_numpy(mask).cuda()])["grids"][0]
    assert_same(out.cpu().numpy(), ok["grids"][0])
