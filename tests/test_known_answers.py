"""
The reference's own known-answer unit tests for the hot path (tests/test_radar_grid_interpolate.py,
tests/test_radar_grid_products.py, tests/test_grid_filters.py in the reference repo), replayed against
  - the CPU oracle              (runs everywhere)
  - the CUDA path via the C ABI (gpu-marked)
with the same hand-built CSR tables and expected values (file:line of the original assertion in comments).
"""
import warnings

import numpy as np
import pytest

import radar_grid_b200 as rg
from oracle import radar_grid_oracle as O

LIMITS = ((0, 1000), (-500, 500), (-500, 500))


class OracleImpl:
    name = "oracle"

    def apply(self, shape, indptr, idx, w, field, masks=(), fill=np.nan):
        return O.apply_geometry(np.asarray(indptr, np.int32), np.asarray(idx, np.int32), np.asarray(w, np.float32), shape, field,
                                masks, fill)

    def colmax(self, g): return O.column_reduce("max", g)
    def colmin(self, g): return O.column_reduce("min", g)
    def colmean(self, g): return O.column_reduce("mean", g)
    def cappi(self, g, shape, lim, alt, interp="linear"): return O.cappi(g, shape, lim, alt, interp)
    def ppi(self, g, shape, lim, elev, interp="linear"): return O.ppi(g, shape, lim, elev, interp)
    def gridfilter(self, kind, plane, *a, **k): return O.grid_filter(kind, plane, *a, **k)


class CudaImpl:
    name = "cuda"

    def _geom(self, shape, indptr, idx, w):
        return rg.GridGeometry(shape, LIMITS, np.asarray(indptr, np.int32), np.asarray(idx, np.int32),
                               np.asarray(w, np.float32), 2000.0)

    def apply(self, shape, indptr, idx, w, field, masks=(), fill=np.nan):
        class _F:                       # stand-in carrying a precomputed exclude mask, as a GateFilter would
            def __init__(self, m): self.gate_excluded = m
        geom = self._geom(shape, indptr, idx, w)
        mask = np.ma.getmask(field)
        for m in masks:
            mask = mask | m
        data = np.ma.getdata(field)
        field = np.ma.array(data, mask=mask if mask is not np.ma.nomask else np.zeros(data.shape, bool))
        return rg.apply_geometry(geom, field, fill_value=fill)

    def _q(self, fn, *a, **k):
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            return fn(*a, **k)

    def colmax(self, g): return self._q(rg.column_max, g)
    def colmin(self, g): return self._q(rg.column_min, g)
    def colmean(self, g): return self._q(rg.column_mean, g)

    def cappi(self, g, shape, lim, alt, interp="linear"):
        return rg.constant_altitude_ppi(g, rg.GridGeometry(shape, lim, None, None, None, 0.0), alt, interp)

    def ppi(self, g, shape, lim, elev, interp="linear"):
        return rg.constant_elevation_ppi(g, rg.GridGeometry(shape, lim, None, None, None, 0.0), elev, interp)

    def gridfilter(self, kind, plane, *a, **k):
        f = rg.GridFilter()
        return {"below": f.apply_below, "above": f.apply_above, "outside": f.apply_outside_range,
                "invalid": f.apply_invalid}[kind](plane, *a, **k)


IMPLS = [pytest.param(OracleImpl(), id="oracle"), pytest.param(CudaImpl(), id="cuda", marks=pytest.mark.gpu)]


def simple_table():
    """2x2x2 grid, two gates per voxel, unit weights (test_radar_grid_interpolate.py:20-38)."""
    return (2, 2, 2), np.arange(0, 17, 2, dtype=np.int32), np.arange(16, dtype=np.int32), np.ones(16, np.float32)


# ---- apply_geometry ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("impl", IMPLS)
def test_basic_shape_dtype_and_values(impl):
    shape, ip, idx, w = simple_table()
    out = impl.apply(shape, ip, idx, w, np.ma.masked_invalid(np.arange(16, dtype=np.float32)))
    assert out.shape == (2, 2, 2) and out.dtype == np.float32                      # :46-48
    np.testing.assert_array_equal(out.ravel(), np.arange(8) * 2 + 0.5)


@pytest.mark.parametrize("impl", IMPLS)
def test_nan_gates_and_masked_gates_give_nan_voxels(impl):
    shape, ip, idx, w = simple_table()
    f = np.ones(16, dtype=np.float32) * 10.0
    f[0:4] = np.nan
    out = impl.apply(shape, ip, idx, w, np.ma.masked_invalid(f))
    assert np.isnan(out.ravel()[0]) and np.isnan(out.ravel()[1]) and out.ravel()[2] == 10.0   # :50-59
    m = np.zeros(16, dtype=bool)
    m[0:4] = True
    out = impl.apply(shape, ip, idx, w, np.ma.array(np.ones(16, dtype=np.float32) * 10.0, mask=m))
    assert np.isnan(out.ravel()[0]) and out.ravel()[3] == 10.0                     # :61-73


@pytest.mark.parametrize("impl", IMPLS)
def test_weighted_average_known_answers(impl):
    one = (1, 1, 1)
    out = impl.apply(one, [0, 2], [0, 1], [0.3, 0.7], np.ma.masked_invalid(np.array([10.0, 20.0], np.float32)))
    np.testing.assert_almost_equal(out[0, 0, 0], 17.0, decimal=5)                  # :75-93
    out = impl.apply(one, [0, 3], [0, 1, 2], [0.2, 0.5, 0.3], np.ma.masked_invalid(np.array([10.0, 20.0, 30.0], np.float32)))
    np.testing.assert_almost_equal(out[0, 0, 0], 21.0, decimal=5)                  # :236-254
    out = impl.apply(one, [0, 3], [0, 1, 2], [0.3, 0.4, 0.3], np.ma.masked_invalid(np.array([10.0, np.nan, 30.0], np.float32)))
    np.testing.assert_almost_equal(out[0, 0, 0], 20.0, decimal=5)                  # :256-277


@pytest.mark.parametrize("impl", IMPLS)
def test_fill_value_empty_rows_all_masked_and_inf(impl):
    one = (1, 1, 1)
    out = impl.apply(one, [0, 0], [], [], np.ma.masked_invalid(np.array([10.0], np.float32)), fill=-9999.0)
    assert out[0, 0, 0] == -9999.0                                                 # :116-132
    out = impl.apply(one, [0, 2], [0, 1], [0.5, 0.5], np.ma.array(np.array([10.0, 20.0], np.float32), mask=[True, True]))
    assert np.isnan(out[0, 0, 0])                                                  # :134-153
    out = impl.apply((2, 2, 2), np.zeros(9, np.int32), [], [], np.ma.masked_invalid(np.array([10.0], np.float32)))
    assert np.all(np.isnan(out))                                                   # :218-234
    out = impl.apply(one, [0, 3], [0, 1, 2], [0.3, 0.4, 0.3], np.ma.masked_invalid(np.array([10.0, np.inf, 30.0], np.float32)))
    assert np.isfinite(out[0, 0, 0])                                               # :283-300
    out = impl.apply(one, [0, 2], [0, 1], [0.5, 0.5], np.ma.masked_invalid(np.array([10.0, -np.inf], np.float32)))
    np.testing.assert_almost_equal(out[0, 0, 0], 10.0, decimal=5)                  # :302-317


@pytest.mark.parametrize("impl", IMPLS)
def test_gate_filter_mask_is_ored_in(impl):
    shape, ip, idx, w = simple_table()
    excl = np.zeros(16, dtype=bool)
    excl[0:4] = True
    out = impl.apply(shape, ip, idx, w, np.ma.array(np.ones(16, dtype=np.float32) * 10.0), masks=[excl])
    assert out.shape == (2, 2, 2) and np.isnan(out.ravel()[0]) and np.isnan(out.ravel()[1]) and out.ravel()[2] == 10.0


@pytest.mark.parametrize("impl", IMPLS)
def test_unmasked_nan_propagates_but_masked_nan_does_not(impl):
    """Only *masked* gates are zeroed (interpolate.py:78-79); an unmasked NaN poisons its voxel."""
    one = (1, 1, 1)
    f = np.ma.array(np.array([10.0, np.nan], np.float32), mask=[False, False])
    assert np.isnan(impl.apply(one, [0, 2], [0, 1], [0.5, 0.5], f)[0, 0, 0])
    f = np.ma.array(np.array([10.0, np.nan], np.float32), mask=[False, True])
    assert impl.apply(one, [0, 2], [0, 1], [0.5, 0.5], f)[0, 0, 0] == 10.0


# ---- products ----------------------------------------------------------------------------------------------------------
def level_grid():
    """(10, 50, 50) grid with value z*10 and an all-NaN corner (test_radar_grid_products.py:281-296)."""
    d = np.zeros((10, 50, 50), dtype=np.float32)
    for z in range(10):
        d[z] = z * 10.0
    d[:, 0:5, 0:5] = np.nan
    return d


PLIM = ((0.0, 10000.0), (-25000.0, 25000.0), (-25000.0, 25000.0))


@pytest.mark.parametrize("impl", IMPLS)
def test_column_aggregations(impl):
    d = level_grid()
    cmax, cmin, cmean = impl.colmax(d), impl.colmin(d), impl.colmean(d)
    assert cmax.shape == (50, 50)
    assert cmax[10, 10] == 90.0 and np.isnan(cmax[0, 0])                           # :299-307
    assert cmin[10, 10] == 0.0 and np.isnan(cmin[0, 0])                            # :309-317
    np.testing.assert_almost_equal(cmean[10, 10], 45.0, decimal=1)                 # :319-329
    assert np.isnan(cmean[0, 0])
    allnan = np.full((10, 50, 50), np.nan, dtype=np.float32)
    assert np.all(np.isnan(impl.colmax(allnan))) and np.all(np.isnan(impl.colmin(allnan))) and np.all(np.isnan(impl.colmean(allnan)))
    part = np.ones((10, 50, 50), dtype=np.float32) * 10.0
    part[0:3] = np.nan
    assert np.all(impl.colmax(part) == 10.0) and np.all(impl.colmin(part) == 10.0) and np.all(impl.colmean(part) == 10.0)   # :342-357


@pytest.mark.parametrize("impl", IMPLS)
def test_cappi_and_ppi_contract(impl):
    rng = np.random.default_rng(5)
    d = (rng.random((10, 50, 50)) * 50).astype(np.float32)
    d[0, 0:10, 0:10] = np.nan
    c = impl.cappi(d, (10, 50, 50), PLIM, 5000.0)
    assert c.shape == (50, 50) and c.dtype == np.float32 and not np.all(np.isnan(c))      # :192-207
    assert np.all(np.isnan(impl.cappi(d, (10, 50, 50), PLIM, 15000.0)))                  # :210-217
    assert impl.cappi(d, (10, 50, 50), PLIM, -1000.0).shape == (50, 50)                  # :219-226
    p = impl.ppi(d, (10, 50, 50), PLIM, 2.0)
    assert p.shape == (50, 50) and p.dtype == np.float64                                  # :257-262
    assert impl.ppi(d, (10, 50, 50), PLIM, 0.0).shape == (50, 50)
    assert impl.ppi(d, (10, 50, 50), PLIM, 45.0).shape == (50, 50)
    assert impl.ppi(d, (10, 50, 50), PLIM, 2.0, "nearest").dtype == np.float32


# ---- GridFilter (reference tests/test_grid_filters.py) -----------------------------------------------------------------
@pytest.mark.parametrize("impl", IMPLS)
def test_grid_filter_thresholds(impl):
    g = np.array([[10.0, 20.0, 30.0, 40.0], [15.0, 25.0, 35.0, 45.0], [12.0, 22.0, 32.0, 42.0]])
    keep = g.copy()
    r = impl.gridfilter("below", g, 15)
    assert np.isnan(r[0, 0]) and np.isnan(r[2, 0]) and r[0, 1] == 20.0 and r[1, 0] == 15.0   # :29-40 (15 is kept)
    np.testing.assert_array_equal(g, keep)                                                   # :42-53 input untouched
    r = impl.gridfilter("above", g, 40)
    assert np.isnan(r[1, 3]) and np.isnan(r[2, 3]) and r[0, 3] == 40.0
    r = impl.gridfilter("outside", g, 15, 35, fill_value=-1.0)
    assert r[0, 0] == -1.0 and r[0, 3] == -1.0 and r[1, 0] == 15.0 and r[1, 2] == 35.0
    h = np.array([[10.0, np.nan, 30.0], [15.0, np.inf, 35.0]], dtype=np.float32)
    r = impl.gridfilter("invalid", h, fill_value=-99.0)
    assert r[0, 1] == -99.0 and r[1, 1] == -99.0 and r[0, 0] == 10.0 and r.dtype == np.float32


# ---- RGBA image of a product (geotiff.py:70-145) -------------------------------------------------------------------
def test_oracle_colormap_index_arithmetic_known_answers():
    """Hand-checked indices of the colour table: vmin -> entry 0, vmax -> the LAST entry (x == 1 is not 'over'), values
    outside [vmin, vmax] clip, NaN -> the 'bad' row with alpha 0, fill_value pixels keep their colour but alpha 0."""
    lut = O.colormap_table(4)                       # 4 colours + under / over / bad
    b = (lut * 255).astype(np.uint8)
    data = np.array([[0.0, 10.0, 2.4, 2.5, 7.49, 7.5, -5.0, 50.0, np.nan, 9.999]], dtype=np.float32)
    got = O.apply_colormap(data, lut, vmin=0.0, vmax=10.0)
    want_idx = [0, 3, 0, 1, 2, 3, 0, 3, 6, 3]
    for k, i in enumerate(want_idx):
        exp = b[i].copy()
        if k == 8:
            exp[3] = 0
        np.testing.assert_array_equal(got[0, k], exp, err_msg=f"pixel {k}")
    got = O.apply_colormap(data, lut, vmin=0.0, vmax=10.0, fill_value=10.0)
    assert got[0, 1, 3] == 0 and np.array_equal(got[0, 1, :3], b[3, :3]) and got[0, 8, 3] == b[6, 3]
    np.testing.assert_array_equal(O.apply_colormap(data, lut, vmin=3.0, vmax=3.0)[0, :8], np.broadcast_to(b[0], (8, 4)))
    auto = O.apply_colormap(data, lut)               # vmin = -5, vmax = 50 from the valid data
    assert np.array_equal(auto[0, 6], b[0]) and np.array_equal(auto[0, 7], b[3])
    with pytest.raises(ValueError):
        O.apply_colormap(data, lut, vmin=2.0, vmax=1.0)


@pytest.mark.gpu
def test_cuda_image_epilogue_matches_the_colormap_restatement_and_gridfilter():
    """The fused image epilogue (GridFilter thresholds, then the colormap) on the CUDA path: the same bytes as
    oracle.apply_colormap applied to the oracle's GridFilter output, for float32 planes (COLMAX, CAPPI), the float64
    PPI plane, an explicit fill_value, and an image-only request (no float plane leaves the device)."""
    from conftest import golden_case
    spec, radar, gates, fields, g = golden_case("tiny")
    geom = rg.GridGeometry(spec.grid_shape, spec.grid_limits, g["indptr"], g["gate_indices"], g["weights"], float(g["toa"][0]))
    dev = geom.device_geometry(n_gates=spec.n_gates)
    lut = O.colormap_table(256)
    dbz = fields["DBZH"]
    table = (lut, "with_extremes")                  # N + 3 rows: the colours plus matplotlib's under / over / bad rows
    img = rg.ImageSpec(table, -10.0, 60.0, filters=[("below", 5.0, np.nan), ("above", 45.0, 45.0)])
    img_fill = rg.ImageSpec(table, 0.0, 50.0, fill_value=-999.0, filters=[("invalid", -999.0), ("outside", 0.0, 40.0, -999.0)],
                            keep_plane=False)
    res = rg.grid_fields(dev, [np.ma.getdata(dbz)], masks=[np.ma.getmaskarray(dbz)], reference_order=True,
                         products=[rg.ColumnMax(image=img), rg.CAPPI(1234.5, image=img_fill), rg.PPI(2.3, image=img)])
    colmax, cappi_plane, ppi = res["products"]
    assert cappi_plane is None and res["images"][1].shape == (1,) + spec.grid_shape[1:] + (4,)
    np.testing.assert_array_equal(colmax[0], g["colmax"])
    f = O.grid_filter("above", O.grid_filter("below", g["colmax"], 5.0), 45.0, fill_value=45.0)
    np.testing.assert_array_equal(res["images"][0][0], O.apply_colormap(f, lut, -10.0, 60.0))
    cap = g["cappi_lin_1234.5"]
    f = O.grid_filter("outside", O.grid_filter("invalid", cap, fill_value=-999.0), 0.0, 40.0, fill_value=-999.0)
    np.testing.assert_array_equal(res["images"][1][0], O.apply_colormap(f, lut, 0.0, 50.0, fill_value=-999.0))
    f = O.grid_filter("above", O.grid_filter("below", g["ppi_lin_2.3"], 5.0), 45.0, fill_value=45.0)
    assert ppi.dtype == np.float64
    np.testing.assert_array_equal(res["images"][2][0], O.apply_colormap(f, lut, -10.0, 60.0))
    # the fast fused path gives the same image wherever its plane equals the exact one (thresholds may flip a pixel that
    # sits within the fast path's 1e-5 of a threshold, so compare through the plane it actually produced)
    fast = rg.grid_fields(dev, [np.ma.getdata(dbz)], masks=[np.ma.getmaskarray(dbz)], want_grid=False, products=[rg.ColumnMax(image=img)])
    f = O.grid_filter("above", O.grid_filter("below", fast["products"][0][0], 5.0), 45.0, fill_value=45.0)
    np.testing.assert_array_equal(fast["images"][0][0], O.apply_colormap(f, lut, -10.0, 60.0))
    # the reference-shaped function on a finished plane, default vmin / vmax
    np.testing.assert_array_equal(rg.apply_colormap_to_array(g["colmax"], table), O.apply_colormap(g["colmax"], lut))
    np.testing.assert_array_equal(rg.apply_colormap_to_array(g["colmax"], lut[:256]), O.apply_colormap(g["colmax"], lut))   # default extremes
    np.testing.assert_array_equal(rg.apply_colormap_to_array(g["colmax"], table, vmin=0.0, vmax=30.0, fill_value=float(np.nanmax(g["colmax"]))),
                                  O.apply_colormap(g["colmax"], lut, 0.0, 30.0, fill_value=float(np.nanmax(g["colmax"]))))
    oob = rg.grid_fields(dev, [np.ma.getdata(dbz)], masks=[np.ma.getmaskarray(dbz)], want_grid=False,
                         products=[rg.CAPPI(1e6, image=img)])                 # outside the grid: all no-data
    np.testing.assert_array_equal(oob["images"][0][0], O.apply_colormap(np.full(spec.grid_shape[1:], np.nan, np.float32), lut, -10.0, 60.0))
