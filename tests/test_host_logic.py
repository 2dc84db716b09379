"""
Host-side logic of the Python mirror (no GPU): GateFilter vs the reference's fixtures, GridGeometry container and
.npz round trip (reference tests/test_radar_grid_geometry.py), product request resolution (which level, which
weights, which dtype), synthetic-data determinism, and the rule that the product never touches the oracle.
"""
import os
import re

import numpy as np
import pytest

import radar_grid_b200 as rg
from conftest import PKG_ROOT, ROOT, assert_same, golden_case, load_golden
from radar_grid_b200 import _native as N
from radar_grid_b200 import synthetic as S


# ---- GateFilter (reference src/radar_grid/filters.py) against outputs of the real reference ---------------------
def test_gatefilter_masks_match_reference_fixture():
    spec, radar, gates, fields, _ = golden_case("tiny")
    f = load_golden("ref_filters_tiny.npz")
    gf = rg.GateFilter(radar)
    assert_same(gf.copy().exclude_below("DBZH", 5.0).gate_excluded, f["below_DBZH_5"])
    assert_same(gf.copy().exclude_above("ZDR", 1.0).gate_excluded, f["above_ZDR_1"])
    assert_same(gf.copy().exclude_outside("RHOHV", 0.8, 0.95).gate_excluded, f["outside_RHOHV"])
    assert_same(gf.copy().exclude_between("KDP", 0.0, 0.5).gate_excluded, f["between_KDP"])
    assert_same(gf.copy().exclude_equal("VRAD", 0.0, atol=1.0).gate_excluded, f["equal_VRAD"])
    assert_same(gf.copy().exclude_invalid("DBZH").gate_excluded, f["invalid_DBZH"])
    assert_same(gf.copy().exclude_masked("DBZH").gate_excluded, f["masked_DBZH"])
    assert_same(gf.copy().exclude_all_invalid("DBZH").gate_excluded, f["all_invalid_DBZH"])
    assert_same(gf.copy().exclude_below_altitude(2000.0).gate_excluded, f["below_alt_2000"])
    assert_same(gf.copy().exclude_above_altitude(5000.0).gate_excluded, f["above_alt_5000"])
    assert_same(gf.copy().exclude_below_range(3000.0).gate_excluded, f["below_range"])
    assert_same(gf.copy().exclude_above_range(15000.0).gate_excluded, f["above_range"])
    assert_same(gf.copy().exclude_below_elevation_angle(2.0).gate_excluded, f["below_elev"])
    assert_same(gf.copy().exclude_above_elevation_angle(10.0).gate_excluded, f["above_elev"])
    assert_same(gf.copy().exclude_outside_elevation_range(1.0, 10.0).gate_excluded, f["outside_elev"])
    chained = gf.copy().exclude_below("RHOHV", 0.8).exclude_above("RHOHV", 1.0).exclude_below("DBZH", 0.0)
    assert_same(chained.gate_excluded, f["chained"])
    d, m = rg.create_mask_from_filter(radar, "DBZH", chained)
    assert_same(d, f["cmff_data"])
    assert_same(m, f["cmff_mask"])
    assert chained.fusable_rules() == [("RHOHV", 0.8, None), ("RHOHV", None, 1.0), ("DBZH", 0.0, None)]
    assert gf.copy().exclude_below("DBZH", 1.0).exclude_invalid("DBZH").fusable_rules() is None


def test_gatefilter_bookkeeping_like_reference_tests():
    spec, radar, *_ = golden_case("tiny")
    gf = rg.GateFilter(radar)
    assert gf.n_gates == radar.nrays * radar.ngates and gf.n_excluded() == 0 and gf.n_included() == gf.n_gates
    gf.exclude_below("NOPE", 1.0)                        # missing field: warning, no-op (filters.py:130-132)
    assert gf.n_excluded() == 0 and gf._filter_history == []
    gf.exclude_below("DBZH", 10.0).exclude_above("DBZH", 40.0)
    assert "GateFilter(excluded=" in repr(gf) and "Filters applied (2)" in gf.summary()
    assert np.array_equal(gf.gate_included, ~gf.gate_excluded)
    with pytest.raises(ValueError):
        gf.exclude_where(np.zeros(3, dtype=bool))
    assert gf.copy().reset().n_excluded() == 0
    assert gf.copy().exclude_all().n_included() == 0


# ---- GridGeometry container (reference tests/test_radar_grid_geometry.py) -----------------------------------------
def make_geometry(n=1000, shape=(10, 10, 10), **kw):
    return rg.GridGeometry(grid_shape=shape, grid_limits=((0, 1000), (-500, 500), (-500, 500)),
                           indptr=np.arange(n + 1, dtype=np.int32), gate_indices=np.zeros(n, dtype=np.int32),
                           weights=np.ones(n, dtype=np.float32), toa=12000.0, **kw)


def test_geometry_helpers():
    g = make_geometry(radar_altitude=100.0)
    assert g.memory_usage_mb() == pytest.approx((1001 * 4 + 1000 * 4 + 1000 * 4) / 1e6)
    assert g.n_grid_points() == 1000 and g.n_pairs() == 1000 and g.avg_neighbors() == 1.0
    z = g.z_levels()
    assert z.shape == (10,) and z[0] == 0 and z[-1] == 1000
    np.testing.assert_array_equal(g.z_levels_absolute(), z + 100.0)
    r = repr(g)
    for token in ("GridGeometry(", "grid_shape=(10, 10, 10)", "toa=12000.0m", "radar_altitude=100.0m", "n_pairs=1,000"):
        assert token in r


def test_npz_round_trip_and_reference_key_set(tmp_path):
    g = make_geometry(radar_altitude=250.0)
    path = str(tmp_path / "geom.npz")
    rg.save_geometry(g, path)
    with np.load(path) as z:
        assert sorted(z.files) == sorted(["grid_shape", "grid_limits_z", "grid_limits_y", "grid_limits_x", "indptr",
                                          "gate_indices", "weights", "toa", "radar_altitude"])
        assert z["indptr"].dtype == np.int32 and z["weights"].dtype == np.float32
    h = rg.load_geometry(path)
    assert tuple(h.grid_shape) == g.grid_shape and h.toa == 12000.0 and h.radar_altitude == 250.0
    for k in ("indptr", "gate_indices", "weights"):
        np.testing.assert_array_equal(getattr(h, k), getattr(g, k))
    assert h.grid_limits == ((0, 1000), (-500, 500), (-500, 500))


def test_legacy_npz_without_altitude_and_toa(tmp_path):
    g = make_geometry()
    path = str(tmp_path / "legacy.npz")
    np.savez_compressed(path, grid_shape=np.array(g.grid_shape), grid_limits_z=np.array(g.grid_limits[0]),
                        grid_limits_y=np.array(g.grid_limits[1]), grid_limits_x=np.array(g.grid_limits[2]),
                        indptr=g.indptr, gate_indices=g.gate_indices, weights=g.weights)
    h = rg.load_geometry(path)
    assert h.radar_altitude == 0.0 and h.toa == np.inf


# ---- product request resolution ----------------------------------------------------------------------------------------
LIM = ((0.0, 10000.0), (-25000.0, 25000.0), (-25000.0, 25000.0))
SHAPE = (10, 50, 50)


def test_cappi_resolution_cases():
    pr, dt = rg.CAPPI(15000.0).resolve(SHAPE, LIM)
    assert pr is None and dt == np.float32                                        # outside -> NaN plane
    pr, _ = rg.CAPPI(0.0).resolve(SHAPE, LIM)
    assert (pr.mode, pr.z_lo) == (N.RG_BLEND_PICK, 0)                             # exact level
    pr, _ = rg.CAPPI(10000.0).resolve(SHAPE, LIM)
    assert (pr.mode, pr.z_lo) == (N.RG_BLEND_PICK, 9)
    pr, _ = rg.CAPPI(5000.0).resolve(SHAPE, LIM)                                  # between 4444.4 and 5555.6
    assert (pr.mode, pr.z_lo, pr.z_hi) == (N.RG_BLEND_F32, 4, 5)
    assert pr.w_lo + pr.w_hi == pytest.approx(1.0) and pr.w_hi == pytest.approx(0.5)
    pr, _ = rg.CAPPI(5000.0, "nearest").resolve(SHAPE, LIM)
    assert pr.mode == N.RG_BLEND_PICK and pr.z_lo in (4, 5)
    lim64 = tuple(tuple(np.float64(v) for v in l) for l in LIM)
    pr, _ = rg.CAPPI(5000.0).resolve(SHAPE, lim64)                                # np.float64 limits -> float64 blend
    assert pr.mode == N.RG_BLEND_F64
    with pytest.raises(ValueError):
        rg.CAPPI(5000.0, "cubic").resolve(SHAPE, LIM)


def test_ppi_and_column_resolution():
    pr, dt = rg.PPI(2.0).resolve(SHAPE, LIM)
    assert dt == np.float64 and pr.kind == N.RG_PROD_BEAM and pr.mode == 0 and pr.earth_curvature == 1
    assert pr.sin_elev == float(np.sin(np.radians(2.0))) and pr.ke_re == 4.0 / 3.0 * 6371000.0
    pr, dt = rg.PPI(89.9, "nearest", earth_curvature=False).resolve(SHAPE, LIM)
    assert dt == np.float32 and pr.mode == 1 and pr.cos_elev_clamped == 0.01
    with pytest.raises(ValueError):
        rg.PPI(1.0, "spline").resolve(SHAPE, LIM)
    pr, _ = rg.ColumnMax().resolve(SHAPE, LIM)
    assert (pr.z_lo, pr.z_hi) == (0, 9)
    pr, _ = rg.ColumnMax(z_min_idx=-3, z_max_idx=99).resolve(SHAPE, LIM)
    assert (pr.z_lo, pr.z_hi) == (0, 9)
    pr, _ = rg.ColumnMin(z_min_alt=1000.0, z_max_alt=8000.0).resolve(SHAPE, LIM)
    zc = np.linspace(0.0, 10000.0, 10)
    assert (pr.z_lo, pr.z_hi) == (int(np.searchsorted(zc, 1000.0)), int(np.searchsorted(zc, 8000.0, side="right") - 1))
    with pytest.raises(ValueError):
        rg.ColumnMean(z_min_alt=1000.0).resolve(SHAPE, LIM, have_geometry=False)


def test_beam_height_helpers_known_answers():
    """reference tests/test_radar_grid_products.py:28-91."""
    d = np.array([10000.0, 20000.0, 50000.0])
    h = rg.compute_beam_height(d, 2.0, 100.0)
    assert h.shape == d.shape and np.all(h > 100.0) and np.all(np.diff(h) > 0)
    assert rg.compute_beam_height(np.array([10000.0]), 45.0, 0.0)[0] > 7000.0
    assert np.all(np.diff(rg.compute_beam_height(np.array([10000.0, 20000.0]), 0.0, 0.0)) > 0)
    assert np.all(rg.compute_beam_height_simple(d, 2.0, 100.0) > 100.0)
    hf = rg.compute_beam_height_flat(np.array([10000.0, 20000.0]), 2.0, 100.0)
    np.testing.assert_almost_equal(hf[1] - hf[0], 10000.0 * np.tan(np.radians(2.0)), decimal=1)
    g = rg.GridGeometry(SHAPE, LIM, np.arange(25001, dtype=np.int32), np.zeros(25000, np.int32), np.ones(25000, np.float32),
                        12000.0, 100.0)
    diff = rg.get_beam_height_difference(g, 2.0)
    assert diff.dtype == np.float64 and np.isfinite(diff).all()
    assert np.isfinite(rg.get_elevation_from_z_level(5, g)).all()
    assert rg.EARTH_RADIUS == 6371000.0 and rg.EFFECTIVE_RADIUS_FACTOR == 4.0 / 3.0


# ---- synthetic data & hygiene ----------------------------------------------------------------------------------------
def test_synthetic_volumes_are_deterministic_and_well_formed():
    spec = S.SPECS["small"]
    a, b = S.gate_coordinates(spec), S.gate_coordinates(spec)
    for x, y in zip(a, b):
        assert x.dtype == np.float32 and np.array_equal(x, y) and x.shape == (spec.n_gates,)
    fa, fb = S.make_fields(spec, 3), S.make_fields(spec, 3)
    for k in spec.fields:
        assert_same(np.ma.getmaskarray(fa[k]), np.ma.getmaskarray(fb[k]))
        assert_same(np.ma.getdata(fa[k]), np.ma.getdata(fb[k]))
        assert fa[k].dtype == np.float32
    assert not np.array_equal(np.ma.getdata(S.make_fields(spec, 4)["ZDR"]), np.ma.getdata(fa["ZDR"]))
    # gate id = (sweep*nrays + ray)*ngates + bin; first gate of ray 0 points north (+y)
    assert abs(a[0][0]) < 1e-3 and a[1][0] > 0
    assert S.CFG3.n_gates == 5_400_000 and S.CFG3.n_voxels == 9_254_440 and S.CFG1.n_voxels == 1_161_620


def test_product_code_never_imports_the_oracle_or_reads_the_reference():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|/root/reference", re.M)
    for base, _, files in os.walk(PKG_ROOT):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(base, f)).read()
                assert not pat.search(text), f"{f} references the oracle / the reference tree"
    hdr = open(os.path.join(ROOT, "include", "radar_grid_b200.h")).read()
    assert "NO CPU fallback" in hdr


def test_geometry_cache_keys_and_lru_eviction():
    """GeometryCache: same inputs -> one build; any change of gates / grid / ROI parameters -> a new table; LRU eviction."""
    built, closed = [], []

    class FakeGeom:
        def __init__(self, n):
            self.info = {"device_bytes": n}
        def close(self):
            closed.append(self)

    def builder(gx, gy, gz, shape, limits, **params):
        g = FakeGeom(100)
        built.append(g)
        return g

    cache = rg.GeometryCache(max_bytes=250, builder=builder)
    gx, gy, gz = (np.arange(10, dtype=np.float32) + k for k in range(3))
    a = cache.get(gx, gy, gz, (2, 3, 4), LIM, min_radius=250.0, weighting="barnes2")
    assert cache.get(gx, gy, gz, (2, 3, 4), LIM, min_radius=250.0, weighting="barnes2") is a
    assert (cache.hits, cache.misses, len(built)) == (1, 1, 1)
    b = cache.get(gx, gy, gz, (2, 3, 4), LIM, min_radius=300.0, weighting="barnes2")          # ROI parameter changed
    c = cache.get(gx + 1, gy, gz, (2, 3, 4), LIM, min_radius=250.0, weighting="barnes2")      # gates changed
    assert len({id(a), id(b), id(c)}) == 3 and len(built) == 3
    assert closed == [] and cache.bytes_held() == 200      # 300 > 250: the oldest is dropped (freed with its last user), never closed under one
    assert cache.get(gx, gy, gz, (2, 3, 5), LIM, min_radius=250.0, weighting="barnes2") is not a   # grid changed
    cache.clear()
    assert cache.bytes_held() == 0


def test_zslab_product_plans():
    """Host half of the z-slab CAPPI / PPI (distributed.py): which slab owns which level, with which weight."""
    from radar_grid_b200 import distributed as D
    shape, limits = (10, 5, 7), ((0.0, 9000.0), (-2000.0, 2000.0), (-3000.0, 3000.0))
    # linear blend between levels 3 and 4, slabs [0,4) and [4,10): one term each, weights of the unsharded resolve
    pr, _ = rg.CAPPI(3250.0).resolve(shape, limits)
    lo, dt = D.cappi_zslab_terms(rg.CAPPI(3250.0), shape, limits, (0, 4))
    hi, _ = D.cappi_zslab_terms(rg.CAPPI(3250.0), shape, limits, (4, 10))
    assert dt == np.float32 and lo == [(3, pr.w_lo)] and hi == [(4, pr.w_hi)]
    assert D.cappi_zslab_terms(rg.CAPPI(3250.0), shape, limits, (5, 10))[0] == []
    # exact level and nearest: a single weight-1 pick on the owner only
    assert D.cappi_zslab_terms(rg.CAPPI(4000.0), shape, limits, (4, 10))[0] == [(4, 1.0)]
    assert D.cappi_zslab_terms(rg.CAPPI(4000.0), shape, limits, (0, 4))[0] == []
    assert D.cappi_zslab_terms(rg.CAPPI(3600.0, "nearest"), shape, limits, (4, 10))[0] == [(4, 1.0)]
    # NumPy-scalar limits promote the blend to float64 (as in CAPPI.resolve); outside the grid: no plan at all
    lims64 = tuple(tuple(np.float64(v) for v in ax) for ax in limits)
    assert D.cappi_zslab_terms(rg.CAPPI(3250.0), shape, lims64, (0, 10))[1] == np.float64
    assert D.cappi_zslab_terms(rg.CAPPI(9500.0), shape, limits, (0, 10)) is None
    with pytest.raises(ValueError):
        D.cappi_zslab_terms(rg.CAPPI(3250.0, "cubic"), shape, limits, (0, 10))
    with pytest.raises(ValueError):
        rg.LevelPick(10).resolve(shape, limits)
    assert rg.LevelPick(9).resolve(shape, limits)[0].z_lo == 9
    # PPI plan: shapes, dtypes, weights sum to one, clipped levels
    plan = D.ppi_zslab_plan(rg.PPI(20.0), shape, limits)
    assert plan["mode"] == "linear" and plan["lo"].shape == (5, 7) and plan["w_lo"].dtype == np.float64
    np.testing.assert_allclose(plan["w_lo"] + plan["w_hi"], 1.0, rtol=0, atol=1e-15)
    assert plan["lo"].min() >= 0 and plan["hi"].max() <= 9 and plan["nan"].dtype == np.bool_
    near = D.ppi_zslab_plan(rg.PPI(89.0, "nearest"), shape, limits)
    assert near["mode"] == "nearest" and near["nan"].any()          # a near-vertical beam leaves the grid top
    with pytest.raises(ValueError):
        D.ppi_zslab_plan(rg.PPI(1.0, "cubic"), shape, limits)
