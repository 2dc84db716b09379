"""
Parity of the BASELINE.json configurations that round 1 only sampled or only ran as example scripts:

  * apply_geometry_multi (reference interpolate.py:107-142): the reference's own two tests
    (tests/test_radar_grid_interpolate.py:171-212) replayed, and the five-field golden made by the real function;
  * cfg1 / cfg2 at FULL size against the table the real reference built (tests/golden/ref_cfg1_digest.npz): every row
    length, the whole set of gate ids (SHA-256 of the row-sorted ids), weights within 1 ulp, and the reference's own
    COLMAX / CAPPI / PPI planes;
  * cfg4 (256-volume time series; 16 seeds here) through VolumePipeline against the oracle;
  * cfg5 (80 x 2001 x 2001) z-slabs: sampled columns of three slabs against a brute-force scan of ALL gates, and the
    slab COLMAX against the oracle applied to those columns' rows.
"""
import hashlib
import warnings
from unittest.mock import Mock

import numpy as np
import pytest

import radar_grid_b200 as rg
from radar_grid_b200 import synthetic as S
from oracle import radar_grid_oracle as O
from conftest import load_golden

pytestmark = pytest.mark.gpu

ATOL, RTOL = 1e-4, 1e-5                      # north_star: 1e-4 dBZ absolute / 1e-5 relative, identical mask


def close_same_mask(a, b, what=""):
    a, b = np.asarray(a), np.asarray(b)
    np.testing.assert_array_equal(np.isnan(a), np.isnan(b), err_msg=f"{what}: NaN mask")
    np.testing.assert_allclose(a, b, rtol=RTOL, atol=ATOL, equal_nan=True, err_msg=what)


# ---- apply_geometry_multi ------------------------------------------------------------------------------------------
def simple_geometry():
    return rg.GridGeometry(grid_shape=(2, 2, 2), grid_limits=((0, 1000), (-500, 500), (-500, 500)),
                           indptr=np.arange(0, 17, 2, dtype=np.int32), gate_indices=np.arange(16, dtype=np.int32),
                           weights=np.ones(16, dtype=np.float32), toa=2000.0)


def test_apply_geometry_multi_replays_the_reference_tests():
    fields = {k: np.ma.masked_invalid(np.ones(16, dtype=np.float32) * v)
              for k, v in (("DBZH", 10.0), ("ZDR", 2.0), ("RHOHV", 0.95))}
    res = rg.apply_geometry_multi(simple_geometry(), fields)                 # test_radar_grid_interpolate.py:171-190
    assert isinstance(res, dict) and list(res) == ["DBZH", "ZDR", "RHOHV"]
    for name, grid in res.items():
        assert grid.shape == (2, 2, 2) and grid.dtype == np.float32
        np.testing.assert_array_equal(grid, np.float32({"DBZH": 10.0, "ZDR": 2.0, "RHOHV": 0.95}[name]))
    radar = Mock()                                                           # :192-212 — a bare GateFilter per field
    radar.nrays = radar.ngates = 4
    radar.fields = {"DBZH": {"data": np.ones((4, 4), dtype=np.float32) * 10.0}}
    gf = rg.GateFilter(radar)
    gf.exclude_below("DBZH", 5.0)
    two = {k: fields[k] for k in ("DBZH", "ZDR")}
    res = rg.apply_geometry_multi(simple_geometry(), two, additional_filters={"DBZH": gf})
    assert len(res) == 2 and np.all(res["DBZH"] == 10.0) and np.all(res["ZDR"] == 2.0)
    gf.exclude_below("DBZH", 50.0)                                           # now every gate of DBZH is excluded
    res = rg.apply_geometry_multi(simple_geometry(), two, additional_filters={"DBZH": [gf]}, fill_value=-1.0)
    assert np.all(res["DBZH"] == -1.0) and np.all(res["ZDR"] == 2.0)
    with pytest.raises(ValueError):
        rg.apply_geometry_multi(simple_geometry(), two, additional_filters={"DBZH": "nope"})


def test_apply_geometry_multi_matches_the_reference_golden():
    spec = S.SPECS["small"]
    radar = S.SyntheticRadar(spec, seed=7)
    z, want = load_golden("ref_small_barnes2_alt0.npz"), load_golden("ref_multi_small.npz")
    geom = rg.GridGeometry(spec.grid_shape, spec.grid_limits, z["indptr"], z["gate_indices"], z["weights"], float(z["toa"][0]))
    fields = {name: rg.get_field_data(radar, name) for name in spec.fields}
    got = rg.apply_geometry_multi(geom, fields)
    assert list(got) == list(spec.fields)
    for name in spec.fields:
        assert got[name].dtype == np.float32 and got[name].shape == spec.grid_shape
        close_same_mask(got[name], want[f"plain_{name}"], f"plain {name}")
        # one field at a time gives the same bits as the shared pass (only the number of fields per index load differs)
        np.testing.assert_array_equal(got[name], rg.apply_geometry(geom, fields[name]), err_msg=name)
    gf_rho = rg.GateFilter(radar).exclude_below("RHOHV", 0.8).exclude_above("RHOHV", 1.0)
    gf_dbz = rg.GateFilter(radar).exclude_below("DBZH", 5.0)
    got = rg.apply_geometry_multi(geom, fields, additional_filters={"DBZH": [gf_rho, gf_dbz], "ZDR": gf_rho}, fill_value=-5.0)
    for name in spec.fields:
        want_g = want[f"filt_{name}"]
        np.testing.assert_array_equal(got[name] == -5.0, want_g == -5.0, err_msg=f"filtered {name}: fill positions")
        close_same_mask(got[name], want_g, f"filtered {name}")


# ---- cfg1 / cfg2 at full size against the reference's own table ------------------------------------------------------
@pytest.fixture(scope="module")
def cfg1_case():
    spec = S.SPECS["cfg2"]
    radar = S.SyntheticRadar(spec, seed=1)
    gates = rg.get_gate_coordinates(radar)
    dev = rg.DeviceGeometry.build(*gates, spec.grid_shape, spec.grid_limits, min_radius=spec.min_radius,
                                  beam_factor=spec.beam_factor, weighting=spec.weighting, toa=spec.toa)
    return spec, radar, dev, load_golden("ref_cfg1_digest.npz")


def test_cfg1_whole_table_equals_the_reference_table(cfg1_case):
    spec, radar, dev, ref = cfg1_case
    nz, ny, nx = spec.grid_shape
    indptr, idx, w = dev.export_csr()
    assert int(indptr[-1]) == int(ref["n_pairs"][0])
    np.testing.assert_array_equal(np.diff(indptr.astype(np.int64)), ref["row_len"].astype(np.int64), err_msg="row lengths")
    lens = np.diff(indptr.astype(np.int64))
    row_of = np.repeat(np.arange(len(lens), dtype=np.int64), lens)
    order = np.lexsort((idx, row_of))
    sha = hashlib.sha256(np.ascontiguousarray(idx[order], dtype=np.int32).tobytes()).hexdigest()
    assert sha == str(ref["gate_ids_sha256"]), "neighbour sets differ from the reference's (all 1.16 M rows)"
    # weights: float64 formula rounded once to float32 on both sides; CUDA exp and libm differ by <= 1 ulp before the
    # rounding, so a few pairs land on the neighbouring float32.  Per column the bit-pattern sums may differ by that count.
    wsum = np.zeros(ny * nx, dtype=np.int64)
    np.add.at(wsum, row_of % (ny * nx), w.view(np.int32).astype(np.int64))
    diff = np.abs(wsum - ref["weight_bits_colsum"])
    assert diff.max() <= 16 and diff.sum() <= 1e-4 * len(w), (int(diff.max()), int(diff.sum()))


def test_cfg1_cfg2_products_equal_the_reference_planes(cfg1_case):
    spec, radar, dev, ref = cfg1_case
    dbz = rg.get_field_data(radar, "DBZH")
    data, mask = np.ma.getdata(dbz), np.ma.getmaskarray(dbz)
    rho = rg.GateFilter(radar)._get_field_data("RHOHV")
    plain = rg.grid_fields(dev, [data], masks=[mask], products=[rg.ColumnMax()])
    assert int(np.isnan(plain["grids"][0]).sum()) == int(ref["grid_nan_count"][0])
    close_same_mask(plain["products"][0][0], ref["colmax"], "cfg1 COLMAX")
    rules = [rg.RangeRule(rho, lo=0.8), rg.RangeRule(rho, hi=1.0)]           # cfg2: RHOHV 0.8-1.0 QC as a fused gate mask
    qc = rg.grid_fields(dev, [data], masks=[mask], rules=rules, products=[rg.ColumnMax(), rg.CAPPI(4000.0), rg.PPI(0.5)])
    assert int(np.isnan(qc["grids"][0]).sum()) == int(ref["grid_nan_count"][1])
    close_same_mask(qc["products"][0][0], ref["colmax_qc"], "cfg2 COLMAX")
    close_same_mask(qc["products"][1][0], ref["cappi_4000_qc"], "cfg2 CAPPI 4000 m")
    assert qc["products"][2].dtype == np.float64
    close_same_mask(qc["products"][2][0], ref["ppi_0.5_qc"], "cfg2 PPI 0.5 deg")
    # the same request through the reference-shaped API (GateFilter object, not RangeRule)
    gf = rg.GateFilter(radar).exclude_below("RHOHV", 0.8).exclude_above("RHOHV", 1.0)
    grid = rg.apply_geometry(rg.GridGeometry(spec.grid_shape, spec.grid_limits, None, None, None, spec.toa,
                                             n_gates=dev.n_gates, _device=dev), dbz, additional_filters=[gf])
    np.testing.assert_array_equal(grid, qc["grids"][0])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        close_same_mask(rg.column_max(grid), ref["colmax_qc"], "column_max(apply_geometry(...))")


# ---- cfg4: time series through VolumePipeline -------------------------------------------------------------------------
def test_cfg4_time_series_products_against_the_oracle(cfg1_case):
    spec = S.SPECS["cfg1"]
    _, _, dev, _ = cfg1_case
    gates = S.gate_coordinates(spec)
    indptr, idx, w = dev.export_csr()
    reqs = [rg.ColumnMax(), rg.CAPPI(4000.0)]
    vols = [S.make_fields(spec, seed=s, gates=gates)["DBZH"] for s in range(16)]
    jobs = [{"fields": [np.ma.getdata(v)], "masks": [np.ma.getmaskarray(v)], "products": reqs, "want_grid": False} for v in vols]
    pipe = rg.VolumePipeline(dev, n_streams=3)
    try:
        out = pipe.map(jobs)
    finally:
        pipe.close()
    for s, (v, res) in enumerate(zip(vols, out)):
        assert res["grids"] == [None]
        grid = O.apply_geometry(indptr, idx, w, spec.grid_shape, v)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            close_same_mask(res["products"][0][0], O.column_reduce("max", grid), f"seed {s} COLMAX")
        close_same_mask(res["products"][1][0], O.cappi(grid, spec.grid_shape, spec.grid_limits, 4000.0), f"seed {s} CAPPI")
        if s % 5 == 0:          # the pipelined call equals the synchronous one bit for bit
            sync = rg.grid_fields(dev, **jobs[s])
            for a, b in zip(res["products"], sync["products"]):
                np.testing.assert_array_equal(a, b)


# ---- cfg5: z-slabs of the large domain ---------------------------------------------------------------------------------
@pytest.mark.parametrize("z_range", [(0, 2), (39, 41), (78, 80)])
def test_cfg5_slab_rows_against_bruteforce_and_colmax_against_the_oracle(z_range):
    spec = S.SPECS["cfg5"]
    nz, ny, nx = spec.grid_shape
    gates = S.gate_coordinates(spec)
    dev = rg.DeviceGeometry.build(*gates, spec.grid_shape, spec.grid_limits, min_radius=spec.min_radius,
                                  beam_factor=spec.beam_factor, weighting=spec.weighting, toa=spec.toa, z_range=z_range)
    nlev = z_range[1] - z_range[0]
    assert dev.n_rows == nlev * ny * nx
    indptr, idx, w = dev.export_csr()
    z_ax, y_ax, x_ax = O.grid_axes(spec.grid_shape, spec.grid_limits)
    rng = np.random.default_rng(z_range[0] + 17)
    cols = [(ny // 2, nx // 2), (ny // 2, nx // 2 + 1), (ny // 2 + 3, nx // 2 - 2), (0, 0), (0, nx - 1), (ny - 1, 0),
            (ny - 1, nx - 1), (ny // 2, 0), (0, nx // 2)]                       # radar origin column, corners, edges
    cols += [(int(rng.integers(0, ny)), int(rng.integers(0, nx))) for _ in range(100 - len(cols))]
    sub_ptr, sub_idx, sub_w = [0], [], []
    for lz in range(nlev):
        for iy, ix in cols:
            row = (lz * ny + iy) * nx + ix
            s, e = int(indptr[row]), int(indptr[row + 1])
            ids, ww = O.neighbours_bruteforce(*gates, (x_ax[ix], y_ax[iy], z_ax[z_range[0] + lz]), min_radius=spec.min_radius,
                                              beam_factor=spec.beam_factor, weighting=spec.weighting, toa=spec.toa)
            order = np.argsort(idx[s:e], kind="stable")
            np.testing.assert_array_equal(idx[s:e][order], ids, err_msg=f"slab {z_range} level {lz} column {(iy, ix)}")
            ulp = np.abs(w[s:e][order].view(np.int32).astype(np.int64) - ww.view(np.int32).astype(np.int64))
            assert ulp.max(initial=0) <= 1
            sub_idx.append(idx[s:e]); sub_w.append(w[s:e]); sub_ptr.append(sub_ptr[-1] + e - s)
    # slab COLMAX of the sampled columns: oracle on exactly those rows (the table as the GPU stores it)
    field = S.make_fields(spec, seed=2, gates=gates)["DBZH"]
    grid = O.apply_geometry(np.asarray(sub_ptr, dtype=np.int64), np.concatenate(sub_idx), np.concatenate(sub_w),
                            (nlev, 1, len(cols)), field)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore", RuntimeWarning)
        want = O.column_reduce("max", grid)[0]
    data, mask = np.ma.getdata(field), np.ma.getmaskarray(field)
    fast = rg.grid_fields(dev, [data], masks=[mask], products=[rg.ColumnMax()], want_grid=False)["products"][0][0]
    exact = rg.grid_fields(dev, [data], masks=[mask], products=[rg.ColumnMax()], want_grid=False,
                           reference_order=True)["products"][0][0]
    iy, ix = np.array(cols).T
    np.testing.assert_array_equal(exact[iy, ix], want, err_msg=f"slab {z_range}: reference-order COLMAX")
    close_same_mask(fast[iy, ix], want, f"slab {z_range}: fused COLMAX")
    dev.close()


# ---- GateFilter fusion and the geometry cache through the reference-shaped API ---------------------------------------
def test_gatefilter_range_rules_are_fused_on_the_device_through_apply_geometry(monkeypatch):
    from radar_grid_b200 import interpolate as I
    spec = S.SPECS["small"]
    radar = S.SyntheticRadar(spec, seed=7)
    z = load_golden("ref_small_barnes2_alt0.npz")
    geom = rg.GridGeometry(spec.grid_shape, spec.grid_limits, z["indptr"], z["gate_indices"], z["weights"], float(z["toa"][0]))
    fields = {name: rg.get_field_data(radar, name) for name in spec.fields}
    seen = []
    real = I.grid_fields

    def spy(dev, datas, **kw):
        seen.append((datas, kw))
        return real(dev, datas, **kw)

    monkeypatch.setattr(I, "grid_fields", spy)
    gf = rg.GateFilter(radar).exclude_below("RHOHV", 0.8).exclude_above("RHOHV", 1.0)
    fused = rg.apply_geometry(geom, fields["DBZH"], additional_filters=[gf])
    assert gf._pending, "the host mask of a range rule must not be computed by apply_geometry"
    datas, kw = seen[-1]
    assert len(kw["rules"]) == 2 and kw["rules"][0].lo == 0.8 and kw["rules"][1].hi == 1.0
    np.testing.assert_array_equal(kw["masks"][0], np.ma.getmaskarray(fields["DBZH"]))     # only the field's own mask
    # against the reference's output for exactly this call, and bit-identical to the host-mask route
    close_same_mask(fused, z["gridqc_DBZH"], "fused QC filter vs reference")
    opaque = rg.GateFilter(radar).exclude_where(gf.gate_excluded.copy())
    assert opaque.fusable_rules() is None
    np.testing.assert_array_equal(fused, rg.apply_geometry(geom, fields["DBZH"], additional_filters=[opaque]))
    # multi: the rule's field is one of the gridded fields -> its device copy is used, nothing extra is uploaded
    gf2 = rg.GateFilter(radar).exclude_outside("RHOHV", 0.8, 1.0)
    multi = rg.apply_geometry_multi(geom, fields, additional_filters={"DBZH": [gf2], "ZDR": gf2})
    datas, kw = seen[-1]
    assert len(kw["rules"]) == 1 and kw["rules"][0].fields == [0, 1]
    assert any(kw["rules"][0].values is np.asarray(d).ravel() or np.shares_memory(kw["rules"][0].values, d) for d in datas)
    np.testing.assert_array_equal(multi["DBZH"], fused)
    np.testing.assert_array_equal(multi["KDP"], rg.apply_geometry(geom, fields["KDP"]))


def test_compute_grid_geometry_builds_a_table_once(tmp_path):
    from radar_grid_b200 import _native as N
    from radar_grid_b200.compute import geometry_cache
    spec = S.SPECS["small"]
    gates = S.gate_coordinates(spec)
    cache = geometry_cache()
    cache.clear()
    kw = dict(min_radius=spec.min_radius, beam_factor=spec.beam_factor, weighting=spec.weighting, toa=spec.toa)
    a = rg.compute_grid_geometry(*gates, spec.grid_shape, spec.grid_limits, str(tmp_path), **kw)
    launches = N.default_context().kernel_launches()
    hits = cache.hits
    b = rg.compute_grid_geometry(*[g.copy() for g in gates], spec.grid_shape, spec.grid_limits, str(tmp_path), **kw)
    assert N.default_context().kernel_launches() == launches, "second call with the same gates must not launch a build"
    assert cache.hits == hits + 1 and b is not a and b.device_geometry() is a.device_geometry()
    c = rg.compute_grid_geometry(*gates, spec.grid_shape, spec.grid_limits, str(tmp_path), **{**kw, "min_radius": spec.min_radius + 1})
    assert c.device_geometry() is not a.device_geometry()
    np.testing.assert_array_equal(a.indptr, b.indptr)
    cache.clear()


def test_gate_coordinates_on_the_device_match_the_float64_transform():
    """rg_gate_coordinates: pyart's 4/3-earth antenna_to_cartesian from (range, azimuth, elevation) on the device, in
    float64, rounded once to float32 -- against the same formula in NumPy on the same float32 inputs (CUDA and libm
    sin / cos / asin differ by <= 1 ulp in float64, so a float32 result may land on the neighbouring value: a handful
    per million at most), and the table built from the device coordinates against the one built from the host ones."""
    import torch
    spec = S.SPECS["cfg1"]
    radar = S.SyntheticRadar(spec, seed=0)
    gx, gy, gz = rg.get_gate_coordinates_device(radar)
    assert gx.is_cuda and gx.dtype == torch.float32 and gx.numel() == spec.n_gates
    R = 4.0 / 3.0 * 6371000.0
    rng = radar.range["data"].astype(np.float64)[None, :]
    el = np.radians(radar.elevation["data"].astype(np.float64))[:, None]
    az = np.radians(radar.azimuth["data"].astype(np.float64))[:, None]
    z = np.sqrt(rng * rng + R * R + 2.0 * rng * R * np.sin(el)) - R
    s = R * np.arcsin(rng * np.cos(el) / (R + z))
    want = [(s * np.sin(az)).ravel().astype(np.float32), (s * np.cos(az)).ravel().astype(np.float32), z.ravel().astype(np.float32)]
    for got, w, name in zip((gx, gy, gz), want, "xyz"):
        g = got.cpu().numpy()
        ulp = np.abs(g.view(np.int32).astype(np.int64) - w.view(np.int32).astype(np.int64))
        near_zero = np.abs(w) < 1e-3                       # sin(180 deg) * s: absolute, not relative, agreement there
        assert ulp[~near_zero].max() <= 1 and (ulp[~near_zero] > 0).mean() < 1e-4, (name, int(ulp[~near_zero].max()))
        assert np.abs(g[near_zero] - w[near_zero]).max(initial=0.0) < 1e-6
    kw = dict(min_radius=spec.min_radius, beam_factor=spec.beam_factor, weighting=spec.weighting, toa=spec.toa)
    dev_d = rg.DeviceGeometry.build(gx, gy, gz, spec.grid_shape, spec.grid_limits, **kw)
    dev_h = rg.DeviceGeometry.build(gx.cpu().numpy(), gy.cpu().numpy(), gz.cpu().numpy(), spec.grid_shape, spec.grid_limits, **kw)
    for a, b in zip(dev_d.export_csr(), dev_h.export_csr()):
        np.testing.assert_array_equal(a, b)
