"""
The interpolation-stage adapter for process_radar_to_cog (radar_grid_b200/adapter.py; reference
src/radar_processor/processor.py:128-163, 480-551, 720-740, utils.py:336-387).

PARITY UNPINNED for the nearest-gate gridding: Py-ART is absent, so the CUDA path is checked against the oracle's
restatement of Py-ART's published map_gates_to_grid algorithm; the collapse and the grid-shape / ROI / filled_DBZH rules
restate NumPy code of the reference and are checked against those restatements and against hand-computed values.
"""
import numpy as np
import pytest

import radar_grid_b200 as rg
from radar_grid_b200 import adapter as A, synthetic as S
from oracle import radar_grid_oracle as O


def test_grid_spec_and_filled_dbzh_rules():
    # processor.py:135-143 with the values process_radar_to_cog uses for a 240 km product at 1 km
    shape, roi = A.grid_spec((0.0, 15000.0), (-240000.0, 240000.0), (-240000.0, 240000.0), 1000.0)
    assert shape == (16, 480, 480) and roi == max(1500.0, 800 + 2.4 * 400)
    shape, roi = A.grid_spec((0.0, 12000.0), (-60000.0, 60000.0), (-60000.0, 60000.0), 300.0)
    assert shape == (int(np.ceil(12000.0 / 300.0)) + 1, 400, 400) and roi == 800 + 0.6 * 400
    f = np.ma.array(np.array([1.0, np.nan, np.inf, -np.inf, 5.0, 7.0], np.float32), mask=[0, 0, 0, 0, 1, 0])
    np.testing.assert_array_equal(A.fill_dbzh(f), np.array([1.0, -30.0, -30.0, -30.0, -30.0, 7.0], np.float32))
    np.testing.assert_array_equal(A.fill_dbzh(np.array([np.nan, 2.0], np.float32)), np.array([-30.0, 2.0], np.float32))


def test_oracle_collapse_known_answers():
    nz, ny, nx = 4, 3, 5
    z = np.linspace(0.0, 3000.0, nz)
    y = np.linspace(-2000.0, 2000.0, ny)
    x = np.linspace(-40000.0, 40000.0, nx)
    data = np.ma.masked_invalid(np.arange(nz * ny * nx, dtype=np.float32).reshape(nz, ny, nx))
    data[2, 1, 2] = np.ma.masked
    cm = O.collapse_field_3d_to_2d(data, "colmax")
    assert cm.dtype == np.float32 and cm[1, 2] == data[3, 1, 2] and not np.ma.getmaskarray(cm).any()
    cap = O.collapse_field_3d_to_2d(data, "cappi", z_levels=z, target_height_m=1900.0)
    np.testing.assert_array_equal(np.ma.getdata(cap)[0], np.ma.getdata(data)[2, 0])
    assert np.ma.getmaskarray(cap)[1, 2]
    ppi = O.collapse_field_3d_to_2d(data, "ppi", x_coords=x, y_coords=y, z_levels=z, elevation_deg=2.0)
    # centre pixel: r = 0 -> level 0; x = 40 km: 40e3 sin(2 deg) + 40e3^2 / (2 * 8.49e6) = 1396 + 94 = 1490 m -> level 1
    assert ppi[1, 2] == data[0, 1, 2] and ppi[1, 4] == data[1, 1, 4] and ppi[1, 3] == data[1, 1, 3]
    assert np.ma.getmaskarray(O.remask_2d(np.ma.array([-30.0, -29.9, np.nan]), "DBZH")).tolist() == [True, False, True]
    assert np.ma.getmaskarray(O.remask_2d(np.ma.array([-30.0, -30.1]), "ZDR", vmin=-30.0)).tolist() == [False, True]


@pytest.mark.gpu
def test_nearest_gate_grid_matches_the_map_gates_to_grid_restatement():
    spec = S.SPECS["tiny"]
    radar = S.SyntheticRadar(spec, seed=5)
    gates = rg.get_gate_coordinates(radar)
    fields = {n: rg.get_field_data(radar, n) for n in ("DBZH", "RHOHV", "KDP")}
    gf = rg.GateFilter(radar).exclude_below("RHOHV", 0.6)
    roi = 1600.0
    want = O.map_gates_to_grid_nearest(*gates, fields, spec.grid_shape, spec.grid_limits, roi, gate_excluded=gf.gate_excluded, toa=9000.0)
    got = A.nearest_gate_grid(*gates, fields, spec.grid_shape, spec.grid_limits, roi, gate_excluded=gf.gate_excluded, toa=9000.0)
    for n in fields:
        np.testing.assert_array_equal(np.ma.getmaskarray(got[n]), np.ma.getmaskarray(want[n]), err_msg=f"{n}: mask")
        np.testing.assert_array_equal(got[n].filled(-999.0), want[n].filled(-999.0), err_msg=n)
        assert 0.05 < np.ma.getmaskarray(got[n]).mean() < 0.95
    # filled reflectivity: nothing masked inside the radar's reach, and the plain field's valid voxels agree where the
    # closest gate of both is the same valid gate
    filled = {"filled_DBZH": np.ma.masked_invalid(A.fill_dbzh(fields["DBZH"]))}
    g2 = A.nearest_gate_grid(*gates, filled, spec.grid_shape, spec.grid_limits, roi)["filled_DBZH"]
    w2 = O.map_gates_to_grid_nearest(*gates, filled, spec.grid_shape, spec.grid_limits, roi)["filled_DBZH"]
    np.testing.assert_array_equal(g2.filled(-999.0), w2.filled(-999.0))
    assert (np.ma.getdata(g2)[~np.ma.getmaskarray(g2)] >= -30.0).all()


@pytest.mark.gpu
def test_collapse_on_the_gpu_matches_the_reference_collapse_and_remask():
    spec = S.SPECS["tiny"]
    radar = S.SyntheticRadar(spec, seed=5)
    grid = A.build_grid3d(radar, "DBZH", gatefilter=None, z_grid_limits=(0.0, 9000.0), y_grid_limits=(-20000.0, 20000.0),
                          x_grid_limits=(-20000.0, 20000.0), grid_resolution=2000.0, toa=9000.0)
    assert grid.grid_shape == (6, 20, 20) and grid.constant_roi == 3000.0
    assert set(grid.fields) == {"DBZH", "RHOHV"} and grid.fields["DBZH"]["data"].shape == (6, 20, 20)
    d3 = grid.fields["DBZH"]["data"]
    x, y, z = grid.x["data"], grid.y["data"], grid.z["data"]
    for product, kw in (("colmax", {}), ("cappi", {"target_height_m": 2900.0}), ("ppi", {"elevation_deg": 6.9}), ("ppi", {"elevation_deg": 0.5})):
        want = O.collapse_field_3d_to_2d(d3, product, x_coords=x, y_coords=y, z_levels=z, **kw)
        got = A.collapse_field_3d_to_2d(d3, product, x_coords=x, y_coords=y, z_levels=z, **kw)
        assert got.dtype == np.float32 and got.shape == (20, 20)
        np.testing.assert_array_equal(np.ma.getmaskarray(got), np.ma.getmaskarray(want), err_msg=product)
        np.testing.assert_array_equal(got.filled(-999.0), want.filled(-999.0), err_msg=product)
    want = O.remask_2d(O.collapse_field_3d_to_2d(d3, "colmax"), "DBZH", vmin=10.0)
    A.collapse_grid_to_2d(grid, "DBZH", "colmax", vmin=10.0)
    out = grid.fields["DBZH"]["data"]
    assert out.shape == (1, 20, 20) and grid.z["data"].tolist() == [0.0]
    np.testing.assert_array_equal(np.ma.getmaskarray(out[0]), np.ma.getmaskarray(want))
    np.testing.assert_array_equal(out[0].filled(-999.0), want.filled(-999.0))
