"""
oracle/_ref — the reference's own gridding modules, byte-compiled by oracle/build_ref.py — against the NumPy
restatement (oracle/radar_grid_oracle.py) on a seeded volume: the same check tests/test_oracle_golden.py makes through
committed fixtures, but live, wherever oracle/_ref travelled to (the GPU box has no /root/reference).
"""
import tempfile
import warnings

import numpy as np
import pytest

from radar_grid_b200 import synthetic as S
from oracle import build_ref, radar_grid_oracle as O

pytestmark = pytest.mark.skipif(not build_ref.available(), reason="oracle/_ref not built (needs /root/reference at build time)")


def test_reference_modules_load_and_agree_with_the_restatement_bit_for_bit():
    ref = build_ref.load()
    spec = S.SPECS["tiny"]
    radar = S.SyntheticRadar(spec, seed=11)
    gx, gy, gz = ref.utils.get_gate_coordinates(radar)
    with tempfile.TemporaryDirectory() as tmp:
        geom = ref.compute.compute_grid_geometry(gx, gy, gz, spec.grid_shape, spec.grid_limits, tmp, min_radius=spec.min_radius,
                                                 beam_factor=spec.beam_factor, weighting="cressman", toa=spec.toa, n_workers=1)
    indptr, idx, w = O.build_geometry(gx, gy, gz, spec.grid_shape, spec.grid_limits, min_radius=spec.min_radius,
                                      beam_factor=spec.beam_factor, weighting="cressman", toa=spec.toa)
    np.testing.assert_array_equal(geom.indptr, indptr)
    np.testing.assert_array_equal(geom.gate_indices, idx)
    np.testing.assert_array_equal(geom.weights, w)
    field = ref.utils.get_field_data(radar, "DBZH")
    gf = ref.filters.GateFilter(radar).exclude_below("RHOHV", 0.8)
    grid = ref.interpolate.apply_geometry(geom, field, additional_filters=[gf])
    np.testing.assert_array_equal(grid, O.apply_geometry(indptr, idx, w, spec.grid_shape, field, extra_masks=[gf.gate_excluded]))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        np.testing.assert_array_equal(ref.products.column_max(grid), O.column_reduce("max", grid))
        np.testing.assert_array_equal(ref.products.constant_altitude_ppi(grid, geom, 1234.5), O.cappi(grid, spec.grid_shape, spec.grid_limits, 1234.5))
        np.testing.assert_array_equal(ref.products.constant_elevation_ppi(grid, geom, 2.3), O.ppi(grid, spec.grid_shape, spec.grid_limits, 2.3))
