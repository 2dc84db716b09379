#!/usr/bin/env python
"""
Generate the golden fixtures in this directory by running the REAL reference implementation
(jgmarti84/radar-processor, mounted read-only at /root/reference) on the seeded synthetic volumes.

Only runnable in the build container (the GPU box has no /root/reference); its outputs are committed so
that the oracle and the CUDA path can be checked against the reference everywhere.

    python tests/golden/make_golden.py            # writes tests/golden/*.npz

``import radar_grid`` fails here (matplotlib/rasterio absent), so the reference's hot-path modules are
imported under a stub package whose __path__ points at the reference sources.
"""

import importlib
import os
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "radar-processor_b200"))

REF_SRC = "/root/reference/src/radar_grid"


def load_reference():
    pkg = types.ModuleType("radar_grid_ref")
    pkg.__path__ = [REF_SRC]
    sys.modules["radar_grid_ref"] = pkg
    mods = {}
    for name in ("geometry", "compute", "interpolate", "products", "filters", "utils"):
        mods[name] = importlib.import_module(f"radar_grid_ref.{name}")
    return types.SimpleNamespace(**mods)


def main():
    from radar_grid_b200 import synthetic as S
    ref = load_reference()
    manifest = []

    for spec_name, weightings, radar_alt in (("tiny", ("barnes2", "cressman", "nearest"), 0.0),
                                             ("small", ("barnes2",), 0.0),
                                             ("tiny", ("barnes2",), 350.0)):
        spec = S.SPECS[spec_name]
        radar = S.SyntheticRadar(spec, seed=7, radar_altitude=radar_alt)
        gx, gy, gz = ref.utils.get_gate_coordinates(radar)
        fields = {name: ref.utils.get_field_data(radar, name) for name in spec.fields}
        for weighting in weightings:
            with tempfile.TemporaryDirectory() as tmp:
                geom = ref.compute.compute_grid_geometry(
                    gx, gy, gz, spec.grid_shape, spec.grid_limits, tmp, radar_altitude=radar_alt,
                    min_radius=spec.min_radius, beam_factor=spec.beam_factor, weighting=weighting,
                    toa=spec.toa if radar_alt == 0.0 else 4000.0, n_workers=1)
            out = {
                "indptr": geom.indptr, "gate_indices": geom.gate_indices, "weights": geom.weights,
                "toa": np.array([geom.toa]), "radar_altitude": np.array([radar_alt]),
            }
            # apply_geometry for every field, plain and with the cfg2-style RHOHV QC filter
            gf = ref.filters.GateFilter(radar)
            gf.exclude_below("RHOHV", 0.8).exclude_above("RHOHV", 1.0)
            out["rhohv_excluded"] = gf.gate_excluded
            for name, data in fields.items():
                grid = ref.interpolate.apply_geometry(geom, data)
                out[f"grid_{name}"] = grid
                out[f"gridqc_{name}"] = ref.interpolate.apply_geometry(geom, data, additional_filters=[gf])
            out["grid_fill_DBZH"] = ref.interpolate.apply_geometry(geom, fields["DBZH"], fill_value=-9999.0)
            g = out["grid_DBZH"]
            gq = out["gridqc_DBZH"]
            zmax = spec.grid_limits[0][1]
            with np.errstate(all="ignore"):
                import warnings
                warnings.simplefilter("ignore")
                out["colmax"] = ref.products.column_max(g)
                out["colmax_qc"] = ref.products.column_max(gq)
                out["colmax_idx_1_3"] = ref.products.column_max(g, z_min_idx=1, z_max_idx=3)
                out["colmax_alt"] = ref.products.column_max(g, z_min_alt=1500.0, z_max_alt=0.6 * zmax, geometry=geom)
                out["colmin"] = ref.products.column_min(g)
                out["colmean"] = ref.products.column_mean(g)
                out["colmax_fill"] = ref.products.column_max(out["grid_fill_DBZH"])
            for alt in (4000.0, 0.0, zmax, 1234.5, spec.grid_limits[0][1] / (spec.grid_shape[0] - 1) * 2):
                tag = f"{alt:.1f}"
                out[f"cappi_lin_{tag}"] = np.array(ref.products.constant_altitude_ppi(g, geom, alt))
                out[f"cappi_near_{tag}"] = np.array(ref.products.constant_altitude_ppi(g, geom, alt, "nearest"))
            out["cappi_oob"] = ref.products.constant_altitude_ppi(g, geom, zmax + 1.0)
            for elev in (0.5, 2.3, 6.9, 25.0):
                tag = f"{elev:.1f}"
                out[f"ppi_lin_{tag}"] = ref.products.constant_elevation_ppi(g, geom, elev)
                out[f"ppi_near_{tag}"] = ref.products.constant_elevation_ppi(g, geom, elev, interpolation="nearest")
                out[f"ppi_flat_{tag}"] = ref.products.constant_elevation_ppi(g, geom, elev, earth_curvature=False)
            # a numpy-scalar grid_limits geometry (as load_geometry returns) flips the CAPPI blend to float64
            geom64 = ref.geometry.GridGeometry(
                grid_shape=spec.grid_shape,
                grid_limits=tuple(tuple(np.float64(v) for v in lim) for lim in spec.grid_limits),
                indptr=geom.indptr, gate_indices=geom.gate_indices, weights=geom.weights, toa=geom.toa)
            out["cappi_lin64_1234.5"] = ref.products.constant_altitude_ppi(g, geom64, 1234.5)
            fname = f"ref_{spec_name}_{weighting}_alt{int(radar_alt)}.npz"
            np.savez_compressed(os.path.join(HERE, fname), **out)
            manifest.append((fname, int(geom.n_pairs()), [k for k in out]))
            print(f"{fname}: pairs={geom.n_pairs():,} max_row={np.diff(geom.indptr).max()} "
                  f"empty={np.mean(np.diff(geom.indptr) == 0):.2f} "
                  f"nan_vox={np.isnan(g).mean():.2f}")

    # GateFilter / GridFilter known answers on the tiny radar
    spec = S.SPECS["tiny"]
    radar = S.SyntheticRadar(spec, seed=7)
    flt = {}
    gf = ref.filters.GateFilter(radar)
    flt["below_DBZH_5"] = gf.copy().exclude_below("DBZH", 5.0).gate_excluded
    flt["above_ZDR_1"] = gf.copy().exclude_above("ZDR", 1.0).gate_excluded
    flt["outside_RHOHV"] = gf.copy().exclude_outside("RHOHV", 0.8, 0.95).gate_excluded
    flt["between_KDP"] = gf.copy().exclude_between("KDP", 0.0, 0.5).gate_excluded
    flt["equal_VRAD"] = gf.copy().exclude_equal("VRAD", 0.0, atol=1.0).gate_excluded
    flt["invalid_DBZH"] = gf.copy().exclude_invalid("DBZH").gate_excluded
    flt["masked_DBZH"] = gf.copy().exclude_masked("DBZH").gate_excluded
    flt["all_invalid_DBZH"] = gf.copy().exclude_all_invalid("DBZH").gate_excluded
    flt["below_alt_2000"] = gf.copy().exclude_below_altitude(2000.0).gate_excluded
    flt["above_alt_5000"] = gf.copy().exclude_above_altitude(5000.0).gate_excluded
    flt["below_range"] = gf.copy().exclude_below_range(3000.0).gate_excluded
    flt["above_range"] = gf.copy().exclude_above_range(15000.0).gate_excluded
    flt["below_elev"] = gf.copy().exclude_below_elevation_angle(2.0).gate_excluded
    flt["above_elev"] = gf.copy().exclude_above_elevation_angle(10.0).gate_excluded
    flt["outside_elev"] = gf.copy().exclude_outside_elevation_range(1.0, 10.0).gate_excluded
    chained = gf.copy().exclude_below("RHOHV", 0.8).exclude_above("RHOHV", 1.0).exclude_below("DBZH", 0.0)
    flt["chained"] = chained.gate_excluded
    d, m = ref.filters.create_mask_from_filter(radar, "DBZH", chained)
    flt["cmff_data"], flt["cmff_mask"] = d, m
    plane = np.load(os.path.join(HERE, "ref_tiny_barnes2_alt0.npz"))["colmax"]
    gfl = ref.filters.GridFilter()
    flt["plane"] = plane
    flt["grid_below"] = gfl.apply_below(plane, 15.0)
    flt["grid_above"] = gfl.apply_above(plane, 30.0)
    flt["grid_outside"] = gfl.apply_outside_range(plane, 10.0, 35.0, fill_value=-1.0)
    flt["grid_invalid"] = gfl.apply_invalid(plane, fill_value=-99.0)
    np.savez_compressed(os.path.join(HERE, "ref_filters_tiny.npz"), **flt)
    print("ref_filters_tiny.npz:", len(flt), "arrays")

    golden_multi(ref)
    golden_cfg1(ref)


def golden_multi(ref):
    """apply_geometry_multi of the real reference (interpolate.py:107-142): five fields, with and without a per-field
    filter dict (a list for one field, a bare GateFilter for another, as the reference's tests pass them)."""
    from radar_grid_b200 import synthetic as S
    spec = S.SPECS["small"]
    radar = S.SyntheticRadar(spec, seed=7)
    z = np.load(os.path.join(HERE, "ref_small_barnes2_alt0.npz"))
    geom = ref.geometry.GridGeometry(grid_shape=spec.grid_shape, grid_limits=spec.grid_limits, indptr=z["indptr"],
                                     gate_indices=z["gate_indices"], weights=z["weights"], toa=float(z["toa"][0]))
    fields = {name: ref.utils.get_field_data(radar, name) for name in spec.fields}
    gf_rho = ref.filters.GateFilter(radar).exclude_below("RHOHV", 0.8).exclude_above("RHOHV", 1.0)
    gf_dbz = ref.filters.GateFilter(radar).exclude_below("DBZH", 5.0)
    out = {}
    for name, g in ref.interpolate.apply_geometry_multi(geom, fields).items():
        out[f"plain_{name}"] = g
    filt = {"DBZH": [gf_rho, gf_dbz], "ZDR": gf_rho}
    for name, g in ref.interpolate.apply_geometry_multi(geom, fields, additional_filters=filt, fill_value=-5.0).items():
        out[f"filt_{name}"] = g
    np.savez_compressed(os.path.join(HERE, "ref_multi_small.npz"), **out)
    print("ref_multi_small.npz:", len(out), "grids")


def row_digest(indptr, idx, w):
    """Order-independent description of a table: SHA-256 of the gate ids with every row sorted by gate id, and per
    voxel column the int64 sum of the float32 bit patterns of its weights (a 1-ulp difference moves a sum by 1)."""
    import hashlib
    indptr = np.asarray(indptr, dtype=np.int64)
    lens = np.diff(indptr)
    row_of = np.repeat(np.arange(len(lens), dtype=np.int64), lens)
    order = np.lexsort((idx, row_of))
    sha = hashlib.sha256(np.ascontiguousarray(idx[order], dtype=np.int32).tobytes()).hexdigest()
    return sha, row_of, lens


def golden_cfg1(ref):
    """BASELINE configs[0]/[1] at FULL size through the real reference: compute_grid_geometry on the cfg1 volume
    (1.16 M voxels, 1.15e7 pairs, ~15 s on 7 workers), then apply_geometry + COLMAX (+ the cfg2 RHOHV filter, PPI 0.5
    deg and CAPPI 4000 m).  The table itself is 100 MB, so what is committed is its digest: row lengths, the SHA-256 of
    the row-sorted gate ids, per-column sums of the weights' bit patterns, and the 2-D products."""
    import warnings
    from radar_grid_b200 import synthetic as S
    spec = S.SPECS["cfg2"]                      # the cfg1 volume + RHOHV
    radar = S.SyntheticRadar(spec, seed=1)
    gx, gy, gz = ref.utils.get_gate_coordinates(radar)
    with tempfile.TemporaryDirectory() as tmp:
        geom = ref.compute.compute_grid_geometry(gx, gy, gz, spec.grid_shape, spec.grid_limits, tmp,
                                                 min_radius=spec.min_radius, beam_factor=spec.beam_factor,
                                                 weighting=spec.weighting, toa=spec.toa, n_workers=max(1, (os.cpu_count() or 2) - 1))
    nz, ny, nx = spec.grid_shape
    sha, row_of, lens = row_digest(geom.indptr, geom.gate_indices, geom.weights)
    col_of = row_of % (ny * nx)
    wsum = np.bincount(col_of, weights=None, minlength=ny * nx) * 0
    wbits = geom.weights.view(np.int32).astype(np.int64)
    wsum = np.zeros(ny * nx, dtype=np.int64)
    np.add.at(wsum, col_of, wbits)
    dbz = ref.utils.get_field_data(radar, "DBZH")
    gf = ref.filters.GateFilter(radar).exclude_below("RHOHV", 0.8).exclude_above("RHOHV", 1.0)
    g = ref.interpolate.apply_geometry(geom, dbz)
    gq = ref.interpolate.apply_geometry(geom, dbz, additional_filters=[gf])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out = {
            "row_len": lens.astype(np.uint16 if lens.max() < 65536 else np.uint32),
            "gate_ids_sha256": np.array(sha), "n_pairs": np.array([geom.n_pairs()], dtype=np.int64),
            "weight_bits_colsum": wsum,
            "colmax": ref.products.column_max(g), "colmax_qc": ref.products.column_max(gq),
            "cappi_4000_qc": np.array(ref.products.constant_altitude_ppi(gq, geom, 4000.0)),
            "ppi_0.5_qc": ref.products.constant_elevation_ppi(gq, geom, 0.5),
            "grid_nan_count": np.array([int(np.isnan(g).sum()), int(np.isnan(gq).sum())], dtype=np.int64),
        }
    np.savez_compressed(os.path.join(HERE, "ref_cfg1_digest.npz"), **out)
    print(f"ref_cfg1_digest.npz: pairs={geom.n_pairs():,} max_row={lens.max()} sha={sha[:16]}")


if __name__ == "__main__":
    main()
