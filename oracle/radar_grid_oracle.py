"""
CPU oracle for the radar_grid gridding hot path  —  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A NumPy/SciPy restatement of the reference algorithm (jgmarti84/radar-processor, ``src/radar_grid``),
written so that it performs the *same floating-point operations in the same order and dtypes* as the
reference does under NumPy >= 2 (NEP 50 promotion), and therefore reproduces its outputs bit for bit.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference`` legs
may import this module; the product package (``radar_grid_b200``) never does.

Parity status: PINNED.  ``tests/golden/make_golden.py`` imports the real reference from
``/root/reference`` in the build container, runs it on the seeded synthetic volumes, and commits its
outputs under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks every function below against
those vectors bit-exactly, and ``tests/test_known_answers.py`` replays the reference's own
known-answer unit tests (tests/test_radar_grid_{interpolate,products,filters}.py) against it.

Each function cites the reference lines it follows (paths relative to /root/reference).
"""

from __future__ import annotations

import numpy as np
from scipy.spatial import cKDTree

EARTH_RADIUS = 6371000.0
KE_DEFAULT = 4.0 / 3.0


# --------------------------------------------------------------------------------------------------
# a2 / a3  neighbour table          reference: src/radar_grid/compute.py:18-103 and :106-284
# --------------------------------------------------------------------------------------------------

def _weights(weighting: str, d2: np.ndarray, r2) -> np.ndarray:
    """compute.py:82-87 — float64 arithmetic, one cast to float32 at the end."""
    if weighting == "barnes2":
        return (np.exp(-d2 / (r2 / 4)) + 1e-5).astype("float32")
    if weighting == "cressman":
        return ((r2 - d2) / (r2 + d2)).astype("float32")
    return np.ones(d2.shape[0], dtype="float32")


def grid_axes(grid_shape, grid_limits):
    """compute.py:184-186 — float32 linspace, both end points included."""
    nz, ny, nx = grid_shape
    z = np.linspace(grid_limits[0][0], grid_limits[0][1], nz, dtype="float32")
    y = np.linspace(grid_limits[1][0], grid_limits[1][1], ny, dtype="float32")
    x = np.linspace(grid_limits[2][0], grid_limits[2][1], nx, dtype="float32")
    return z, y, x


def build_geometry(gate_x, gate_y, gate_z, grid_shape, grid_limits, radar_altitude=0.0,
                   min_radius=250.0, beam_factor=0.01746, weighting="barnes2", toa=17000.0,
                   z_range=None, kd_order=True):
    """
    Neighbour table as (indptr int64, gate_indices int32, weights float32).

    Follows compute.py:182-193 (relative gate height, float32 axes, float64 voxel coordinates, TOA mask)
    and compute.py:38-91 per level: KD-tree over the TOA-valid gates as candidate generator (:38-39,:60),
    exact float64 test ``(dx*dx + dy*dy) + dz*dz < r*r`` (:69-74), weights (:82-87).
    With ``kd_order=True`` each row keeps the KD-tree traversal order the reference stores; otherwise
    rows are sorted by gate index (the canonical form used for set comparison).
    ``z_range=(z0, z1)`` restricts the table to a z-slab (rows of levels z0..z1-1), which is how the
    reference itself decomposes the work (compute.py:203-207).
    """
    if weighting not in ("barnes2", "cressman", "nearest"):
        raise ValueError(f"Unknown weighting function: {weighting}")
    nz, ny, nx = grid_shape
    gate_x = np.asarray(gate_x)
    gate_y = np.asarray(gate_y)
    gate_z_rel = np.asarray(gate_z) - radar_altitude                      # :182 (stays float32)
    z_ax, y_ax, x_ax = grid_axes(grid_shape, grid_limits)
    yy, xx = np.meshgrid(y_ax, x_ax, indexing="ij")
    vy = yy.ravel().astype("float64")                                     # :189-190
    vx = xx.ravel().astype("float64")
    valid = gate_z_rel <= toa                                             # :193
    valid_ids = np.where(valid)[0]
    gxv, gyv, gzv = gate_x[valid], gate_y[valid], gate_z_rel[valid]
    tree = cKDTree(np.column_stack([gxv, gyv, gzv]).astype("float64"))   # :38-39

    z0, z1 = (0, nz) if z_range is None else z_range
    indptr = [0]
    idx_chunks, w_chunks = [], []
    for iz in range(z0, z1):
        gz = z_ax[iz]                                                     # float32 scalar
        vz = np.full(ny * nx, gz, dtype="float64")                        # :43
        roi = np.maximum(min_radius, np.sqrt(vx ** 2 + vy ** 2 + vz ** 2) * beam_factor)   # :46-47
        pts = np.column_stack([vx, vy, vz])
        cand_lists = tree.query_ball_point(pts, roi, return_sorted=False)  # :60, traversal order
        for i in range(ny * nx):
            cand = cand_lists[i]
            if not cand:
                indptr.append(indptr[-1])
                continue
            loc = np.array(cand, dtype="int32")
            r = roi[i]
            r2 = r * r
            dx = gxv[loc] - vx[i]                                         # float32 - float64 -> float64
            dy = gyv[loc] - vy[i]
            dz = gzv[loc] - vz[i]
            d2 = dx * dx + dy * dy + dz * dz                              # :72
            keep = d2 < r2                                                # :74 strict
            if not np.any(keep):
                indptr.append(indptr[-1])
                continue
            ids = valid_ids[loc][keep].astype("int32")
            w = _weights(weighting, d2[keep], r2)
            if not kd_order:
                order = np.argsort(ids, kind="stable")
                ids, w = ids[order], w[order]
            idx_chunks.append(ids)
            w_chunks.append(w)
            indptr.append(indptr[-1] + ids.shape[0])
    gate_indices = np.concatenate(idx_chunks) if idx_chunks else np.zeros(0, "int32")
    weights = np.concatenate(w_chunks) if w_chunks else np.zeros(0, "float32")
    return np.asarray(indptr, dtype="int64"), gate_indices.astype("int32"), weights.astype("float32")


def neighbours_bruteforce(gate_x, gate_y, gate_z, voxel_xyz, radar_altitude=0.0, min_radius=250.0,
                          beam_factor=0.01746, weighting="barnes2", toa=17000.0):
    """
    Exact neighbour set of ONE voxel by testing every gate — no spatial index, so it is independent of
    the KD-tree candidate generator (used to spot-check full-size configs on sampled voxels).
    ``voxel_xyz`` are the float32 axis values (widened to float64 exactly as compute.py:43,189-190 do).
    Returns (sorted gate ids int32, weights float32 in that order).
    """
    vx, vy, vz = (np.float64(np.float32(c)) for c in voxel_xyz)
    gz_rel = np.asarray(gate_z) - radar_altitude
    valid_ids = np.where(gz_rel <= toa)[0]
    r = np.maximum(min_radius, np.sqrt(vx ** 2 + vy ** 2 + vz ** 2) * beam_factor)
    r2 = r * r
    dx = np.asarray(gate_x)[valid_ids] - vx
    dy = np.asarray(gate_y)[valid_ids] - vy
    dz = gz_rel[valid_ids] - vz
    d2 = dx * dx + dy * dy + dz * dz
    keep = d2 < r2
    return valid_ids[keep].astype("int32"), _weights(weighting, d2[keep], r2)


def canonical_rows(indptr, gate_indices, weights):
    """Sort every CSR row by gate id (stable) so that two tables can be compared as sets."""
    indptr = np.asarray(indptr, dtype=np.int64)
    n_rows = indptr.shape[0] - 1
    row_of = np.repeat(np.arange(n_rows, dtype=np.int64), np.diff(indptr))
    order = np.lexsort((gate_indices, row_of))
    return gate_indices[order], weights[order]


# --------------------------------------------------------------------------------------------------
# a4 / a5  CSR gather-weighted mean      reference: src/radar_grid/interpolate.py:15-104, 107-142
# --------------------------------------------------------------------------------------------------

def apply_geometry(indptr, gate_indices, weights, grid_shape, field_data, extra_masks=(), fill_value=np.nan):
    """
    interpolate.py:59-104.  ``field_data`` is a (masked) array of length n_gates, ``extra_masks`` the
    ``gate_excluded`` arrays of any GateFilters (OR-ed into the field mask, :60-61).
    Masked gates contribute weight 0 and value 0 (:78-79); products are float32 (:82); segment sums use
    ``np.add.reduceat`` over the non-empty rows (:85-93), i.e. ``v[0] + pairwise_sum(v[1:])`` per row;
    a voxel is ``fill_value`` when it has no gates or its effective weight sum is not > 0 (:99-102).
    """
    n_grid = int(np.prod(grid_shape))
    mask = np.ma.getmask(field_data)
    for m in extra_masks:
        mask = mask | m
    data = np.ma.getdata(field_data)
    vals = data[gate_indices]
    msk = mask[gate_indices]
    safe_v = np.where(msk, 0.0, vals)
    eff_w = np.where(msk, 0.0, weights)
    wv = eff_w * safe_v
    seg_len = np.diff(indptr)
    rows = np.where(seg_len > 0)[0]
    starts = np.asarray(indptr)[:-1][seg_len > 0]
    out = np.full(n_grid, fill_value, dtype="float32")
    if starts.shape[0]:
        s_wv = np.add.reduceat(wv, starts)
        s_w = np.add.reduceat(eff_w, starts)
        ok = s_w > 0
        out[rows[ok]] = s_wv[ok] / s_w[ok]
    return out.reshape(grid_shape)


# --------------------------------------------------------------------------------------------------
# a6  gate filters                        reference: src/radar_grid/filters.py:91-237
# --------------------------------------------------------------------------------------------------

def filter_values(field_2d) -> np.ndarray:
    """filters.py:91-102 — raw float32 values, NaN/Inf left in place (so they compare False)."""
    return np.ma.getdata(np.ma.masked_invalid(field_2d)).ravel().astype("float32")


def exclude_below(field_2d, thr):
    return filter_values(field_2d) < thr          # filters.py:133-134


def exclude_above(field_2d, thr):
    return filter_values(field_2d) > thr          # filters.py:156-157


def exclude_outside(field_2d, lo, hi):
    v = filter_values(field_2d)                   # filters.py:208-209
    return (v < lo) | (v > hi)


# --------------------------------------------------------------------------------------------------
# a7  column reductions                   reference: src/radar_grid/products.py:420-580
# --------------------------------------------------------------------------------------------------

def _z_slice(grid, z_min_idx, z_max_idx, z_min_alt, z_max_alt, grid_limits):
    nz = grid.shape[0]
    if z_min_alt is not None or z_max_alt is not None:
        if grid_limits is None:
            raise ValueError("geometry is required when using altitude-based limits")
        zc = np.linspace(grid_limits[0][0], grid_limits[0][1], nz)       # float64, :469
        if z_min_alt is not None:
            z_min_idx = np.searchsorted(zc, z_min_alt)
        if z_max_alt is not None:
            z_max_idx = np.searchsorted(zc, z_max_alt, side="right") - 1
    z_min_idx = 0 if z_min_idx is None else z_min_idx
    z_max_idx = nz - 1 if z_max_idx is None else z_max_idx
    return max(0, z_min_idx), min(nz - 1, z_max_idx)


def column_reduce(kind, grid, z_min_idx=None, z_max_idx=None, z_min_alt=None, z_max_alt=None, grid_limits=None):
    """products.py:462-490 (max), :509-535 (min), :554-580 (mean): nan-ignoring reductions over z."""
    lo, hi = _z_slice(grid, z_min_idx, z_max_idx, z_min_alt, z_max_alt, grid_limits)
    sl = grid[lo:hi + 1]
    with np.errstate(all="ignore"):
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore", RuntimeWarning)
            return {"max": np.nanmax, "min": np.nanmin, "mean": np.nanmean}[kind](sl, axis=0)


# --------------------------------------------------------------------------------------------------
# a8  CAPPI                               reference: src/radar_grid/products.py:317-415
# --------------------------------------------------------------------------------------------------

def cappi(grid, grid_shape, grid_limits, altitude, interpolation="linear"):
    nz, ny, nx = grid_shape
    z_min, z_max = grid_limits[0]
    zc = np.linspace(z_min, z_max, nz, dtype="float32")                  # :367
    if altitude < z_min or altitude > z_max:                              # :370-372
        return np.full((ny, nx), np.nan, dtype="float32")
    if interpolation == "nearest":
        return grid[np.argmin(np.abs(zc - altitude))]                     # :377-378
    if interpolation != "linear":
        raise ValueError(f"Unknown interpolation method: {interpolation}")
    hit = np.isclose(zc, altitude, rtol=1e-6)                             # :382-386
    if np.any(hit):
        return grid[np.where(hit)[0][0]]
    z_step = (z_max - z_min) / (nz - 1) if nz > 1 else 1.0
    z_frac = (altitude - z_min) / z_step
    lo = int(np.floor(z_frac))
    hi = lo + 1
    if lo < 0:
        return grid[0]
    if hi >= nz:
        return grid[nz - 1]
    w_hi = z_frac - lo
    w_lo = 1.0 - w_hi
    return (w_lo * grid[lo] + w_hi * grid[hi]).astype("float32")          # :411-412


# --------------------------------------------------------------------------------------------------
# a9  constant-elevation PPI              reference: src/radar_grid/products.py:23-89, 139-165, 168-314
# --------------------------------------------------------------------------------------------------

def beam_height(horizontal_distance, elevation_angle, radar_altitude=0.0, ke=KE_DEFAULT, re=EARTH_RADIUS):
    """products.py:70-89."""
    e = np.radians(elevation_angle)
    ke_re = ke * re
    sr = horizontal_distance / np.maximum(np.cos(e), 0.01)
    return np.sqrt(sr ** 2 + ke_re ** 2 + 2 * sr * ke_re * np.sin(e)) - ke_re + radar_altitude


def beam_height_flat(horizontal_distance, elevation_angle, radar_altitude=0.0):
    """products.py:164-165."""
    return horizontal_distance * np.tan(np.radians(elevation_angle)) + radar_altitude


def ppi(grid, grid_shape, grid_limits, elevation_angle, interpolation="linear", earth_curvature=True,
        ke=KE_DEFAULT):
    nz, ny, nx = grid_shape
    z_min, z_max = grid_limits[0]
    yc = np.linspace(grid_limits[1][0], grid_limits[1][1], ny, dtype="float32")   # :227-228
    xc = np.linspace(grid_limits[2][0], grid_limits[2][1], nx, dtype="float32")
    yy, xx = np.meshgrid(yc, xc, indexing="ij")
    hdist = np.sqrt(xx ** 2 + yy ** 2)                                    # float32, :235
    tz = (beam_height(hdist, elevation_angle, 0.0, ke) if earth_curvature
          else beam_height_flat(hdist, elevation_angle, 0.0))             # :238-251
    z_step = (z_max - z_min) / (nz - 1) if nz > 1 else 1.0
    yi, xi = np.meshgrid(np.arange(ny), np.arange(nx), indexing="ij")
    if interpolation == "nearest":                                        # :256-272
        zi = np.round((tz - z_min) / z_step).astype(int)
        ok = (zi >= 0) & (zi < nz)
        out = grid[np.clip(zi, 0, nz - 1), yi, xi]
        out[~ok] = np.nan
        return out
    if interpolation != "linear":
        raise ValueError(f"Unknown interpolation method: {interpolation}")
    z_frac = (tz - z_min) / z_step                                        # :279-309
    lo = np.floor(z_frac).astype(int)
    hi = lo + 1
    w_hi = z_frac - lo
    w_lo = 1.0 - w_hi
    v_lo = grid[np.clip(lo, 0, nz - 1), yi, xi]
    v_hi = grid[np.clip(hi, 0, nz - 1), yi, xi]
    out = w_lo * v_lo + w_hi * v_hi
    out[tz < z_min] = np.nan
    out[tz > z_max] = np.nan
    return out


# --------------------------------------------------------------------------------------------------
# a10  GridFilter                         reference: src/radar_grid/filters.py:631-780
# --------------------------------------------------------------------------------------------------

def grid_filter(kind, plane, a=None, b=None, fill_value=np.nan):
    out = plane.copy()
    if kind == "below":
        out[out < a] = fill_value
    elif kind == "above":
        out[out > a] = fill_value
    elif kind == "outside":
        out[(out < a) | (out > b)] = fill_value
    elif kind == "invalid":
        out[np.isnan(out) | np.isinf(out)] = fill_value
    else:
        raise ValueError(kind)
    return out


# --------------------------------------------------------------------------------------------------
# f4  RGBA image of a 2-D product      reference: src/radar_grid/geotiff.py:70-145
# --------------------------------------------------------------------------------------------------

def colormap_table(n=256):
    """A deterministic (n + 3, 4) float table in matplotlib's ``Colormap._lut`` layout (n colours, then the under / over /
    bad rows): test input only.  matplotlib is not installed where these tests run, so colour tables are inputs here."""
    t = np.linspace(0.0, 1.0, n)
    lut = np.stack([t, np.sin(np.pi * t) ** 2, 1.0 - t, np.full(n, 1.0)], axis=1)
    return np.concatenate([lut, lut[:1], lut[-1:], np.zeros((1, 4))])


def apply_colormap(data, lut, vmin=None, vmax=None, fill_value=None):
    """
    geotiff.py:105-143 with matplotlib's own steps written out (parity UNPINNED for the matplotlib half: the package is
    absent from both containers, so Normalize / Colormap.__call__ are restated from matplotlib 3.8-3.10's sources):

      nodata = (data == fill_value) if fill_value is not None else isnan(data)            geotiff.py:112-116
      vmin / vmax default to nanmin / nanmax of the valid pixels, 0 / 1 if there are none  geotiff.py:118-129
      Normalize(vmin, vmax, clip=True)(data): process_value keeps float32 data float32, but np.clip against the
        float64 limits promotes to float64 (NEP 50), so x = (clip(data) - vmin) / (vmax - vmin) in float64;
        vmin == vmax -> 0; vmin > vmax -> ValueError                                      colors.py Normalize.__call__
      Colormap.__call__: xa = x * N; xa[xa == N] = N - 1; under = xa < 0; over = xa >= N; bad = isnan(xa);
        xa.astype(int) (truncation); rgba = lut.take(xa)                                  colors.py Colormap.__call__
      (rgba * 255).astype(uint8); alpha = 0 where nodata                                  geotiff.py:137-141

    ``lut`` is the (N + 3, 4) float table (Colormap._lut).
    """
    data = np.array(data, copy=True)
    lut = np.asarray(lut, dtype=np.float64)
    n = lut.shape[0] - 3
    nodata = (data == fill_value) if fill_value is not None else np.isnan(data)
    valid = data[~nodata]
    if len(valid) > 0:
        if vmin is None:
            vmin = np.nanmin(valid)
        if vmax is None:
            vmax = np.nanmax(valid)
    else:
        vmin = 0.0 if vmin is None else vmin
        vmax = 1.0 if vmax is None else vmax
    vmin64, vmax64 = np.float64(vmin), np.float64(vmax)
    if vmin64 > vmax64:
        raise ValueError("minvalue must be less than or equal to maxvalue")
    if vmin64 == vmax64:
        x = np.zeros(data.shape, dtype=data.dtype)
    else:
        x = np.clip(data, vmin64, vmax64)          # float64 result
        x -= vmin64
        x /= (vmax64 - vmin64)
    xa = np.array(x, copy=True)
    xa *= n
    xa[xa == n] = n - 1
    under, over, bad = xa < 0, xa >= n, np.isnan(xa)
    with np.errstate(invalid="ignore"):
        idx = xa.astype(int)
    idx[under], idx[over], idx[bad] = n, n + 1, n + 2
    rgba = (lut.take(idx, axis=0, mode="clip") * 255).astype(np.uint8)
    rgba[nodata, 3] = 0
    return rgba


# --------------------------------------------------------------------------------------------------
# f1  process_radar_to_cog's interpolation stage      reference: src/radar_processor/processor.py:128-163, 480-551
# --------------------------------------------------------------------------------------------------

def map_gates_to_grid_nearest(gate_x, gate_y, gate_z, fields, grid_shape, grid_limits, roi, gate_excluded=None, toa=17000.0):
    """
    PARITY UNPINNED.  What the reference asks Py-ART for (processor.py:152-163: pyart.map.grid_from_radars(...,
    gridding_algo='map_gates_to_grid', weighting_function='nearest', roi_func='constant', constant_roi=roi,
    gatefilters=gf)), restated from the published algorithm of arm-pyart's GateToGridMapper (pyart/map/_gate_to_grid_map.pyx,
    2.x): arm-pyart is not vendored under /root/reference nor installed, and the reference's tests mock it.

      for every gate (in gate order) not excluded by the gate filter and with z <= toa:
          for every grid point with dist2 = dx^2 + dy^2 + dz^2 < roi^2:
              for every field whose value at the gate is not masked:
                  if dist2 < min_dist2[point, field]:  min_dist2 = dist2; value[point, field] = gate value
      points no gate reached are masked.

    Distances here are float64 on the float32 gate coordinates and the float32 grid axes the rest of this package uses
    (Py-ART evaluates them in float32 on float64 axes: near-ties may resolve differently — part of "unpinned").
    Brute force over grid points: small cases only.  fields: {name: masked float32[n_gates]} -> {name: masked (nz, ny, nx)}.
    """
    z_ax, y_ax, x_ax = grid_axes(grid_shape, grid_limits)
    gx, gy, gz = (np.asarray(a, dtype=np.float32).astype(np.float64) for a in (gate_x, gate_y, gate_z))
    ok = gz.astype(np.float32) <= np.float32(toa)
    if gate_excluded is not None:
        ok &= ~np.asarray(gate_excluded, dtype=bool).ravel()
    nz, ny, nx = grid_shape
    out = {n: np.ma.masked_all((nz, ny, nx), dtype=np.float32) for n in fields}
    data = {n: np.ma.getdata(f).astype(np.float32) for n, f in fields.items()}
    valid = {n: ok & ~np.ma.getmaskarray(f).ravel() for n, f in fields.items()}
    r2 = float(roi) * float(roi)
    for iz in range(nz):
        dz2 = (gz - float(z_ax[iz])) ** 2
        near_z = dz2 < r2
        for iy in range(ny):
            dy2 = (gy - float(y_ax[iy])) ** 2
            cand = np.nonzero(near_z & (dy2 < r2))[0]
            if cand.size == 0:
                continue
            for ix in range(nx):
                dx = gx[cand] - float(x_ax[ix])
                d2 = (dx * dx + dy2[cand]) + dz2[cand]
                inside = d2 < r2
                if not inside.any():
                    continue
                idx, dd = cand[inside], d2[inside].astype(np.float32)     # the table stores float32(d2)
                for n in fields:
                    v = valid[n][idx]
                    if v.any():
                        k = np.argmin(np.where(v, dd, np.float32(np.inf)))   # first minimum = lowest gate id
                        out[n][iz, iy, ix] = data[n][idx[k]]
    return out


def collapse_field_3d_to_2d(data3d, product, x_coords=None, y_coords=None, z_levels=None, elevation_deg=None,
                            target_height_m=None):
    """radar_processor/utils.py:336-387 (pure NumPy in the reference; the module cannot be imported because it imports
    pyart at the top): 'ppi' level closest to r sin(el) + r^2 / (2 * 8.49e6), 'cappi' closest level, 'colmax' masked max."""
    if data3d.ndim == 2:
        arr2d = data3d
    elif product == "ppi":
        X, Y = np.meshgrid(x_coords, y_coords, indexing="xy")
        r = np.sqrt(X ** 2 + Y ** 2)
        Re = 8.49e6
        z_target = r * np.sin(np.deg2rad(elevation_deg)) + (r ** 2) / (2.0 * Re)
        iz = np.abs(z_target[..., None] - z_levels[None, None, :]).argmin(axis=2)
        yy = np.arange(len(y_coords))[:, None]
        xx = np.arange(len(x_coords))[None, :]
        arr2d = data3d[iz, yy, xx]
    elif product == "cappi":
        iz = np.abs(z_levels - float(target_height_m)).argmin()
        arr2d = data3d[iz, :, :]
    elif product == "colmax":
        arr2d = data3d.max(axis=0)
    else:
        raise ValueError("Producto inválido")
    return np.ma.array(arr2d.astype(np.float32), mask=np.ma.getmaskarray(arr2d))


def remask_2d(arr2d, field, vmin=-30.0):
    """processor.py:541-546."""
    arr2d = np.ma.masked_invalid(arr2d)
    if field in ["filled_DBZH", "DBZH", "DBZV", "DBZHF", "composite_reflectivity"]:
        arr2d = np.ma.masked_less_equal(arr2d, vmin)
    elif field in ["KDP", "ZDR"]:
        arr2d = np.ma.masked_less(arr2d, vmin)
    return arr2d
