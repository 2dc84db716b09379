"""
Interpolation-stage adapter for ``radar_processor.process_radar_to_cog``.

``process_radar_to_cog`` does not go through ``radar_grid``: it asks Py-ART for a nearest-gate grid with a constant
radius of influence (reference src/radar_processor/processor.py:128-163) and collapses it with its own PPI / CAPPI /
COLMAX rules (processor.py:480-551, utils.py:336-387).  This module provides that stage on the CUDA engine:

  * ``grid_spec``            the grid-shape and constant-ROI rule of processor.py:135-143
  * ``fill_dbzh``            the filled_DBZH preparation of processor.py:720-740
  * ``build_grid3d``         shaped like ``_get_or_build_grid3d`` (processor.py:33-43): nearest-gate 3-D grids of the
                             requested field plus the QC fields, as a Grid-like object (``.fields[name]['data']`` masked
                             (nz, ny, nx), ``.x/.y/.z['data']``)
  * ``collapse_field_3d_to_2d`` / ``collapse_grid_to_2d``   the reference's collapse, including the ``<= vmin`` re-mask

PARITY UNPINNED for the gridding half: Py-ART (arm-pyart >= 2.1.1) is not installed in either container and the
reference's tests mock it, so ``build_grid3d`` is checked against a restatement of Py-ART's published
``map_gates_to_grid`` algorithm (oracle.map_gates_to_grid_nearest), not against Py-ART itself.  The collapse half
restates pure-NumPy reference code and is checked against that restatement.
Rules taken from Py-ART's GateToGridMapper: a gate contributes to the grid points with squared distance < roi^2;
per field, masked gates and gates excluded by the gate filter are skipped; with weighting_function='nearest' a grid point
keeps the value of the closest contributing gate, the first one in gate order among equals; gates above ``toa`` are
skipped; a grid point no gate reaches is masked.
"""

from __future__ import annotations

from types import SimpleNamespace
from typing import Dict, Optional, Sequence

import numpy as np

from . import _native as N
from .engine import PPI, ColumnMax, DeviceGeometry, GeometryCache, LevelPick, grid_fields, run_products

REFLECTIVITY_LIKE = ("filled_DBZH", "DBZH", "DBZV", "DBZHF", "composite_reflectivity")   # processor.py:543
AFFECTS_INTERP_FIELDS = ("RHOHV",)          # QC fields gridded next to the requested one when the radar has them

_cache = GeometryCache(max_bytes=32 << 30)


def grid_spec(z_grid_limits, y_grid_limits, x_grid_limits, grid_resolution: float):
    """(grid_shape, constant_roi) — processor.py:135-143."""
    range_max_m = (y_grid_limits[1] - y_grid_limits[0]) / 2
    constant_roi = max(grid_resolution * 1.5, 800 + (range_max_m / 100000) * 400)
    z_points = int(np.ceil(z_grid_limits[1] / grid_resolution)) + 1
    y_points = int((y_grid_limits[1] - y_grid_limits[0]) / grid_resolution)
    x_points = int((x_grid_limits[1] - x_grid_limits[0]) / grid_resolution)
    return (z_points, y_points, x_points), constant_roi


def fill_dbzh(field_data, fill: float = -30.0) -> np.ndarray:
    """filled_DBZH (processor.py:720-740): non-finite and masked reflectivities become -30 dBZ, nothing stays masked."""
    data = np.array(np.ma.getdata(field_data), copy=True)
    data[~np.isfinite(data)] = fill
    mask = np.ma.getmask(field_data)
    if mask is not np.ma.nomask and not isinstance(mask, (bool, np.bool_)):
        data[mask] = fill
    return data


def nearest_gate_grid(gate_x, gate_y, gate_z, fields: Dict[str, np.ndarray], grid_shape, grid_limits, roi: float,
                      gate_excluded: Optional[np.ndarray] = None, toa: float = 17000.0, radar_altitude: float = 0.0,
                      products: Sequence = (), ctx: Optional[N.Context] = None):
    """
    Nearest-gate gridding of masked gate fields (``{name: masked float32 array of n_gates}``) with a constant radius of
    influence: ``{name: masked (nz, ny, nx) float32}``.  The table (gate ids + squared distances) is cached per scan
    geometry like every other table.  ``products``: optional engine product requests evaluated on the grids.
    """
    dev = _cache.get(gate_x, gate_y, gate_z, grid_shape, grid_limits, radar_altitude=radar_altitude, min_radius=float(roi),
                     beam_factor=0.0, weighting="dist2", toa=toa)
    names = list(fields)
    out = {}
    planes = {}
    for start in range(0, len(names), N.RG_MAX_FIELDS):
        chunk = names[start:start + N.RG_MAX_FIELDS]
        datas = [np.ma.getdata(fields[n]) for n in chunk]
        masks = []
        for n in chunk:
            m = np.ma.getmaskarray(fields[n]).ravel()
            masks.append(m | gate_excluded if gate_excluded is not None else m)
        res = grid_fields(dev, datas, masks=masks, reference_order="nearest_gate", products=products, ctx=ctx)
        for k, n in enumerate(chunk):
            g = res["grids"][k].reshape(grid_shape)
            out[n] = np.ma.masked_invalid(g)              # fill value NaN = no gate within the ROI
            planes[n] = [p[k] for p in res["products"]]
    return (out, planes) if products else out


def build_grid3d(radar, field_to_use: str, gatefilter=None, z_grid_limits=(0.0, 15000.0), y_grid_limits=(-240000.0, 240000.0),
                 x_grid_limits=(-240000.0, 240000.0), grid_resolution: float = 1000.0, toa: float = 17000.0):
    """
    The interpolation stage of ``_get_or_build_grid3d`` (processor.py:33-181) without its caches and Py-ART containers:
    grid ``field_to_use`` and the QC fields the radar carries onto ``grid_spec(...)`` with nearest-gate weighting and the
    constant ROI.  ``gatefilter``: anything with a ``gate_excluded`` array (pyart.filters.GateFilter, radar_grid GateFilter).
    Returns a Grid-like namespace: ``fields[name]['data']`` masked (nz, ny, nx), ``x/y/z['data']`` float64 axes.
    """
    from .utils import get_field_data, get_gate_coordinates
    shape, roi = grid_spec(z_grid_limits, y_grid_limits, x_grid_limits, grid_resolution)
    limits = (tuple(z_grid_limits), tuple(y_grid_limits), tuple(x_grid_limits))
    names = [field_to_use] + [q for q in AFFECTS_INTERP_FIELDS if q in radar.fields and q != field_to_use]
    fields = {n: get_field_data(radar, n) for n in names}
    excluded = None if gatefilter is None else np.asarray(gatefilter.gate_excluded).ravel()
    gx, gy, gz = get_gate_coordinates(radar)
    grids = nearest_gate_grid(gx, gy, gz, fields, shape, limits, roi, gate_excluded=excluded, toa=toa)
    axis = lambda lim, n: {"data": np.linspace(lim[0], lim[1], n)}
    return SimpleNamespace(fields={n: {"data": g, "_FillValue": -9999.0} for n, g in grids.items()},
                           z=axis(limits[0], shape[0]), y=axis(limits[1], shape[1]), x=axis(limits[2], shape[2]),
                           grid_shape=shape, grid_limits=limits, constant_roi=roi)


def collapse_field_3d_to_2d(data3d, product: str, *, x_coords=None, y_coords=None, z_levels=None,
                            elevation_deg: Optional[float] = None, target_height_m: Optional[float] = None):
    """
    radar_processor's own collapse (utils.py:336-387) on the GPU: 'ppi' = the level closest to
    r sin(el) + r^2 / (2 * 8.49e6) per pixel, 'cappi' = the level closest to the target height, 'colmax' = masked max
    over z.  ``data3d`` is a (masked) (nz, ny, nx) array on regular axes; returns a masked float32 (ny, nx) array.
    """
    if np.ndim(data3d) == 2:
        arr = np.ma.masked_invalid(np.ma.filled(np.ma.asarray(data3d, dtype=np.float32), np.nan))
        return np.ma.array(arr.astype(np.float32), mask=np.ma.getmaskarray(arr))
    grid = np.ma.filled(np.ma.asarray(data3d).astype(np.float32), np.nan)
    nz, ny, nx = grid.shape
    if product == "ppi":
        assert elevation_deg is not None and x_coords is not None and y_coords is not None and z_levels is not None
        req = PPI(float(elevation_deg), "closest_level")
    elif product == "cappi":
        assert target_height_m is not None and z_levels is not None
        req = LevelPick(int(np.abs(np.asarray(z_levels) - float(target_height_m)).argmin()))
    elif product == "colmax":
        req = ColumnMax()
    else:
        raise ValueError("Producto inválido")
    ax = lambda c, n: (float(c[0]), float(c[-1])) if c is not None and len(c) > 1 else (0.0, float(max(n - 1, 1)))
    limits = (ax(z_levels, nz), ax(y_coords, ny), ax(x_coords, nx))
    plane = run_products([grid], (nz, ny, nx), limits, [req])[0][0]
    return np.ma.masked_invalid(plane)


def collapse_grid_to_2d(grid, field: str, product: str, *, elevation_deg=None, target_height_m=None, vmin: float = -30.0):
    """processor.py:480-551: collapse ``grid.fields[field]`` in place to one level and re-mask (``<= vmin`` for the
    reflectivity-like fields, ``< vmin`` for KDP / ZDR)."""
    arr2d = collapse_field_3d_to_2d(grid.fields[field]["data"], product, x_coords=grid.x["data"], y_coords=grid.y["data"],
                                    z_levels=grid.z["data"], elevation_deg=elevation_deg, target_height_m=target_height_m)
    if field in REFLECTIVITY_LIKE:
        arr2d = np.ma.masked_less_equal(arr2d, vmin)
    elif field in ("KDP", "ZDR"):
        arr2d = np.ma.masked_less(arr2d, vmin)
    grid.fields[field]["data"] = arr2d[np.newaxis, ...]
    grid.fields[field]["_FillValue"] = -9999.0
    grid.z["data"] = np.array([0.0], dtype=float)
    return grid
