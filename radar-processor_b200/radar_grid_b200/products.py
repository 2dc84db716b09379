"""
Radar products with the reference's names, arguments and dtypes (src/radar_grid/products.py):
constant_elevation_ppi (:168), constant_altitude_ppi (:317), column_max/min/mean (:420-580), beam-height
helpers (:23-165, :583-697).

The grid reductions run on the GPU (csrc/rg_apply.cu, ``products_kernel``) on NumPy or torch CUDA grids; the
same per-column code is the epilogue of the fused interpolation (``engine.grid_fields(products=...)``), which
is the fast path when the 3-D grid itself is not wanted.  The helper formulas on coordinate arrays
(compute_beam_height & co.) are scalar-parameter NumPy expressions, kept as such.

Differences from the reference, by design: results are always fresh arrays (the reference returns a *view*
of ``grid`` for an exact/nearest CAPPI level, products.py:378,386), and grids are processed as float32 (what
apply_geometry produces).
"""

from __future__ import annotations

import logging
from typing import Optional

import numpy as np

from . import _native as N
from .engine import (CAPPI, EARTH_RADIUS, EFFECTIVE_RADIUS_FACTOR, PPI, ColumnMax, ColumnMean, ColumnMin,
                     run_products, warn_all_nan)
from .geometry import GridGeometry

logger = logging.getLogger(__name__)


# ---- beam geometry helpers (host formulas on coordinate arrays) ------------------------------------------
def compute_beam_height(horizontal_distance, elevation_angle: float, radar_altitude: float = 0.0,
                        ke: float = EFFECTIVE_RADIUS_FACTOR, re: float = EARTH_RADIUS):
    """h = sqrt(r^2 + (ke Re)^2 + 2 r ke Re sin(theta)) - ke Re + h0 with r = s / max(cos(theta), 0.01)."""
    theta = np.radians(elevation_angle)
    a = ke * re
    r = horizontal_distance / np.maximum(np.cos(theta), 0.01)
    return np.sqrt(r ** 2 + a ** 2 + 2 * r * a * np.sin(theta)) - a + radar_altitude


def compute_beam_height_simple(horizontal_distance, elevation_angle: float, radar_altitude: float = 0.0,
                               ke: float = EFFECTIVE_RADIUS_FACTOR, re: float = EARTH_RADIUS):
    """h = r sin(theta) + r^2 / (2 ke Re) + h0."""
    theta = np.radians(elevation_angle)
    r = horizontal_distance / np.maximum(np.cos(theta), 0.01)
    return r * np.sin(theta) + (r ** 2) / (2 * (ke * re)) + radar_altitude


def compute_beam_height_flat(horizontal_distance, elevation_angle: float, radar_altitude: float = 0.0):
    """h = s tan(theta) + h0."""
    return horizontal_distance * np.tan(np.radians(elevation_angle)) + radar_altitude


def _horizontal_distance(geometry: GridGeometry):
    _, ny, nx = geometry.grid_shape
    y = np.linspace(geometry.grid_limits[1][0], geometry.grid_limits[1][1], ny)
    x = np.linspace(geometry.grid_limits[2][0], geometry.grid_limits[2][1], nx)
    yy, xx = np.meshgrid(y, x, indexing="ij")
    return np.sqrt(xx ** 2 + yy ** 2)


def get_beam_height_difference(geometry: GridGeometry, elevation_angle: float, radar_altitude: float = 0.0,
                               ke: float = EFFECTIVE_RADIUS_FACTOR) -> np.ndarray:
    """(curved - flat) beam height over the grid's (y, x) plane, float64."""
    s = _horizontal_distance(geometry)
    return (compute_beam_height(s, elevation_angle, radar_altitude, ke=ke)
            - compute_beam_height_flat(s, elevation_angle, radar_altitude))


def get_elevation_from_z_level(z_level: float, geometry: GridGeometry, radar_altitude: float = 0.0,
                               earth_curvature: bool = True, ke: float = EFFECTIVE_RADIUS_FACTOR) -> np.ndarray:
    """Elevation angle (degrees) whose beam reaches altitude `z_level` at each (y, x)."""
    s = np.maximum(_horizontal_distance(geometry), 1.0)
    theta = np.arctan((z_level - radar_altitude) / s)
    if earth_curvature:
        a = ke * EARTH_RADIUS
        for _ in range(5):                      # fixed-point refinement of the flat-earth guess
            r = s / np.maximum(np.cos(theta), 0.01)
            h = np.sqrt(r ** 2 + a ** 2 + 2 * r * a * np.sin(theta)) - a + radar_altitude
            theta = np.clip(theta + (z_level - h) / (r + 1), -np.pi / 2, np.pi / 2)
    return np.degrees(theta)


# ---- grid products -----------------------------------------------------------------------------------------
def _as_grid(grid):
    if N.is_device_array(grid):
        return grid
    if isinstance(grid, np.ma.MaskedArray):
        grid = np.ma.getdata(grid)
    return np.asarray(grid)


def _one(grid, shape, limits, request, have_geometry=True):
    out = run_products([grid], shape, limits, [request], have_geometry=have_geometry)[0]
    return out[0]


def constant_elevation_ppi(grid, geometry: GridGeometry, elevation_angle: float, interpolation: str = "linear",
                           earth_curvature: bool = True, ke: float = EFFECTIVE_RADIUS_FACTOR):
    """Beam-following slice; float64 for 'linear' (as NumPy promotes it), grid dtype for 'nearest'."""
    return _one(_as_grid(grid), geometry.grid_shape, geometry.grid_limits,
                PPI(elevation_angle, interpolation, earth_curvature, ke))


def constant_altitude_ppi(grid, geometry: GridGeometry, altitude: float, interpolation: str = "linear"):
    """CAPPI: level pick or 2-level blend; all-NaN (with a logged warning) outside the grid's z range."""
    return _one(_as_grid(grid), geometry.grid_shape, geometry.grid_limits, CAPPI(altitude, interpolation))


def _column(kind_cls, kind_name, grid, z_min_idx, z_max_idx, z_min_alt, z_max_alt, geometry):
    grid = _as_grid(grid)
    nz, ny, nx = (int(v) for v in grid.shape)
    limits = geometry.grid_limits if geometry is not None else ((0.0, 1.0), (0.0, 1.0), (0.0, 1.0))
    out = _one(grid, (nz, ny, nx), limits, kind_cls(z_min_idx, z_max_idx, z_min_alt, z_max_alt),
               have_geometry=geometry is not None)
    if not N.is_device_array(out):
        warn_all_nan(kind_name, out)
    return out


def column_max(grid, z_min_idx: Optional[int] = None, z_max_idx: Optional[int] = None,
               z_min_alt: Optional[float] = None, z_max_alt: Optional[float] = None,
               geometry: Optional[GridGeometry] = None):
    """COLMAX: np.nanmax over z (index limits inclusive; altitude limits need `geometry`)."""
    return _column(ColumnMax, "max", grid, z_min_idx, z_max_idx, z_min_alt, z_max_alt, geometry)


def column_min(grid, z_min_idx: Optional[int] = None, z_max_idx: Optional[int] = None,
               z_min_alt: Optional[float] = None, z_max_alt: Optional[float] = None,
               geometry: Optional[GridGeometry] = None):
    return _column(ColumnMin, "min", grid, z_min_idx, z_max_idx, z_min_alt, z_max_alt, geometry)


def column_mean(grid, z_min_idx: Optional[int] = None, z_max_idx: Optional[int] = None,
                z_min_alt: Optional[float] = None, z_max_alt: Optional[float] = None,
                geometry: Optional[GridGeometry] = None):
    return _column(ColumnMean, "mean", grid, z_min_idx, z_max_idx, z_min_alt, z_max_alt, geometry)
