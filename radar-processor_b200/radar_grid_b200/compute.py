"""
compute_grid_geometry — same signature as the reference (src/radar_grid/compute.py:106-119), built on the GPU.

The reference's per-level worker pool, KD-tree rebuilds and temp-file exchange (compute.py:203-272) have no
equivalent here: one binning pass and two warp-per-voxel passes build the whole table in HBM
(csrc/rg_geometry.cu).  ``temp_dir`` and ``n_workers`` are accepted and validated for drop-in
compatibility but unused.
"""

from __future__ import annotations

import logging
import os
from typing import Optional, Tuple

import numpy as np

from . import _native as N
from .engine import DeviceGeometry, GeometryCache
from .geometry import GridGeometry

logger = logging.getLogger(__name__)

# Tables are built once per radar and scan strategy and reused for every volume (the reference gets that from
# save_geometry / load_geometry on disk, geometry.py:94-150): compute_grid_geometry looks the request up here first.
# RADAR_GRID_B200_GEOMETRY_CACHE_GB=0 disables it.
_CACHE = GeometryCache(max_bytes=int(float(os.environ.get("RADAR_GRID_B200_GEOMETRY_CACHE_GB", "48")) * (1 << 30)))


def geometry_cache() -> GeometryCache:
    """The module-level device cache of neighbour tables used by compute_grid_geometry."""
    return _CACHE


def compute_grid_geometry(gate_x, gate_y, gate_z, grid_shape: Tuple[int, int, int], grid_limits, temp_dir: str,
                          radar_altitude: float = 0.0, min_radius: float = 250.0, beam_factor: float = 0.01746,
                          weighting: str = "barnes2", toa: float = 17000.0,
                          n_workers: Optional[int] = None) -> GridGeometry:
    """
    Sparse mapping from grid points to radar gates, as a GridGeometry whose table lives on the GPU.

    Neighbour sets equal the reference's ({gate: d2 < r2} in float64, strict, r = max(min_radius,
    |voxel| * beam_factor), TOA cull in float32); weights are its float64 formulas rounded to float32.
    Rows are ordered cell-major / gate id, not in KD-tree order.  Like the reference (compute.py:277-284)
    the returned geometry does not carry ``radar_altitude`` (it stays 0.0).
    """
    if not os.path.isdir(temp_dir):
        raise ValueError(f"temp_dir does not exist: {temp_dir}")
    if weighting not in ("barnes2", "cressman", "nearest"):
        raise ValueError(f"Unknown weighting function: {weighting}")
    params = dict(radar_altitude=radar_altitude, min_radius=min_radius, beam_factor=beam_factor, weighting=weighting, toa=toa)
    if _CACHE.max_bytes > 0 and not N.is_device_array(gate_x):
        dev = _CACHE.get(gate_x, gate_y, gate_z, grid_shape, grid_limits, **params)
    else:
        dev = DeviceGeometry.build(gate_x, gate_y, gate_z, grid_shape, grid_limits, **params)
    info = dev.info
    logger.info(f"Radar altitude: {radar_altitude:.1f} m")
    logger.info(f"TOA filter: {info['n_gates_binned']:,} gates binned below {toa}m, "
                f"{info['n_gates'] - info['n_gates_binned']:,} excluded")
    logger.info(f"Built {info['n_pairs']:,} pairs for {info['n_rows']:,} grid points on the GPU in "
                f"{info['build_ms']:.1f} ms ({info['n_candidates']:,} distance tests)")
    return GridGeometry(grid_shape=grid_shape, grid_limits=grid_limits, toa=toa, n_gates=info["n_gates"], _device=dev)
