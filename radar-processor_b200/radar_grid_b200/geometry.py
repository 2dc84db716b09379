"""
GridGeometry container and .npz persistence — same constructor, attributes, helper methods and file
format as the reference (src/radar_grid/geometry.py:14-150), plus a lazily created device twin.

A geometry can be born on either side:
  * from NumPy CSR arrays (constructor, ``load_geometry``): uploaded to the GPU on first use;
  * on the GPU (``compute_grid_geometry``): ``indptr`` / ``gate_indices`` / ``weights`` are then exported
    to NumPy only when somebody reads them (``save_geometry``, tests), so a build-then-apply pipeline never
    moves the table over PCIe.
"""

from __future__ import annotations

import logging
import os
from typing import Optional, Tuple

import numpy as np

from . import _native as N
from .engine import DeviceGeometry

logger = logging.getLogger(__name__)


class GridGeometry:
    """Precomputed gate-to-grid mapping in CSR form (row i = voxel i, z-major / y / x order)."""

    def __init__(self, grid_shape, grid_limits, indptr=None, gate_indices=None, weights=None, toa=np.inf,
                 radar_altitude: float = 0.0, *, n_gates: Optional[int] = None, _device: Optional[DeviceGeometry] = None):
        self.grid_shape = grid_shape
        self.grid_limits = grid_limits
        self._indptr = indptr
        self._gate_indices = gate_indices
        self._weights = weights
        self.toa = toa
        self.radar_altitude = radar_altitude
        self._n_gates = n_gates
        self._devices = {}
        if _device is not None:
            self._devices[_device.ctx.device] = _device

    # -- CSR arrays (exported from the device on demand)
    def _export(self):
        if self._indptr is None:
            dev = next(iter(self._devices.values()))
            self._indptr, self._gate_indices, self._weights = dev.export_csr()

    @property
    def indptr(self) -> np.ndarray:
        self._export()
        return self._indptr

    @property
    def gate_indices(self) -> np.ndarray:
        self._export()
        return self._gate_indices

    @property
    def weights(self) -> np.ndarray:
        self._export()
        return self._weights

    # -- device twin
    def device_geometry(self, n_gates: Optional[int] = None, ctx: Optional[N.Context] = None) -> DeviceGeometry:
        ctx = ctx or N.default_context()
        dev = self._devices.get(ctx.device)
        need = n_gates if n_gates is not None else self._n_gates
        if dev is not None and (need is None or dev.n_gates == need or self._indptr is None):
            return dev
        if self._indptr is None:       # built on another device: go through the host arrays
            self._export()
        if need is None:
            need = int(self._gate_indices.max()) + 1 if len(self._gate_indices) else 0
        dev = DeviceGeometry.from_csr(self._indptr, self._gate_indices, self._weights, self.grid_shape,
                                      self.grid_limits, need, ctx=ctx)
        self._devices[ctx.device] = dev
        return dev

    # -- reference helpers (geometry.py:54-78)
    def memory_usage_mb(self) -> float:
        if self._indptr is None:
            dev = next(iter(self._devices.values()))
            isz = 4 if dev.n_pairs <= 0x7FFFFFFF else 8
            return ((dev.n_rows + 1) * isz + dev.n_pairs * 8) / 1e6
        return (self._indptr.nbytes + self._gate_indices.nbytes + self._weights.nbytes) / 1e6

    def n_grid_points(self) -> int:
        return int(np.prod(self.grid_shape))

    def n_pairs(self) -> int:
        if self._gate_indices is None:
            return next(iter(self._devices.values())).n_pairs
        return len(self._gate_indices)

    def avg_neighbors(self) -> float:
        return self.n_pairs() / self.n_grid_points()

    def z_levels(self) -> np.ndarray:
        z_min, z_max = self.grid_limits[0]
        return np.linspace(z_min, z_max, self.grid_shape[0])

    def z_levels_absolute(self) -> np.ndarray:
        return self.z_levels() + self.radar_altitude

    def __repr__(self) -> str:
        return ("GridGeometry(\n"
                f"  grid_shape={self.grid_shape},\n"
                f"  grid_limits={self.grid_limits},\n"
                f"  toa={self.toa}m,\n"
                f"  radar_altitude={self.radar_altitude}m,\n"
                f"  n_pairs={self.n_pairs():,},\n"
                f"  avg_neighbors={self.avg_neighbors():.1f},\n"
                f"  memory={self.memory_usage_mb():.1f} MB\n"
                ")")


def save_geometry(geometry: GridGeometry, filepath: str) -> None:
    """Compressed .npz with the reference's keys (geometry.py:94-118), readable by its load_geometry."""
    np.savez_compressed(
        filepath,
        grid_shape=np.array(geometry.grid_shape),
        grid_limits_z=np.array(geometry.grid_limits[0]),
        grid_limits_y=np.array(geometry.grid_limits[1]),
        grid_limits_x=np.array(geometry.grid_limits[2]),
        indptr=geometry.indptr,
        gate_indices=geometry.gate_indices,
        weights=geometry.weights,
        toa=np.array([geometry.toa]),
        radar_altitude=np.array([geometry.radar_altitude]),
    )
    logger.info(f"Saved geometry to {filepath} ({os.path.getsize(filepath) / 1e6:.1f} MB on disk)")


def load_geometry(filepath: str) -> GridGeometry:
    """Read a file written by this package or by the reference (geometry.py:121-150; legacy files without
    ``toa`` / ``radar_altitude`` get inf / 0.0)."""
    data = np.load(filepath)
    geometry = GridGeometry(
        grid_shape=tuple(data["grid_shape"]),
        grid_limits=(tuple(data["grid_limits_z"]), tuple(data["grid_limits_y"]), tuple(data["grid_limits_x"])),
        indptr=data["indptr"],
        gate_indices=data["gate_indices"],
        weights=data["weights"],
        toa=float(data["toa"][0]) if "toa" in data else np.inf,
        radar_altitude=float(data["radar_altitude"][0]) if "radar_altitude" in data else 0.0,
    )
    logger.info(f"Loaded geometry: {geometry.memory_usage_mb():.1f} MB in memory, toa={geometry.toa}m")
    return geometry
