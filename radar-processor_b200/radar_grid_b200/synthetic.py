"""
Seeded synthetic radar volumes.

The reference ships no data files, so every parity and bench input is generated here, identically
for the oracle and for the CUDA path.  Geometry mirrors pyart's 4/3-earth ``antenna_to_cartesian``
(the transform that produces ``radar.gate_x/y/z`` which the reference consumes through
``get_gate_coordinates``, reference ``src/radar_grid/utils.py:12-38``):

    z = sqrt(r^2 + R^2 + 2 r R sin(e)) - R,   s = R asin(r cos(e) / (R + z)),
    x = s sin(a),  y = s cos(a),              R = 4/3 * 6 371 000 m

all in float64, cast to float32 at the end.  Gate id = (sweep * nrays + ray) * ngates + bin, i.e. the
C-order flattening of a (nsweeps*nrays, ngates) array — the same flattening the reference uses.

Pure NumPy, no device code: this module is input generation, not part of the gridding path.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Tuple

import numpy as np

_R_EFF = 4.0 / 3.0 * 6371000.0

ELEV_10 = (0.5, 1.3, 2.3, 3.5, 5.0, 6.9, 9.1, 11.8, 15.1, 19.2)
ELEV_15 = (0.5, 0.9, 1.3, 1.9, 2.3, 3.0, 3.5, 5.0, 6.9, 9.1, 11.8, 15.1, 19.2, 24.0, 30.0)


@dataclass(frozen=True)
class VolumeSpec:
    """Scan strategy + target grid for one synthetic configuration."""
    name: str
    elevations: Tuple[float, ...]
    nrays: int
    ngates: int
    gate_spacing: float
    grid_shape: Tuple[int, int, int]
    grid_limits: Tuple[Tuple[float, float], Tuple[float, float], Tuple[float, float]]
    fields: Tuple[str, ...] = ("DBZH",)
    min_radius: float = 250.0
    beam_factor: float = 0.01746
    weighting: str = "barnes2"
    toa: float = 17000.0

    @property
    def n_gates(self) -> int:
        return len(self.elevations) * self.nrays * self.ngates

    @property
    def n_voxels(self) -> int:
        nz, ny, nx = self.grid_shape
        return nz * ny * nx


ALL_FIELDS = ("DBZH", "ZDR", "RHOHV", "KDP", "VRAD")

# BASELINE.json configs.
CFG1 = VolumeSpec("cfg1", ELEV_10, 360, 480, 250.0, (20, 241, 241),
                  ((0.0, 19000.0), (-120000.0, 120000.0), (-120000.0, 120000.0)), ("DBZH",))
CFG2 = VolumeSpec("cfg2", ELEV_10, 360, 480, 250.0, (20, 241, 241),
                  ((0.0, 19000.0), (-120000.0, 120000.0), (-120000.0, 120000.0)), ("DBZH", "RHOHV"))
CFG3 = VolumeSpec("cfg3", ELEV_15, 360, 1000, 120.0, (40, 481, 481),
                  ((0.0, 19500.0), (-120000.0, 120000.0), (-120000.0, 120000.0)), ALL_FIELDS)
CFG5 = VolumeSpec("cfg5", ELEV_15, 360, 2000, 125.0, (80, 2001, 2001),
                  ((0.0, 19750.0), (-250000.0, 250000.0), (-250000.0, 250000.0)), ("DBZH",))
# Small shapes the oracle finishes in seconds (parity tests, golden fixtures, smoke()).
TINY = VolumeSpec("tiny", (0.5, 2.3, 6.9, 15.1), 36, 40, 500.0, (6, 21, 21),
                  ((0.0, 10000.0), (-20000.0, 20000.0), (-20000.0, 20000.0)), ALL_FIELDS,
                  min_radius=1200.0, beam_factor=0.09)
SMALL = VolumeSpec("small", (0.5, 1.3, 2.3, 3.5, 6.9, 11.8), 90, 120, 500.0, (10, 61, 61),
                   ((0.0, 13500.0), (-60000.0, 60000.0), (-60000.0, 60000.0)), ALL_FIELDS,
                   min_radius=800.0, beam_factor=0.06)

SPECS = {s.name: s for s in (CFG1, CFG2, CFG3, CFG5, TINY, SMALL)}


def gate_coordinates(spec: VolumeSpec) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """float32 gate_x, gate_y, gate_z of shape (n_gates,), relative to the radar."""
    elev = np.radians(np.asarray(spec.elevations, dtype=np.float64))[:, None, None]
    azim = np.radians(np.arange(spec.nrays, dtype=np.float64) * (360.0 / spec.nrays))[None, :, None]
    rng = ((np.arange(spec.ngates, dtype=np.float64) + 0.5) * spec.gate_spacing)[None, None, :]
    z = np.sqrt(rng * rng + _R_EFF * _R_EFF + 2.0 * rng * _R_EFF * np.sin(elev)) - _R_EFF
    s = _R_EFF * np.arcsin(rng * np.cos(elev) / (_R_EFF + z))
    x = s * np.sin(azim)
    y = s * np.cos(azim)
    z = np.broadcast_to(z, x.shape)
    return (x.ravel().astype(np.float32), y.ravel().astype(np.float32),
            np.ascontiguousarray(z).ravel().astype(np.float32))


def _storm_field(rng: np.random.Generator, gx, gy, gz) -> np.ndarray:
    """Sum of 6 Gaussian storm cells + noise, in dBZ (float64)."""
    x = gx.astype(np.float64)
    y = gy.astype(np.float64)
    z = gz.astype(np.float64)
    extent = float(max(np.abs(x).max(), np.abs(y).max(), 1.0))
    dbz = np.full(x.shape, -10.0)
    for _ in range(6):
        cx, cy = rng.uniform(-0.8 * extent, 0.8 * extent, size=2)
        sigma = rng.uniform(8000.0, 25000.0) * min(1.0, extent / 120000.0)
        top = rng.uniform(6000.0, 12000.0)
        peak = rng.uniform(35.0, 55.0)
        horiz = np.exp(-((x - cx) ** 2 + (y - cy) ** 2) / (2.0 * sigma * sigma))
        vert = np.exp(-np.maximum(z - 0.4 * top, 0.0) ** 2 / (2.0 * (0.35 * top) ** 2))
        dbz = np.maximum(dbz, -10.0 + (peak + 10.0) * horiz * vert)
    dbz += rng.normal(0.0, 2.0, size=x.shape)
    return np.clip(dbz, -20.0, 70.0)


def make_fields(spec: VolumeSpec, seed: int = 0, gates=None) -> Dict[str, np.ma.MaskedArray]:
    """
    Masked float32 field arrays of shape (n_gates,), as ``get_field_data`` would return them
    (``np.ma.masked_invalid(field).ravel().astype('float32')``, reference utils.py:64-66).
    """
    rng = np.random.default_rng(seed)
    gx, gy, gz = gates if gates is not None else gate_coordinates(spec)
    n = gx.shape[0]
    out: Dict[str, np.ma.MaskedArray] = {}
    for name in spec.fields:
        if name == "DBZH":
            v = _storm_field(rng, gx, gy, gz)
            v[v < -10.0] = np.nan
        elif name == "ZDR":
            v = rng.normal(0.5, 1.0, size=n)
        elif name == "RHOHV":
            v = rng.uniform(0.5, 1.0, size=n)
        elif name == "KDP":
            v = rng.normal(0.3, 0.8, size=n)
        elif name == "VRAD":
            v = rng.uniform(-30.0, 30.0, size=n)
        else:
            v = rng.normal(0.0, 1.0, size=n)
        v[rng.random(n) < 0.05] = np.nan  # speckle
        out[name] = np.ma.masked_invalid(v).astype(np.float32)
    return out


class SyntheticRadar:
    """
    Duck-typed stand-in for ``pyart.core.Radar`` carrying exactly the attributes the reference's
    ``GateFilter`` / ``utils`` read: nrays, ngates, nsweeps, fields[name]['data'] (2-D),
    gate_x/gate_y/gate_z/gate_altitude['data'], range['data'], elevation['data'], altitude['data'].
    """

    def __init__(self, spec: VolumeSpec, seed: int = 0, radar_altitude: float = 0.0):
        self.spec = spec
        self.nsweeps = len(spec.elevations)
        self.nrays = self.nsweeps * spec.nrays
        self.ngates = spec.ngates
        gx, gy, gz = gate_coordinates(spec)
        shape2d = (self.nrays, self.ngates)
        self.gate_x = {"data": gx.reshape(shape2d)}
        self.gate_y = {"data": gy.reshape(shape2d)}
        self.gate_z = {"data": gz.reshape(shape2d) + np.float32(radar_altitude)}
        self.gate_altitude = {"data": self.gate_z["data"]}
        self.range = {"data": ((np.arange(spec.ngates) + 0.5) * spec.gate_spacing).astype(np.float32)}
        self.elevation = {"data": np.repeat(np.asarray(spec.elevations, dtype=np.float32), spec.nrays)}
        self.azimuth = {"data": np.tile(np.arange(spec.nrays, dtype=np.float32) * (360.0 / spec.nrays),
                                        self.nsweeps)}
        self.altitude = {"data": np.array([radar_altitude])}
        self.latitude = {"data": np.array([-31.44])}
        self.longitude = {"data": np.array([-64.19])}
        self.metadata = {"instrument_name": "SYNTH", "scan_id": spec.name, "volume_number": 1}
        flat = make_fields(spec, seed, gates=(gx, gy, gz))
        self.fields = {k: {"data": v.reshape(shape2d)} for k, v in flat.items()}
