"""
Adapters from a (duck-typed) pyart Radar to the flat arrays the gridding path consumes — same functions
as the reference's src/radar_grid/utils.py.  Gate id = ray * ngates + bin (C-order flattening).
"""

from __future__ import annotations

from typing import Tuple

import numpy as np


def get_gate_coordinates(radar) -> Tuple[np.ndarray, np.ndarray, np.ndarray]:
    """float32 (gate_x, gate_y, gate_z), each of shape (nrays * ngates,), metres relative to the radar."""
    return tuple(getattr(radar, k)["data"].ravel().astype("float32") for k in ("gate_x", "gate_y", "gate_z"))


def get_field_data(radar, field_name: str) -> np.ndarray:
    """Flattened float32 masked array of one field; NaN/Inf and pre-existing masks are masked."""
    return np.ma.masked_invalid(radar.fields[field_name]["data"]).ravel().astype("float32")


def get_available_fields(radar) -> list:
    return list(radar.fields.keys())


def get_radar_altitude(radar) -> float:
    return float(radar.altitude["data"][0])


def get_radar_info(radar) -> dict:
    md = radar.metadata
    return {
        "radar_name": md.get("instrument_name", "UNKNOWN"),
        "strategy": md.get("scan_id", "UNKNOWN"),
        "volume_nr": f"{int(md.get('volume_number', 0)):02d}",
        "nrays": radar.nrays,
        "ngates": radar.ngates,
        "nsweeps": radar.nsweeps,
        "total_gates": radar.nrays * radar.ngates,
        "fields": list(radar.fields.keys()),
        "range_min": float(radar.range["data"][0]),
        "range_max": float(radar.range["data"][-1]),
        "latitude": float(radar.latitude["data"][0]),
        "longitude": float(radar.longitude["data"][0]),
        "altitude": float(radar.altitude["data"][0]),
    }
