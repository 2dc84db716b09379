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


def get_gate_coordinates_device(radar, ctx=None):
    """
    The same three arrays as torch CUDA tensors, computed ON the device from the scan's polar description
    (radar.range, radar.azimuth, radar.elevation: a few kilobytes) instead of being read from radar.gate_x / gate_y /
    gate_z and shipped (12 bytes per gate): pyart's 4/3-earth antenna_to_cartesian in float64, rounded to float32
    (``rg_gate_coordinates``).  Feed them to ``DeviceGeometry.build``; radar.gate_x & co. are never touched, so pyart's
    lazy evaluation of them never runs.
    """
    import ctypes as C
    import torch
    from . import _native as N
    ctx = ctx or N.default_context()
    rng = np.ascontiguousarray(radar.range["data"], dtype=np.float32)
    az = np.ascontiguousarray(radar.azimuth["data"], dtype=np.float32)
    el = np.ascontiguousarray(radar.elevation["data"], dtype=np.float32)
    if az.shape != el.shape:
        raise ValueError("azimuth and elevation must have one entry per ray")
    n = az.shape[0] * rng.shape[0]
    dev = torch.device("cuda", ctx.device)
    with torch.cuda.device(dev):
        polar = torch.from_numpy(np.concatenate([rng, az, el])).to(dev)
        out = torch.empty((3, n), dtype=torch.float32, device=dev)
    nb, nr = rng.shape[0], az.shape[0]
    base = polar.data_ptr()
    with N.torch_stream_order(ctx, True):
        N.check(N.lib().rg_gate_coordinates(ctx.handle, base, base + 4 * nb, base + 4 * (nb + nr), nr, nb, N.RG_DEVICE,
                                            out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), N.RG_DEVICE))
    del polar
    return out[0], out[1], out[2]


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
