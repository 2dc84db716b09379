"""
apply_geometry / apply_geometry_multi — the reference's interpolation API (src/radar_grid/interpolate.py:15-142)
on the CUDA engine.  ``apply_geometry_multi`` grids all fields in ONE pass over the neighbour table (up to 8
fields per index load) instead of the reference's per-field loop (interpolate.py:138-140).
"""

from __future__ import annotations

import logging
from typing import Dict, List, Optional

import numpy as np

from . import _native as N
from .engine import RangeRule, grid_fields
from .filters import GateFilter
from .geometry import GridGeometry

logger = logging.getLogger(__name__)


def _normalise_filters(additional_filters):
    """interpolate.py:50-56."""
    if not isinstance(additional_filters, list):
        if additional_filters is None:
            return []
        if isinstance(additional_filters, GateFilter):
            return [additional_filters]
        raise ValueError("additional_filters must be a list of GateFilter objects")
    return additional_filters


def _field_inputs(field_data, filters: List[GateFilter]):
    """(float32 values, uint8 mask or None) — interpolate.py:59-64: mask = getmask(field) | gate_excluded..."""
    mask = np.ma.getmask(field_data)
    for gf in filters:
        mask = mask | gf.gate_excluded
    data = np.ma.getdata(field_data)
    if mask is np.ma.nomask:
        if np.ndim(mask) == 0 and not filters:
            # the reference indexes the scalar `nomask` with the gate indices and fails (interpolate.py:75)
            raise IndexError("too many indices for array: field_data carries no mask array "
                             "(pass np.ma.masked_invalid(field) as get_field_data does)")
        mask = None
    return np.asarray(data), (None if mask is None else np.asarray(mask, dtype=bool))


def apply_geometry(geometry: GridGeometry, field_data, additional_filters: Optional[List[GateFilter]] = None,
                   fill_value: float = np.nan) -> np.ndarray:
    """Interpolate one field onto the grid: float32 array of shape geometry.grid_shape."""
    filters = _normalise_filters(additional_filters)
    data, mask = _field_inputs(field_data, filters)
    dev = geometry.device_geometry(n_gates=int(np.asarray(data).size))
    res = grid_fields(dev, [data], masks=[mask], fill_value=fill_value, want_grid=True)
    return res["grids"][0].reshape(geometry.grid_shape)


def apply_geometry_multi(geometry: GridGeometry, fields: Dict[str, np.ndarray],
                         additional_filters: Optional[Dict[str, List[GateFilter]]] = None,
                         fill_value: float = np.nan) -> Dict[str, np.ndarray]:
    """Interpolate several fields sharing one pass over the table (chunks of 8 fields)."""
    if additional_filters is None:
        additional_filters = {}
    names = list(fields.keys())
    results: Dict[str, np.ndarray] = {}
    for start in range(0, len(names), N.RG_MAX_FIELDS):
        chunk = names[start:start + N.RG_MAX_FIELDS]
        datas, masks = [], []
        for name in chunk:
            d, m = _field_inputs(fields[name], _normalise_filters(additional_filters.get(name, None)))
            datas.append(d)
            masks.append(m)
        dev = geometry.device_geometry(n_gates=int(np.asarray(datas[0]).size))
        res = grid_fields(dev, datas, masks=masks, fill_value=fill_value, want_grid=True)
        for name, g in zip(chunk, res["grids"]):
            results[name] = g.reshape(geometry.grid_shape)
    return results
