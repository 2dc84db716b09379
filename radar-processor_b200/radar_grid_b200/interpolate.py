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


def _field_inputs(field_data, filters: List[GateFilter], fuse: bool = True):
    """
    (float32 values, bool mask or None, fusable rules) — interpolate.py:59-64: mask = getmask(field) | gate_excluded...

    A GateFilter made only of exclude_below / exclude_above / exclude_outside calls is not turned into a host mask:
    its (field, lo, hi) rules go to the device, where the pack kernel evaluates them on the raw field values while
    it builds the gate records (``GateFilter.fusable_rules``; same comparisons as filters.py:133-134, 156-157,
    208-209, NaN compares False).  Any other filter contributes its ``gate_excluded`` array as before.
    """
    mask = np.ma.getmask(field_data)
    rules = []
    for gf in filters:
        fr = gf.fusable_rules() if fuse and hasattr(gf, "fusable_rules") else None
        if fr is None:
            mask = mask | gf.gate_excluded
        else:
            rules += [(gf, name, lo, hi) for name, lo, hi in fr]
    data = np.ma.getdata(field_data)
    if mask is np.ma.nomask:
        if np.ndim(mask) == 0 and not filters:
            # the reference indexes the scalar `nomask` with the gate indices and fails (interpolate.py:75)
            raise IndexError("too many indices for array: field_data carries no mask array "
                             "(pass np.ma.masked_invalid(field) as get_field_data does)")
        mask = None
    return np.asarray(data), (None if mask is None else np.asarray(mask, dtype=bool)), rules


def _range_rules(per_field_rules, datas) -> List[RangeRule]:
    """RangeRule list for one fused pass.  The values of a rule are the raw values of the radar field it tests
    (GateFilter._get_field_data); when that array is bit-identical to one of the fields being gridded, the rule reads
    the gridded field's device copy and nothing extra is uploaded."""
    merged = {}
    for i, rules in enumerate(per_field_rules):
        for gf, name, lo, hi in rules:
            key = (id(gf.radar), name, lo, hi)
            if key not in merged:
                merged[key] = [gf, name, lo, hi, []]
            merged[key][4].append(i)
    out, values_of = [], {}
    for gf, name, lo, hi, idxs in merged.values():
        vk = (id(gf.radar), name)
        if vk not in values_of:
            vals = gf._get_field_data(name)
            for d in datas:                                   # alias: same gates, same bits (NaN == NaN)
                d32 = np.asarray(d).ravel()
                if d32.dtype == np.float32 and d32.shape == vals.shape and np.array_equal(d32.view(np.uint32), vals.view(np.uint32)):
                    vals = d32
                    break
            values_of[vk] = vals
        out.append(RangeRule(values_of[vk], lo=lo, hi=hi, fields=sorted(set(idxs))))
    return out


def apply_geometry(geometry: GridGeometry, field_data, additional_filters: Optional[List[GateFilter]] = None,
                   fill_value: float = np.nan) -> np.ndarray:
    """Interpolate one field onto the grid: float32 array of shape geometry.grid_shape."""
    filters = _normalise_filters(additional_filters)
    data, mask, rules = _field_inputs(field_data, filters)
    fused = _range_rules([rules], [data])
    if len(fused) > N.RG_MAX_RULES:                        # more rules than one pass takes: host masks instead
        data, mask, _ = _field_inputs(field_data, filters, fuse=False)
        fused = []
    dev = geometry.device_geometry(n_gates=int(np.asarray(data).size))
    res = grid_fields(dev, [data], masks=[mask], rules=fused, fill_value=fill_value, want_grid=True)
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
        datas, masks, rules = [], [], []
        for name in chunk:
            d, m, r = _field_inputs(fields[name], _normalise_filters(additional_filters.get(name, None)))
            datas.append(d)
            masks.append(m)
            rules.append(r)
        fused = _range_rules(rules, datas)
        if len(fused) > N.RG_MAX_RULES:                    # more rules than one pass takes: fall back to host masks
            masks = [_field_inputs(fields[name], _normalise_filters(additional_filters.get(name, None)), fuse=False)[1]
                     for name in chunk]
            fused = []
        dev = geometry.device_geometry(n_gates=int(np.asarray(datas[0]).size))
        res = grid_fields(dev, datas, masks=masks, rules=fused, fill_value=fill_value, want_grid=True)
        for name, g in zip(chunk, res["grids"]):
            results[name] = g.reshape(geometry.grid_shape)
    return results
