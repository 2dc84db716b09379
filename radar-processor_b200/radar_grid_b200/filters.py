"""
GateFilter / GridFilter / create_mask_from_filter with the reference's names and semantics
(src/radar_grid/filters.py).

GateFilter stays host-side Python, as in the reference: it accumulates an OR of boolean exclude masks over
the flattened gates.  In addition every threshold-type excluder is remembered as a *range rule*
(field, lo, hi), which the engine can evaluate on the GPU while it packs the gate records
(``GateFilter.fusable_rules``; ``apply_geometry[_multi]`` do that), so the cfg2-style RHOHV QC filter never has to
exist as a host array: the host mask of a range rule is only computed when ``gate_excluded`` is read.
GridFilter thresholds run on the GPU (``rg_plane_filter``); only ``apply_custom`` calls back into Python.
"""

from __future__ import annotations

import ctypes as C
import logging
from typing import Callable, List, Optional, Tuple

import numpy as np

from . import _native as N

logger = logging.getLogger(__name__)


class GateFilter:
    """Boolean exclude-mask over radar gates (True = excluded); conditions combine with OR."""

    def __init__(self, radar):
        self.radar = radar
        self.n_gates = radar.nrays * radar.ngates
        self._gate_excluded = np.zeros(self.n_gates, dtype=bool)
        self._filter_history: List[str] = []
        self._rules: List[Tuple[str, Optional[float], Optional[float]]] = []   # fusable (field, lo, hi)
        self._opaque = False            # True once a non-range excluder contributed to the mask
        self._pending: List[Callable[[], np.ndarray]] = []   # range rules whose host mask nobody has asked for yet

    # ---- views
    @property
    def gate_excluded(self) -> np.ndarray:
        # Range rules are kept as rules: apply_geometry evaluates them on the GPU (interpolate._field_inputs), so their
        # host masks are only computed if somebody reads this attribute, as the reference's callers may.
        while self._pending:
            self._gate_excluded = self._gate_excluded | self._pending.pop(0)()
        return self._gate_excluded

    @property
    def gate_included(self) -> np.ndarray:
        return ~self.gate_excluded

    def n_excluded(self) -> int:
        return self.gate_excluded.sum()

    def n_included(self) -> int:
        return (~self.gate_excluded).sum()

    def summary(self) -> str:
        ne, ni, n = self.n_excluded(), self.n_included(), self.n_gates
        head = ["GateFilter Summary:", f"  Total gates: {n:,}",
                f"  Excluded: {ne:,} ({100 * ne / n:.1f}%)", f"  Included: {ni:,} ({100 * ni / n:.1f}%)",
                f"  Filters applied ({len(self._filter_history)}):"]
        return "\n".join(head + [f"    - {h}" for h in self._filter_history])

    def __repr__(self) -> str:
        return f"GateFilter(excluded={self.n_excluded():,}/{self.n_gates:,}, filters={len(self._filter_history)})"

    # ---- plumbing
    def _get_field_data(self, field_name: str) -> np.ndarray:
        """Raw float32 values of a field, NaN/Inf left in place (filters.py:91-102)."""
        raw = np.ma.masked_invalid(self.radar.fields[field_name]["data"])
        return np.ma.getdata(raw).ravel().astype("float32")

    def _add_filter(self, mask, description: str) -> "GateFilter":
        if callable(mask):
            self._pending.append(mask)
        else:
            self._gate_excluded = self.gate_excluded | mask
        self._filter_history.append(description)
        return self

    def _has(self, field_name: str) -> bool:
        if field_name in self.radar.fields:
            return True
        logger.warning(f"Field '{field_name}' not found in radar. No gates excluded.")
        return False

    def fusable_rules(self):
        """[(field_name, lo, hi)] if this filter consists only of exclude_below/above/outside calls, else None."""
        return None if self._opaque else list(self._rules)

    # ---- thresholds on field values (NaN compares False, so invalid gates are never excluded here)
    def exclude_below(self, field_name: str, threshold: float) -> "GateFilter":
        if not self._has(field_name):
            return self
        self._rules.append((field_name, threshold, None))
        return self._add_filter(lambda: self._get_field_data(field_name) < threshold, f"{field_name} < {threshold}")

    def exclude_above(self, field_name: str, threshold: float) -> "GateFilter":
        if not self._has(field_name):
            return self
        self._rules.append((field_name, None, threshold))
        return self._add_filter(lambda: self._get_field_data(field_name) > threshold, f"{field_name} > {threshold}")

    def exclude_outside(self, field_name: str, low: float, high: float) -> "GateFilter":
        if not self._has(field_name):
            return self
        self._rules.append((field_name, low, high))

        def mask():
            v = self._get_field_data(field_name)
            return (v < low) | (v > high)
        return self._add_filter(mask, f"{field_name} outside [{low}, {high}]")

    def exclude_between(self, field_name: str, low: float, high: float) -> "GateFilter":
        if not self._has(field_name):
            return self
        v = self._get_field_data(field_name)
        self._opaque = True
        return self._add_filter((v > low) & (v < high), f"{low} < {field_name} < {high}")

    def exclude_equal(self, field_name: str, value: float, atol: float = 1e-5) -> "GateFilter":
        if not self._has(field_name):
            return self
        self._opaque = True
        return self._add_filter(np.abs(self._get_field_data(field_name) - value) < atol, f"{field_name} == {value}")

    # ---- invalid / masked data
    def exclude_invalid(self, field_name: str) -> "GateFilter":
        if not self._has(field_name):
            return self
        v = self._get_field_data(field_name)
        self._opaque = True
        return self._add_filter(np.isnan(v) | np.isinf(v), f"{field_name} invalid (NaN/Inf)")

    def exclude_masked(self, field_name: str) -> "GateFilter":
        if not self._has(field_name):
            return self
        field = self.radar.fields[field_name]["data"]
        mask = (np.ma.getmaskarray(field).ravel() if isinstance(field, np.ma.MaskedArray)
                else np.zeros(self.n_gates, dtype=bool))
        self._opaque = True
        return self._add_filter(mask, f"{field_name} masked")

    def exclude_all_invalid(self, field_name: str) -> "GateFilter":
        if not self._has(field_name):
            return self
        mask = np.ma.getmaskarray(np.ma.masked_invalid(self.radar.fields[field_name]["data"])).ravel()
        self._opaque = True
        return self._add_filter(mask, f"{field_name} all invalid (NaN/Inf/masked)")

    # ---- geometry of the scan
    def _per_gate(self, per_ray=None, per_bin=None) -> np.ndarray:
        if per_ray is not None:
            return np.repeat(per_ray, self.radar.ngates)
        return np.broadcast_to(per_bin, (self.radar.nrays, self.radar.ngates)).ravel()

    def exclude_below_altitude(self, altitude: float) -> "GateFilter":
        self._opaque = True
        return self._add_filter(self.radar.gate_altitude["data"].ravel() < altitude, f"altitude < {altitude}m")

    def exclude_above_altitude(self, altitude: float) -> "GateFilter":
        self._opaque = True
        return self._add_filter(self.radar.gate_altitude["data"].ravel() > altitude, f"altitude > {altitude}m")

    def exclude_below_range(self, range_min: float) -> "GateFilter":
        self._opaque = True
        return self._add_filter(self._per_gate(per_bin=self.radar.range["data"]) < range_min, f"range < {range_min}m")

    def exclude_above_range(self, range_max: float) -> "GateFilter":
        self._opaque = True
        return self._add_filter(self._per_gate(per_bin=self.radar.range["data"]) > range_max, f"range > {range_max}m")

    def exclude_below_elevation_angle(self, min_elev: float) -> "GateFilter":
        self._opaque = True
        return self._add_filter(self._per_gate(per_ray=self.radar.elevation["data"]) < min_elev,
                                f"elevation angle < {min_elev}°")

    def exclude_above_elevation_angle(self, max_elev: float) -> "GateFilter":
        self._opaque = True
        return self._add_filter(self._per_gate(per_ray=self.radar.elevation["data"]) > max_elev,
                                f"elevation angle > {max_elev}°")

    def exclude_outside_elevation_range(self, min_elev: float, max_elev: float) -> "GateFilter":
        e = self._per_gate(per_ray=self.radar.elevation["data"])
        self._opaque = True
        return self._add_filter((e < min_elev) | (e > max_elev), f"elevation angle outside [{min_elev}°, {max_elev}°]")

    # ---- user supplied
    def exclude_where(self, mask: np.ndarray, description: str = "custom") -> "GateFilter":
        flat = mask.ravel().astype(bool)
        if len(flat) != self.n_gates:
            raise ValueError(f"Mask size {len(flat)} doesn't match n_gates {self.n_gates}")
        self._opaque = True
        return self._add_filter(flat, description)

    def exclude_by_function(self, field_name: str, func: Callable[[np.ndarray], np.ndarray],
                            description: str = "custom function") -> "GateFilter":
        self._opaque = True
        return self._add_filter(func(self._get_field_data(field_name)), f"{field_name}: {description}")

    # ---- housekeeping
    def copy(self) -> "GateFilter":
        other = GateFilter(self.radar)
        other._gate_excluded = self._gate_excluded.copy()
        other._pending = list(self._pending)
        other._filter_history = self._filter_history.copy()
        other._rules = list(self._rules)
        other._opaque = self._opaque
        return other

    def reset(self) -> "GateFilter":
        self._gate_excluded = np.zeros(self.n_gates, dtype=bool)
        self._filter_history = []
        self._rules = []
        self._pending = []
        self._opaque = False
        return self

    def include_all(self) -> "GateFilter":
        return self.reset()

    def exclude_all(self) -> "GateFilter":
        self._pending = []
        self._gate_excluded = np.ones(self.n_gates, dtype=bool)
        self._filter_history.append("exclude all")
        self._opaque = True
        return self


def create_mask_from_filter(radar, field_name: str, gatefilter: Optional[GateFilter] = None):
    """(float32 values, combined bool mask) of one field — reference filters.py:560-598."""
    masked = np.ma.masked_invalid(radar.fields[field_name]["data"])
    data = np.ma.getdata(masked).ravel().astype("float32")
    mask = np.ma.getmaskarray(masked).ravel()
    if gatefilter is not None:
        mask = mask | gatefilter.gate_excluded
    return data, mask


class GridFilter:
    """Threshold filters on 2-D products after interpolation (reference filters.py:609-780); GPU-evaluated."""

    @staticmethod
    def _run(grid, kind: int, a: float = 0.0, b: float = 0.0, fill_value: float = np.nan):
        ctx = N.default_context()
        if N.is_device_array(grid):
            src = grid.contiguous()
            import torch
            if src.dtype not in (torch.float32, torch.float64):
                src = src.to(torch.float32)
            out = torch.empty_like(src)
            bits = 32 if src.dtype == torch.float32 else 64
            with N.torch_stream_order(ctx, True):
                N.check(N.lib().rg_plane_filter(ctx.handle, N.device_ptr(src), N.device_ptr(out), src.numel(), bits, kind,
                                                float(a), float(b), float(fill_value), N.RG_DEVICE))
            return out
        src = np.asarray(grid)
        if src.dtype not in (np.float32, np.float64):
            src = src.astype(np.float64)
        src = np.ascontiguousarray(src)
        out = np.empty_like(src)
        N.check(N.lib().rg_plane_filter(ctx.handle, N.host_ptr(src), N.host_ptr(out), src.size, src.dtype.itemsize * 8,
                                        kind, float(a), float(b), float(fill_value), N.RG_HOST))
        return out

    def apply_below(self, grid, threshold: float, fill_value: float = np.nan):
        return self._run(grid, N.RG_PF_BELOW, threshold, 0.0, fill_value)

    def apply_above(self, grid, threshold: float, fill_value: float = np.nan):
        return self._run(grid, N.RG_PF_ABOVE, threshold, 0.0, fill_value)

    def apply_outside_range(self, grid, vmin: float, vmax: float, fill_value: float = np.nan):
        return self._run(grid, N.RG_PF_OUTSIDE, vmin, vmax, fill_value)

    def apply_invalid(self, grid, fill_value: float = np.nan):
        return self._run(grid, N.RG_PF_INVALID, 0.0, 0.0, fill_value)

    def apply_custom(self, grid: np.ndarray, func: Callable[[np.ndarray], np.ndarray], fill_value: float = np.nan):
        result = grid.copy()
        result[func(result)] = fill_value
        return result
