"""
radar_grid_b200 — B200-native (sm_100a) drop-in for the gridding hot path of jgmarti84/radar-processor.

Exports the names of the reference's ``radar_grid`` package (src/radar_grid/__init__.py:39-82) that belong to
the hot path — geometry build, CSR interpolation, PPI/CAPPI/COLMAX products, gate/grid filters, pyart adapters —
with the same signatures and return types, computed by hand-written CUDA kernels behind a C-ABI library
(include/radar_grid_b200.h).  Visualisation and GeoTIFF/COG export stay with the reference (out of scope).

Extras beyond the reference API: ``DeviceGeometry`` / ``grid_fields`` (fused multi-field interpolation with
products in the epilogue, host or device buffers), ``distributed`` (volume-batch and z-slab sharding).
"""

from ._native import get_device, pinned_empty, set_device
from .geometry import GridGeometry, load_geometry, save_geometry
from .compute import compute_grid_geometry
from .interpolate import apply_geometry, apply_geometry_multi
from .utils import (get_available_fields, get_field_data, get_gate_coordinates, get_gate_coordinates_device, get_radar_altitude,
                    get_radar_info)
from .filters import GateFilter, GridFilter, create_mask_from_filter
from .products import (EARTH_RADIUS, EFFECTIVE_RADIUS_FACTOR, column_max, column_mean, column_min,
                       compute_beam_height, compute_beam_height_flat, compute_beam_height_simple,
                       constant_altitude_ppi, constant_elevation_ppi, get_beam_height_difference,
                       get_elevation_from_z_level)
from .geotiff import apply_colormap_to_array
from . import adapter
from .engine import (PreparedCall, prepare_grid_fields, ImageSpec, colormap_lut_bytes, CAPPI, PPI, LevelPick, ColumnMax, ColumnMean, ColumnMin, DeviceGeometry, GeometryCache, RangeRule, VolumePipeline, grid_fields,
                     run_products)

__version__ = "0.1.0"

__all__ = [
    "GridGeometry", "save_geometry", "load_geometry", "compute_grid_geometry",
    "apply_geometry", "apply_geometry_multi",
    "get_gate_coordinates", "get_gate_coordinates_device", "get_field_data", "get_available_fields", "get_radar_info", "get_radar_altitude",
    "GateFilter", "GridFilter", "create_mask_from_filter",
    "constant_altitude_ppi", "constant_elevation_ppi", "column_max", "column_min", "column_mean",
    "get_elevation_from_z_level", "get_beam_height_difference", "compute_beam_height",
    "compute_beam_height_flat", "compute_beam_height_simple", "EARTH_RADIUS", "EFFECTIVE_RADIUS_FACTOR",
    # engine-level API
    "DeviceGeometry", "grid_fields", "prepare_grid_fields", "PreparedCall", "run_products", "RangeRule", "VolumePipeline", "GeometryCache",
    "ColumnMax", "ColumnMin", "ColumnMean", "CAPPI", "PPI", "LevelPick", "ImageSpec", "colormap_lut_bytes",
    "apply_colormap_to_array", "adapter",
    "set_device", "get_device", "pinned_empty",
]
