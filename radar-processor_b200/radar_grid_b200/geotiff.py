"""
apply_colormap_to_array — the one function of the reference's ``radar_grid/geotiff.py`` (:70-145) that sits between the
gridding path and the host-side warp / GeoTIFF writer, on the GPU: a 2-D product becomes an RGBA uint8 image before it
leaves the device (4 bytes per pixel instead of a float plane that the host would colour afterwards).  Everything else
in that module (rasterio / GDAL writing, Web-Mercator warp) stays with the reference: host I/O, out of scope.

The same arithmetic is available fused into the product epilogue: ``ColumnMax(image=ImageSpec(...))`` etc. in
``engine.grid_fields`` / ``engine.run_products``.
"""

from __future__ import annotations

from typing import Optional

import numpy as np

from . import _native as N
from .engine import ImageSpec, LevelPick, run_products


def apply_colormap_to_array(data, cmap, vmin: Optional[float] = None, vmax: Optional[float] = None,
                            fill_value: Optional[float] = None):
    """
    RGBA image (ny, nx, 4) uint8 of a 2-D array: RGB from the colormap after Normalize(vmin, vmax, clip=True), alpha 0
    for no-data pixels (``== fill_value``, or NaN when fill_value is None) and 255 otherwise (the colormap's own alpha).

    ``cmap``: matplotlib Colormap, colormap name (both need matplotlib), or the colormap's table as an (N, 4) array
    (float in [0, 1] or uint8).  ``data``: NumPy array or torch CUDA tensor, float32 or float64.
    vmin / vmax default to the minimum / maximum of the valid data (geotiff.py:117-129), 0 / 1 when there is none.
    """
    device = N.is_device_array(data)
    if device:
        import torch
        plane = data.contiguous()
        if plane.dtype != torch.float32:
            plane = plane.to(torch.float32)        # the epilogue's planes are float32; float64 inputs are colour-mapped as float32
        nodata = (plane == fill_value) if fill_value is not None else torch.isnan(plane)
        valid = plane[~nodata]
        valid = valid[~torch.isnan(valid)]
        lo = float(valid.min()) if valid.numel() else None
        hi = float(valid.max()) if valid.numel() else None
    else:
        plane = np.ascontiguousarray(np.asarray(data), dtype=np.float32)
        nodata = (plane == fill_value) if fill_value is not None else np.isnan(plane)
        valid = plane[~nodata]
        lo = float(np.nanmin(valid)) if valid.size and not np.isnan(valid).all() else None
        hi = float(np.nanmax(valid)) if valid.size and not np.isnan(valid).all() else None
    if plane.ndim != 2:
        raise ValueError("data must be a 2-D array")
    if vmin is None:
        vmin = lo if lo is not None else 0.0
    if vmax is None:
        vmax = hi if hi is not None else 1.0
    ny, nx = (int(v) for v in plane.shape)
    spec = ImageSpec(cmap, float(vmin), float(vmax), fill_value, keep_plane=False)
    _, images = run_products([plane.reshape(1, ny, nx)], (1, ny, nx), ((0.0, 1.0), (0.0, 1.0), (0.0, 1.0)),
                             [LevelPick(0, image=spec)], with_images=True)
    return images[0][0]
