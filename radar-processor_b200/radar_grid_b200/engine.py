"""
Device-side engine: the thin host layer between the reference-shaped Python API and the C ABI.

``DeviceGeometry`` owns a device-resident neighbour table (one z-slab); ``grid_fields`` runs one fused
pass — gate-mask fusion + record packing (K4), CSR gather-weighted mean for up to 8 fields per index load
(K5) and the COLMAX / CAPPI / PPI epilogue (K6) — on host (NumPy) or device (torch CUDA) buffers.

The product classes below hold only the *host decisions* the reference makes in Python before touching
the grid (which level, which weights, which dtype); they mirror the reference line by line so that the
kernels receive exactly the scalars NumPy would have used.
"""

from __future__ import annotations

import ctypes as C
import logging
import warnings
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _native as N

logger = logging.getLogger(__name__)

EARTH_RADIUS = 6371000.0
EFFECTIVE_RADIUS_FACTOR = 4.0 / 3.0


def _grid_spec(grid_shape, grid_limits, z_range=None) -> N.GridSpec:
    nz, ny, nx = (int(v) for v in grid_shape)
    z0, z1 = (0, nz) if z_range is None else (int(z_range[0]), int(z_range[1]))
    (zmin, zmax), (ymin, ymax), (xmin, xmax) = grid_limits
    return N.GridSpec(nz, ny, nx, z0, z1, 0, float(zmin), float(zmax), float(ymin), float(ymax),
                      float(xmin), float(xmax))


class DeviceGeometry:
    """Device-resident CSR neighbour table of one z-slab (``rg_geometry``)."""

    def __init__(self, handle, ctx: N.Context, grid_shape, grid_limits, z_range):
        self._h = handle
        self.ctx = ctx
        self.grid_shape = tuple(int(v) for v in grid_shape)
        self.grid_limits = grid_limits
        self.z_range = (0, self.grid_shape[0]) if z_range is None else (int(z_range[0]), int(z_range[1]))
        self._info = None

    # -- constructors
    @classmethod
    def build(cls, gate_x, gate_y, gate_z, grid_shape, grid_limits, radar_altitude=0.0, min_radius=250.0,
              beam_factor=0.01746, weighting="barnes2", toa=17000.0, z_range=None,
              ctx: Optional[N.Context] = None) -> "DeviceGeometry":
        """GPU build (K1-K3); replaces reference compute.py:106-284."""
        if weighting not in N.RG_W:
            raise ValueError(f"Unknown weighting function: {weighting}")
        ctx = ctx or N.default_context()
        spec = _grid_spec(grid_shape, grid_limits, z_range)
        h = C.c_void_p()
        if N.is_device_array(gate_x):
            ptrs = [N.device_ptr(a) for a in (gate_x, gate_y, gate_z)]
            n = int(gate_x.numel()) if hasattr(gate_x, "numel") else int(np.prod(gate_x.shape))
            space = N.RG_DEVICE
            keep = (gate_x, gate_y, gate_z)
        else:
            keep = tuple(np.ascontiguousarray(np.asarray(a).ravel(), dtype=np.float32) for a in (gate_x, gate_y, gate_z))
            if not (keep[0].shape == keep[1].shape == keep[2].shape):
                raise ValueError("gate_x, gate_y and gate_z must have the same length")
            ptrs = [N.host_ptr(a) for a in keep]
            n = keep[0].shape[0]
            space = N.RG_HOST
        with N.torch_stream_order(ctx, space == N.RG_DEVICE):
            N.check(N.lib().rg_geometry_build(ctx.handle, ptrs[0], ptrs[1], ptrs[2], n, space, C.byref(spec),
                                              float(radar_altitude), float(min_radius), float(beam_factor),
                                              N.RG_W[weighting], float(toa), C.byref(h)))
        del keep
        return cls(h, ctx, grid_shape, grid_limits, z_range)

    @staticmethod
    def level_pairs(gate_x, gate_y, gate_z, grid_shape, grid_limits, radar_altitude=0.0, min_radius=250.0,
                    beam_factor=0.01746, toa=17000.0, column_stride: int = 4, z_range=None,
                    ctx: Optional[N.Context] = None) -> np.ndarray:
        """Pair count of every z-level (int64, scaled from every ``column_stride``-th column; exact for stride 1) without
        building a table: the weights ``distributed.zslab_ranges`` balances z-slabs with."""
        ctx = ctx or N.default_context()
        spec = _grid_spec(grid_shape, grid_limits, z_range)
        out = np.zeros(spec.z_end - spec.z_begin, dtype=np.int64)
        if N.is_device_array(gate_x):
            keep, space = (gate_x, gate_y, gate_z), N.RG_DEVICE
            ptrs = [N.device_ptr(a) for a in keep]
            n = int(gate_x.numel())
        else:
            keep = tuple(np.ascontiguousarray(np.asarray(a).ravel(), dtype=np.float32) for a in (gate_x, gate_y, gate_z))
            ptrs, space, n = [N.host_ptr(a) for a in keep], N.RG_HOST, keep[0].shape[0]
        with N.torch_stream_order(ctx, space == N.RG_DEVICE):
            N.check(N.lib().rg_geometry_level_pairs(ctx.handle, ptrs[0], ptrs[1], ptrs[2], n, space, C.byref(spec),
                                                    float(radar_altitude), float(min_radius), float(beam_factor), float(toa),
                                                    int(column_stride), out.ctypes.data_as(C.POINTER(C.c_int64))))
        del keep
        return out

    @classmethod
    def from_csr(cls, indptr, gate_indices, weights, grid_shape, grid_limits, n_gates: int, z_range=None,
                 ctx: Optional[N.Context] = None) -> "DeviceGeometry":
        """Upload an existing table (row order preserved), e.g. one built and saved by the reference."""
        ctx = ctx or N.default_context()
        spec = _grid_spec(grid_shape, grid_limits, z_range)
        indptr = np.ascontiguousarray(indptr)
        if indptr.dtype.itemsize == 8:
            indptr = indptr.astype(np.int64, copy=False)
            bits = 64
        else:
            indptr = indptr.astype(np.int32, copy=False)
            bits = 32
        n_rows = (spec.z_end - spec.z_begin) * spec.ny * spec.nx
        if indptr.shape[0] != n_rows + 1:
            raise ValueError(f"indptr has {indptr.shape[0]} entries, expected {n_rows + 1}")
        idx = np.ascontiguousarray(gate_indices, dtype=np.int32)
        w = np.ascontiguousarray(weights, dtype=np.float32)
        if idx.shape[0] != w.shape[0] or idx.shape[0] != int(indptr[-1]):
            raise ValueError("gate_indices / weights length must equal indptr[-1]")
        h = C.c_void_p()
        N.check(N.lib().rg_geometry_from_csr(ctx.handle, C.byref(spec), N.host_ptr(indptr), bits,
                                             N.host_ptr(idx) if idx.size else None,
                                             N.host_ptr(w) if w.size else None, int(n_gates), N.RG_HOST, C.byref(h)))
        return cls(h, ctx, grid_shape, grid_limits, z_range)

    # -- queries
    @property
    def info(self) -> dict:
        if self._info is None:
            gi = N.GeometryInfo()
            N.check(N.lib().rg_geometry_get_info(self._h, C.byref(gi)))
            self._info = {k: getattr(gi, k) for k, _ in N.GeometryInfo._fields_ if k != "grid"}
        return self._info

    @property
    def duo_slots(self) -> int:
        """> 0: slots of the column-pair copy the multi-field pass reads; 0: not built yet; < 0: not available for this table."""
        n = C.c_int64(0)
        N.check(N.lib().rg_geometry_duo_slots(self._h, C.byref(n)))
        return int(n.value)

    @property
    def n_pairs(self) -> int:
        return int(self.info["n_pairs"])

    @property
    def n_rows(self) -> int:
        return int(self.info["n_rows"])

    @property
    def n_gates(self) -> int:
        return int(self.info["n_gates"])

    @property
    def slab_shape(self) -> Tuple[int, int, int]:
        return (self.z_range[1] - self.z_range[0], self.grid_shape[1], self.grid_shape[2])

    def export_csr(self, index_dtype=None):
        """(indptr, gate_indices int32, weights float32) as NumPy arrays — the reference's GridGeometry arrays."""
        n_rows, n_pairs = self.n_rows, self.n_pairs
        if index_dtype is None:
            index_dtype = np.int32 if n_pairs <= 0x7FFFFFFF else np.int64
        indptr = np.empty(n_rows + 1, dtype=index_dtype)
        idx = np.empty(n_pairs, dtype=np.int32)
        w = np.empty(n_pairs, dtype=np.float32)
        N.check(N.lib().rg_geometry_export_csr(self.ctx.handle, self._h, N.host_ptr(indptr), indptr.dtype.itemsize * 8,
                                               N.host_ptr(idx) if n_pairs else None,
                                               N.host_ptr(w) if n_pairs else None, N.RG_HOST))
        return indptr, idx, w

    def close(self):
        if getattr(self, "_h", None):
            N.lib().rg_geometry_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ------------------------------------------------------------------------------------------------------
# 2-D product requests: host-side decisions, mirrored from the reference
# ------------------------------------------------------------------------------------------------------
def resolve_z_limits(nz, grid_limits, z_min_idx=None, z_max_idx=None, z_min_alt=None, z_max_alt=None,
                     have_geometry=True):
    """Index limits of column_max/min/mean — reference products.py:462-487."""
    if z_min_alt is not None or z_max_alt is not None:
        if not have_geometry:
            raise ValueError("geometry is required when using altitude-based limits")
        z_coords = np.linspace(grid_limits[0][0], grid_limits[0][1], nz)
        if z_min_alt is not None:
            z_min_idx = np.searchsorted(z_coords, z_min_alt)
        if z_max_alt is not None:
            z_max_idx = np.searchsorted(z_coords, z_max_alt, side="right") - 1
    if z_min_idx is None:
        z_min_idx = 0
    if z_max_idx is None:
        z_max_idx = nz - 1
    return max(0, int(z_min_idx)), min(nz - 1, int(z_max_idx))


def colormap_lut_bytes(cmap) -> np.ndarray:
    """
    (N + 3, 4) uint8 table the image epilogue indexes: rows 0..N-1 the colours, then matplotlib's under / over / bad
    rows, each ``(lut * 255).astype(uint8)`` exactly as reference geotiff.py:137 scales the colormap's output.

    ``cmap`` may be a matplotlib Colormap (its ``_lut`` is used), a colormap name (needs matplotlib), or an array: float
    in [0, 1] or uint8, with N rows (under / over default to the first / last colour, bad to transparent black, as
    matplotlib's defaults) or N + 3 rows when ``cmap`` is a ``(table, "with_extremes")`` tuple.
    """
    with_ext = False
    if isinstance(cmap, tuple) and len(cmap) == 2 and cmap[1] == "with_extremes":
        cmap, with_ext = cmap[0], True
    if isinstance(cmap, str):
        try:
            import matplotlib.pyplot as plt
        except ImportError as e:                            # the reference calls plt.get_cmap(cmap) (geotiff.py:106-107)
            raise ImportError("colormap names need matplotlib; pass the colormap's table (an (N, 4) array) instead") from e
        cmap = plt.get_cmap(cmap)
    if hasattr(cmap, "_lut") and hasattr(cmap, "N"):
        if not getattr(cmap, "_isinit", True):
            cmap._init()
        table, with_ext = np.asarray(cmap._lut), True
    else:
        table = np.asarray(cmap)
    if table.ndim != 2 or table.shape[1] != 4:
        raise ValueError("colormap table must have shape (N, 4) or (N + 3, 4)")
    if table.dtype != np.uint8:
        table = (table.astype(np.float64) * 255).astype(np.uint8)
    if not with_ext:
        table = np.concatenate([table, table[:1], table[-1:], np.zeros((1, 4), np.uint8)])
    if table.shape[0] < 4:
        raise ValueError("colormap table needs at least one colour")
    return np.ascontiguousarray(table)


@dataclass
class ImageSpec:
    """RGBA form of a 2-D product, produced by the epilogue that finishes the product (rg_image): GridFilter thresholds
    (reference filters.py:631-746; ``filters`` = [("below", thr, fill) | ("above", thr, fill) | ("outside", lo, hi, fill)
    | ("invalid", fill)], applied in order), then reference geotiff.py:70-145 ``apply_colormap_to_array`` with explicit
    vmin / vmax.  ``keep_plane=False`` drops the float plane: only ny*nx*4 bytes per field leave the GPU."""
    cmap: object
    vmin: float
    vmax: float
    fill_value: Optional[float] = None
    filters: Sequence[tuple] = ()
    keep_plane: bool = True

    def struct(self, lut_ptr: int, n_entries: int, out_ptr: int) -> "N.Image":
        if self.vmin > self.vmax:
            raise ValueError("minvalue must be less than or equal to maxvalue")
        if len(self.filters) > N.RG_MAX_IMAGE_FILTERS:
            raise ValueError(f"at most {N.RG_MAX_IMAGE_FILTERS} grid filters per image")
        im = N.Image()
        im.n_filters = len(self.filters)
        for i, f in enumerate(self.filters):
            kind = {"below": N.RG_PF_BELOW, "above": N.RG_PF_ABOVE, "outside": N.RG_PF_OUTSIDE, "invalid": N.RG_PF_INVALID,
                    "below_equal": N.RG_PF_BELOW_EQUAL}[f[0]]
            im.filter_kind[i] = kind
            if kind == N.RG_PF_OUTSIDE:
                im.filter_a[i], im.filter_b[i], im.filter_fill[i] = float(f[1]), float(f[2]), float(f[3]) if len(f) > 3 else np.nan
            elif kind == N.RG_PF_INVALID:
                im.filter_fill[i] = float(f[1]) if len(f) > 1 else np.nan
            else:
                im.filter_a[i], im.filter_fill[i] = float(f[1]), float(f[2]) if len(f) > 2 else np.nan
        im.vmin, im.vmax = float(self.vmin), float(self.vmax)
        im.has_fill_value = int(self.fill_value is not None)
        im.fill_value = 0.0 if self.fill_value is None else float(self.fill_value)
        im.lut_entries = n_entries
        im.lut = lut_ptr
        im.out = out_ptr
        return im


@dataclass
class ColumnMax:
    z_min_idx: Optional[int] = None
    z_max_idx: Optional[int] = None
    z_min_alt: Optional[float] = None
    z_max_alt: Optional[float] = None
    partial: bool = False      # z-slab term: "no data in this slab" is -inf (+inf for ColumnMin), see distributed.zslab_products
    image: Optional[ImageSpec] = None
    kind = N.RG_PROD_COLMAX
    name = "column_max"

    def resolve(self, grid_shape, grid_limits, have_geometry=True):
        lo, hi = resolve_z_limits(grid_shape[0], grid_limits, self.z_min_idx, self.z_max_idx, self.z_min_alt,
                                  self.z_max_alt, have_geometry)
        if hi < lo:
            raise ValueError("zero-size array to reduction operation has no identity (empty z range)")
        return N.Product(kind=self.kind, z_lo=lo, z_hi=hi, partial=int(self.partial)), np.float32


class ColumnMin(ColumnMax):
    kind = N.RG_PROD_COLMIN
    name = "column_min"


class ColumnMean(ColumnMax):
    kind = N.RG_PROD_COLMEAN
    name = "column_mean"

    def resolve(self, grid_shape, grid_limits, have_geometry=True):
        lo, hi = resolve_z_limits(grid_shape[0], grid_limits, self.z_min_idx, self.z_max_idx, self.z_min_alt,
                                  self.z_max_alt, have_geometry)
        if hi < lo:
            # np.nanmean of an empty slice (products.py:578-580) is an all-NaN plane plus a RuntimeWarning, not an error
            warnings.warn("Mean of empty slice", RuntimeWarning, stacklevel=3)
            return None, np.float32
        return N.Product(kind=self.kind, z_lo=lo, z_hi=hi), np.float32


@dataclass
class CAPPI:
    """constant_altitude_ppi — reference products.py:317-415."""
    altitude: float
    interpolation: str = "linear"
    partial: bool = False      # z-slab term of the blend (sum over the owned levels of weight * level), see distributed.zslab_products
    image: Optional[ImageSpec] = None
    name = "cappi"

    def resolve(self, grid_shape, grid_limits, have_geometry=True):
        pr, dt = self._resolve(grid_shape, grid_limits)
        if pr is not None and self.partial:
            pr.partial = 1
            if pr.mode == N.RG_BLEND_F64:
                dt = np.float64          # the float64 blend is rounded to float32 only after the ranks' terms are added
        return pr, dt

    def _resolve(self, grid_shape, grid_limits):
        nz = grid_shape[0]
        z_min, z_max = grid_limits[0]
        altitude = self.altitude
        if self.interpolation not in ("linear", "nearest"):
            raise ValueError(f"Unknown interpolation method: {self.interpolation}")
        z_coords = np.linspace(z_min, z_max, nz, dtype="float32")
        if altitude < z_min or altitude > z_max:                      # :370-372
            logger.warning(f"Altitude {altitude}m is outside grid range [{z_min}, {z_max}]m")
            return None, np.float32                                   # all-NaN plane, no kernel needed
        if self.interpolation == "nearest":                           # :377-378
            return N.Product(kind=N.RG_PROD_LEVEL, mode=N.RG_BLEND_PICK,
                             z_lo=int(np.argmin(np.abs(z_coords - altitude)))), np.float32
        hit = np.isclose(z_coords, altitude, rtol=1e-6)               # :382-386
        if np.any(hit):
            return N.Product(kind=N.RG_PROD_LEVEL, mode=N.RG_BLEND_PICK, z_lo=int(np.where(hit)[0][0])), np.float32
        z_step = (z_max - z_min) / (nz - 1) if nz > 1 else 1.0        # :389
        z_frac = (altitude - z_min) / z_step
        z_low = int(np.floor(z_frac))
        z_high = z_low + 1
        if z_low < 0:                                                 # :397-400
            return N.Product(kind=N.RG_PROD_LEVEL, mode=N.RG_BLEND_PICK, z_lo=0), np.float32
        if z_high >= nz:
            return N.Product(kind=N.RG_PROD_LEVEL, mode=N.RG_BLEND_PICK, z_lo=nz - 1), np.float32
        weight_high = z_frac - z_low
        weight_low = 1.0 - weight_high
        # NumPy (NEP 50): a Python-float weight times a float32 grid stays float32, a np.float64 weight (limits
        # that came out of load_geometry) promotes the blend to float64 before the final float32 cast (:411-412)
        strong = isinstance(weight_high, np.floating)
        return N.Product(kind=N.RG_PROD_LEVEL, mode=N.RG_BLEND_F64 if strong else N.RG_BLEND_F32, z_lo=z_low,
                         z_hi=z_high, w_lo=float(weight_low), w_hi=float(weight_high)), np.float32


@dataclass
class LevelPick:
    """One level of the grid by its GLOBAL index (0 .. nz-1).  Not a reference product by itself: it is the term a
    z-slab contributes to a CAPPI whose two levels live in different slabs (distributed.cappi_zslab)."""
    level: int
    partial: bool = False      # z-slab term: the level where the slab owns it, -0.0 elsewhere
    image: Optional[ImageSpec] = None
    name = "level"

    def resolve(self, grid_shape, grid_limits, have_geometry=True):
        if not (0 <= int(self.level) < grid_shape[0]):
            raise ValueError(f"level {self.level} outside the grid (nz = {grid_shape[0]})")
        return N.Product(kind=N.RG_PROD_LEVEL, mode=N.RG_BLEND_PICK, z_lo=int(self.level), partial=int(self.partial)), np.float32


@dataclass
class PPI:
    """constant_elevation_ppi — reference products.py:168-314."""
    elevation_angle: float
    interpolation: str = "linear"
    earth_curvature: bool = True
    ke: float = EFFECTIVE_RADIUS_FACTOR
    partial: bool = False      # z-slab term of the per-pixel blend, see distributed.zslab_products
    image: Optional[ImageSpec] = None
    name = "ppi"

    def resolve(self, grid_shape, grid_limits, have_geometry=True):
        if self.interpolation == "closest_level":
            # radar_processor's own collapse (processor.py:513-528): np.sin(np.deg2rad(elevation_deg)), Re = 8.49e6 in the kernel
            return N.Product(kind=N.RG_PROD_BEAM, mode=2, partial=int(self.partial),
                             sin_elev=float(np.sin(np.deg2rad(self.elevation_angle)))), np.float32
        if self.interpolation not in ("linear", "nearest"):
            raise ValueError(f"Unknown interpolation method: {self.interpolation}")
        elevation_rad = np.radians(self.elevation_angle)              # products.py:70
        ke_re = self.ke * EARTH_RADIUS
        pr = N.Product(kind=N.RG_PROD_BEAM, mode=0 if self.interpolation == "linear" else 1,
                       earth_curvature=1 if self.earth_curvature else 0, partial=int(self.partial),
                       sin_elev=float(np.sin(elevation_rad)),
                       cos_elev_clamped=float(np.maximum(np.cos(elevation_rad), 0.01)),
                       tan_elev=float(np.tan(elevation_rad)), ke_re=float(ke_re), ke_re_sq=float(ke_re ** 2))
        return pr, (np.float64 if self.interpolation == "linear" else np.float32)


# ------------------------------------------------------------------------------------------------------
# fused gate-mask rules
# ------------------------------------------------------------------------------------------------------
@dataclass
class RangeRule:
    """Exclude gates of `fields` where values < lo or values > hi (either bound optional)."""
    values: object                       # float32 array of n_gates (host or device)
    lo: Optional[float] = None
    hi: Optional[float] = None
    fields: Optional[Sequence[int]] = None   # indices into the gridded field list; None = all


def _as_f32_host(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a).ravel(), dtype=np.float32)


def _check_buffer(x, device: bool, n: int, dtype, what: str):
    """A caller-supplied buffer goes to the library as a raw pointer: refuse anything that is not exactly the C-contiguous
    array of `n` elements of `dtype` on the side (host / device) the call runs on."""
    if device:
        import torch
        want = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64, np.dtype(np.uint8): torch.uint8}[np.dtype(dtype)]
        if not (hasattr(x, "is_cuda") and x.is_cuda):
            raise ValueError(f"{what}: expected a CUDA tensor (the fields of this call are device buffers)")
        ok_dtype = x.dtype == want or (want == torch.uint8 and x.dtype == torch.bool)
        if not ok_dtype or not x.is_contiguous() or x.numel() != n:
            raise ValueError(f"{what}: expected a contiguous {want} tensor of {n} elements, got {x.dtype} {tuple(x.shape)}"
                             f"{'' if x.is_contiguous() else ' (not contiguous)'}")
    else:
        if N.is_device_array(x) or not isinstance(x, np.ndarray):
            raise ValueError(f"{what}: expected a NumPy array (the fields of this call are host buffers)")
        if x.dtype != np.dtype(dtype) or not x.flags.c_contiguous or x.size != n or not x.flags.writeable:
            raise ValueError(f"{what}: expected a writable C-contiguous {np.dtype(dtype)} array of {n} elements, got {x.dtype} {x.shape}")


def _alloc_like(device: bool, shape, dtype, ref=None):
    if not device:
        return np.empty(shape, dtype=dtype)
    import torch
    tdt = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}[np.dtype(dtype)]
    return torch.empty(shape, dtype=tdt, device=ref.device)


def _ptr(x, device: bool) -> int:
    return N.device_ptr(x) if device else N.host_ptr(x)


def _fill_nan(x, device):
    if device:
        x.fill_(float("nan"))
    else:
        x.fill(np.nan)


def _attach_image(request, pr, device: bool, F: int, ny: int, nx: int, ref, keep: list):
    """Allocate the RGBA output of an image request, stage its LUT and hook both onto the product struct."""
    spec = getattr(request, "image", None)
    if spec is None or pr is None:
        return None
    table = colormap_lut_bytes(spec.cmap)
    if device:
        import torch
        lut = torch.from_numpy(table).to(ref.device)
        out = torch.empty((F, ny, nx, 4), dtype=torch.uint8, device=ref.device)
    else:
        lut, out = table, np.empty((F, ny, nx, 4), dtype=np.uint8)
    im = spec.struct(_ptr(lut, device), table.shape[0] - 3, _ptr(out, device))
    pr.image = C.pointer(im)
    keep += [lut, im]
    return out


def _nodata_image(request, device: bool, F: int, ny: int, nx: int, ref):
    """Image of an all-NaN plane (a CAPPI outside the grid): every pixel is the colormap's "bad" entry."""
    spec = getattr(request, "image", None)
    if spec is None:
        return None
    table = colormap_lut_bytes(spec.cmap)
    px = table[-1].copy()
    if spec.fill_value is None:
        px[3] = 0                                            # NaN is the no-data value: transparent
    img = np.broadcast_to(px, (F, ny, nx, 4)).copy()
    if device:
        import torch
        return torch.from_numpy(img).to(ref.device)
    return img


def run_products(grids: Sequence, grid_shape, grid_limits, products: Sequence, z_range=None,
                 ctx: Optional[N.Context] = None, have_geometry=True, with_images: bool = False):
    """Stand-alone K6 on existing 3-D grids (NumPy or torch CUDA float32, shape slab x ny x nx).  With ``with_images``
    the return value is ``(planes, images)``: images[i] is the (F, ny, nx, 4) uint8 RGBA form of request i or None."""
    ctx = ctx or N.default_context()
    device = N.is_device_array(grids[0])
    F = len(grids)
    nz, ny, nx = grid_shape
    spec = _grid_spec(grid_shape, grid_limits, z_range)
    if device:
        held = [g.contiguous() for g in grids]
    else:
        held = [np.ascontiguousarray(np.asarray(g), dtype=np.float32) for g in grids]
    n_rows = (spec.z_end - spec.z_begin) * ny * nx
    for g in held:
        if int(np.prod(g.shape)) != n_rows:
            raise ValueError(f"grid has {int(np.prod(g.shape))} voxels, expected {n_rows}")
    outs, structs, images, keep = [], [], [], []
    for p in products:
        pr, dt = p.resolve(grid_shape, grid_limits, have_geometry)
        img_spec = getattr(p, "image", None)
        out = None if (img_spec is not None and not img_spec.keep_plane) else _alloc_like(device, (F, ny, nx), dt, held[0])
        if pr is None:
            if out is not None:
                _fill_nan(out, device)
            images.append(_nodata_image(p, device, F, ny, nx, held[0]))
        else:
            pr.out = None if out is None else _ptr(out, device)
            images.append(_attach_image(p, pr, device, F, ny, nx, held[0], keep))
            structs.append(pr)
        outs.append(out)
    if structs:
        gp = (C.c_void_p * F)(*[_ptr(g, device) for g in held])
        arr = (N.Product * len(structs))(*structs)
        with N.torch_stream_order(ctx, device):
            N.check(N.lib().rg_products(ctx.handle, C.byref(spec), F, gp, len(structs), arr,
                                        N.RG_DEVICE if device else N.RG_HOST))
    del keep
    return (outs, images) if with_images else outs


def grid_fields(geom: DeviceGeometry, fields: Sequence, masks: Optional[Sequence] = None,
                mask_invalid: Sequence[bool] | bool = False, rules: Sequence[RangeRule] = (),
                fill_value: float = np.nan, want_grid: Sequence[bool] | bool = True, products: Sequence = (),
                reference_order=False, ctx: Optional[N.Context] = None,
                out_grids: Optional[Sequence] = None, out_products: Optional[Sequence] = None,
                _prepare_only: bool = False) -> Dict[str, object]:
    """
    One fused pass over the neighbour table for ``len(fields)`` (<= 8) fields.

    fields        list of float32 arrays of n_gates (NumPy -> the library copies H2D/D2H itself; torch CUDA
                  tensors -> zero-copy, asynchronous on the context's stream)
    masks         per-field boolean/uint8 arrays (True = excluded) or None
    mask_invalid  per-field flag: treat NaN/Inf as masked (np.ma.masked_invalid semantics)
    rules         RangeRule list, evaluated on the device and OR-ed into the masks (GateFilter fusion)
    want_grid     per-field flag: materialise the 3-D grid
    products      ColumnMax / ColumnMin / ColumnMean / CAPPI / PPI requests, fused in the epilogue
    reference_order  False: fast fused path; True: the reference's summation order (bit-exact, un-fused);
                  "nearest_gate": nearest-gate gridding on a table built with weighting="dist2" (adapter.py)
    out_grids / out_products  optional preallocated outputs (e.g. pinned host arrays) of the right shape/dtype

    Returns {"grids": [array or None per field], "products": [array (F, ny, nx) per request, None for an image-only
    request], "images": [uint8 array (F, ny, nx, 4) per request with an ``image`` spec, else None]}.
    """
    ctx = ctx or geom.ctx
    F = len(fields)
    if F < 1 or F > N.RG_MAX_FIELDS:
        raise ValueError(f"between 1 and {N.RG_MAX_FIELDS} fields per pass")
    device = N.is_device_array(fields[0])
    G = geom.n_gates
    if device:
        import torch
        fheld = [f.contiguous() for f in fields]
        for f in fheld:
            if not (hasattr(f, "is_cuda") and f.is_cuda) or f.dtype != torch.float32:
                raise ValueError("device fields must be float32 CUDA tensors (convert before the call: the library reads raw float32)")
            if f.numel() != G:
                raise ValueError(f"field has {f.numel()} gates, geometry expects {G}")
    else:
        fheld = [_as_f32_host(f) for f in fields]
        for f in fheld:
            if f.shape[0] != G:
                raise ValueError(f"field has {f.shape[0]} gates, geometry expects {G}")
    mheld = [None] * F
    if masks is not None:
        for i, m in enumerate(masks):
            if m is None:
                continue
            if device:
                import torch
                if not (hasattr(m, "is_cuda") and m.is_cuda) or m.dtype not in (torch.bool, torch.uint8) or m.numel() != G:
                    raise ValueError("device masks must be bool / uint8 CUDA tensors with one entry per gate")
                mheld[i] = m.contiguous().view(torch.uint8) if m.dtype == torch.bool else m.contiguous()
            else:
                mheld[i] = np.ascontiguousarray(np.asarray(m).ravel()).astype(np.uint8, copy=False) \
                    if np.asarray(m).dtype != np.bool_ else np.ascontiguousarray(np.asarray(m).ravel()).view(np.uint8)
                if mheld[i].shape[0] != G:
                    raise ValueError("mask length does not match the number of gates")
    if isinstance(mask_invalid, (bool, np.bool_)):
        mask_invalid = [bool(mask_invalid)] * F
    inv_bits = sum(1 << i for i, b in enumerate(mask_invalid) if b)
    if isinstance(want_grid, (bool, np.bool_)):
        want_grid = [bool(want_grid)] * F

    nzs, ny, nx = geom.slab_shape
    grids = []
    for i in range(F):
        if not want_grid[i]:
            grids.append(None)
        elif out_grids is not None and out_grids[i] is not None:
            g = out_grids[i]
            _check_buffer(g, device, nzs * ny * nx, np.float32, f"out_grids[{i}]")
            grids.append(g)
        else:
            grids.append(_alloc_like(device, (nzs, ny, nx), np.float32, fheld[0]))

    rheld, rstructs = [], []
    for r in rules:
        if r.lo is None and r.hi is None:
            continue
        vals = r.values.contiguous() if device else _as_f32_host(r.values)
        for i, f in enumerate(fheld):            # alias detection keeps a single device copy
            if (not device) and vals is not f and vals.shape == f.shape and np.shares_memory(vals, f):
                vals = f
        rheld.append(vals)
        bits = (1 << F) - 1 if r.fields is None else sum(1 << int(i) for i in r.fields)
        rstructs.append(N.QcRule(values=_ptr(vals, device), lo=0.0 if r.lo is None else float(r.lo),
                                 hi=0.0 if r.hi is None else float(r.hi), use_lo=int(r.lo is not None),
                                 use_hi=int(r.hi is not None), field_bits=bits))
    if len(rstructs) > N.RG_MAX_RULES:
        raise ValueError(f"at most {N.RG_MAX_RULES} fused range rules per pass")

    pouts, pstructs, pimages, ikeep = [], [], [], []
    for k, p in enumerate(products):
        pr, dt = p.resolve(geom.grid_shape, geom.grid_limits, True)
        img_spec = getattr(p, "image", None)
        if out_products is not None and out_products[k] is not None:
            out = out_products[k]
            _check_buffer(out, device, F * ny * nx, dt, f"out_products[{k}] ({type(p).__name__}: {np.dtype(dt)})")
        elif img_spec is not None and not img_spec.keep_plane:
            out = None
        else:
            out = _alloc_like(device, (F, ny, nx), dt, fheld[0])
        if pr is None:
            if out is not None:
                _fill_nan(out, device)
            pimages.append(_nodata_image(p, device, F, ny, nx, fheld[0]))
        else:
            pr.out = None if out is None else _ptr(out, device)
            pimages.append(_attach_image(p, pr, device, F, ny, nx, fheld[0], ikeep))
            pstructs.append(pr)
        pouts.append(out)

    args = N.ApplyArgs()
    args.n_fields = F
    args.n_rules = len(rstructs)
    args.n_products = len(pstructs)
    args.reference_order = 2 if reference_order == "nearest_gate" else int(bool(reference_order))
    args.mask_invalid_bits = inv_bits
    args.fill_value = float(fill_value)
    fptr = (C.c_void_p * F)(*[_ptr(f, device) for f in fheld])
    mptr = (C.c_void_p * F)(*[(None if m is None else _ptr(m, device)) for m in mheld])
    gptr = (C.c_void_p * F)(*[(None if g is None else _ptr(g, device)) for g in grids])
    args.fields = fptr
    args.masks = mptr
    args.grid_out = gptr
    rarr = (N.QcRule * max(len(rstructs), 1))(*rstructs)
    parr = (N.Product * max(len(pstructs), 1))(*pstructs)
    args.rules = rarr
    args.products = parr
    result = {"grids": grids, "products": pouts, "images": pimages, "_keep": (fheld, mheld, rheld, ikeep)}
    call = PreparedCall(ctx, geom, args, device, result, (fptr, mptr, gptr, rarr, parr, pstructs, rstructs))
    if _prepare_only:
        return call
    call.launch()
    return result


class PreparedCall:
    """A ``grid_fields`` call with every argument resolved and marshalled: ``launch()`` is one ``rg_apply`` and nothing
    else.  For loops that grid volume after volume through the SAME buffers (a device-resident time series: new data
    is written into the field tensors, the outputs are consumed, ``launch()`` again), where the host-side preparation of
    ``grid_fields`` (product resolution, pointer marshalling: ~0.2 ms) would otherwise be paid per volume.
    ``result`` is the dictionary ``grid_fields`` returns; its arrays are overwritten by every launch."""

    def __init__(self, ctx, geom, args, device, result, keep):
        self.ctx, self.geom, self.args, self.device, self.result, self._keep = ctx, geom, args, device, result, keep
        self._space = N.RG_DEVICE if device else N.RG_HOST
        self._fn = N.lib().rg_apply

    def launch(self):
        with N.torch_stream_order(self.ctx, self.device):
            N.check(self._fn(self.ctx.handle, self.geom._h, C.byref(self.args), self._space))
        return self.result


def prepare_grid_fields(geom: DeviceGeometry, fields: Sequence, **kwargs) -> PreparedCall:
    """``grid_fields`` with the same arguments, returning the marshalled call instead of running it."""
    return grid_fields(geom, fields, _prepare_only=True, **kwargs)


class GeometryCache:
    """
    Device-side cache of neighbour tables keyed by (gate coordinates, grid, ROI parameters) — the reuse the reference
    gets from ``save_geometry`` / ``load_geometry`` (geometry.py:94-150): a table is built once per radar and scan
    strategy and applied to every volume.  Least-recently-used tables are freed when ``max_bytes`` of HBM is exceeded.
    """

    def __init__(self, max_bytes: int = 64 << 30, builder=None):
        from collections import OrderedDict
        self.max_bytes = int(max_bytes)
        self._items = OrderedDict()
        self._builder = builder or DeviceGeometry.build
        self.hits = self.misses = 0

    @staticmethod
    def key(gate_x, gate_y, gate_z, grid_shape, grid_limits, **params) -> str:
        import hashlib
        h = hashlib.sha256()                 # SHA-NI: ~1.2 GB/s, 55 ms for the 65 MB of cfg3 gate coordinates
        for a in (gate_x, gate_y, gate_z):
            h.update(memoryview(np.ascontiguousarray(np.asarray(a).ravel(), dtype=np.float32)))
        h.update(repr((tuple(int(v) for v in grid_shape), tuple(tuple(float(v) for v in l) for l in grid_limits),
                       sorted((k, (float(v) if isinstance(v, (int, float, np.floating, np.integer)) else v)) for k, v in params.items()))).encode())
        return h.hexdigest()

    def get(self, gate_x, gate_y, gate_z, grid_shape, grid_limits, **params) -> "DeviceGeometry":
        k = self.key(gate_x, gate_y, gate_z, grid_shape, grid_limits, **params)
        if k in self._items:
            self._items.move_to_end(k)
            self.hits += 1
            return self._items[k]
        self.misses += 1
        geom = self._builder(gate_x, gate_y, gate_z, grid_shape, grid_limits, **params)
        self._items[k] = geom
        self._evict()
        return geom

    def bytes_held(self) -> int:
        return sum(int(g.info["device_bytes"]) for g in self._items.values())

    def _evict(self):
        # dropped, not closed: a GridGeometry handed out earlier may still hold the table; it is freed with its last user
        while len(self._items) > 1 and self.bytes_held() > self.max_bytes:
            self._items.popitem(last=False)

    def clear(self):
        self._items.clear()


class VolumePipeline:
    """
    Overlap the host<->device copies of consecutive volumes with the kernels of their neighbours.

    ``grid_fields`` on host buffers is synchronous (H2D, pack, apply, D2H, wait).  A time series of volumes that
    share one neighbour table (BASELINE config 4; reference examples/batch_processing.py does this with a
    ThreadPoolExecutor around the CPU path) is instead fed through ``n_streams`` library contexts, each with its
    own CUDA stream and staging buffers, driven by one Python thread each (the ctypes call releases the GIL):
    while volume i runs its kernels, volume i+1 is on the H2D copy engine and volume i-1 on the D2H one.
    Use pinned host arrays (``pinned_empty``) for inputs and outputs, otherwise the copies serialise.
    """

    def __init__(self, geom: DeviceGeometry, n_streams: int = 3):
        from concurrent.futures import ThreadPoolExecutor
        self.geom = geom
        self.ctxs = [N.Context(geom.ctx.device) for _ in range(max(1, int(n_streams)))]
        self._pool = ThreadPoolExecutor(len(self.ctxs))

    def _run(self, slot, kwargs):
        return grid_fields(self.geom, ctx=self.ctxs[slot], **kwargs)

    def map(self, jobs: Sequence[dict]) -> List[Dict[str, object]]:
        """Run ``grid_fields(geom, **job)`` for every job; job i uses context i % n_streams.  Results in order."""
        n = len(self.ctxs)
        lanes = [[] for _ in range(n)]
        for i, job in enumerate(jobs):
            lanes[i % n].append((i, job))

        def work(slot):
            return [(i, self._run(slot, job)) for i, job in lanes[slot]]

        out: List[Optional[dict]] = [None] * len(jobs)
        for fut in [self._pool.submit(work, s) for s in range(n)]:
            for i, res in fut.result():
                out[i] = res
        return out

    def kernel_launches(self) -> int:
        return sum(c.kernel_launches() for c in self.ctxs)

    def close(self):
        self._pool.shutdown(wait=True)
        for c in self.ctxs:
            c.close()
        self.ctxs = []


def warn_all_nan(kind: str, plane: np.ndarray):
    """NumPy's nan-reductions warn when a column has no valid level; keep that observable behaviour."""
    if np.isnan(plane).any():
        msg = "Mean of empty slice" if kind == "mean" else "All-NaN slice encountered"
        warnings.warn(msg, RuntimeWarning, stacklevel=3)
