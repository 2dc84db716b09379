"""
ctypes binding of libradargrid_b200.so (C ABI: include/radar_grid_b200.h).

There is no CPU fallback: if the shared library is missing or a CUDA device is not available, every
compute call raises.  Loading the library itself does not need a GPU (the CPU test-suite checks that all
declared symbols are exported).
"""

from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PKG_ROOT = os.path.dirname(_HERE)

RG_OK, RG_ERR_INVALID, RG_ERR_CUDA, RG_ERR_NOMEM, RG_ERR_UNSUPPORTED = 0, 1, 2, 3, 4
RG_HOST, RG_DEVICE = 0, 1
RG_W = {"barnes2": 0, "cressman": 1, "nearest": 2, "dist2": 3}
RG_PROD_COLMAX, RG_PROD_COLMIN, RG_PROD_COLMEAN, RG_PROD_LEVEL, RG_PROD_BEAM = 1, 2, 3, 4, 5
RG_BLEND_PICK, RG_BLEND_F32, RG_BLEND_F64, RG_BLEND_F64_OUT64 = 0, 1, 2, 3
RG_PF_BELOW, RG_PF_ABOVE, RG_PF_OUTSIDE, RG_PF_INVALID, RG_PF_BELOW_EQUAL = 1, 2, 3, 4, 5
RG_MAX_FIELDS, RG_MAX_RULES, RG_MAX_SLICES = 8, 8, 4


class GridSpec(C.Structure):
    _fields_ = [("nz", C.c_int32), ("ny", C.c_int32), ("nx", C.c_int32),
                ("z_begin", C.c_int32), ("z_end", C.c_int32), ("reserved_", C.c_int32),
                ("z_min", C.c_double), ("z_max", C.c_double), ("y_min", C.c_double),
                ("y_max", C.c_double), ("x_min", C.c_double), ("x_max", C.c_double)]


class GeometryInfo(C.Structure):
    _fields_ = [("n_rows", C.c_int64), ("n_pairs", C.c_int64), ("n_gates", C.c_int64),
                ("n_empty_rows", C.c_int64), ("max_row_len", C.c_int64), ("n_gates_binned", C.c_int64),
                ("n_candidates", C.c_int64), ("device_bytes", C.c_int64),
                ("build_ms", C.c_double), ("cell_size", C.c_double), ("grid", GridSpec)]


RG_MAX_IMAGE_FILTERS = 4


class Image(C.Structure):
    _fields_ = [("n_filters", C.c_int32), ("filter_kind", C.c_int32 * RG_MAX_IMAGE_FILTERS),
                ("filter_a", C.c_double * RG_MAX_IMAGE_FILTERS), ("filter_b", C.c_double * RG_MAX_IMAGE_FILTERS),
                ("filter_fill", C.c_double * RG_MAX_IMAGE_FILTERS), ("vmin", C.c_double), ("vmax", C.c_double),
                ("fill_value", C.c_double), ("has_fill_value", C.c_int32), ("lut_entries", C.c_int32),
                ("lut", C.c_void_p), ("out", C.c_void_p)]


class Product(C.Structure):
    _fields_ = [("kind", C.c_int32), ("mode", C.c_int32), ("z_lo", C.c_int32), ("z_hi", C.c_int32),
                ("earth_curvature", C.c_int32), ("partial", C.c_int32),
                ("w_lo", C.c_double), ("w_hi", C.c_double), ("sin_elev", C.c_double),
                ("cos_elev_clamped", C.c_double), ("tan_elev", C.c_double), ("ke_re", C.c_double),
                ("ke_re_sq", C.c_double), ("out", C.c_void_p), ("image", C.POINTER(Image))]


class QcRule(C.Structure):
    _fields_ = [("values", C.c_void_p), ("lo", C.c_float), ("hi", C.c_float),
                ("use_lo", C.c_int32), ("use_hi", C.c_int32), ("field_bits", C.c_uint32),
                ("reserved_", C.c_uint32)]


class ApplyArgs(C.Structure):
    _fields_ = [("n_fields", C.c_int32), ("n_rules", C.c_int32), ("n_products", C.c_int32),
                ("reference_order", C.c_int32), ("mask_invalid_bits", C.c_uint32), ("fill_value", C.c_float),
                ("fields", C.POINTER(C.c_void_p)), ("masks", C.POINTER(C.c_void_p)),
                ("rules", C.POINTER(QcRule)), ("grid_out", C.POINTER(C.c_void_p)),
                ("products", C.POINTER(Product))]


# name -> (restype, argtypes); mirrors include/radar_grid_b200.h one to one
_PROTOTYPES = {
    "rg_abi_version": (C.c_int, []),
    "rg_last_error": (C.c_char_p, []),
    "rg_device_count": (C.c_int, [C.POINTER(C.c_int32)]),
    "rg_context_create": (C.c_int, [C.c_int32, C.c_void_p, C.POINTER(C.c_void_p)]),
    "rg_context_destroy": (C.c_int, [C.c_void_p]),
    "rg_context_set_stream": (C.c_int, [C.c_void_p, C.c_void_p]),
    "rg_context_synchronize": (C.c_int, [C.c_void_p]),
    "rg_context_get_stream": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p)]),
    "rg_context_kernel_launches": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "rg_context_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int64]),
    "rg_context_kernel_time": (C.c_int, [C.c_void_p, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_int64), C.c_int32]),
    "rg_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_int64]),
    "rg_host_free": (C.c_int, [C.c_void_p]),
    "rg_linspace_f32": (C.c_int, [C.c_double, C.c_double, C.c_int32, C.c_void_p]),
    "rg_geometry_build": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                    C.POINTER(GridSpec), C.c_double, C.c_double, C.c_double, C.c_int32,
                                    C.c_double, C.POINTER(C.c_void_p)]),
    "rg_geometry_from_csr": (C.c_int, [C.c_void_p, C.POINTER(GridSpec), C.c_void_p, C.c_int32, C.c_void_p,
                                       C.c_void_p, C.c_int64, C.c_int32, C.POINTER(C.c_void_p)]),
    "rg_geometry_get_info": (C.c_int, [C.c_void_p, C.POINTER(GeometryInfo)]),
    "rg_geometry_duo_slots": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64)]),
    "rg_geometry_export_csr": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p,
                                         C.c_int32]),
    "rg_geometry_destroy": (C.c_int, [C.c_void_p]),
    "rg_gate_coordinates": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, C.c_int32,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]),
    "rg_geometry_level_pairs": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32,
                                          C.POINTER(GridSpec), C.c_double, C.c_double, C.c_double, C.c_double, C.c_int32,
                                          C.POINTER(C.c_int64)]),
    "rg_products": (C.c_int, [C.c_void_p, C.POINTER(GridSpec), C.c_int32, C.POINTER(C.c_void_p), C.c_int32,
                              C.POINTER(Product), C.c_int32]),
    "rg_plane_filter": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_int32,
                                  C.c_double, C.c_double, C.c_double, C.c_int32]),
    "rg_apply": (C.c_int, [C.c_void_p, C.c_void_p, C.POINTER(ApplyArgs), C.c_int32]),
}

_lib = None
_lib_lock = threading.Lock()


def library_path() -> str:
    # RADAR_GRID_B200_LIB is a DEVELOPMENT switch (A/B builds made with build.py --variant, the CPU emulator of
    # tools/cpu_emu): it must name a build of this library and is checked against the ABI version on load; it is never
    # set by the package, the tests' defaults, bench.py or __graft_entry__.py.
    override = os.environ.get("RADAR_GRID_B200_LIB")
    if override:
        return override
    return os.path.join(_PKG_ROOT, "lib", "libradargrid_b200.so")


def lib():
    """Load (building first if the sources are newer and nvcc is present) and return the CDLL."""
    global _lib
    if _lib is not None:
        return _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        path = library_path()
        if "RADAR_GRID_B200_LIB" not in os.environ:
            import importlib.util
            spec = importlib.util.spec_from_file_location("_rg_b200_build", os.path.join(_PKG_ROOT, "build.py"))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            path = mod.build()
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: run `python radar-processor_b200/build.py` "
                               "(radar_grid_b200 has no CPU fallback)")
        handle = C.CDLL(path)
        for name, (restype, argtypes) in _PROTOTYPES.items():
            fn = getattr(handle, name)       # AttributeError here = header / library mismatch
            fn.restype = restype
            fn.argtypes = argtypes
        if handle.rg_abi_version() != 2:
            raise RuntimeError("libradargrid_b200.so ABI version mismatch")
        _lib = handle
        return _lib


class RadarGridError(RuntimeError):
    pass


def check(status: int):
    if status == RG_OK:
        return
    msg = lib().rg_last_error().decode("utf-8", "replace")
    if status == RG_ERR_INVALID:
        raise ValueError(msg)
    if status == RG_ERR_NOMEM:
        raise MemoryError(msg)
    if status == RG_ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise RadarGridError(msg)


def device_count() -> int:
    n = C.c_int32(0)
    st = lib().rg_device_count(C.byref(n))
    return int(n.value) if st == RG_OK else 0


# ---- contexts: one per (thread, device, stream) --------------------------------------------------------
class Context:
    def __init__(self, device: int = 0, stream: int | None = None):
        self.device = int(device)
        h = C.c_void_p()
        check(lib().rg_context_create(self.device, C.c_void_p(stream) if stream else None, C.byref(h)))
        self.handle = h
        # kernel A/B knobs for tests and experiments (rg_context_set_option); unset = the library's own choice
        for env, key in (("RG_GROUP_WIDTH", "group_width"), ("RG_APPLY_VARIANT", "apply_variant")):
            if os.environ.get(env):
                self.set_option(key, int(os.environ[env]))

    def set_stream(self, stream: int | None):
        check(lib().rg_context_set_stream(self.handle, C.c_void_p(stream) if stream else None))

    def synchronize(self):
        check(lib().rg_context_synchronize(self.handle))

    def stream_ptr(self) -> int:
        p = C.c_void_p()
        check(lib().rg_context_get_stream(self.handle, C.byref(p)))
        return int(p.value or 0)

    def kernel_launches(self) -> int:
        n = C.c_int64(0)
        check(lib().rg_context_kernel_launches(self.handle, C.byref(n)))
        return int(n.value)

    def kernel_time(self, which: int, reset: bool = True):
        """(total device ms, launch count) of the timed pack (0) / apply (1) launches; needs option timing=1."""
        ms, n = C.c_double(0), C.c_int64(0)
        check(lib().rg_context_kernel_time(self.handle, int(which), C.byref(ms), C.byref(n), int(reset)))
        return float(ms.value), int(n.value)

    def set_option(self, key: str, value: int):
        check(lib().rg_context_set_option(self.handle, key.encode(), int(value)))

    def close(self):
        if getattr(self, "handle", None):
            lib().rg_context_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_tls = threading.local()
_default_device = 0


def set_device(device: int):
    global _default_device
    _default_device = int(device)


def get_device() -> int:
    return _default_device


def default_context(device: int | None = None) -> Context:
    """The calling thread's context for `device` (private stream); created on first use."""
    dev = _default_device if device is None else int(device)
    cache = getattr(_tls, "ctx", None)
    if cache is None:
        cache = _tls.ctx = {}
    ctx = cache.get(dev)
    if ctx is None:
        ctx = cache[dev] = Context(dev)
    return ctx


# ---- pinned host arrays -----------------------------------------------------------------------------------
class _PinnedBlock:
    def __init__(self, nbytes: int):
        p = C.c_void_p()
        check(lib().rg_host_alloc(C.byref(p), int(nbytes)))
        self.ptr = p
        self.nbytes = int(nbytes)

    def __del__(self):
        try:
            if self.ptr:
                lib().rg_host_free(self.ptr)
                self.ptr = None
        except Exception:
            pass


def pinned_empty(shape, dtype=np.float32) -> np.ndarray:
    """NumPy array backed by page-locked host memory (for the end-to-end H2D/D2H path)."""
    dtype = np.dtype(dtype)
    n = int(np.prod(shape)) if np.ndim(shape) else int(shape)
    block = _PinnedBlock(max(n * dtype.itemsize, 1))
    buf = (C.c_char * block.nbytes).from_address(block.ptr.value)
    buf._rg_block = block            # numpy keeps `buf` alive through .base, `buf` keeps the allocation
    return np.frombuffer(buf, dtype=dtype, count=n).reshape(shape)


# ---- buffer helpers -------------------------------------------------------------------------------------------
def is_device_array(x) -> bool:
    """torch CUDA tensor or anything exposing __cuda_array_interface__."""
    if hasattr(x, "__cuda_array_interface__") and not isinstance(x, np.ndarray):
        return bool(getattr(x, "is_cuda", True))
    return False


class torch_stream_order:
    """
    Stream-order a device-buffer call with the caller's torch stream: the library's stream first waits for
    everything already queued on torch's current stream (the inputs), and torch's current stream then waits for
    the library's work (the outputs) — so torch code before and after the call needs no manual synchronisation.
    No host synchronisation happens.
    """

    def __init__(self, ctx: "Context", active: bool):
        self.ext = self.cur = None
        if active:
            import torch
            self.cur = torch.cuda.current_stream(ctx.device)
            if self.cur.cuda_stream != ctx.stream_ptr():
                self.ext = torch.cuda.ExternalStream(ctx.stream_ptr(), device=ctx.device)

    def __enter__(self):
        if self.ext is not None:
            self.ext.wait_stream(self.cur)
        return self

    def __exit__(self, *exc):
        if self.ext is not None:
            self.cur.wait_stream(self.ext)
        return False


def device_ptr(x) -> int:
    if hasattr(x, "data_ptr"):
        return int(x.data_ptr())
    return int(x.__cuda_array_interface__["data"][0])


def host_ptr(a: np.ndarray) -> int:
    return int(a.ctypes.data)


def linspace_f32(start: float, stop: float, num: int) -> np.ndarray:
    out = np.empty(int(num), dtype=np.float32)
    check(lib().rg_linspace_f32(float(start), float(stop), int(num), C.c_void_p(host_ptr(out)) if num else None))
    return out
