"""pytest configuration: marker registration, import paths, shared fixtures."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_ROOT = os.path.join(ROOT, "radar-processor_b200")
for p in (ROOT, PKG_ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def golden():
    return load_golden


def golden_case(spec_name, weighting="barnes2", alt=0):
    """(spec, radar, gates, fields, golden dict) for one committed reference fixture."""
    from radar_grid_b200 import synthetic as S
    spec = S.SPECS[spec_name]
    radar = S.SyntheticRadar(spec, seed=7, radar_altitude=float(alt))
    gx = radar.gate_x["data"].ravel().astype("float32")
    gy = radar.gate_y["data"].ravel().astype("float32")
    gz = radar.gate_z["data"].ravel().astype("float32")
    fields = {k: np.ma.masked_invalid(v["data"]).ravel().astype("float32") for k, v in radar.fields.items()}
    g = load_golden(f"ref_{spec_name}_{weighting}_alt{int(alt)}.npz")
    return spec, radar, (gx, gy, gz), fields, g


def assert_same(a, b, what=""):
    """Bit-level equality including NaN positions (NaN payloads are not compared)."""
    a = np.asarray(a)
    b = np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} != {b.shape}"
    assert a.dtype == b.dtype, f"{what}: dtype {a.dtype} != {b.dtype}"
    np.testing.assert_array_equal(a, b, err_msg=what)


def _has_device():
    try:
        from radar_grid_b200 import _native as N
        return N.device_count() > 0
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests need a CUDA device (or, for kernel debugging only, the dev emulator with RG_EMU=1)."""
    have = None
    for item in items:
        if "gpu" in item.keywords:
            if have is None:
                have = _has_device()
            if not have:
                item.add_marker(pytest.mark.skip(reason="no CUDA device"))


def canonical(indptr, idx, w):
    """Sort every CSR row by gate id so two tables compare as sets."""
    from oracle import radar_grid_oracle as O
    return O.canonical_rows(indptr, idx, w)


def ulp_diff_f32(a, b):
    a = np.ascontiguousarray(a, dtype=np.float32).view(np.int32).astype(np.int64)
    b = np.ascontiguousarray(b, dtype=np.float32).view(np.int32).astype(np.int64)
    return np.abs(a - b)


@pytest.fixture(autouse=True, scope="session")
def _apply_variant_from_env():
    """RG_APPLY_VARIANT_TEST=N runs the gpu tests on apply kernel variant N (see rg_context_set_option) instead of the default."""
    v = os.environ.get("RG_APPLY_VARIANT_TEST")
    if v and _has_device():
        from radar_grid_b200 import _native as N
        N.default_context().set_option("apply_variant", int(v))
    yield
