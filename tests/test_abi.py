"""
C-ABI checks that need no GPU: the shared library loads, exports every symbol include/radar_grid_b200.h
declares, the ctypes mirrors of the public structs have the C layout, and the host-side helpers compute.
"""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT
from radar_grid_b200 import _native as N

HEADER = os.path.join(ROOT, "include", "radar_grid_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rg_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    lib = N.lib()
    names = declared_functions()
    assert len(names) >= 20
    for name in names:
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"
    assert sorted(N._PROTOTYPES) == names, "ctypes prototypes and header declarations differ"
    assert lib.rg_abi_version() == 2


def test_struct_layouts_match_the_c_compiler(tmp_path):
    structs = {"rg_grid_spec": N.GridSpec, "rg_geometry_info": N.GeometryInfo, "rg_product": N.Product, "rg_image": N.Image,
               "rg_qc_rule": N.QcRule, "rg_apply_args": N.ApplyArgs}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "radar_grid_b200.h"', "int main(void) {"]
    for cname, ct in structs.items():
        lines.append(f'printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in ct._fields_:
            lines.append(f'printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines.append("return 0; }")
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(l.split() for l in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, ct in structs.items():
        assert int(out[cname]) == C.sizeof(ct), cname
        for fname, _ in ct._fields_:
            assert int(out[f"{cname}.{fname}"]) == getattr(ct, fname).offset, f"{cname}.{fname}"


def test_header_is_plain_c(tmp_path):
    src = tmp_path / "inc.c"
    src.write_text('#include "radar_grid_b200.h"\nint main(void){return RG_OK;}\n')
    subprocess.run(["gcc", "-std=c99", "-pedantic", "-Werror", "-I", os.path.join(ROOT, "include"), "-c", str(src),
                    "-o", str(tmp_path / "inc.o")], check=True)


@pytest.mark.parametrize("a,b,n", [(0, 19500, 40), (-120000, 120000, 481), (0.0, 1000.0, 1), (3.3, 3.3, 5),
                                   (-7.7, 1234.56, 97), (0, 19000, 20), (-250000.0, 250000.0, 2001), (10, 0, 7)])
def test_linspace_matches_numpy_bit_for_bit(a, b, n):
    np.testing.assert_array_equal(N.linspace_f32(a, b, n), np.linspace(a, b, n, dtype="float32"))


def test_error_reporting_without_device():
    if N.device_count() > 0:
        pytest.skip("a device is present")
    with pytest.raises(N.RadarGridError):
        N.Context(0)
    st = N.lib().rg_linspace_f32(0.0, 1.0, -1, None)
    assert st == N.RG_ERR_INVALID and b"bad argument" in N.lib().rg_last_error()
