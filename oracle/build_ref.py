#!/usr/bin/env python
"""
Recipe for oracle/_ref: the reference's OWN gridding modules, compiled from where they lie.

The reference's hot path is six pure-Python modules (src/radar_grid/{geometry,compute,interpolate,products,filters,
utils}.py).  "Compiling" them means byte-compiling: this script runs ``py_compile`` on the sources under
/root/reference and writes only the resulting byte code to ``oracle/_ref/radar_grid_ref/<module>.rgc`` — build output,
like a .so: git-ignored, never the sources themselves, but it travels to the GPU box with the gpurun snapshot (same
image, same CPython), where /root/reference does not exist.  (The files are ordinary ``.pyc`` files under another
suffix: snapshot tools tend to drop ``*.pyc``.)  ``load()`` imports them as a package with a sourceless loader, without
the reference's ``__init__`` (which needs matplotlib / rasterio).

Users: tests (to cross-check the NumPy restatement in radar_grid_oracle.py) and ``bench.py --impl reference`` (to time
the genuine reference code on the box's host cores, ``cpu_baseline.kind = "reference"``).  TEST INFRASTRUCTURE ONLY.

    python oracle/build_ref.py          # in the build container; __graft_entry__.build() does it too
"""
import importlib
import os
import py_compile
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src/radar_grid"
OUT = os.path.join(HERE, "_ref", "radar_grid_ref")
MODULES = ("geometry", "compute", "interpolate", "products", "filters", "utils")


def build() -> bool:
    """Byte-compile the reference modules into oracle/_ref.  False (nothing done) where the reference is absent."""
    if not os.path.isdir(REF_SRC):
        return False
    os.makedirs(OUT, exist_ok=True)
    for name in MODULES:
        py_compile.compile(os.path.join(REF_SRC, name + ".py"), cfile=os.path.join(OUT, name + ".rgc"),
                           dfile=f"reference/src/radar_grid/{name}.py", doraise=True)
    with open(os.path.join(OUT, "PYTHON_TAG"), "w") as fh:
        fh.write(sys.implementation.cache_tag)
    return True


def available() -> bool:
    try:
        with open(os.path.join(OUT, "PYTHON_TAG")) as fh:
            tag = fh.read().strip()
    except OSError:
        return False
    return tag == sys.implementation.cache_tag and all(os.path.exists(os.path.join(OUT, m + ".rgc")) for m in MODULES)


def load():
    """The reference modules as a namespace (``ref.compute.compute_grid_geometry`` ...), from oracle/_ref."""
    if not available():
        raise RuntimeError("oracle/_ref is missing or was built by another CPython: run oracle/build_ref.py in the build container")
    import importlib.machinery
    import importlib.util
    pkg = types.ModuleType("radar_grid_ref")
    pkg.__path__ = [OUT]
    sys.modules["radar_grid_ref"] = pkg
    mods, todo = {}, list(MODULES)
    for _ in range(len(MODULES) + 1):          # the modules import one another relatively: load until all resolve
        for name in list(todo):
            full = f"radar_grid_ref.{name}"
            loader = importlib.machinery.SourcelessFileLoader(full, os.path.join(OUT, name + ".rgc"))
            spec = importlib.util.spec_from_loader(full, loader)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[full] = mod
            try:
                loader.exec_module(mod)
            except ImportError:
                del sys.modules[full]
                continue
            setattr(pkg, name, mod)
            mods[name] = mod
            todo.remove(name)
        if not todo:
            break
    if todo:
        raise RuntimeError(f"could not import reference modules {todo} from oracle/_ref")
    return types.SimpleNamespace(**mods)


if __name__ == "__main__":
    print("built oracle/_ref" if build() else "reference sources not found: nothing built")
