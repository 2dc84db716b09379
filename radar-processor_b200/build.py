#!/usr/bin/env python
"""
Build libradargrid_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python radar-processor_b200/build.py [--force] [--variant NAME -DFLAG ...]

The library has no torch / Python dependency: it is a plain C-ABI shared object (include/radar_grid_b200.h)
that the Python mirror loads with ctypes.  Built files live under radar-processor_b200/lib/ (git-ignored,
but they travel to the GPU box with the gpurun snapshot).
"""
import argparse
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = [os.path.join(HERE, "csrc", f) for f in ("rg_api.cu", "rg_apply.cu", "rg_duo.cu", "rg_geometry.cu")]
HDR = [os.path.join(HERE, "csrc", "rg_internal.cuh"), os.path.join(HERE, "csrc", "rg_device.cuh"), os.path.join(ROOT, "include", "radar_grid_b200.h")]
LIBDIR = os.path.join(HERE, "lib")

NVCC_FLAGS = [
    "-O3", "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-fmad=false",            # parity: no silent FMA contraction; the fast kernel asks for fmaf explicitly
    "-Xcompiler", "-fPIC", "-shared",
    "-I", os.path.join(ROOT, "include"),
]


def lib_path(variant=""):
    return os.path.join(LIBDIR, f"libradargrid_b200{('_' + variant) if variant else ''}.so")


def _digest(extra):
    h = hashlib.sha256()
    for f in SRC + HDR + [os.path.abspath(__file__)]:
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(extra).encode())
    return h.hexdigest()


def build(force=False, variant="", extra_flags=(), verbose=False):
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    os.makedirs(LIBDIR, exist_ok=True)
    out = lib_path(variant)
    stamp = out + ".sha256"
    digest = _digest(list(extra_flags))
    if not force and os.path.exists(out) and os.path.exists(stamp) and open(stamp).read().strip() == digest:
        return out
    if not os.path.exists(nvcc):
        if os.path.exists(out):
            return out          # GPU box without the sources' toolchain: use the shipped binary
        raise RuntimeError("nvcc not found and no prebuilt libradargrid_b200.so")
    if not force and os.path.exists(out) and os.environ.get("GRAFT_REPO_ROOT") and not os.environ.get("RADAR_GRID_B200_REBUILD"):
        # on a leased GPU box the shipped binary is what was meant to run: rebuilding ~190 kernels there would burn five
        # GPU-minutes.  Say so loudly instead; build before taking the snapshot (or set RADAR_GRID_B200_REBUILD=1).
        sys.stderr.write(f"radar_grid_b200: {out} is older than its sources; using it as shipped (GPU box)\n")
        return out
    # one nvcc per source file, in parallel (rg_apply.cu alone instantiates ~190 kernels), then one link step.  Objects are
    # cached under /tmp by the digest of (source, headers, flags that reach the file): -DRG_DUO_* switches reach rg_duo.cu only,
    # so an A/B variant of that kernel recompiles one file.
    import tempfile
    from concurrent.futures import ThreadPoolExecutor
    base_flags = [f for f in NVCC_FLAGS if f != "-shared"] + (["-Xptxas", "-v"] if verbose else [])
    duo_flags = [f for f in extra_flags if f.startswith("-DRG_DUO_")]
    common_flags = [f for f in extra_flags if not f.startswith("-DRG_DUO_")]
    cache = os.environ.get("RG_OBJ_CACHE", os.path.join(tempfile.gettempdir(), "rg_objcache"))
    os.makedirs(cache, exist_ok=True)
    hdr_digest = hashlib.sha256(b"".join(open(h, "rb").read() for h in HDR)).hexdigest()
    with tempfile.TemporaryDirectory(prefix="rg_build_") as tmp:
        # rg_apply.cu is compiled as two translation units (field counts 1..4 + common code, and 5..8: see RG_PART there)
        jobs = []
        for f in SRC:
            base = os.path.basename(f)[:-3]
            fl = common_flags + (duo_flags if base == "rg_duo" else [])
            parts = [("_lo", ["-DRG_PART=1"]), ("_hi", ["-DRG_PART=2"])] if base == "rg_apply" else [("", [])]
            for suffix, part_flags in parts:
                key = hashlib.sha256((open(f, "rb").read().hex() + hdr_digest + " ".join(base_flags + fl + part_flags)).encode()).hexdigest()[:24]
                jobs.append((f, os.path.join(cache, f"{base}{suffix}_{key}.o"), fl + part_flags))
        objs = [j[1] for j in jobs]

        def compile_one(job):
            src, obj, extra = job
            if os.path.exists(obj) and not verbose and not force:
                return subprocess.CompletedProcess([], 0, "", "")
            res = subprocess.run([nvcc] + base_flags + extra + ["-c", src, "-o", obj + ".tmp"], capture_output=True, text=True)
            if res.returncode == 0:
                os.replace(obj + ".tmp", obj)
            return res

        with ThreadPoolExecutor(max_workers=len(jobs)) as pool:
            results = list(pool.map(compile_one, jobs))
        for res in results:
            if res.returncode != 0:
                sys.stderr.write(res.stdout + res.stderr)
                raise RuntimeError("nvcc failed")
            if verbose:
                sys.stderr.write(res.stderr)
        res = subprocess.run([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", out] + objs,
                             capture_output=True, text=True)
        if res.returncode != 0:
            sys.stderr.write(res.stdout + res.stderr)
            raise RuntimeError("nvcc link failed")
    with open(stamp, "w") as fh:
        fh.write(digest)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--variant", default="")
    ap.add_argument("--verbose", action="store_true")
    args, extra = ap.parse_known_args()
    print(build(force=args.force, variant=args.variant, extra_flags=extra, verbose=args.verbose))
