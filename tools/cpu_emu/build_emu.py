#!/usr/bin/env python
"""Build the DEV-ONLY CPU emulation of libradargrid_b200 (see cuda_runtime.h in this directory).

    python tools/cpu_emu/build_emu.py        ->  /tmp/rg_emu/libradargrid_b200_emu.so
    RADAR_GRID_B200_LIB=/tmp/rg_emu/libradargrid_b200_emu.so RG_EMU=1 python -m pytest tests -m gpu -k tiny

The output is written outside the repository on purpose: it must never ship or be picked up as a fallback.
"""
import os
import re
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "radar-processor_b200", "csrc")
OUT = os.environ.get("RG_EMU_DIR", "/tmp/rg_emu")


def rewrite_launches(src: str) -> str:
    out, i = [], 0
    while True:
        j = src.find("<<<", i)
        if j < 0:
            out.append(src[i:])
            break
        # kernel expression: identifier plus optional <template args>, scanning backwards
        k = j
        if src[k - 1] == ">":
            depth = 0
            while True:
                k -= 1
                if src[k] == ">":
                    depth += 1
                elif src[k] == "<":
                    depth -= 1
                    if depth == 0:
                        break
        while k > 0 and (src[k - 1].isalnum() or src[k - 1] in "_:"):
            k -= 1
        name = src[k:j]
        e = src.find(">>>", j)
        cfg = src[j + 3:e]
        # split launch config at top-level commas
        parts, depth, cur = [], 0, ""
        for ch in cfg:
            if ch in "([":
                depth += 1
            elif ch in ")]":
                depth -= 1
            if ch == "," and depth == 0:
                parts.append(cur)
                cur = ""
            else:
                cur += ch
        parts.append(cur)
        p = e + 3
        while src[p].isspace():
            p += 1
        assert src[p] == "(", src[p:p + 40]
        depth, q = 0, p
        while True:
            if src[q] == "(":
                depth += 1
            elif src[q] == ")":
                depth -= 1
                if depth == 0:
                    break
            q += 1
        args = src[p + 1:q].strip()
        out.append(src[i:k])
        smem = parts[2].strip() if len(parts) > 2 else "0"
        out.append(f"emu_launch({name}, dim3({parts[0].strip()}), dim3({parts[1].strip()}), (size_t)({smem})" + (", " + args if args else "") + ")")
        i = q + 1
    return "".join(out)


def main():
    os.makedirs(OUT, exist_ok=True)
    cpps = []
    for f in ("rg_api.cu", "rg_apply.cu", "rg_duo.cu", "rg_geometry.cu"):
        src = open(os.path.join(CSRC, f)).read()
        dst = os.path.join(OUT, f.replace(".cu", ".cpp"))
        text = rewrite_launches(src)
        text = re.sub(r"extern\s+__shared__\s+(\w+)\s+(\w+)\[\];", r"\1* \2 = (\1*)emu_dynamic_smem;", text)
        open(dst, "w").write(text)
        cpps.append(dst)
    lib = os.path.join(OUT, "libradargrid_b200_emu.so")
    cmd = ["g++", "-std=c++20", "-O1", "-g", "-fPIC", "-shared", "-pthread", "-ffp-contract=off", "-Wno-unknown-pragmas", "-DRG_EMU", *os.environ.get("RG_EMU_FLAGS", "").split(),
           "-I", HERE, "-I", CSRC, "-I", os.path.join(ROOT, "include"), "-o", lib] + cpps
    res = subprocess.run(cmd, capture_output=True, text=True)
    sys.stderr.write(res.stderr[-6000:])
    if res.returncode != 0:
        raise SystemExit("emulator build failed")
    print(lib)


if __name__ == "__main__":
    main()
