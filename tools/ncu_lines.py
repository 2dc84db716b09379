#!/usr/bin/env python
"""Per-source-line view of an ncu capture: joins `ncu --page source --csv` (per SASS instruction: executed count,
stall samples) with the line table of the same kernel from `nvdisasm -g`.

    python tools/ncu_lines.py <rep.ncu-rep | source-page.csv> <lib.so> <mangled kernel name> [top N]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile


def sh(cmd, **kw):
    return subprocess.run(cmd, capture_output=True, text=True, **kw)


def line_table(lib, kernel):
    tmp = tempfile.mkdtemp(prefix="ncu_lines_")
    sh(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp)
    for f in sorted(os.listdir(tmp)):
        path = os.path.join(tmp, f)
        sym = sh(["readelf", "-sW", path]).stdout
        m = re.search(r"^\s*(\d+):.*FUNC.*\s" + re.escape(kernel) + r"$", sym, re.M)
        if not m:
            continue
        text = sh(["nvdisasm", "-g", "-fun", m.group(1), path]).stdout
        table, cur, inl = {}, None, None
        for ln in text.splitlines():
            m2 = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
            if m2:
                cur = (os.path.basename(m2.group(1)), int(m2.group(2)))
                continue
            m3 = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m3 and cur:
                table[int(m3.group(1), 16)] = cur
        return table
    raise SystemExit("kernel not found in " + lib)


def main():
    rep, lib, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    table = line_table(lib, kernel)
    text = open(rep).read() if rep.endswith(".csv") else sh(["ncu", "-i", rep, "--page", "source", "--csv"]).stdout
    rows = list(csv.reader(io.StringIO(text)))
    hdr_i = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    hdr = rows[hdr_i]
    ia, isrc, iex, ismp = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    ilsb = hdr.index("stall_long_sb")
    base = None
    per_line = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
    tot_ex = tot_s = 0
    for r in rows[hdr_i + 1:]:
        if len(r) <= ismp or not r[ia].startswith("0x"):
            continue
        a = int(r[ia], 16)
        base = a if base is None else base
        key = table.get(a - base, ("?", 0))
        ex, smp = int(r[iex] or 0), int(r[ismp] or 0)
        e = per_line[key]
        e[0] += ex; e[1] += smp; e[2] += int(r[ilsb] or 0)
        e[3][r[isrc].split()[0].lstrip("@!P0123456789 ") if r[isrc].startswith("@") else r[isrc].split()[0]] += ex
        tot_ex += ex; tot_s += smp
    print(f"total warp-instructions {tot_ex:,}, samples {tot_s:,}")
    print("| file:line | warp-instr | % | samples % | long_sb % of line | top opcodes |\n|---|---|---|---|---|---|")
    for key, e in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        ops = ", ".join(f"{k} {v / 1e6:.1f}M" for k, v in e[3].most_common(4))
        print(f"| {key[0]}:{key[1]} | {e[0]:,} | {100 * e[0] / tot_ex:.1f} | {100 * e[1] / max(tot_s, 1):.1f} | "
              f"{100 * e[2] / max(e[1], 1):.0f} | {ops} |")


if __name__ == "__main__":
    main()
