#!/usr/bin/env python
"""
Box-wide host<->device copy ceiling: N concurrent processes (one per GPU) of tools/pcie_ceiling (bare
cudaMemcpyAsync, no library code) for N = 1, 2, 4, 8 up to the GPUs present; every (staging kind, direction) test
starts in the same wall-clock slot in all processes.  Writes one JSON document (default
gpurun_out/pcie_ceiling.json; copy it to profiles/ to have it judged).

    nvcc -O2 -o tools/pcie_ceiling tools/pcie_ceiling.cu      # here (the binary travels with the gpurun snapshot)
    python tools/pcie_ceiling.py [--mib 256] [--pieces 1] [--out gpurun_out/pcie_ceiling.json]
"""
import argparse
import json
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))


def gpu_count():
    try:
        out = subprocess.run(["nvidia-smi", "-L"], capture_output=True, text=True).stdout
        return sum(1 for l in out.splitlines() if l.startswith("GPU "))
    except Exception:
        return 0


def run_n(n, mib, pieces):
    start = int(time.time()) + 6                       # CUDA context + pinned allocations of every process fit in here
    procs = [subprocess.Popen([os.path.join(HERE, "pcie_ceiling"), str(i), str(mib), "2", str(start), str(pieces)],
                              stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True) for i in range(n)]
    rows = []
    for p in procs:
        out, err = p.communicate(timeout=300)
        if p.returncode != 0:
            sys.stderr.write(err[-400:])
        rows += [json.loads(l) for l in out.splitlines() if l.startswith("{")]
    table = {}
    for r in rows:
        key = f"{r['kind']}/{r['dir']}"
        if r.get("gbs") is None:
            table.setdefault(key, None)
            continue
        e = table.setdefault(key, {"sum_gbs": 0.0, "per_gpu": []}) or {"sum_gbs": 0.0, "per_gpu": []}
        e["sum_gbs"] = round(e["sum_gbs"] + r["gbs"], 2)
        e["per_gpu"].append(r["gbs"])
        table[key] = e
    huge = [r.get("anon_huge_kb") for r in rows if r.get("kind") == "thp" and r.get("gbs") is not None]
    return {"processes": n, "tests": table, "thp_anon_huge_kb": huge[:1]}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=256)
    ap.add_argument("--pieces", type=int, default=1)
    ap.add_argument("--out", default="gpurun_out/pcie_ceiling.json")
    args = ap.parse_args()
    n_gpu = gpu_count()
    if n_gpu < 1:
        raise SystemExit("no GPU")
    doc = {"what": "bare cudaMemcpyAsync, one process per GPU, GB/s (decimal); 'both' = h2d + d2h summed", "gpus": n_gpu,
           "mib_per_copy": args.mib, "pieces": args.pieces, "host_cpus": os.cpu_count(), "runs": []}
    try:
        doc["thp_enabled"] = open("/sys/kernel/mm/transparent_hugepage/enabled").read().strip()
        doc["nr_hugepages"] = int(open("/proc/sys/vm/nr_hugepages").read())
    except Exception:
        pass
    for n in (1, 2, 4, 8):
        if n <= n_gpu:
            doc["runs"].append(run_n(n, args.mib, args.pieces))
    os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
    with open(args.out, "w") as fh:
        json.dump(doc, fh, indent=1)
    for r in doc["runs"]:
        print(r["processes"], {k: (v and v["sum_gbs"]) for k, v in r["tests"].items()})


if __name__ == "__main__":
    main()
