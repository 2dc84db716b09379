#!/usr/bin/env python
"""Summarise an .ncu-rep (one `ncu --set full` capture) into the handful of numbers the roofline discussion uses.

    python tools/ncu_summary.py gpurun_out/prof_apply.ncu-rep [launch_index] > profiles/<name>.md
    python tools/ncu_summary.py gpurun_out/<name>            # reads <name>.raw.csv + <name>.source.csv (tools/gpu_ncu_apply.sh)
    python tools/ncu_summary.py gpurun_out/<name> --traffic cfg3   # also rewrites profiles/apply_traffic.json, stamped with the kernel sources' hash
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "l1tex__t_requests_pipe_lsu_mem_local_op_st.sum",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum", "lts__t_sectors_srcunit_tex_op_read.sum",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__grid_size", "launch__block_size",
    "launch__waves_per_multiprocessor", "sm__cycles_elapsed.max",
]


def run(args):
    return subprocess.run(["ncu"] + args, capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    which = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 0
    from_csv = not rep.endswith(".ncu-rep")
    raw_text = open(rep + ".raw.csv").read() if from_csv else run(["-i", rep, "--page", "raw", "--csv"])
    rows = list(csv.reader(io.StringIO(raw_text)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    r = data[which]
    name = r[hdr.index("Kernel Name")]
    print(f"# ncu summary: `{rep}` launch {which}\n\nkernel: `{name}`\n")
    print("| metric | value | unit |\n|---|---|---|")
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"| {k} | {r[i]} | {units[i]} |")
    # instruction mix + stall reasons from the SASS page
    if "--traffic" in sys.argv:
        import hashlib, json, os
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        hh = hashlib.sha256()
        for f in ("rg_apply.cu", "rg_duo.cu", "rg_geometry.cu", "rg_api.cu", "rg_internal.cuh", "rg_device.cuh"):
            hh.update(open(os.path.join(root, "radar-processor_b200", "csrc", f), "rb").read())
        to_bytes = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
        tot = sum(float(r[hdr.index(k)]) * to_bytes[units[hdr.index(k)]] for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        doc = {"workload": sys.argv[sys.argv.index("--traffic") + 1], "kernel": name, "traffic_bytes_per_launch": tot,
               "source": os.path.basename(rep), "source_sha256": hh.hexdigest(),
               "how": "dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of the apply kernel"}
        json.dump(doc, open(os.path.join(root, "profiles", "apply_traffic.json"), "w"), indent=1)
    sass_text = open(rep + ".source.csv").read() if from_csv else run(["-i", rep, "--page", "source", "--csv", "--print-source", "sass"])
    sass = list(csv.reader(io.StringIO(sass_text)))
    hi = [i for i, x in enumerate(sass) if x and x[0] == "Address"]
    if hi:
        h = sass[hi[0]]
        ci = {n: i for i, n in enumerate(h)}
        body = []
        for x in sass[hi[0] + 1:]:
            if x and x[0] == "Kernel Name":
                break
            if len(x) > 10:
                body.append(x)
        tot = sum(int(x[ci["Instructions Executed"]] or 0) for x in body)
        ops = collections.Counter()
        for x in body:
            m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", x[ci["Source"]])
            ops[(m.group(2).split(".")[0] if m else "?")] += int(x[ci["Instructions Executed"]] or 0)
        print(f"\nSASS: {len(body)} instructions, {tot:,} warp-instructions executed\n")
        print("| opcode | executed | share |\n|---|---|---|")
        for o, c in ops.most_common(14):
            print(f"| {o} | {c:,} | {100 * c / max(tot, 1):.1f}% |")
        st = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
        agg = {n: sum(int(x[ci[n]] or 0) for x in body) for n in st}
        ts = sum(agg.values()) or 1
        print("\n| stall reason (all samples) | share |\n|---|---|")
        for n, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]:
            print(f"| {n} | {100 * v / ts:.1f}% |")


if __name__ == "__main__":
    main()
