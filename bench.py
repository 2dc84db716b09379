#!/usr/bin/env python
"""
bench.py — headline benchmark of the gridding hot path (BASELINE.json metric: voxels/s for interpolation +
COLMAX(+CAPPI), fraction of the HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg3|cfg1|small]

Workload (N=1): BASELINE.json configs[2] — the configuration the north-star target is quoted on: a 15-sweep x
360 x 1000-gate volume with five fields (DBZH, ZDR, RHOHV, KDP, VRAD) gridded onto 40x481x481 at 0.5 km through
ONE shared neighbour table; a step = one volume: gate-mask fusion + record packing (K4), fused 5-field CSR
interpolation writing the five 3-D grids, with COLMAX and CAPPI(4000 m) in the epilogue (K5+K6).
The neighbour table is built once on the GPU before the timed region (it is amortised over every volume of a
scan strategy; its build time is reported in `config`).  Data are synthetic, seeded.

N>1: weak scaling, no data-path collective — every rank holds a replica of the table and grids its own
volumes (the reference's time-series batch shard, BASELINE.json configs[3]); value = all ranks' voxels /
max-over-ranks device time.

--impl reference times the reference's own CPU path — the genuine modules byte-compiled into oracle/_ref where they
travelled with the snapshot, else the NumPy oracle port — on the SAME workload (all levels, all fields) on the box's
host cores, one worker process per field.
"""

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "radar-processor_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = "voxels/sec (interp+COLMAX+CAPPI, voxel-fields gridded per second)"
UNIT = "voxels/s"
CAPPI_ALT = 4000.0


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as fh:
            return float(json.load(fh)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def kernel_source_hash():
    """SHA-256 over the CUDA sources: profiles/apply_traffic.json is stamped with it when it is regenerated from an ncu
    capture (tools/ncu_summary.py --traffic), so a stale DRAM-traffic figure is never reported for a changed kernel."""
    import hashlib
    h = hashlib.sha256()
    for f in ("rg_apply.cu", "rg_duo.cu", "rg_geometry.cu", "rg_api.cu", "rg_internal.cuh", "rg_device.cuh"):
        with open(os.path.join(ROOT, "radar-processor_b200", "csrc", f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def measured_traffic(workload):
    """dram__bytes_read.sum + dram__bytes_write.sum of the apply kernel from the committed ncu capture (per launch);
    None when the capture belongs to another workload or to other kernel sources."""
    try:
        with open(os.path.join(ROOT, "profiles", "apply_traffic.json")) as fh:
            t = json.load(fh)
        if t.get("workload") != workload or t.get("source_sha256") != kernel_source_hash():
            return None
        return float(t["traffic_bytes_per_launch"])
    except Exception:
        return None


def algorithmic_bytes(P, V, G, F, ncol, grids=True, n_planes=2):
    """8P + 4(V+1) + 5FG + B_out, B_out = 4FV for the 3-D grids + 4F*ny*nx per 2-D product."""
    return 8 * P + 4 * (V + 1) + 5 * F * G + (4 * F * V if grids else 0) + n_planes * 4 * F * ncol


# ---------------------------------------------------------------------------------------------------------
# clocks: sampled DURING the timed region
# ---------------------------------------------------------------------------------------------------------
class ClockSampler:
    def __init__(self, device_index):
        self.idx = device_index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._th = None
        self._nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(device_index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nv = None

    def _loop(self):
        nv = self._nv
        names = {}
        for n in dir(nv):
            if n.startswith("nvmlClocksEventReason") or n.startswith("nvmlClocksThrottleReason"):
                try:
                    names[int(getattr(nv, n))] = n.replace("nvmlClocksEventReason", "").replace("nvmlClocksThrottleReason", "")
                except Exception:
                    pass
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                try:
                    bits = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                except Exception:
                    bits = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
                for b, n in names.items():
                    if b and bits & b and n not in ("None", "GpuIdle", "All"):
                        self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.002)

    def start(self):
        if self._nv is not None:
            self._th = threading.Thread(target=self._loop, daemon=True)
            self._th.start()

    def stop(self):
        self._stop.set()
        if self._th is not None:
            self._th.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["unsampled"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------
def raw_fields(spec, seed, gates):
    from radar_grid_b200 import synthetic as S
    fields = S.make_fields(spec, seed=seed, gates=gates)
    out = []
    for name in spec.fields:
        v = np.ma.getdata(fields[name]).astype(np.float32).copy()
        v[np.ma.getmaskarray(fields[name])] = np.nan       # the device re-derives the masks (masked_invalid)
        out.append(v)
    return fields, out


def run_b200(args):
    import torch
    import radar_grid_b200 as rg
    from radar_grid_b200 import _native as N, synthetic as S

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if N.device_count() < 1 or not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: radar_grid_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    affinity = {"bound": False, "why": "single rank"}
    if world > 1:
        import torch.distributed as dist
        from radar_grid_b200 import distributed as D
        affinity = D.bind_host_to_gpu(local_rank)       # before any pinned buffer exists (first-touch placement)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    spec = S.SPECS[args.workload]
    if os.environ.get("RG_BENCH_FIELDS"):                      # kernel experiments only: first k fields of the workload
        import dataclasses
        spec = dataclasses.replace(spec, fields=spec.fields[:int(os.environ["RG_BENCH_FIELDS"])])
    F = len(spec.fields)
    nz, ny, nx = spec.grid_shape
    V, ncol, G = nz * ny * nx, ny * nx, spec.n_gates

    stream = torch.cuda.Stream()
    ctx = N.Context(local_rank, stream.cuda_stream)
    if os.environ.get("RG_APPLY_VARIANT"):
        ctx.set_option("apply_variant", int(os.environ["RG_APPLY_VARIANT"]))
    if os.environ.get("RG_SORT_ROWS"):
        ctx.set_option("sort_rows", int(os.environ["RG_SORT_ROWS"]))
    if os.environ.get("RG_GROUP_WIDTH"):
        ctx.set_option("group_width", int(os.environ["RG_GROUP_WIDTH"]))
    gates = S.gate_coordinates(spec)
    t0 = time.perf_counter()
    dev = rg.DeviceGeometry.build(*gates, spec.grid_shape, spec.grid_limits, min_radius=spec.min_radius,
                                  beam_factor=spec.beam_factor, weighting=spec.weighting, toa=spec.toa, ctx=ctx)
    build_wall = time.perf_counter() - t0
    info = dev.info
    P = dev.n_pairs

    fields_ma, raw = raw_fields(spec, seed=rank, gates=gates)
    products = [rg.ColumnMax(), rg.CAPPI(CAPPI_ALT)]

    with torch.cuda.stream(stream):
        dfields = [torch.from_numpy(r).cuda(non_blocking=False) for r in raw]
        out_grids = [torch.empty((nz, ny, nx), dtype=torch.float32, device="cuda") for _ in range(F)]
        out_prods = [torch.empty((F, ny, nx), dtype=torch.float32, device="cuda") for _ in products]

        # every step grids the volume through the same device buffers: marshal the call once, launch it per step
        call = rg.prepare_grid_fields(dev, dfields, mask_invalid=True, products=products, want_grid=True, ctx=ctx,
                                      out_grids=out_grids, out_products=out_prods)
        step = call.launch

        for _ in range(args.warmup):
            step()
        stream.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        launches0 = ctx.kernel_launches()
        ctx.set_option("timing", 1)
        sampler = ClockSampler(local_rank)
        sampler.start()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record(stream)
        for _ in range(args.steps):
            step()
        ev1.record(stream)
        stream.synchronize()
        torch.cuda.synchronize()
        clocks = sampler.stop()
        elapsed_ms = ev0.elapsed_time(ev1)
        if world > 1:
            dist.barrier()
        ctx.set_option("timing", 0)
        launches = ctx.kernel_launches() - launches0
        apply_ms, n_apply = ctx.kernel_time(1)
        pack_ms, n_pack = ctx.kernel_time(0)
        if world > 1:
            t = torch.tensor([elapsed_ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            elapsed_ms = float(t.item())

        # ---- end-to-end: host buffers in, host buffers out, through the public API (copies inside the calls).
        # Consecutive volumes go through VolumePipeline (3 contexts/streams), as a time series would: the H2D copy of
        # volume i+1 and the D2H copy of volume i-1 overlap the kernels of volume i.  Every volume is copied in full.
        e2e_steps = max(3, min(args.steps, args.e2e_steps))
        n_slots = max(1, args.e2e_streams)
        pipe = rg.VolumePipeline(dev, n_streams=n_slots)
        slots = []
        for _ in range(n_slots):
            pin_in = []
            for r in raw:
                a = rg.pinned_empty(r.shape, np.float32)
                a[:] = r
                pin_in.append(a)
            slots.append({"fields": pin_in, "mask_invalid": True, "products": products, "want_grid": True,
                          "out_grids": [rg.pinned_empty((nz, ny, nx), np.float32) for _ in range(F)],
                          "out_products": [rg.pinned_empty((F, ny, nx), np.float32) for _ in products]})
        pipe.map([slots[i % n_slots] for i in range(2 * n_slots)])          # warm-up: staging buffers, first touches
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        pipe.map([slots[i % n_slots] for i in range(e2e_steps)])            # returns when all results are in host memory
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        # the same volume through one synchronous call (no overlap), for the record
        e2e_single_ms = float("inf")
        for _ in range(3):                                                  # best of three: the first one pays host-side one-offs
            t0 = time.perf_counter()
            rg.grid_fields(dev, ctx=ctx, **slots[0])
            e2e_single_ms = min(e2e_single_ms, (time.perf_counter() - t0) * 1e3)
        if world > 1:
            t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_s = float(t.item())
        # the same pipeline for a products-only request (COLMAX + CAPPI planes, no 3-D grid leaves the GPU): what an
        # operational product server asks for; 194 MB of the D2H traffic above become 9 MB
        pslots = [{"fields": sl["fields"], "mask_invalid": True, "products": products, "want_grid": False,
                   "out_products": sl["out_products"]} for sl in slots]
        pipe.map([pslots[i % n_slots] for i in range(2 * n_slots)])
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        pipe.map([pslots[i % n_slots] for i in range(e2e_steps)])
        torch.cuda.synchronize()
        e2e_prod_s = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([e2e_prod_s], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_prod_s = float(t.item())
        pipe.close()
        pin_grids = slots[0]["out_grids"]
        h2d = F * G * 4
        d2h = F * V * 4 + len(products) * F * ncol * 4
        # parity spot check of what the timed path produced (device outputs vs host-path outputs)
        same = all(torch.equal(out_grids[f].cpu().nan_to_num(-1e30), torch.from_numpy(pin_grids[f]).nan_to_num(-1e30))
                   for f in range(F))

    duo_slots = dev.duo_slots
    ms_per_step = elapsed_ms / args.steps
    value = world * F * V / (ms_per_step * 1e-3)
    peak, peak_src = load_peaks()
    b_alg = algorithmic_bytes(P, V, G, F, ncol)
    apply_avg_ms = apply_ms / max(n_apply, 1)
    achieved = b_alg / (apply_avg_ms * 1e-3) / 1e9 if apply_avg_ms > 0 else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic (seeded storm-cell volume, SURVEY.md 8d)",
        "config": {
            "workload": f"{spec.name}: {len(spec.elevations)} sweeps x {spec.nrays} x {spec.ngates} gates, fields "
                        f"{'+'.join(spec.fields)} -> {nz}x{ny}x{nx} grid, one shared neighbour table ({spec.weighting}); "
                        f"per step: mask+pack, {F}-field 3-D grids + COLMAX + CAPPI {CAPPI_ALT:.0f} m",
            "volumes_per_s": world / (ms_per_step * 1e-3),
            "pairs": P, "voxels": V, "gates": G, "fields": F,
            "l2_policy": "inputs larger than L2 (8P-byte pair stream = %.2f GB per step)" % (8 * P / 1e9),
            "parallelism": f"volume-batch shard x{world}, table replicated, no data-path collective",
            "host_affinity_rank0": affinity,
            "geometry_build_ms_device": info["build_ms"], "geometry_build_s_wall": build_wall,
            "geometry_candidates_per_pair": info["n_candidates"] / max(P, 1),
            "pack_ms_per_step": pack_ms / max(n_pack, 1), "apply_ms_per_step": apply_avg_ms,
            # > 0: the pass read the column-pair copy of the table (apply_duo_kernel); lane-slots per pair = 12-byte entries
            # (padding included) per 8-byte pair of the CSR table
            "duo_slots": duo_slots, "duo_lane_slots_per_pair": (32.0 * duo_slots / P) if duo_slots > 0 else None,
            "device_vs_host_path_identical": bool(same),
        },
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": (achieved / peak) if achieved else None, "traffic": measured_traffic(spec.name),
                     "kernel": "apply_duo_kernel" if duo_slots > 0 else "apply_columns_kernel", "algorithmic_bytes": b_alg, "peak_source": peak_src,
                     "frac_of_nominal_8TBs": (achieved / 8000.0) if achieved else None},
        "clocks": clocks,
        "e2e": {"value": world * F * V * e2e_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "steps": e2e_steps, "ms_per_step": e2e_s / e2e_steps * 1e3,
                "single_call_ms": e2e_single_ms, "streams": n_slots,
                "note": "pinned host fields in, 3-D grids + COLMAX + CAPPI planes back to pinned host memory, every volume "
                        "copied in full; VolumePipeline -> grid_fields() -> rg_apply(RG_HOST) on several streams (see \"streams\") so that copies of "
                        "neighbouring volumes overlap the kernels"},
        "e2e_products_only": {"value": world * F * V * e2e_steps / e2e_prod_s, "unit": UNIT, "h2d_bytes_per_step": h2d,
                              "d2h_bytes_per_step": len(products) * F * ncol * 4, "steps": e2e_steps,
                              "ms_per_step": e2e_prod_s / e2e_steps * 1e3, "streams": n_slots,
                              "note": "same pipeline and inputs, want_grid=False: COLMAX + CAPPI planes only come back"},
        "gpu_launches": launches,
    }
    if world > 1:
        line["zslab"] = zslab_record(rg, N, S, spec, gates, raw, dev, ctx, world, rank, local_rank)

    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_from_table(dev, spec, fields_ma)
    if world > 1:
        dist.destroy_process_group()
    return line if rank == 0 else None


def zslab_record(rg, N, S, spec, gates, raw, dev, ctx, world, rank, local_rank):
    """
    The z-slab shard of the SAME volume on the same ranks (N > 1 only): the grid is cut into `world` slabs balanced by
    pair count (level census, DeviceGeometry.level_pairs), every rank builds and grids only its slab (COLMAX + CAPPI
    4000 m as fused z-slab terms, no 3-D grid) and ONE all-reduce(MAX) + ONE all-reduce(SUM) over NCCL finish the
    planes.  Reported: wall ms per volume including the collectives (max over ranks), the collectives alone, the
    unsharded products-only pass on one GPU, and whether the planes are bit-identical to the unsharded ones.  (A cfg3
    volume takes 0.66 ms on ONE GPU: the two collectives and the host-side call overhead weigh as much as the kernels, so
    this record shows what the exchange costs rather than a speed-up; the shard is meant for grids like cfg5, see
    examples/zslab_colmax.py.)
    """
    import torch
    import torch.distributed as dist
    from radar_grid_b200 import distributed as D
    F = len(spec.fields)
    nz = spec.grid_shape[0]
    kw = dict(min_radius=spec.min_radius, beam_factor=spec.beam_factor, toa=spec.toa)
    census = rg.DeviceGeometry.level_pairs(*gates, spec.grid_shape, spec.grid_limits, column_stride=4, ctx=ctx, **kw)
    ranges = D.zslab_ranges(nz, world, weights=census)
    t0 = time.perf_counter()
    slab = rg.DeviceGeometry.build(*gates, spec.grid_shape, spec.grid_limits, weighting=spec.weighting, z_range=ranges[rank], ctx=ctx, **kw)
    build_s = time.perf_counter() - t0
    products = [rg.ColumnMax(), rg.CAPPI(CAPPI_ALT)]
    if rank != 0:
        _, raw = raw_fields(spec, seed=0, gates=gates)        # a z-slab shard grids ONE volume: rank 0's, on every rank
    dfields = [torch.from_numpy(r).cuda() for r in raw]
    ctx.set_option("group_width", 4)                       # same summation order for the slabs and the unsharded table
    try:
        want = rg.grid_fields(dev, dfields, mask_invalid=True, products=products, want_grid=False, ctx=ctx)["products"]
        got = D.zslab_products(slab, dfields, products, mask_invalid=True, ctx=ctx)
        identical = all(torch.equal(a.nan_to_num(-1e30), b.nan_to_num(-1e30)) for a, b in zip(got, want))
        reps, wall, coll, apply_ms = 20, [], [], []
        ctx.kernel_time(0), ctx.kernel_time(1)                # reset the per-kernel device timers
        ctx.set_option("timing", 1)
        for _ in range(3 + reps):
            dist.barrier()
            torch.cuda.synchronize()
            tm = {}
            t0 = time.perf_counter()
            D.zslab_products(slab, dfields, products, mask_invalid=True, ctx=ctx, timings=tm)
            torch.cuda.synchronize()
            wall.append((time.perf_counter() - t0) * 1e3)
            coll.append(tm["allreduce_ms"])
            apply_ms.append(tm["apply_ms"])
        ctx.set_option("timing", 0)
        dev_ms = (ctx.kernel_time(0)[0] + ctx.kernel_time(1)[0]) / (3 + reps)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            rg.grid_fields(dev, dfields, mask_invalid=True, products=products, want_grid=False, ctx=ctx)
        torch.cuda.synchronize()
        unsharded_ms = (time.perf_counter() - t0) * 1e3 / reps
    finally:
        ctx.set_option("group_width", int(os.environ.get("RG_GROUP_WIDTH") or 0))
    stats = torch.tensor([float(np.median(wall[3:])), float(np.median(coll[3:])), float(np.median(apply_ms[3:])), build_s,
                          float(slab.n_pairs), float(identical), dev_ms], dtype=torch.float64, device="cuda")
    allr = [torch.zeros_like(stats) for _ in range(world)]
    dist.all_gather(allr, stats)
    allr = torch.stack(allr).cpu().numpy()
    slab.close()
    return {"what": f"{spec.name} in {world} z-slabs balanced by pair count, COLMAX + CAPPI {CAPPI_ALT:.0f} m as fused z-slab terms, "
                    "all-reduce(MAX) + all-reduce(SUM) over NCCL; wall clock per volume incl. the collectives, median of 20",
            "slab_levels": [list(r) for r in ranges], "slab_pairs": [int(v) for v in allr[:, 4]],
            "wall_ms": float(allr[:, 0].max()), "allreduce_ms": float(allr[:, 1].max()),
            "slab_apply_wall_ms_per_rank": [float(v) for v in allr[:, 2]],
            "slab_kernels_device_ms_per_rank": [float(v) for v in allr[:, 6]],
            "slab_build_s_per_rank": [float(v) for v in allr[:, 3]], "unsharded_ms_one_gpu": unsharded_ms,
            "speedup_vs_one_gpu": unsharded_ms / float(allr[:, 0].max()), "identical": bool(allr[:, 5].min() == 1.0)}


def cpu_baseline_from_table(dev, spec, fields_ma):
    """The oracle port (NumPy, same arithmetic as the reference) on ONE field over the full table, 1 core."""
    from oracle import radar_grid_oracle as O
    indptr, idx, w = dev.export_csr()
    nz, ny, nx = spec.grid_shape
    name = spec.fields[0]
    t0 = time.perf_counter()
    grid = O.apply_geometry(indptr, idx, w, spec.grid_shape, fields_ma[name])
    O.column_reduce("max", grid)
    O.cappi(grid, spec.grid_shape, spec.grid_limits, CAPPI_ALT)
    dt = time.perf_counter() - t0
    return {"value": nz * ny * nx / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"1 of {len(spec.fields)} fields ({name}) over the full {nz}x{ny}x{nx} grid and table: apply_geometry + "
                      f"column_max + CAPPI, single-threaded NumPy as in the reference ({dt:.2f} s)",
            "host_cpus": os.cpu_count()}


# ---------------------------------------------------------------------------------------------------------
# reference arm: the reference's own CPU implementation of the path on the host cores — the genuine modules from
# oracle/_ref (byte-compiled from /root/reference by oracle/build_ref.py) when they travelled with the snapshot, else
# the NumPy oracle port.  Same workload as our arm: all levels, all fields, COLMAX + the real CAPPI blend.
# ---------------------------------------------------------------------------------------------------------
_REF = {}


def _ref_build_level(iz):
    from oracle import radar_grid_oracle as O
    s = _REF["spec"]
    return iz, O.build_geometry(*_REF["gates"], s.grid_shape, s.grid_limits, min_radius=s.min_radius,
                                beam_factor=s.beam_factor, weighting=s.weighting, toa=s.toa, z_range=(iz, iz + 1))


def _ref_field(name):
    """One field through the CPU path: apply_geometry + column_max + constant_altitude_ppi (single-threaded NumPy, as the
    reference runs it); the fields of a volume are independent, so they run in parallel worker processes."""
    import warnings
    s = _REF["spec"]
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        if _REF["ref"] is not None:
            ref, geom = _REF["ref"], _REF["geom"]
            grid = ref.interpolate.apply_geometry(geom, _REF["fields"][name])
            cmax = ref.products.column_max(grid)
            cap = ref.products.constant_altitude_ppi(grid, geom, CAPPI_ALT)
        else:
            from oracle import radar_grid_oracle as O
            indptr, idx, w = _REF["table"]
            grid = O.apply_geometry(indptr, idx, w, s.grid_shape, _REF["fields"][name])
            cmax = O.column_reduce("max", grid)
            cap = O.cappi(grid, s.grid_shape, s.grid_limits, CAPPI_ALT)
    return name, float(np.nansum(cmax)), float(np.nansum(cap))


def _ref_table(spec, gates, cores, genuine):
    """The neighbour table of the whole grid, built by the reference itself (or the port), cached in /tmp for the other
    invocations of the same round-end run (the reference's own save_geometry / load_geometry idea)."""
    import multiprocessing as mp
    import tempfile
    tag = "reference" if genuine else "port"
    path = os.path.join(tempfile.gettempdir(), f"rg_bench_{tag}_table_{spec.name}.npz")
    t0 = time.perf_counter()
    if os.path.exists(path):
        with np.load(path) as z:
            return (z["indptr"], z["gate_indices"], z["weights"]), time.perf_counter() - t0, "loaded from " + path
    if genuine:
        from oracle import build_ref
        ref = build_ref.load()
        with tempfile.TemporaryDirectory() as tmp:
            g = ref.compute.compute_grid_geometry(*gates, spec.grid_shape, spec.grid_limits, tmp, min_radius=spec.min_radius,
                                                  beam_factor=spec.beam_factor, weighting=spec.weighting, toa=spec.toa,
                                                  n_workers=max(1, cores - 1))
        table = (g.indptr, g.gate_indices, g.weights)
    else:
        with mp.get_context("fork").Pool(cores) as pool:
            parts = dict(pool.map(_ref_build_level, range(spec.grid_shape[0])))
        ptr, idx, w, base = [np.zeros(1, dtype=np.int64)], [], [], 0
        for iz in range(spec.grid_shape[0]):
            p, i, ww = parts[iz]
            ptr.append(np.asarray(p[1:], dtype=np.int64) + base)
            base += int(p[-1])
            idx.append(i)
            w.append(ww)
        table = (np.concatenate(ptr), np.concatenate(idx), np.concatenate(w))
    try:
        np.savez(path, indptr=table[0], gate_indices=table[1], weights=table[2])
    except OSError:
        pass
    return table, time.perf_counter() - t0, "built"


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return None
    import multiprocessing as mp
    import logging
    from radar_grid_b200 import synthetic as S
    from oracle import build_ref
    logging.disable(logging.INFO)
    spec = S.SPECS[args.workload]
    nz, ny, nx = spec.grid_shape
    F = len(spec.fields)
    cores = os.cpu_count() or 1
    genuine = build_ref.available()
    gates = S.gate_coordinates(spec)
    _REF.update(spec=spec, gates=gates, ref=None)
    table, table_s, how = _ref_table(spec, gates, cores, genuine)
    if genuine:
        ref = build_ref.load()
        _REF["ref"] = ref
        _REF["geom"] = ref.geometry.GridGeometry(grid_shape=spec.grid_shape, grid_limits=spec.grid_limits, indptr=table[0],
                                                 gate_indices=table[1], weights=table[2], toa=spec.toa)
        radar = S.SyntheticRadar(spec, seed=0)
        _REF["fields"] = {name: ref.utils.get_field_data(radar, name) for name in spec.fields}
    else:
        _REF["table"] = table
        _REF["fields"] = S.make_fields(spec, seed=0, gates=gates)
    # apply_geometry is single-threaded NumPy per field, so one volume keeps F cores busy; a time series keeps the rest
    # busy with further volumes in flight (what the reference's examples/batch_processing.py does with its executor):
    # cores // F volumes per step, bounded by host memory (every worker holds ~7 GB of temporaries at cfg3)
    try:
        with open("/proc/meminfo") as fh:
            avail_gb = next(int(l.split()[1]) for l in fh if l.startswith("MemAvailable")) / 1e6
    except Exception:
        avail_gb = 64.0
    per_worker_gb = max(0.5, 26.0 * int(table[0][-1]) / 1e9)          # six P-long float32 temporaries + the gathers
    n_vol = int(max(1, min(cores // F, (0.7 * avail_gb) // (per_worker_gb * F))))
    workers = min(cores, F * n_vol)
    jobs = list(spec.fields) * n_vol
    with mp.get_context("fork").Pool(workers) as pool:
        for _ in range(args.warmup):
            pool.map(_ref_field, jobs, chunksize=1)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            sums = pool.map(_ref_field, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    value = n_vol * F * nz * ny * nx * args.steps / dt
    kind = "reference" if genuine else "port"
    sample = (f"the full workload, nothing sampled ({n_vol} volume(s) in flight per step): all {nz} levels x {F} fields per volume through "
              f"{'the reference modules (oracle/_ref)' if genuine else 'the NumPy oracle port'}: apply_geometry + column_max + "
              f"constant_altitude_ppi({CAPPI_ALT:.0f} m) per field, {workers} worker processes (one field each, single-threaded "
              f"NumPy as in the reference); table of {int(table[0][-1])} pairs {how} in {table_s:.0f} s by "
              f"{'compute_grid_geometry, ' + str(max(1, cores - 1)) + ' workers' if genuine else 'the port'} (untimed, as on our arm)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps / n_vol * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic (seeded storm-cell volume, SURVEY.md 8d)",
        "config": {"workload": f"{spec.name}: {len(spec.elevations)} sweeps x {spec.nrays} x {spec.ngates} gates, fields "
                               f"{'+'.join(spec.fields)} -> {nz}x{ny}x{nx} grid ({spec.weighting}); per volume: {F}-field 3-D grids + "
                               f"COLMAX + CAPPI {CAPPI_ALT:.0f} m on the CPU path of the reference",
                   "pairs": int(table[0][-1]), "voxels": nz * ny * nx, "fields": F, "volumes_per_step": n_vol,
                   "colmax_checksums": [s[1] for s in sums[:F]]},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": workers, "kind": kind, "sample": sample, "host_cpus": cores},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    return line


class _QuietStdout:
    """Route stdout to stderr while the benchmark runs (NCCL and friends print banners to fd 1), so that the ONLY
    thing on stdout is the final JSON line."""

    def __enter__(self):
        sys.stdout.flush()
        self._saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self._saved, 1)
        os.close(self._saved)
        return False


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=["cfg3", "cfg1", "cfg2", "small", "tiny"])
    ap.add_argument("--e2e-steps", type=int, default=48)
    ap.add_argument("--e2e-streams", type=int, default=3, help="contexts/streams of the end-to-end VolumePipeline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = 3 if args.steps is None else max(1, args.steps)
        args.warmup = 1 if args.warmup is None else max(0, args.warmup)
        with _QuietStdout():
            line = run_reference(args)
    else:
        args.steps = 300 if args.steps is None else max(1, args.steps)
        args.warmup = 10 if args.warmup is None else max(3, args.warmup)
        with _QuietStdout():
            line = run_b200(args)
    if line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
