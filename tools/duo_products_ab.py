#!/usr/bin/env python
"""Column-pair against column-group kernel for other product requests than the bench's (cfg3, five fields, device-resident):
the generic product path (COLMAX + COLMIN + COLMEAN + CAPPI + PPI), products only, no products.   usage: python tools/duo_products_ab.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "radar-processor_b200")]
import numpy as np
import torch

import radar_grid_b200 as rg
from radar_grid_b200 import _native as N, synthetic as S
import bench

spec = S.SPECS["cfg3"]
nz, ny, nx = spec.grid_shape
gates = S.gate_coordinates(spec)
_, raw = bench.raw_fields(spec, seed=0, gates=gates)
F = len(raw)
stream = torch.cuda.Stream()
ctx = N.Context(0, stream.cuda_stream)
dev = rg.DeviceGeometry.build(*gates, spec.grid_shape, spec.grid_limits, min_radius=spec.min_radius, beam_factor=spec.beam_factor,
                              weighting=spec.weighting, toa=spec.toa, ctx=ctx)
REQS = {
    "colmax+cappi (bench)": ([rg.ColumnMax(), rg.CAPPI(4000.0)], True),
    "colmax+cappi+ppi": ([rg.ColumnMax(), rg.CAPPI(4000.0), rg.PPI(1.0)], True),
    "max+min+mean+cappi+ppi": ([rg.ColumnMax(), rg.ColumnMin(), rg.ColumnMean(), rg.CAPPI(4000.0), rg.PPI(1.0)], True),
    "colmax+cappi+ppi, products only": ([rg.ColumnMax(), rg.CAPPI(4000.0), rg.PPI(1.0)], False),
    "colmax only, products only": ([rg.ColumnMax()], False),
    "grids only": ([], True),
}
out = {}
with torch.cuda.stream(stream):
    dfields = [torch.from_numpy(r).cuda() for r in raw]
    for name, (reqs, want_grid) in REQS.items():
        row = {}
        for duo in (0, 1):
            ctx.set_option("duo", duo)
            call = rg.prepare_grid_fields(dev, dfields, mask_invalid=True, products=reqs, want_grid=want_grid, ctx=ctx)
            for _ in range(5):
                call.launch()
            stream.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(30):
                call.launch()
            e1.record(stream)
            stream.synchronize()
            row["duo" if duo else "column_group"] = round(e0.elapsed_time(e1) / 30, 4)
        out[name] = row
        print(name, row, flush=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "duo_products_ab.json"), "w"), indent=1)
