#!/usr/bin/env python3
"""Per-slab carve times of an N-GPU run, measured on ONE GPU: plans the balanced z-slabs for N = 2, 4, 8 and carves every slab
on its own (the slabs are independent, so the max over the slabs is what an N-GPU step costs; SCALE runs confirm it).

  python tools/slab_timing.py [--config C4] [--parts 2,4,8] [--reps 7]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import ar_voxel_project_b200 as A
from ar_voxel_project_b200.synth import Workload, CONFIGS

ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C4")
ap.add_argument("--parts", default="1,2,4,8")
ap.add_argument("--reps", type=int, default=7)
ap.add_argument("--profiling", action="store_true", help="plain launches with the classify / per-voxel split instead of the cached CUDA graph")
a = ap.parse_args()
w = Workload(**CONFIGS[a.config])
out = {"config": a.config}
with A.VoxelEngine(w.X, w.Y, w.Z, w.s) as e:
    e.set_views(w.P, w.W, w.H, w.M)
    e.set_masks_bits(w.mask_bits)
    e.set_profiling(a.profiling)
    for n in [int(x) for x in a.parts.split(",")]:
        e.set_slab(0, w.Z)
        bounds = e.plan_slabs(n) if n > 1 else [0, w.Z]
        rows = []
        for r in range(n):
            e.set_slab(bounds[r], bounds[r + 1])
            ts, cs = [], []
            for _ in range(a.reps):
                e.reset()
                e.carve(0)
                e.synchronize()
                st = e.stats()
                ts.append(st["last_carve_ms"]), cs.append(st["last_classify_ms"])
            rows.append((float(np.median(ts)), float(np.median(cs))))
        out[f"x{n}"] = {"bounds": bounds, "carve_ms": [round(t, 4) for t, _ in rows], "classify_ms": [round(c, 4) for _, c in rows],
                        "max_carve_ms": max(t for t, _ in rows)}
        print(f"N={n}: max carve {out[f'x{n}']['max_carve_ms']:.4f} ms  carve {out[f'x{n}']['carve_ms']}  classify {out[f'x{n}']['classify_ms']}", flush=True)
print(json.dumps(out))
