#!/usr/bin/env python3
"""ONE process driving one engine per visible GPU (what a C++ host built on include/voxcarve_host.hpp does): balanced slabs,
carve, then vc_gather_peer - every engine pulls the other slabs into its whole-grid buffers with peer copies over NVLink.
Prints the gather time (wall clock around enqueue + synchronise of all engines, best / median of --reps), the bytes every GPU
receives and checks the gathered grid of the first and last engine against a single-engine carve.

  python tools/peer_gather_bench.py [--config C4] [--gpus N] [--reps 7]"""
import argparse
import hashlib
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import ar_voxel_project_b200 as A
from ar_voxel_project_b200.engine import gather_peer, exchange_halos_peer
from ar_voxel_project_b200.synth import Workload, CONFIGS


def h(a):
    return hashlib.blake2b(np.ascontiguousarray(a).view(np.uint8), digest_size=8).hexdigest()


ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C4")
ap.add_argument("--gpus", type=int, default=0)
ap.add_argument("--reps", type=int, default=7)
a = ap.parse_args()
import torch
n = a.gpus or torch.cuda.device_count()
w = Workload(**CONFIGS[a.config])
with A.VoxelEngine(w.X, w.Y, w.Z, w.s, device=0) as ref:
    ref.set_views(w.P, w.W, w.H, w.M)
    ref.set_masks_bits(w.mask_bits)
    bounds = ref.plan_slabs(n) if n > 1 else [0, w.Z]
    ref.carve()
    ref_hash = (h(ref.download_occupied()), h(ref.download_seen()))
    ref.mc_classify()
    ref_hist = ref.download_mc()[0]
engines = []
for r in range(n):
    e = A.VoxelEngine(w.X, w.Y, w.Z, w.s, z_begin=bounds[r], z_end=bounds[r + 1], device=r)
    e.alloc_full_volumes()
    e.set_views(w.P, w.W, w.H, w.M)
    e.set_masks_bits(w.mask_bits)
    engines.append(e)
out = {"config": a.config, "n_gpus": n, "bounds": bounds}
for what, key in (((True, False), "occupied"), ((True, True), "occupied_and_seen")):
    ts = []
    for _ in range(a.reps + 1):
        for e in engines:
            e.reset(), e.carve()
        for e in engines:
            e.synchronize()
        t0 = time.perf_counter()
        gather_peer(engines, occupied=what[0], seen=what[1])
        for e in engines:
            e.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    ts = ts[1:]  # the first round enables peer access
    recv = (w.Z - min(b1 - b0 for b0, b1 in zip(bounds[:-1], bounds[1:]))) * w.Y * ((w.X + 31) // 32) * 4 * (2 if what[1] else 1)
    out[key] = {"ms_best": min(ts), "ms_median": float(np.median(ts)), "max_bytes_received_per_gpu": recv, "GBps_received_per_gpu_at_best": recv / (min(ts) * 1e-3) / 1e9}
ok = True
for e in (engines[0], engines[-1]):
    ok = ok and (h(e.download_full(0)), h(e.download_full(1))) == ref_hash
# halos instead of the gather: consumers on the slabs
for e in engines:
    e.reset(), e.carve()
ts = []
for _ in range(a.reps):
    for e in engines:
        e.synchronize()
    t0 = time.perf_counter()
    exchange_halos_peer(engines)
    for e in engines:
        e.synchronize()
    ts.append((time.perf_counter() - t0) * 1e3)
tot = np.zeros(256, np.uint64)
for e in engines:
    e.mc_classify()
    tot += e.download_mc()[0]
out["halo_exchange_peer_ms_best"] = min(ts)
out["gathered_grid_equals_single_engine"] = bool(ok)
out["slab_histograms_sum_to_single_engine"] = bool(np.array_equal(tot, ref_hist))
for e in engines:
    e.close()
print(json.dumps(out))
sys.exit(0 if ok and out["slab_histograms_sum_to_single_engine"] else 1)
