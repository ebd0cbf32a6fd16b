#!/usr/bin/env python3
"""Device time of vc_set_masks (device-resident bit masks -> copy + summed-area tables) for library variants.
  python tools/sat_build_timing.py [--config C4] lib1.so lib2.so ..."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
ap = argparse.ArgumentParser()
ap.add_argument("--config", default="C4")
ap.add_argument("libs", nargs="*")
a = ap.parse_args()
import ar_voxel_project_b200._lib as L
from ar_voxel_project_b200.synth import Workload, CONFIGS
w = Workload(**CONFIGS[a.config])
default = L.LIB_PATH
for lib in (a.libs or [default]):
    L.LIB_PATH, L._lib = os.path.abspath(lib), None
    import ar_voxel_project_b200 as A
    st = torch.cuda.Stream()
    torch.cuda.set_stream(st)
    with A.VoxelEngine(w.X, w.Y, w.Z, w.s) as e:
        e.set_stream(st.cuda_stream)
        e.set_views(w.P, w.W, w.H, w.M)
        e.set_masks_bits(w.mask_bits)
        d_bits = torch.from_numpy(w.mask_bits.view(np.int32)).cuda()
        ts = []
        for _ in range(8):
            t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0.record()
            e.set_masks_bits_device(d_bits.data_ptr())
            t1.record()
            torch.cuda.synchronize()
            ts.append(t0.elapsed_time(t1))
        e.reset(), e.carve()
        n_occ = e.count_occupied()
    print(f"{os.path.basename(lib)}: set_masks (D2D copy + tables) {min(ts[2:]):.4f} ms best, {np.median(ts[2:]):.4f} median; occupied {n_occ}", flush=True)
