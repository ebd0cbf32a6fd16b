#!/usr/bin/env python3
"""time vc_mc_classify and vc_color on the C4 workload (1 GPU)"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import time, numpy as np, ar_voxel_project_b200 as A
from ar_voxel_project_b200.synth import Workload, CONFIGS
w = Workload(**CONFIGS["C4"])
with A.VoxelEngine(w.X, w.Y, w.Z, w.s) as e:
    e.set_views(w.P, w.W, w.H, w.M); e.set_masks_bits(w.mask_bits); e.set_images(w.images_bgr())
    e.carve(); e.synchronize()
    for _ in range(3):
        t=time.perf_counter(); e.mc_classify(); e.synchronize(); print("mc_classify ms", (time.perf_counter()-t)*1e3)
    print(e.download_mc()[1:])
    for mode in (2, 1, 2, 1):
        t=time.perf_counter(); e.color(mode); e.synchronize(); dt=(time.perf_counter()-t)*1e3
        idx,rgbn=e.download_colors(); print("color mode",mode,"ms %.3f"%dt,"surface voxels",len(idx), "obs mean", rgbn[:,3].mean(), flush=True)
