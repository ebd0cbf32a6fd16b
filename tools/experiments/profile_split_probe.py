#!/usr/bin/env python3
"""Why does bench.py's profiling-mode split add up to more than the graph-mode step?  C4/C5, engine-owned whole-grid volumes."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import ar_voxel_project_b200 as A
from ar_voxel_project_b200.synth import Workload, CONFIGS

cfg = sys.argv[1] if len(sys.argv) > 1 else "C4"
dev = torch.device("cuda:0")
stream = torch.cuda.Stream(device=dev)
torch.cuda.set_stream(stream)
w = Workload(**CONFIGS[cfg])
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
with A.VoxelEngine(w.X, w.Y, w.Z, w.s) as e:
    e.set_stream(torch.cuda.current_stream().cuda_stream)
    e.set_views(w.P, w.W, w.H, w.M)
    e.set_masks_bits(w.mask_bits)
    if "--slab-volumes" not in sys.argv:
        e.alloc_full_volumes()
    for profiling in (False, True):
        e.set_profiling(profiling)
        for do_flush in (False, True):
            for sync_after_flush in (False, True):
                ts, ks, cs = [], [], []
                for i in range(8):
                    if do_flush:
                        flush.fill_(1)
                    if sync_after_flush:
                        torch.cuda.synchronize()
                    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    a.record()
                    e.reset(); e.carve(0)
                    b.record()
                    torch.cuda.synchronize()
                    st = e.stats()
                    if i >= 3:
                        ts.append(a.elapsed_time(b)); ks.append(st["last_carve_ms"]); cs.append(st["last_classify_ms"])
                print(f"{cfg} profiling={profiling} flush={do_flush} sync_after_flush={sync_after_flush}: torch events {np.median(ts):.4f} ms, "
                      f"engine events {np.median(ks):.4f} ms of which classification {np.median(cs):.4f}", flush=True)
