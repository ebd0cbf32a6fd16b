"""Worker of tests/test_gpu_parity.py::test_nccl_two_ranks_halos_gather_bitwise (launched under torch.distributed.run, one rank
per GPU).  Every rank: plans balanced slabs, carves its own, exchanges one-plane halos over NCCL inside libvoxcarve.so, runs the
cube-index and colour passes on the slab, all-reduces the histogram, gathers both volumes, and compares every word with a
single-GPU carve of the whole grid done on the same rank and with oracle planes around the slab boundary.
usage: nccl_slab_worker.py <out dir> [X Y Z V W H]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import ar_voxel_project_b200 as A
    from ar_voxel_project_b200.dist import init_engine_comm
    from ar_voxel_project_b200.synth import Workload
    from oracle import oracle as O

    out_dir = sys.argv[1]
    dims = [int(a) for a in sys.argv[2:8]] if len(sys.argv) >= 8 else [200, 96, 160, 10, 320, 240]
    X, Y, Z, V, W, H = dims
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")  # carries the 128-byte NCCL id only; the data path is NCCL inside libvoxcarve.so
    verdict = {"rank": rank, "world": world, "ok": False}
    try:
        w = Workload(max(X, Y, Z), V, W, H, seed=6, dims=(X, Y, Z))
        ref = A.VoxelEngine(X, Y, Z, w.s, device=local)
        ref.set_views(w.P, w.W, w.H, w.M), ref.set_masks_bits(w.mask_bits), ref.set_images(w.images_bgr())
        ref.carve()
        occ, seen = ref.download_occupied(), ref.download_seen()
        ref.mc_classify()
        hist = ref.download_mc()[0]
        ref.color(2)
        ridx, rrgbn = ref.download_colors()
        bounds = ref.plan_slabs(world)
        ref.close()
        z0, z1 = bounds[rank], bounds[rank + 1]
        e = A.VoxelEngine(X, Y, Z, w.s, z_begin=z0, z_end=z1, device=local)
        e.alloc_full_volumes()
        e.set_views(w.P, w.W, w.H, w.M), e.set_masks_bits(w.mask_bits), e.set_images(w.images_bgr())
        init_engine_comm(e)
        info = e.comm_info()
        verdict.update(nccl_version=info["nccl_version"], bounds=bounds)
        checks = {}
        for it in range(2):  # twice: the second round runs on warm communicators and re-validates the halo invalidation
            e.reset()
            e.carve()
            e.exchange_halos()
            e.mc_classify()
            tot = e.allreduce_u64(e.download_mc()[0])
            checks[f"hist{it}"] = bool(np.array_equal(tot, hist))
            e.color(2)
            idx, rgbn = e.download_colors()
            sel = (ridx >= np.uint64(X * Y * z0)) & (ridx < np.uint64(X * Y * z1))
            checks[f"colors{it}"] = bool(np.array_equal(idx, ridx[sel]) and np.array_equal(rgbn, rrgbn[sel]))
            checks[f"slab{it}"] = bool(np.array_equal(e.download_occupied(), occ[z0:z1]) and np.array_equal(e.download_seen(), seen[z0:z1]))
            e.gather(bounds, occupied=True, seen=True)
            e.synchronize()
            e.mc_classify()   # on the gathered grid now: same slab histogram
            checks[f"hist_gathered{it}"] = bool(np.array_equal(e.allreduce_u64(e.download_mc()[0]), hist))
        full_occ, full_seen = e.download_full(0), e.download_full(1)
        checks["gathered_occupied"] = bool(np.array_equal(full_occ, occ))
        checks["gathered_seen"] = bool(np.array_equal(full_seen, seen))
        # oracle planes around my slab's boundaries
        for zb in {max(z0 - 1, 0), min(z1 - 1, Z - 2)}:
            ro, rs = O.carve(X, Y, Z, w.s, w.P, w.W, w.H, mask_bits=w.mask_bits, z0=zb, z1=zb + 2)
            checks[f"oracle_planes_{zb}"] = bool(np.array_equal(full_occ[zb:zb + 2], ro) and np.array_equal(full_seen[zb:zb + 2], rs))
        e.close()
        verdict["checks"] = checks
        verdict["ok"] = all(checks.values())
    except Exception as ex:  # the verdict file must exist either way
        import traceback
        verdict["error"] = f"{type(ex).__name__}: {ex}\n{traceback.format_exc()}"
    json.dump(verdict, open(os.path.join(out_dir, f"rank{rank}.json"), "w"))
    dist.barrier()
    dist.destroy_process_group()
    return 0 if verdict["ok"] else 1


if __name__ == "__main__":
    sys.exit(main())
