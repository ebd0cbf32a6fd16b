"""N > 1 host logic on CPU: world_size-2 (and 3, ragged) gloo groups assemble z-slabs into the whole grid in place
and reduce per-slab cube-index histograms.  The slab contents come from the oracle (test infrastructure)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, Z, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ar_voxel_project_b200.dist import all_gather_slabs, all_reduce_counts, slab_range
        from ar_voxel_project_b200.synth import Workload
        from oracle import oracle as O
        w = Workload(48, 5, 160, 120, seed=4, dims=(70, 20, Z))
        z0, z1 = slab_range(rank, world, Z)
        occ_s, seen_s = O.carve(70, 20, Z, w.s, w.P, w.W, w.H, mask_bits=w.mask_bits, z0=z0, z1=z1)
        full = torch.zeros((Z, 20, 3), dtype=torch.int32)
        full[z0:z1] = torch.from_numpy(occ_s.view(np.int32))
        all_gather_slabs(full, Z)
        ref_occ, _ = O.carve(70, 20, Z, w.s, w.P, w.W, w.H, mask_bits=w.mask_bits)
        ok = np.array_equal(full.numpy().view(np.uint32), ref_occ)
        # per-slab cube-index histograms (cells whose lower plane is in the slab; plane -1 on rank 0) sum to the global one
        occ_full = full.numpy().view(np.uint32)
        tc = O.tri_counts()

        def cells(zlo, zhi):  # histogram of cells with lower plane z in [zlo, zhi)
            b = O.unpack(occ_full, 70)
            pad = np.zeros((Z + 2, 22, 72), bool)
            pad[1:-1, 1:-1, 1:-1] = b
            h = np.zeros(256, np.uint64)
            for z in range(zlo, zhi):
                lo, hi = pad[z + 1], pad[z + 2]
                c = [lo[:-1, 1:], lo[:-1, :-1], lo[1:, :-1], lo[1:, 1:], hi[:-1, 1:], hi[:-1, :-1], hi[1:, :-1], hi[1:, 1:]]
                idx = sum(((~c[i]).astype(np.int64) << i) for i in range(8))
                h += np.bincount(idx.ravel(), minlength=256).astype(np.uint64)
            return h
        mine = cells(-1 if rank == 0 else z0, z1)
        tot = all_reduce_counts(mine)
        rh, _, rnt = O.mc_classify(70, 20, Z, ref_occ)
        ok = ok and np.array_equal(tot, rh) and int((tot * tc).sum()) == rnt
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,Z", [(2, 16), (3, 16), (2, 7)])
def test_slab_gather_and_histogram_reduce(world, Z):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, Z, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(res) == [(r, True) for r in range(world)]


def _id_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ar_voxel_project_b200.dist import broadcast_comm_id
        q.put((rank, broadcast_comm_id().hex()))
    finally:
        dist.destroy_process_group()


def test_nccl_id_reaches_every_rank():
    """the plumbing of init_engine_comm without a GPU: rank 0 makes the 128-byte NCCL id inside libvoxcarve.so (NCCL is loaded with
    dlopen, no device needed for that), gloo carries it, every rank holds the same bytes"""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_id_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert len(res[0]) == 256 and res[0] == res[1] and set(res[0]) != {"0"}


def test_slab_ranges_tile_the_grid():
    from ar_voxel_project_b200.dist import slab_range
    for Z in (1, 7, 100, 1024, 2048):
        for world in (1, 2, 3, 4, 8):
            edges = [slab_range(r, world, Z) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == Z
            assert all(a[1] == b[0] for a, b in zip(edges[:-1], edges[1:]))
            sizes = [b - a for a, b in edges]
            assert max(sizes) - min(sizes) <= 1
