"""z-slab sharding of the grid over the GPUs of one box.

Voxels are independent, so each rank carves the contiguous slab z in [z0, z1) with every view; there
is no exchange during carve.  When a consumer needs the whole bit-packed grid (marching cubes, the
host Model), the slabs are assembled IN PLACE in a whole-grid buffer with one all-gather over
NVLink (torch.distributed = plumbing; NCCL on GPUs, gloo in the CPU tests).  Reductions of the
per-slab cube-index histograms / voxel counts are plain all-reduces.
"""
import numpy as np
import torch
import torch.distributed as dist


def slab_range(rank, world, Z):
    """contiguous z-range of `rank`; sizes differ by at most one plane"""
    return (rank * Z) // world, ((rank + 1) * Z) // world


def all_gather_slabs(full, Z, world=None, group=None, bounds=None):
    """full: tensor of shape (Z, ...) whose own slab is already filled in. Fills the other slabs in place.
    bounds: optional world+1 slab boundaries (e.g. from VoxelEngine.plan_slabs); default = slab_range."""
    world = dist.get_world_size(group) if world is None else world
    if world == 1:
        return full
    rank = dist.get_rank(group)
    flat = full.view(Z, -1)
    if bounds is None:
        bounds = [slab_range(r, world, Z)[0] for r in range(world)] + [Z]
    sizes = {bounds[r + 1] - bounds[r] for r in range(world)}
    if len(sizes) == 1:
        dist.all_gather_into_tensor(flat.view(-1), flat[bounds[rank]:bounds[rank + 1]].reshape(-1), group=group)
    elif dist.get_backend(group) == "nccl":
        # ragged slabs: torch's NCCL all_gather takes per-rank output views of different sizes and issues the per-owner
        # broadcasts inside one ncclGroup (one launch, all links busy at once)
        outs = [flat[bounds[r]:bounds[r + 1]] for r in range(world)]
        dist.all_gather(outs, flat[bounds[rank]:bounds[rank + 1]], group=group)
    else:  # gloo (CPU tests): one broadcast per owner
        for r in range(world):
            a, b = bounds[r], bounds[r + 1]
            if b > a:
                dist.broadcast(flat[a:b], src=dist.get_global_rank(group, r) if group is not None else r, group=group)
    return full


def broadcast_comm_id(group=None):
    """rank 0 makes the 128-byte NCCL id (vc_comm_unique_id), every rank of `group` returns the same bytes; any backend"""
    from .engine import comm_unique_id
    rank = dist.get_rank(group)
    box = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
    return box[0]


def init_engine_comm(engine, group=None):
    """Give `engine` its NCCL communicator inside libvoxcarve.so (vc_comm_init): torch.distributed (any backend) only carries
    the 128 bytes of the id to the other ranks.  From then on the data path - vc_exchange_halos, vc_gather,
    vc_comm_allreduce_u64 - runs in the library on the engine's stream; torch is out of it."""
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    engine.comm_init(rank, world, broadcast_comm_id(group))
    return rank, world


def all_reduce_counts(values, device=None, group=None):
    """sum of uint64 counters (histograms, voxel counts) over ranks -> numpy uint64"""
    t = torch.from_numpy(np.ascontiguousarray(values, np.uint64).view(np.int64).copy())
    if device is not None:
        t = t.to(device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t.cpu().numpy().view(np.uint64)
