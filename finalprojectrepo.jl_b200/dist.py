"""One-process-per-GPU plumbing over torch.distributed (NCCL on the GPU box, gloo in CPU tests).

The data path has no NCCL collective: z-slab halos are pushed by the stencil kernel straight into the neighbour's
memory (CUDA IPC mapping over NVLink) and the per-iteration 8-byte norm exchange is a peer store + flag
(csrc/diffusion3d_kernels.cuh). torch.distributed only carries the out-of-band set-up (IPC handles), barriers and the
max-over-ranks of timings.  Replaces: init_global_grid / MPI.Allreduce! / gather! plumbing of
scripts-part1/part1_kernel_programming.jl:100-101,223 and part1_utils.jl:36-40.
"""
import os


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def slab_layout(rank, world):
    """dims = (1, 1, world): rank r hosts z-slab r.  Returns (nslabs_total, slab_begin, slab_count)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    return world, rank, 1


def global_nz(nz_local, world):
    """nz_g = dims*(nz-2)+2 (ImplicitGlobalGrid, overlap 2)."""
    return world * (nz_local - 2) + 2


def z_offset(rank, nz_local):
    """First global z index of slab `rank`'s local plane 0."""
    return rank * (nz_local - 2)


def all_gather_blobs(blob, dist=None):
    """Every rank contributes one opaque bytes blob; returns the list in rank order (identical on all ranks)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [blob]
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, blob)
    if any(len(b) != len(blob) for b in out):
        raise RuntimeError("ranks exported IPC blobs of different sizes (mismatched library builds?)")
    return out


def max_over_ranks(x, dist=None, device="cpu"):
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(x)
    import torch
    t = torch.tensor([float(x)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def connect(handle, dist=None):
    """Exchange the CUDA IPC handles of all ranks' arenas and map the neighbours (b2s_diff3d_ipc_connect)."""
    blobs = all_gather_blobs(handle.ipc_export(), dist)
    if len(blobs) > 1:
        handle.ipc_connect(blobs)
        dist.barrier()
    return len(blobs)


def gather_global(handle, dist=None):
    """gather!(Array(Ht), H_g) across processes: rank 0 receives (nx, ny, nz*world) -- (nx*dimx, ny*dimy, nz*dimz) for a
    general decomposition --, others None."""
    import numpy as np
    loc = handle.gather()
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return loc
    parts = [None] * dist.get_world_size() if dist.get_rank() == 0 else None
    dist.gather_object(loc, parts, dst=0)
    if dist.get_rank() != 0:
        return None
    if handle.dims[0] * handle.dims[1] > 1:  # general decomposition: every rank filled its own block of the global array
        out = parts[0].copy(order="F")
        for q in parts[1:]:
            out += q  # the blocks are disjoint, everything else is zero
        return out
    return np.asfortranarray(np.concatenate(parts, axis=2))
