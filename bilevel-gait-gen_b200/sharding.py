"""Host-side plumbing of the multi-GPU runs: MPC instances are independent, so a batch is cut into contiguous slices,
one per rank (one process per GPU), there is NO collective on the solve path, and the only exchanges are the barrier,
the max-over-ranks of the device time and the final gather of per-instance results.
torch.distributed is the transport (NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np


def shard_range(total, rank, world):
    """Contiguous slice [rank * total / world, (rank + 1) * total / world) of a batch of `total` instances."""
    if world <= 0 or not 0 <= rank < world:
        raise ValueError("bad rank / world size")
    return (rank * total) // world, ((rank + 1) * total) // world


def max_over_ranks(values, dist=None, device="cpu"):
    """Element-wise maximum of a list of floats over all ranks (every rank gets the result)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [float(v) for v in values]
    import torch
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.cpu()]


def gather_to_root(local, total, dist=None, device="cpu"):
    """Gather per-instance results (1-D numpy array over this rank's slice, slices as in shard_range) to rank 0 in global
    instance order.  Returns the concatenated array on rank 0 and None elsewhere."""
    local = np.ascontiguousarray(local)
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(total, r, world)[1] - shard_range(total, r, world)[0] for r in range(world)]
    if len(local) != sizes[rank]:
        raise ValueError(f"rank {rank}: slice has {len(local)} entries, expected {sizes[rank]}")
    pad = max(sizes)
    buf = torch.zeros(pad, dtype=torch.from_numpy(local).dtype, device=device)
    buf[:len(local)] = torch.from_numpy(local).to(device)
    out = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, out, dst=0)
    if rank != 0:
        return None
    return np.concatenate([o.cpu().numpy()[:n] for o, n in zip(out, sizes)])
