"""Multi-GPU plumbing: the state batch shards across ranks with no data-path collective;
the only exchange is one replication of the built distance field per rebuild, and the gather of per-state verdicts.
One process per GPU, torch.distributed for the rendezvous (NCCL on GPUs, gloo in the CPU tests)."""
import numpy as np


def shard_range(n, rank, world):
    """Contiguous N/G split (the shape of evaluate_collision_checking_speed.cpp's per-thread shards)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def gather_verdicts(local, n_total, rank, world, dist=None, device="cpu"):
    """Rank 0 receives the full verdict vector ([n_total] uint8); other ranks get None."""
    if world == 1:
        return np.asarray(local, dtype=np.uint8)
    import torch
    sizes = [shard_range(n_total, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    buf = torch.zeros(width, dtype=torch.uint8, device=device)
    loc = torch.as_tensor(np.asarray(local, dtype=np.uint8), device=device)
    buf[:len(loc)] = loc
    out = [torch.zeros(width, dtype=torch.uint8, device=device) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, out, dst=0)
    if rank != 0:
        return None
    full = np.empty(n_total, dtype=np.uint8)
    for r, (lo, hi) in enumerate(sizes):
        full[lo:hi] = out[r][:hi - lo].cpu().numpy()
    return full


def broadcast_field_arrays(arrays, dist, src=0):
    """Replicates the built field (list of torch tensors: occupancy, d2[, neg_d2]) from `src` to every rank:
    the one exchange step of this path."""
    for a in arrays:
        dist.broadcast(a, src=src)
    return arrays


def replicate_field(field, dist, rank, src=0, device=None):
    """NCCL broadcast of a built engine.Field from rank `src` into the same-geometry field of every other rank."""
    import torch
    _, neg, nbytes, elem = field.device_arrays()
    nvox = field.dims[0] * field.dims[1] * field.dims[2]
    occ = torch.empty(nvox, dtype=torch.uint8, device=device)
    d2 = torch.empty(nbytes, dtype=torch.uint8, device=device)
    ng = torch.empty(nbytes, dtype=torch.uint8, device=device) if neg else None
    stream = torch.cuda.current_stream().cuda_stream
    if rank == src:
        field.export_device(occ.data_ptr(), d2.data_ptr(), ng.data_ptr() if neg else None, stream)
    torch.cuda.current_stream().synchronize()
    broadcast_field_arrays([occ, d2] + ([ng] if neg else []), dist, src)
    torch.cuda.current_stream().synchronize()
    if rank != src:
        field.import_device(occ.data_ptr(), d2.data_ptr(), ng.data_ptr() if neg else None, stream)
    return nvox + nbytes * (2 if neg else 1)
