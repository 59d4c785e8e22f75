"""Multi-GPU plumbing: the path shards by IMAGE, one process per GPU, no data-path
collective.  torch.distributed (NCCL on GPUs, gloo in the CPU tests) is used for exactly three
things: the start barrier, the max-over-ranks step time and the final gather of per-image headers."""
import torch
import torch.distributed as dist


def shard_images(n_images, rank, world):
    """Contiguous block of image indices owned by `rank` (image i -> GPU i // ceil(n/world))."""
    per = (n_images + world - 1) // world
    lo = min(rank * per, n_images)
    return list(range(lo, min(lo + per, n_images)))


def max_over_ranks(values, device):
    """Element-wise max of a list of floats over all ranks (step times measured per rank)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


def gather_headers(header, device):
    """All ranks contribute one fixed-size header (list of floats); every rank gets the list of all."""
    h = torch.tensor(list(header), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        out = [torch.empty_like(h) for _ in range(dist.get_world_size())]
        dist.all_gather(out, h)
        return [[float(v) for v in o] for o in out]
    return [[float(v) for v in h]]
