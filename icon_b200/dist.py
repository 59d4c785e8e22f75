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


def gather_meshes(meshes, image_ids, device, dst=0):
    """The final gather of the path: every rank holds the meshes of ITS images
    (`meshes` = list of (verts [nv,3] float, faces [nf,3] int64), `image_ids` = their global image indices);
    rank `dst` receives all of them.

    Two phases: (1) all_gather of a fixed-size header per rank -- (image id, nv, nf) for up to the largest
    per-rank image count -- so that the receiver can allocate; (2) one grouped batch of point-to-point
    sends / receives of the variable-length vertex and face buffers (ncclSend / ncclRecv inside one group on the
    GPU, gloo send / recv in the CPU tests).  Vertices travel as float64 or float32 exactly as produced (no
    conversion), faces as int64.  Returns on rank `dst`: {image id: (verts, faces)} for all images, tensors on
    `device`; on other ranks: {} .  Also returns the number of payload bytes this rank sent or received."""
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank() if world > 1 else 0
    if world == 1:
        return {i: m for i, m in zip(image_ids, meshes)}, 0
    n_local = torch.tensor([len(meshes)], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n_local) for _ in range(world)]
    dist.all_gather(counts, n_local)
    max_n = max(int(c.item()) for c in counts)
    hdr = torch.full((max(max_n, 1), 4), -1, dtype=torch.int64, device=device)
    for j, (i, (v, f)) in enumerate(zip(image_ids, meshes)):
        hdr[j, 0], hdr[j, 1], hdr[j, 2] = int(i), v.shape[0], f.shape[0]
        hdr[j, 3] = 8 if v.dtype == torch.float64 else 4
    hdrs = [torch.empty_like(hdr) for _ in range(world)]
    dist.all_gather(hdrs, hdr)
    ops, out, nbytes = [], {}, 0
    if rank == dst:
        for r in range(world):
            for j in range(int(counts[r].item())):
                i, nv, nf, vb = (int(x) for x in hdrs[r][j])
                if r == dst:
                    out[i] = meshes[j]
                    continue
                v = torch.empty(nv, 3, dtype=torch.float64 if vb == 8 else torch.float32, device=device)
                f = torch.empty(nf, 3, dtype=torch.int64, device=device)
                out[i] = (v, f)
                if nv:
                    ops.append(dist.P2POp(dist.irecv, v, r))
                if nf:
                    ops.append(dist.P2POp(dist.irecv, f, r))
                nbytes += nv * 3 * vb + nf * 24
    else:
        for v, f in meshes:
            v, f = v.contiguous(), f.contiguous()
            if v.shape[0]:
                ops.append(dist.P2POp(dist.isend, v, dst))
            if f.shape[0]:
                ops.append(dist.P2POp(dist.isend, f, dst))
            nbytes += v.numel() * v.element_size() + f.numel() * 8
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return out, nbytes
