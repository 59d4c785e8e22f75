"""N>1 plumbing on CPU: world_size-2 gloo (the GPU path uses the same functions over NCCL)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from icon_b200 import dist as D


def _free_port():
    """A port the OS says is free right now (a fixed pid-derived port collided once in a while)."""
    with socket.socket(socket.AF_INET, socket.SOCK_STREAM) as sk:
        sk.bind(("127.0.0.1", 0))
        return sk.getsockname()[1]


def test_shard_images_covers_every_image_once():
    for n, w in [(8, 1), (8, 2), (32, 8), (64, 8), (5, 4), (3, 8)]:
        got = [i for r in range(w) for i in D.shard_images(n, r, w)]
        assert got == list(range(n))
        sizes = [len(D.shard_images(n, r, w)) for r in range(w)]
        assert max(sizes) - min(s for s in sizes if s or True) <= (n + w - 1) // w


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = D.shard_images(5, rank, world)
        ms = D.max_over_ranks([10.0 + rank, 3.0 - rank], torch.device("cpu"))
        hdrs = D.gather_headers([float(rank), float(len(mine)), float(sum(mine))], torch.device("cpu"))
        dist.barrier()
        q.put((rank, mine, ms, hdrs))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_max_and_gather():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, m0, ms0, h0), (r1, m1, ms1, h1) = res
    assert m0 == [0, 1, 2] and m1 == [3, 4]
    assert ms0 == ms1 == [11.0, 3.0]                      # max over ranks, element-wise
    assert h0 == h1 == [[0.0, 3.0, 3.0], [1.0, 2.0, 7.0]]  # every rank sees every header, in rank order


def _fake_mesh(i):
    """Deterministic variable-length 'mesh' for image i (image 2 is empty: engine returned None)."""
    g = torch.Generator().manual_seed(100 + i)
    nv, nf = (0, 0) if i == 2 else (50 + 17 * i, 90 + 31 * i)
    dt = torch.float64 if i % 2 else torch.float32          # both export_mesh branches (PyMCubes f64 / kaolin f32)
    return torch.rand(nv, 3, generator=g, dtype=dt), torch.randint(0, max(nv, 1), (nf, 3), generator=g)


def _mesh_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = D.shard_images(5, rank, world)
        got, nbytes = D.gather_meshes([_fake_mesh(i) for i in mine], mine, torch.device("cpu"))
        dist.barrier()
        # numpy: pickled by value.  Tensors travel through an mp.Queue as file descriptors, which the parent can only
        # receive while this process is alive -- it exits right after the put.
        q.put((rank, {i: (v.numpy().copy(), f.numpy().copy()) for i, (v, f) in got.items()}, nbytes))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_mesh_gather_equals_single_process():
    """Per-image results after the sharded run + gather are identical to producing all images in one process."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_mesh_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(2)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, got0, b0), (_, got1, b1) = res
    assert got1 == {} and sorted(got0) == [0, 1, 2, 3, 4]
    for i in range(5):
        v, f = _fake_mesh(i)
        gv, gf = torch.from_numpy(got0[i][0]), torch.from_numpy(got0[i][1])
        assert gv.dtype == v.dtype and torch.equal(gv, v) and torch.equal(gf, f)
    assert b0 == b1 > 0                                      # bytes received on rank 0 == bytes sent by rank 1
    single, _ = D.gather_meshes([_fake_mesh(i) for i in range(5)], list(range(5)), torch.device("cpu"))
    assert sorted(single) == [0, 1, 2, 3, 4]
