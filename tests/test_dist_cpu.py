"""world_size-2 gloo test (CPU) of the data-parallel gradient exchange host logic."""
import os
import socket

import torch
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    from flow2gan_b200.dist import GradBuckets, broadcast_module_state, cleanup_dist, setup_dist
    setup_dist(rank, world, backend="gloo")
    # construction-time sync (what the reference's DDP wrapper does): differently seeded replicas end
    # up with rank 0's parameters and buffers
    torch.manual_seed(100 + rank)
    mod = torch.nn.Sequential(torch.nn.Linear(5, 3), torch.nn.BatchNorm1d(3))
    mod[1].running_mean.fill_(float(rank + 1))
    nb = broadcast_module_state(mod)
    q.put(("sync", rank, nb, [t.detach().clone() for t in list(mod.parameters()) + list(mod.buffers())]))
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(s)) for s in ((7, 5), (3,), (), (1000,), (64, 9, 3))]
    for i, p in enumerate(params):
        p.grad = None if i == 1 and rank == 1 else torch.full_like(p, float(rank + 1) * (i + 1))
    b = GradBuckets(params, bucket_bytes=4096)         # forces several buckets
    nbytes = b.allreduce_mean()
    out = [p.grad.clone() for p in params]
    q.put((rank, nbytes, len(b.buckets), out))
    cleanup_dist()


def _run_once(world):
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        got = [q.get(timeout=180) for _ in range(2 * world)]
        sync = sorted([g for g in got if g[0] == "sync"], key=lambda t: t[1])
        assert sync[0][2] == sync[1][2] == len(sync[0][3]) > 0
        for a, b in zip(sync[0][3], sync[1][3]):
            assert torch.equal(a, b)
        assert float(sync[1][3][-2].mean()) == 1.0 or any(float(t.float().mean()) == 1.0 for t in sync[1][3])
        res = sorted([g for g in got if g[0] != "sync"], key=lambda t: t[0])
    finally:
        for p in procs:
            p.join(timeout=60)
            if p.is_alive():
                p.kill()
    assert all(p.exitcode == 0 for p in procs)
    return res


def test_grad_buckets_allreduce_mean_gloo():
    world = 2
    res = None
    for attempt in range(3):     # rendezvous port race / slow spawn on a busy box: retry on a fresh port
        try:
            res = _run_once(world)
            break
        except Exception:
            if attempt == 2:
                raise
    (_, nb0, k0, g0), (_, nb1, k1, g1) = res
    assert nb0 == nb1 == sum(t.numel() for t in g0) * 4 and k0 == k1 > 1
    for i, (a, b) in enumerate(zip(g0, g1)):
        want = (1.5 if i != 1 else 0.5) * (i + 1)       # rank 1 had no grad for tensor 1 -> treated as 0
        assert torch.equal(a, b)
        assert torch.allclose(a, torch.full_like(a, want))
