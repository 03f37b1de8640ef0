"""CPU suite: the N>1 host logic (row sharding, flattened gradient all-reduce) with world_size 2 on gloo."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ibl_nerf_b200 import training
    # (1) contiguous row tiles cover [0, n) exactly once
    n = 1001
    lo, hi = training.shard_rows(n, rank, world)
    cover = torch.zeros(n)
    cover[lo:hi] = 1
    dist.all_reduce(cover)
    ok_cover = bool((cover == 1).all())
    # (2) flattened gradient all-reduce == mean of per-rank gradients (TrainStep.allreduce_grads without a GPU)
    torch.manual_seed(0)
    params = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7))]
    for i, p in enumerate(params):
        p.grad = torch.full_like(p, float(rank + 1 + i))
    ts = training.TrainStep.__new__(training.TrainStep)
    ts.params, ts.world = params, world
    ts.allreduce_grads()
    want = [sum(r + 1 + i for r in range(world)) / world for i in range(2)]
    ok_grad = all(torch.allclose(p.grad, torch.full_like(p, w)) for p, w in zip(params, want))
    # (3) the fused path: both networks' gradients live in ONE flat buffer (training.FlatParameters) that is
    # all-reduced in place; the modules' p.grad views see the reduced values and the parameters stay views
    torch.manual_seed(1)
    nets = [training.IBLNeRF(**training.KITCHEN_ARCH) for _ in range(2)]
    ref = [p.detach().clone() for net in nets for p in net.ordered_params()]
    flat = training.FlatParameters(nets)
    ok_flat = flat.n == 2 * 798994 and all(torch.equal(p.detach(), r) for p, r in
                                           zip([p for net in nets for p in net.ordered_params()], ref))
    flat.grad.fill_(float(rank + 1))
    dist.all_reduce(flat.grad, op=dist.ReduceOp.SUM)
    tot = float(sum(range(1, world + 1)))
    ok_flat = ok_flat and all(bool((p.grad == tot).all()) and p.grad.data_ptr() >= flat.grad.data_ptr()
                              for net in nets for p in net.parameters())
    ok_flat = ok_flat and nets[1]._grad_sink.data_ptr() == flat.grad.data_ptr() + 4 * 798994
    # (4) the final gather of the tile-sharded render: every map of this rank's row tile packed into one buffer, ONE
    # all_gather, unpacked to full-size maps (ragged last tile: 1001 rays over 2 ranks)
    g = torch.Generator().manual_seed(3)
    full = {"color_map": torch.rand(n, 3, generator=g), "depth_map": torch.rand(n, generator=g), "weights": torch.rand(n, 5, generator=g)}
    mine = {k: v[lo:hi].clone() for k, v in full.items()}
    calls = []
    orig = dist.all_gather_into_tensor
    dist.all_gather_into_tensor = lambda *a, **k: (calls.append(1), orig(*a, **k))[1]
    got = training.gather_maps(mine, sorted(full), n, world)
    dist.all_gather_into_tensor = orig
    ok_gather = len(calls) == 1 and all(torch.equal(got[k], full[k]) for k in full)
    q.put((rank, ok_cover, ok_grad and ok_flat and ok_gather))
    dist.destroy_process_group()


def test_shard_rows_and_flat_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert len(res) == 2 and all(a and b for _, a, b in res), res


def test_shard_rows_edge_cases():
    sys.path.insert(0, ROOT)
    from ibl_nerf_b200 import training
    for n in (0, 1, 7, 8, 4096):
        for w in (1, 2, 4, 8):
            seen = []
            for r in range(w):
                lo, hi = training.shard_rows(n, r, w)
                assert 0 <= lo <= hi <= n
                seen += list(range(lo, hi))
            assert seen == list(range(n))
