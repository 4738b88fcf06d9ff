"""World-size-2 gloo tests of the batch sharding / gradient all-reduce plumbing (CPU, no GPU)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dc_vic_b200 import parallel as P


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def test_shard_bounds_cover_the_batch():
    for batch in (0, 1, 5, 6, 64):
        for world in (1, 2, 4, 8):
            spans = [P.shard_bounds(batch, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == batch
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c and a <= b and c <= d
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    x = torch.arange(10).view(5, 2)
    assert torch.equal(torch.cat([P.shard_batch(x, r, 2) for r in range(2)]), x)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        # a stand-in for "whatever stays trainable": codebook [16,4], entropy params, one frozen tensor
        params = [torch.nn.Parameter(torch.randn(16, 4)), torch.nn.Parameter(torch.randn(7)),
                  torch.nn.Parameter(torch.randn(3, 3), requires_grad=False), torch.nn.Parameter(torch.randn(5))]
        g = torch.Generator().manual_seed(100 + rank)
        params[0].grad = torch.randn(16, 4, generator=g)
        params[1].grad = torch.randn(7, generator=g)
        if rank == 0:
            params[3].grad = torch.randn(5, generator=g)      # rank 1 has no gradient for this one
        expect = []
        for i, p in enumerate(params):
            acc = torch.zeros_like(p)
            for r in range(world):
                gg = torch.Generator().manual_seed(100 + r)
                g0, g1 = torch.randn(16, 4, generator=gg), torch.randn(7, generator=gg)
                g3 = torch.randn(5, generator=gg) if r == 0 else torch.zeros(5)
                acc += {0: g0, 1: g1, 2: torch.zeros(3, 3), 3: g3}[i]
            expect.append(acc / world)
        n = P.allreduce_gradients(params, bucket_bytes=16 * 4 * 4)   # forces several buckets
        ok = n >= 2 and params[2].grad is None
        for i in (0, 1, 3):
            ok = ok and torch.allclose(params[i].grad, expect[i], atol=1e-6)
        ok = ok and P.max_over_ranks(float(rank)) == float(world - 1)
        ok = ok and P.sum_over_ranks([1.0, float(rank)]) == [float(world), float(sum(range(world)))]
        # sharded rate: per-rank bits summed == whole-batch bits
        lik = torch.rand(6, 8, generator=torch.Generator().manual_seed(5)) * 0.9 + 0.05
        mine = P.shard_batch(lik, rank, world)
        tot = P.sum_over_ranks([float(-torch.log2(mine).sum())])[0]
        ok = ok and abs(tot - float(-torch.log2(lik).sum())) < 1e-9 * abs(tot) + 1e-9
        # the flat bucket the gradient kernels write into (codebook dE + entropy-parameter grads), one collective
        b = P.GradBucket("cpu", [("codebook", (16, 4)), ("entropy", (7,)), ("quantiles", (5, 1, 3))])
        gb = torch.Generator().manual_seed(300 + rank)
        b.view("codebook").copy_(torch.randn(16, 4, generator=gb))
        b.view("entropy").copy_(torch.randn(7, generator=gb))
        b.view("quantiles").copy_(torch.randn(5, 1, 3, generator=gb))
        b.allreduce_async()
        b.wait()
        exp = {k: torch.zeros(v[2]) for k, v in b.offsets.items()}
        for r in range(world):
            ge = torch.Generator().manual_seed(300 + r)
            exp["codebook"] += torch.randn(16, 4, generator=ge)
            exp["entropy"] += torch.randn(7, generator=ge)
            exp["quantiles"] += torch.randn(5, 1, 3, generator=ge)
        for k in exp:
            ok = ok and torch.allclose(b.view(k), exp[k] / world, atol=1e-6)
        ok = ok and b.offsets["entropy"][0] % 4 == 0 and b.offsets["quantiles"][0] % 4 == 0
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_gradient_allreduce_world2():
    world = 2
    port = _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}
