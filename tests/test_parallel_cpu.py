"""CPU suite, part 3: the N>1 host logic (stream sharding + max-over-ranks timing) under gloo, world_size 2."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from deft_b200.parallel import aggregate_throughput, stream_assignment


def test_stream_assignment_covers_all_streams_once():
    for n, w in ((6, 2), (6, 4), (6, 8), (1, 8), (17, 3)):
        owned = [stream_assignment(n, w, r) for r in range(w)]
        flat = sorted(s for o in owned for s in o)
        assert flat == list(range(n))
        assert max(len(o) for o in owned) - min(len(o) for o in owned) <= 1
    assert stream_assignment(6, 4, 0) == [0, 4] and stream_assignment(6, 8, 7) == []


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # rank 0: 10 frames in 100 ms, rank 1: 10 frames in 250 ms -> job = 20 frames / 0.25 s = 80 fps
    fps, ms = aggregate_throughput(10.0, 100.0 if rank == 0 else 250.0)
    dist.barrier()
    q.put((rank, fps, ms, stream_assignment(6, world, rank)))
    dist.destroy_process_group()


def test_two_rank_gloo_aggregation():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, fps, ms, streams in res:
        assert abs(fps - 80.0) < 1e-9 and abs(ms - 250.0) < 1e-9
    assert res[0][3] == [0, 2, 4] and res[1][3] == [1, 3, 5]


def test_single_process_passthrough():
    fps, ms = aggregate_throughput(30.0, 500.0)
    assert fps == 60.0 and ms == 500.0


def _ddp_worker(rank, world, port, q):
    import os
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from deft_b200.train import BucketedGradAllReduce
    torch.manual_seed(0)                                        # identical replicas
    net = torch.nn.Sequential(torch.nn.Linear(16, 64), torch.nn.ReLU(), torch.nn.Linear(64, 64), torch.nn.ReLU(), torch.nn.Linear(64, 3))
    unused = torch.nn.Parameter(torch.ones(5))                  # a parameter no loss touches: still reduced (as zeros)
    red = BucketedGradAllReduce(list(net.parameters()) + [unused], bucket_bytes=6000)
    assert len(red.buckets) >= 3
    g = torch.Generator(); g.manual_seed(100 + rank)            # different data per rank
    x = torch.randn(8, 16, generator=g)
    net(x).square().mean().backward()
    red.finish()
    q.put((rank, [p.grad.clone() for p in net.parameters()], unused.grad.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_bucketed_grad_allreduce_world2_gloo():
    """Config C5's collective: the bucketed, hook-driven gradient all-reduce equals the average of the per-rank gradients
    (world_size 2, gloo on CPU; the same class runs over NCCL on the GPUs)."""
    import torch
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29671
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # reference: the two local gradients averaged
    grads = []
    for rank in range(2):
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Linear(16, 64), torch.nn.ReLU(), torch.nn.Linear(64, 64), torch.nn.ReLU(), torch.nn.Linear(64, 3))
        g = torch.Generator(); g.manual_seed(100 + rank)
        net(torch.randn(8, 16, generator=g)).square().mean().backward()
        grads.append([p.grad.clone() for p in net.parameters()])
    want = [(a + b) / 2 for a, b in zip(*grads)]
    for rank in range(2):
        for got, w in zip(res[rank][1], want):
            assert torch.allclose(got, w, rtol=1e-6, atol=1e-7)
        assert float(res[rank][2].abs().max()) == 0.0
