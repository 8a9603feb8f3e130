"""world_size-2 gloo test (CPU) of the multi-GPU host logic: round-robin image sharding with no
data-path collective, max-over-ranks timing, sum of processed units."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n_items, out):
    sys.path.insert(0, ROOT)
    import isr_b200  # noqa: F401
    from isr_b200 import dist as D
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = D.shard_indices(n_items, rank, world)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    t = D.max_over_ranks(10.0 + rank)             # pretend rank r took 10+r ms
    total = D.sum_over_ranks(float(len(mine)))
    if rank == 0:
        out.put((gathered, t, total))
    dist.barrier()
    dist.destroy_process_group()


def test_round_robin_sharding_and_timing_reduction_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    n_items, world = 101, 2
    procs = [ctx.Process(target=_worker, args=(r, world, 29641, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered, t, total = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(i for g in gathered for i in g) == list(range(n_items))     # every image exactly once
    assert gathered[0] == list(range(0, n_items, 2)) and gathered[1] == list(range(1, n_items, 2))
    assert abs(len(gathered[0]) - len(gathered[1])) <= 1                      # balanced to one unit
    assert t == 11.0 and total == float(n_items)


def test_shard_edge_cases():
    sys.path.insert(0, ROOT)
    import isr_b200  # noqa: F401
    from isr_b200 import dist as D
    assert D.shard_indices(0, 0, 4) == []
    assert D.shard_indices(3, 3, 8) == []                                     # more ranks than images
    assert D.shard(list("abcde"), 1, 2) == ["b", "d"]
    assert D.max_over_ranks(3.5) == 3.5                                       # no process group: identity


def test_job_schedule_covers_every_image_once_and_tiles_the_tail():
    sys.path.insert(0, ROOT)
    import isr_b200  # noqa: F401
    from isr_b200 import dist as D
    for n, world in [(100, 1), (100, 2), (100, 4), (100, 8), (7, 3), (5, 8), (0, 4), (9, 8)]:
        whole, tail = D.job_schedule(n, world)
        assert len(whole) == world
        seen = sorted([i for w in whole for i in w] + [img for img, _, _ in tail])
        assert seen == list(range(n)), (n, world)
        for img, ranks, grid in tail:
            assert grid[0] * grid[1] == len(ranks) > 1
        used = [r for _, ranks, _ in tail for r in ranks]
        assert len(used) == len(set(used))                                    # a rank works on one tail image at most
    whole, tail = D.job_schedule(100, 8)
    assert all(len(w) == 12 for w in whole) and whole[3][:2] == [3, 11]      # items[rank::world]
    assert [(img, ranks, grid) for img, ranks, grid in tail] == [(96, [0, 1], (1, 2)), (97, [2, 3], (1, 2)),
                                                                  (98, [4, 5], (1, 2)), (99, [6, 7], (1, 2))]
    whole, tail = D.job_schedule(9, 8)                                        # one left-over image over all 8 ranks
    assert tail == [(8, list(range(8)), (2, 4))]


def _train_worker(rank, world, port, out):
    """N-rank flat-bucket step == single-process gradient accumulation over the same N micro-batches
    (what the reference's accumulation_steps does, train.py:331-357).  CPU tensors + gloo: exercises
    FlatBucket / allreduce_mean_ / sync_bn_buffers, the host logic of the NCCL path."""
    sys.path.insert(0, ROOT)
    import isr_b200  # noqa: F401
    from isr_b200.trainer import FlatBucket, allreduce_mean_, sync_bn_buffers
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.BatchNorm2d(8), torch.nn.ReLU(),
                              torch.nn.Conv2d(8, 3, 1))
    g = torch.Generator().manual_seed(42)
    xs = [torch.randn(2, 3, 8, 8, generator=g) for _ in range(world)]
    params = list(net.parameters())
    bucket = FlatBucket(params)
    grads = bucket.new_like()
    for i, p in enumerate(params):
        p.data = bucket.view(i)
        o = bucket.offsets[i]
        p.grad = grads[o:o + p.numel()].view(p.shape)
    net(xs[rank]).square().mean().backward()
    assert params[0].grad.data_ptr() == grads.data_ptr()            # autograd accumulated in place into the bucket
    allreduce_mean_(grads)
    sync_bn_buffers(net)
    if rank == 0:
        out.put((grads.clone(), net[1].running_mean.clone()))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_bucket_allreduce_equals_gradient_accumulation_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world = 2
    procs = [ctx.Process(target=_train_worker, args=(r, world, 29643, q)) for r in range(world)]
    for p in procs:
        p.start()
    grads, rmean = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference: accumulate loss/world over the same micro-batches
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.BatchNorm2d(8), torch.nn.ReLU(),
                              torch.nn.Conv2d(8, 3, 1))
    g = torch.Generator().manual_seed(42)
    xs = [torch.randn(2, 3, 8, 8, generator=g) for _ in range(world)]
    rms = []
    for x in xs:
        net[1].running_mean.zero_()
        (net(x).square().mean() / world).backward()
        rms.append(net[1].running_mean.clone())
    ref = torch.cat([torch.nn.functional.pad(p.grad.reshape(-1), (0, (-p.numel()) % 4)) for p in net.parameters()])
    assert torch.allclose(grads, ref, atol=1e-7)
    assert torch.allclose(rmean, sum(rms) / world, atol=1e-7)
