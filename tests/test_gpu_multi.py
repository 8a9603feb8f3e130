"""Multi-process NCCL tests of the inference partitioning (need >= 2 GPUs on the box; one process per GPU).  The host
logic (round-robin / tail schedule / timing reduction) is covered on the CPU with gloo in tests/test_dist_gloo.py."""
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, precision, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import isr_b200
    from isr_b200.serving import fuse_tiled
    from isr_b200.dist import job_schedule
    from oracle import fusion_oracle as O
    from oracle.perturb import perturb_state_dict
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    torch.manual_seed(0)
    m = isr_b200.CompleteEnhancedFusionSR(None)
    m.load_state_dict(perturb_state_dict(m.state_dict(), seed=3))
    m.eval().to(dev)
    m.precision = precision
    H, W = 48, 104
    lr, imgs, fts, _ = O.synthetic_inputs(1, H, W)                       # the same image on both ranks
    lr, imgs, fts = lr.to(dev), {k: v.to(dev) for k, v in imgs.items()}, {k: v.to(dev) for k, v in fts.items()}
    with torch.no_grad():
        whole = m.forward_with_precomputed(lr, imgs, fts)
        # (a) one image over both ranks: halo tiles, ONE all-reduce assembles the cores
        tiled = fuse_tiled(m, lr, imgs, fts, grid=(1, 2), rank=rank, world=world)
        err_tiled = float((tiled - whole).abs().max())
        # (b) the job schedule of bench.py on 3 images: one whole image per rank + the left-over image split over a group
        sched_whole, tail = job_schedule(3, world)
        groups = {img: dist.new_group(ranks) for img, ranks, grid in tail}
        n_whole = len(sched_whole[rank])
        errs = []
        for img, ranks, grid in tail:
            if rank in ranks:
                t = fuse_tiled(m, lr, imgs, fts, grid=grid, rank=ranks.index(rank), world=len(ranks), group=groups[img])
                errs.append(float((t - whole).abs().max()))
    torch.cuda.synchronize()
    dist.barrier()
    q.put((rank, err_tiled, n_whole, errs, [(img, ranks, tuple(grid)) for img, ranks, grid in tail]))
    dist.destroy_process_group()


@pytest.mark.parametrize("precision,tol", [("fp32", 1e-5), ("bf16", 4e-3)])
def test_fuse_tiled_across_two_processes_matches_whole_image(precision, tol):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices (one process per GPU; NCCL ranks must not share a GPU)")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, port = 2, 29733 + (0 if precision == "fp32" else 1)
    procs = [ctx.Process(target=_worker, args=(r, world, port, precision, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, err_tiled, n_whole, errs, tail in res:
        assert err_tiled <= tol, (rank, err_tiled)               # every rank holds the assembled image after the all-reduce
        assert n_whole == 1 and tail == [(2, [0, 1], (1, 2))]
        assert errs and max(errs) <= tol, (rank, errs)
