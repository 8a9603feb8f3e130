"""Device batch loader against the restated reference loader: same samples, same augmentation draws."""
import ctypes as C
import os
import random
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import isr_b200  # noqa: E402,F401
from isr_b200 import _cabi as K  # noqa: E402
from isr_b200 import cache as CA  # noqa: E402
from oracle import cache_oracle as CO  # noqa: E402


def _collate(samples):
    out = {"lr": torch.stack([s["lr"] for s in samples]), "hr": torch.stack([s["hr"] for s in samples]),
           "expert_imgs": {k: torch.stack([s["expert_imgs"][k] for s in samples]) for k in samples[0]["expert_imgs"]},
           "filename": [s["filename"] for s in samples]}
    if "expert_feats" in samples[0]:
        out["expert_feats"] = {k: torch.stack([s["expert_feats"][k] for s in samples]) for k in samples[0]["expert_feats"]}
    return out


def _check(batch, want, cast=lambda t: t):
    assert batch["filename"] == want["filename"]
    for k in ("lr", "hr"):
        assert torch.equal(batch[k].cpu(), want[k]), k
    for grp in ("expert_imgs", "expert_feats"):
        assert (grp in batch) == (grp in want)
        if grp in want:
            assert list(batch[grp].keys()) == ["drct", "grl", "nafnet", "mamba"]
            for k in want[grp]:
                assert batch[grp][k].dtype == torch.float32
                assert torch.equal(batch[grp][k].cpu(), cast(want[grp][k])), (grp, k)


@pytest.mark.gpu
@pytest.mark.parametrize("mode,load_features", [("source", True), ("source", False), ("fp16", True)])
def test_device_batches_equal_the_reference_loader(tmp_path, mode, load_features):
    dev = torch.device("cuda:0")
    d = tmp_path / "cache"
    CO.write_mock_cache(d, n=7, lr_hw=(16, 16), seed=21, mamba_missing=(4,))
    shard = tmp_path / "s.ffsrc"
    CA.pack_cache(str(d), str(shard), dtype=mode)
    ref = CO.OracleCachedDataset(str(d), augment=True, repeat_factor=2, load_features=load_features)
    loader = CA.DeviceBatchLoader(str(shard), 3, dev, augment=True, shuffle=False, drop_last=False, repeat_factor=2,
                                  load_features=load_features, rng=random.Random(5), depth=2)
    assert len(loader) == 5
    random.seed(5)                                   # the same Mersenne stream the loader's rng walks
    n = 0
    cast = (lambda t: t.half().float()) if mode == "fp16" else (lambda t: t)
    for b, batch in enumerate(loader):
        idxs = list(range(3 * b, min(3 * b + 3, 14)))
        want = _collate([ref[i] for i in idxs])
        _check(batch, want, cast)
        n += 1
    assert n == 5 and loader.launches == 5           # one kernel per batch
    # second epoch keeps going (staging slots are recycled), shuffled order is a permutation split over ranks
    la = CA.DeviceBatchLoader(str(shard), 2, dev, augment=False, shuffle=True, rank=0, world=2, seed=3)
    lb = CA.DeviceBatchLoader(str(shard), 2, dev, augment=False, shuffle=True, rank=1, world=2, seed=3)
    names = [f for bt in la for f in bt["filename"]] + [f for bt in lb for f in bt["filename"]]
    # 7 samples, 2 ranks x batch 2: the shared permutation is cut to 4 so that both ranks run the SAME number of steps
    # (every step is a collective; epoch_batches) -> one full batch per rank, all distinct
    assert len(names) == 4 and len(set(names)) == 4
    e2 = [f for bt in la for f in bt["filename"]]
    assert len(e2) == 2 and la.epoch == 2


@pytest.mark.gpu
def test_non_square_samples_and_bf16_output(tmp_path):
    dev = torch.device("cuda:0")
    d = tmp_path / "cache"
    CO.write_mock_cache(d, n=4, lr_hw=(9, 14), seed=8)
    shard = tmp_path / "s.ffsrc"
    CA.pack_cache(str(d), str(shard))
    ref = CO.OracleCachedDataset(str(d), augment=True)
    loader = CA.DeviceBatchLoader(str(shard), 1, dev, augment=True, shuffle=False, rng=random.Random(2))
    random.seed(2)
    shapes = set()
    for i, batch in enumerate(loader):
        want = _collate([ref[i]])
        _check(batch, want)
        shapes.add(tuple(batch["lr"].shape[2:]))
    assert shapes <= {(9, 14), (14, 9)}
    with pytest.raises(ValueError, match="non-square"):
        for _ in CA.DeviceBatchLoader(str(shard), 4, dev, augment=True, shuffle=False, rng=random.Random(2)):
            pass
    lb = CA.DeviceBatchLoader(str(shard), 2, dev, augment=False, shuffle=False, out_dtype=torch.bfloat16)
    nb = next(iter(lb))
    refn = CO.OracleCachedDataset(str(d), augment=False)
    assert nb["expert_feats"]["grl"].dtype == torch.bfloat16
    assert torch.equal(nb["expert_feats"]["grl"].cpu(), torch.stack([refn[0]["expert_feats"]["grl"], refn[1]["expert_feats"]["grl"]]).bfloat16())


@pytest.mark.gpu
def test_loader_feeds_the_fusion_module(tmp_path):
    dev = torch.device("cuda:0")
    d = tmp_path / "cache"
    CO.write_mock_cache(d, n=2, lr_hw=(16, 16), seed=4)
    CA.pack_cache(str(d), str(tmp_path / "s.ffsrc"))
    torch.manual_seed(0)
    m = isr_b200.CompleteEnhancedFusionSR(None).eval().to(dev)
    batch = next(iter(CA.DeviceBatchLoader(str(tmp_path / "s.ffsrc"), 2, dev, augment=False, shuffle=False)))
    sr = m.forward_with_precomputed(batch["lr"], batch["expert_imgs"], batch["expert_feats"])
    ref = CO.OracleCachedDataset(str(d), augment=False)
    want = _collate([ref[0], ref[1]])
    sr2 = m.forward_with_precomputed(want["lr"].to(dev), {k: v.to(dev) for k, v in want["expert_imgs"].items()},
                                     {k: v.to(dev) for k, v in want["expert_feats"].items()})
    assert tuple(sr.shape) == (2, 3, 64, 64) and torch.equal(sr, sr2)


def test_cache_unpack_rejects_bad_arguments():
    """Argument validation happens before any launch: checkable without a GPU."""
    lib = K.load()
    seg = (K.CacheSegment * 1)()
    seg[0].src_offset, seg[0].dst, seg[0].C, seg[0].h, seg[0].w = 0, 4096, 3, 8, 8
    seg[0].src_dtype, seg[0].dst_dtype = K.DT_F32, K.DT_F32
    assert lib.ffsr_cache_unpack(4096, 1024, 1, seg, 0, None, 148, None) == -1            # no segments
    assert lib.ffsr_cache_unpack(4096, 512, 1, seg, 1, None, 148, None) == -1             # tensor leaves the record
    assert lib.ffsr_cache_unpack(4100, 1024, 1, seg, 1, None, 148, None) == -2            # misaligned records
    seg[0].src_dtype = K.DT_BF16
    assert lib.ffsr_cache_unpack(4096, 1024, 1, seg, 1, None, 148, None) == -1            # bf16 is not a storage dtype
    assert b"cache_unpack" in lib.ffsr_last_error()
    assert C.sizeof(K.CacheSegment) == lib.ffsr_cache_segment_size() == 40


@pytest.mark.gpu
def test_validate_epoch_and_ema_swap(tmp_path):
    """trainer.validate_epoch (train.py:415-515, cached mode) over a device loader of per-image-sized samples: metrics equal
    the metric oracle applied to the same outputs; the EMA exchange puts the shadow in, and the weights back, bit-exactly."""
    import numpy as np
    from isr_b200 import losses as FL
    from isr_b200.trainer import FusionTrainer, validate_epoch
    from oracle import loss_oracle as L
    dev = torch.device("cuda:0")
    d = tmp_path / "val"
    d.mkdir()
    for i, hw in enumerate([(16, 24), (24, 16), (16, 16)]):               # val caches: one full image per sample
        CO.write_mock_cache(d / f"p{i}", n=1, lr_hw=hw, seed=30 + i)
        for f in os.listdir(d / f"p{i}"):
            os.rename(d / f"p{i}" / f, d / f.replace("img_000", f"img_{i:03d}"))
    CA.pack_cache(str(d), str(tmp_path / "val.ffsrc"), dtype="fp16")
    torch.manual_seed(0)
    m = isr_b200.CompleteEnhancedFusionSR(None).to(dev)
    crit = FL.CombinedLoss()
    crit.set_weights({"charbonnier": 0, "l2": 0, "vgg": 0, "edge": 0, "clip": 0, "l1": 1.0, "swt": 0, "fft": 0, "ssim": 0})
    tr = FusionTrainer(m, crit, lr=1e-2, ema_decay=0.9, cuda_graph=False)
    lr, imgs, fts, hr = __import__("oracle.fusion_oracle", fromlist=["x"]).synthetic_inputs(2, 12, 12)
    for _ in range(2):                                                      # move the weights away from their EMA
        tr.step(lr.to(dev), {k: v.to(dev) for k, v in imgs.items()}, {k: v.to(dev) for k, v in fts.items()}, hr.to(dev))
    opt = tr.optimizer
    w0, e0 = opt.bucket.flat.clone(), opt.ema.clone()
    assert not torch.equal(w0, e0)
    names = [n for n, _ in m.named_parameters()]
    with tr.ema_weights():
        assert torch.equal(opt.bucket.flat, e0) and torch.equal(opt.ema, w0)
        o = opt.bucket.offsets[5]
        p5 = dict(m.named_parameters())[names[5]]
        assert torch.equal(p5.data.reshape(-1), e0[o:o + p5.numel()])        # parameters are views of the bucket
    assert torch.equal(opt.bucket.flat, w0) and torch.equal(opt.ema, e0)

    loader = CA.DeviceBatchLoader(str(tmp_path / "val.ffsrc"), 1, dev, augment=False, shuffle=False, drop_last=False)
    got = tr.validate(loader, crop_border=4, test_y_channel=True, use_ema=True)
    assert m.training and torch.equal(opt.bucket.flat, w0)                  # mode and weights restored
    # the same thing by hand: EMA weights, eval forward, oracle metrics on the host
    opt.swap_ema()
    m.eval()
    ps, ss = [], []
    for batch in CA.DeviceBatchLoader(str(tmp_path / "val.ffsrc"), 1, dev, augment=False, shuffle=False, drop_last=False):
        sr = m.forward_with_precomputed(batch["lr"], batch["expert_imgs"], batch["expert_feats"]).clamp(0, 1)
        ps.append(L.metric_psnr(sr[0].cpu(), batch["hr"][0].cpu(), 4, True))
        ss.append(L.metric_ssim(sr[0].cpu(), batch["hr"][0].cpu(), 4, True))
    opt.swap_ema()
    m.train()
    assert abs(got["psnr"] - float(np.mean(ps))) < 2e-3 and abs(got["ssim"] - float(np.mean(ss))) < 2e-5
    no_ema = validate_epoch(m, loader, dev, 4, True, None)
    assert no_ema["psnr"] != got["psnr"]
