"""Loader throughput on the GPU box: the restated reference loader (three pickles per sample through
torch.load in DataLoader workers, CPU up-cast / flips, collate, pinned .to(device)) against pack_cache + DeviceBatchLoader
on the SAME synthetic cache of C2-sized samples (64x64 LR, 13.9 MB fp32 each).

    python tools/bench_loader.py [--samples 48] [--batch 32] [--epochs 3]

Prints one JSON line per arm.  The cache lives in a temporary directory (page-cache warm for BOTH arms: the arms
differ in parsing / copying work, not in disk speed)."""
import argparse
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", type=int, default=48)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--epochs", type=int, default=3)
    ap.add_argument("--workers", type=int, default=4)
    ap.add_argument("--skip-reference", action="store_true", help="only the shard arms (e.g. under ncu)")
    args = ap.parse_args()
    import isr_b200  # noqa: F401
    from isr_b200 import cache as CA
    from oracle import cache_oracle as CO
    dev = torch.device("cuda:0")
    repeat = max(1, (args.batch * 6) // args.samples)

    with tempfile.TemporaryDirectory() as tmp:
        d = os.path.join(tmp, "cache")
        CO.write_mock_cache(d, n=args.samples, lr_hw=(64, 64), seed=0)
        sample_mb = sum(os.path.getsize(os.path.join(d, f)) for f in os.listdir(d)) / args.samples / 1e6

        # ---- reference-style arm ------------------------------------------------------------------
        for workers in (() if args.skip_reference else (0, args.workers)):
            ds = CO.OracleCachedDataset(d, augment=True, repeat_factor=repeat)
            dl = torch.utils.data.DataLoader(ds, batch_size=args.batch, shuffle=True, num_workers=workers, pin_memory=True,
                                             drop_last=True, persistent_workers=workers > 0,
                                             prefetch_factor=4 if workers > 0 else None)
            n, t0 = 0, None
            for ep in range(args.epochs):
                for batch in dl:
                    lr = batch["lr"].to(dev, non_blocking=True)
                    hr = batch["hr"].to(dev, non_blocking=True)
                    imgs = {k: v.to(dev, non_blocking=True) for k, v in batch["expert_imgs"].items()}
                    fts = {k: v.to(dev, non_blocking=True) for k, v in batch["expert_feats"].items()}
                    if t0 is not None:
                        n += lr.shape[0]
                if ep == 0:
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()             # first epoch = warm-up (worker start, page cache)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            print(json.dumps({"arm": "reference-style loader (torch.load x3, CPU augment, collate, pinned H2D)",
                              "workers": workers, "samples_per_s": n / dt, "GB_per_s_fp32": n * 13.9e-3 / dt,
                              "sample_MB_on_disk": sample_mb, "batch": args.batch}), flush=True)
            del dl

        # ---- shard arm ----------------------------------------------------------------------------
        for mode in ("source", "fp16"):
            shard = os.path.join(tmp, f"train_{mode}.ffsrc")
            t0 = time.perf_counter()
            CA.pack_cache(d, shard, dtype=mode)
            pack_s = time.perf_counter() - t0
            loader = CA.DeviceBatchLoader(shard, args.batch, dev, augment=True, shuffle=True, repeat_factor=repeat)
            n, t0 = 0, None
            for ep in range(args.epochs):
                for batch in loader:
                    if t0 is not None:
                        n += batch["lr"].shape[0]
                if ep == 0:
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            # the kernel alone: records resident on the device, CUDA events on the stream it is launched on
            import ctypes as C
            from isr_b200 import _cabi as K
            slot = loader._slots[0]
            segs, rbytes = loader.cache.layout(0)
            outs = [torch.empty(args.batch, Cc, hh, ww, device=dev) for _, Cc, hh, ww, _, _ in segs]
            arr = (K.CacheSegment * len(segs))()
            rd = wr = 0
            for i, (key, Cc, hh, ww, sdt, off) in enumerate(segs):
                arr[i].src_offset, arr[i].dst, arr[i].C, arr[i].h, arr[i].w = off, outs[i].data_ptr(), Cc, hh, ww
                arr[i].src_dtype, arr[i].dst_dtype = (K.DT_F16 if sdt == "f16" else K.DT_F32), K.DT_F32
                rd += args.batch * Cc * hh * ww * (2 if sdt == "f16" else 4)
                wr += args.batch * Cc * hh * ww * 4
            st = torch.cuda.Stream(dev)
            base = loader._codes_bytes
            ptr = slot["dev"].data_ptr()
            with torch.cuda.stream(st):
                for _ in range(3):
                    K.check(loader.lib.ffsr_cache_unpack(ptr + base, rbytes, args.batch, arr, len(segs), ptr, 148, C.c_void_p(st.cuda_stream)))
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                for _ in range(20):
                    K.check(loader.lib.ffsr_cache_unpack(ptr + base, rbytes, args.batch, arr, len(segs), ptr, 148, C.c_void_p(st.cuda_stream)))
                e1.record(st)
            torch.cuda.synchronize()
            k_ms = e0.elapsed_time(e1) / 20
            print(json.dumps({"kernel": "k_cache_unpack", "mode": mode, "batch": args.batch, "ms": k_ms, "bytes_read": rd, "bytes_written": wr,
                              "GB_per_s": (rd + wr) / k_ms / 1e6, "note": "augmentation codes of the last batch (mixed flips / quarter turns); "
                              "working set > L2 (126 MB)"}), flush=True)
            print(json.dumps({"arm": f"pack_cache({mode}) + DeviceBatchLoader", "samples_per_s": n / dt,
                              "GB_per_s_fp32_equivalent": n * 13.9e-3 / dt, "shard_MB_per_sample": os.path.getsize(shard) / args.samples / 1e6,
                              "pack_seconds": pack_s, "kernel_launches_per_batch": 1, "batch": args.batch,
                              "staging_MB": slot["host"].numel() / 1e6}), flush=True)


if __name__ == "__main__":
    main()
