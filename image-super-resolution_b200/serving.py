"""Host <-> device pipelining around ``forward_with_precomputed`` for inference streams.

At B200 speed one 2040x1356 fusion forward takes ~20 ms while its 553 MB of fp32 cached inputs
need ~10 ms of PCIe time, so a serving loop has to overlap the copies with the previous image's
compute ("per-GPU input staging must be overlapped with compute").  This is the
caller-side loop of models/team29_FreqFusionSR/io.py:330-345 /
scripts/generate_fast_submission.py:190-256, restated with three CUDA streams:

    copy-in stream : pinned host -> device input set k%depth   (waits until set k%depth is free)
    compute stream : fusion forward on set k%depth              (waits for its copy)
    copy-out stream: SR image -> pinned host buffer             (waits for the forward)

No collective, no extra kernels: only cudaMemcpyAsync + events around the same model call.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch


class PackedImage:
    """All cached inputs of one image in ONE pinned host buffer, so that staging it is a single cudaMemcpyAsync
    (nine separate copies per image serialise on the copy engine's launch latency and, with several GPUs behind one
    host, on the host's memory system).  ``dtype=torch.float16`` is the format the reference's val / TTA caches are
    stored in (scripts/extract_val_cache.py:167-209, scripts/extract_test_tta_cache.py:296-326); the LR image stays
    fp32.  Segments are 256-byte aligned; ``views(device_buffer)`` returns typed tensors over a device copy."""

    def __init__(self, lr: torch.Tensor, imgs: Dict[str, torch.Tensor], feats: Optional[Dict[str, torch.Tensor]] = None,
                 dtype: torch.dtype = torch.float16, pin: bool = True):
        self.layout = []          # (kind, name, shape, dtype, offset, nbytes)
        off = 0

        def add(kind, name, t, dt):
            nonlocal off
            nb = t.numel() * torch.empty((), dtype=dt).element_size()
            self.layout.append((kind, name, tuple(t.shape), dt, off, nb))
            off = (off + nb + 255) // 256 * 256

        add("lr", "lr", lr, torch.float32)
        for k, v in imgs.items():
            add("img", k, v, dtype)
        for k, v in (feats or {}).items():
            add("feat", k, v, dtype)
        self.nbytes = off
        self.buf = torch.empty(off, dtype=torch.uint8)
        if pin:
            self.buf = self.buf.pin_memory()
        src = {"lr": {"lr": lr}, "img": imgs, "feat": feats or {}}
        for kind, name, shape, dt, o, nb in self.layout:
            self.buf[o:o + nb].view(dt).view(shape).copy_(src[kind][name])
        self.payload_bytes = sum(e[5] for e in self.layout)

    def key(self):
        return tuple((e[0], e[1], e[2], e[3]) for e in self.layout)

    def views(self, buf: torch.Tensor):
        lr, imgs, feats = None, {}, {}
        for kind, name, shape, dt, o, nb in self.layout:
            t = buf[o:o + nb].view(dt).view(shape)
            if kind == "lr":
                lr = t
            elif kind == "img":
                imgs[name] = t
            else:
                feats[name] = t
        return lr, imgs, (feats or None)


class PipelinedFusion:
    """``cuda_graph=True`` (default): the forward of each input set is captured once into a CUDA graph (the device
    buffers of a set are static, so are the engine's workspaces) and replayed; a change of weights, precision or
    shapes re-captures.  The ~100 launches of a forward then cost one graph launch of host time and no inter-kernel
    launch gaps, which matters once a 2040x1356 forward is down to ~16 ms."""

    def __init__(self, model, depth: int = 2, device: Optional[torch.device] = None, cuda_graph: bool = True):
        self.model = model
        self.depth = depth
        self.cuda_graph = cuda_graph
        self._graphs = [None] * depth         # (key, CUDAGraph, static output)
        self._out_free = [None] * depth       # event: the copy-out of this set's static output finished
        self.dev = device or next(model.parameters()).device
        self.s_in = torch.cuda.Stream(self.dev)
        self.s_out = torch.cuda.Stream(self.dev)
        self.s_compute = torch.cuda.current_stream(self.dev)
        self._sets = [None] * depth           # device input buffers
        self._free = [None] * depth           # event: compute on this set finished
        self._outs = [None] * depth           # device output kept alive until copied out
        self._k = 0

    def _alloc_like(self, host_lr, host_imgs, host_feats):
        d = self.dev
        return (torch.empty_like(host_lr, device=d), {k: torch.empty_like(v, device=d) for k, v in host_imgs.items()},
                {k: torch.empty_like(v, device=d) for k, v in (host_feats or {}).items()})

    def submit(self, lr: torch.Tensor, imgs: Dict[str, torch.Tensor], feats: Optional[Dict[str, torch.Tensor]],
               out_host: torch.Tensor) -> None:
        """Enqueue one image: pinned host inputs -> SR written into the pinned ``out_host``."""
        i = self._k % self.depth
        self._k += 1
        if self._sets[i] is None or len(self._sets[i]) != 3 or self._sets[i][0].shape != lr.shape:
            self._sets[i] = self._alloc_like(lr, imgs, feats)
            self._graphs[i] = None
        d_lr, d_imgs, d_feats = self._sets[i]
        with torch.cuda.stream(self.s_in):
            if self._free[i] is not None:
                self.s_in.wait_event(self._free[i])          # the forward that last read this set is done
            d_lr.copy_(lr, non_blocking=True)
            for k, v in imgs.items():
                d_imgs[k].copy_(v, non_blocking=True)
            for k, v in (feats or {}).items():
                d_feats[k].copy_(v, non_blocking=True)
            copied = torch.cuda.Event()
            copied.record(self.s_in)
        self.s_compute.wait_event(copied)
        sr, static = self._forward(i, d_lr, d_imgs, d_feats if feats else None)
        done = torch.cuda.Event()
        done.record(self.s_compute)
        self._free[i] = done
        self._outs[i] = sr
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(done)
            out_host.copy_(sr, non_blocking=True)
            if static:
                ev = torch.cuda.Event()
                ev.record(self.s_out)
                self._out_free[i] = ev
            else:
                sr.record_stream(self.s_out)

    def submit_packed(self, packed: "PackedImage", out_host: torch.Tensor) -> None:
        """Like ``submit`` for a ``PackedImage``: ONE host->device copy of the whole input set."""
        i = self._k % self.depth
        self._k += 1
        cur = self._sets[i]
        if cur is None or not isinstance(cur, tuple) or len(cur) != 2 or cur[0] != packed.key():
            dbuf = torch.empty(packed.nbytes, dtype=torch.uint8, device=self.dev)
            self._sets[i] = (packed.key(), (dbuf,) + tuple(packed.views(dbuf)))
            self._graphs[i] = None
        dbuf, d_lr, d_imgs, d_feats = self._sets[i][1]
        with torch.cuda.stream(self.s_in):
            if self._free[i] is not None:
                self.s_in.wait_event(self._free[i])
            dbuf.copy_(packed.buf, non_blocking=True)
            copied = torch.cuda.Event()
            copied.record(self.s_in)
        self.s_compute.wait_event(copied)
        sr, static = self._forward(i, d_lr, d_imgs, d_feats)
        done = torch.cuda.Event()
        done.record(self.s_compute)
        self._free[i] = done
        self._outs[i] = sr
        with torch.cuda.stream(self.s_out):
            self.s_out.wait_event(done)
            out_host.copy_(sr, non_blocking=True)
            if static:
                ev = torch.cuda.Event()
                ev.record(self.s_out)
                self._out_free[i] = ev
            else:
                sr.record_stream(self.s_out)

    def _graph_key(self, d_lr, d_imgs, d_feats):
        m = self.model
        return (tuple(d_lr.shape), d_lr.dtype, tuple(sorted((k, v.dtype) for k, v in d_imgs.items())),
                tuple(sorted((k, tuple(v.shape), v.dtype) for k, v in (d_feats or {}).items())), m.precision, m.training,
                tuple(p._version for p in m.parameters()), tuple(b._version for b in m.buffers()))

    def _forward(self, i, d_lr, d_imgs, d_feats):
        """-> (SR tensor, is_static_graph_output)"""
        m = self.model
        if not self.cuda_graph or m.training:
            return m.forward_with_precomputed(d_lr, d_imgs, d_feats), False
        key = self._graph_key(d_lr, d_imgs, d_feats)
        entry = self._graphs[i]
        if entry is not None and entry[0] == key:
            if self._out_free[i] is not None:
                self.s_compute.wait_event(self._out_free[i])     # the previous copy-out of this static output is done
            entry[1].replay()
            return entry[2], True
        # first use of this set (or weights / shapes changed): one eager forward serves this submission and warms the
        # engine up (workspaces, packed weights); then the same call is captured for the following submissions
        sr = m.forward_with_precomputed(d_lr, d_imgs, d_feats)
        eng = m._engine
        prev = eng.overlap_routing
        eng.overlap_routing = False                              # no stream switching inside a capture
        try:
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = m.forward_with_precomputed(d_lr, d_imgs, d_feats)
            self._graphs[i] = (key, g, out)
        finally:
            eng.overlap_routing = prev
        return sr, False

    def finish(self) -> None:
        self.s_in.synchronize()
        self.s_compute.synchronize()
        self.s_out.synchronize()


# ---------------------------------------------------------------------------------------------------
# 8x test-time augmentation over cached variants
# ---------------------------------------------------------------------------------------------------
class ConcurrentFusion:
    """``ways`` fusion forwards in flight on ``ways`` CUDA streams of ONE device: each way has its own ``FusionEngine`` (own
    workspaces, the model's weights shared), images go to the ways round-robin.  The tile-resident kernels of a forward run at
    one CTA per SM and a third of the issue slots, and every launch ends with a tail of idle SMs: a second, independent image
    on another stream fills part of that (measured: 10.5 -> 10.1 ms per C3 image with two ways, tools/experiments/two_in_flight.py).
    Same results as ``model.forward_with_precomputed`` image by image (same kernels, same order within an image)."""

    def __init__(self, model, ways: int = 2, device: Optional[torch.device] = None):
        from .pipeline import FusionEngine
        if ways < 1:
            raise ValueError("ways must be >= 1")
        if model.training:
            raise RuntimeError("ConcurrentFusion is an inference helper: call model.eval() first")
        self.m = model
        self.dev = device if device is not None else next(model.parameters()).device
        if model._engine is None:
            model._engine = FusionEngine(model)
        self.engines = [model._engine] + [FusionEngine(model) for _ in range(ways - 1)]
        self.streams = [torch.cuda.Stream(self.dev) for _ in range(ways)]
        self.launches = 0

    @torch.no_grad()
    def run(self, items, sink=None):
        """items: iterable of (lr, expert_imgs dict, expert_feats dict or None).  ``sink(i, sr)`` is called with the SR tensor of
        item i while its stream is current (copy it out / reduce it there); without a sink the outputs are dropped.  Returns
        after the CALLER's stream has been made to wait for every way (no host synchronisation)."""
        from .pipeline import EXPERT_ORDER
        cur = torch.cuda.current_stream(self.dev)
        ev = torch.cuda.Event()
        ev.record(cur)
        for st in self.streams:
            st.wait_event(ev)
        self.launches = 0
        up = self.m.upscale
        for i, (lr, imgs, feats) in enumerate(items):
            k = i % len(self.engines)
            img_list = [imgs[n] for n in EXPERT_ORDER if n in imgs]
            fts = {n: feats[n] for n in EXPERT_ORDER if n in feats} if feats is not None else {}
            with torch.cuda.stream(self.streams[k]):
                sr, _ = self.engines[k].forward(lr, img_list, fts, lr.shape[2] * up, lr.shape[3] * up, False)
                self.launches += self.engines[k].launches
                if sink is not None:
                    sink(i, sr)
        for st in self.streams:
            cur.wait_stream(st)


def reverse_tta(t: torch.Tensor, hflip: bool, rot: int) -> torch.Tensor:
    """Undo the geometric TTA transform on an SR output (scripts/generate_fast_submission.py:55-61)."""
    if rot > 0:
        t = torch.rot90(t, -rot, [2, 3])
    if hflip:
        t = torch.flip(t, [3])
    return t


@torch.no_grad()
def fuse_tta(model, variants) -> torch.Tensor:
    """Mean of the reverse-transformed fusion outputs of the cached TTA variants of ONE image, clamped to [0,1]
    (scripts/generate_fast_submission.py:190-250), computed entirely on the device.

    ``variants``: iterable of ``(lr [1,3,h,w], expert_imgs {name: [1,3,4h,4w]}, expert_feats {name: [1,C,h,w]} | None,
    hflip, rot)`` as ``scripts/extract_test_tta_cache.py:253-256`` enumerates them (hflip in {F,T} x rot in 0..3).
    Variants of equal LR shape (rot 0/2 vs rot 1/3) are batched into one forward each, so the 8 variants cost two
    B=4 forwards instead of eight B=1 forwards with a device->host copy and an ``empty_cache()`` in between, which is
    what the reference does."""
    groups = {}
    for lr, imgs, feats, hflip, rot in variants:
        groups.setdefault(tuple(lr.shape[2:]), []).append((lr, imgs, feats, bool(hflip), int(rot)))
    total, count = None, 0
    for items in groups.values():
        lr = torch.cat([it[0] for it in items], 0)
        imgs = {k: torch.cat([it[1][k] for it in items], 0) for k in items[0][1]}
        feats = None
        if items[0][2]:
            feats = {k: torch.cat([it[2][k] for it in items], 0) for k in items[0][2]}
        sr = model.forward_with_precomputed(lr, imgs, feats)
        for j, it in enumerate(items):
            r = reverse_tta(sr[j:j + 1], it[3], it[4]).float()
            total = r.clone() if total is None else total.add_(r)
            count += 1
    if total is None:
        raise ValueError("fuse_tta: no variants")
    return (total / count).clamp_(0, 1)


# ---------------------------------------------------------------------------------------------------
# One image across GPUs: halo tiles, bands of the whole image, no halo exchange
# ---------------------------------------------------------------------------------------------------
TILE_HALO_LR = 40       # LR px.  Receptive field of phases 3-7 behind the bands: LKA 12 + selector 3 + bilinear 1 LR px,
                        # then refine 6 + Laplacian pyramid / edge nets <= 46 HR px  =>  <= 110 HR px < 4 * 40
                        # (measured on the reference: a feature/image perturbation reaches 94 HR px).


def tile_grid(H: int, W: int, ty: int, tx: int):
    """Split an H x W LR image into ty x tx core rectangles (y0, y1, x0, x1); inner boundaries are multiples of 8 LR px
    (the DCT block grid; also keeps every HR / half / quarter resolution boundary aligned)."""
    def cuts(n, k):
        c = [0] + [min(n, max(8, int(round(n * i / k / 8.0)) * 8)) for i in range(1, k)] + [n]
        if any(b <= a for a, b in zip(c, c[1:])):
            raise ValueError(f"cannot cut {n} LR px into {k} tiles on the 8-px grid")
        return c
    ys, xs = cuts(H, ty), cuts(W, tx)
    return [(ys[i], ys[i + 1], xs[j], xs[j + 1]) for i in range(ty) for j in range(tx)]


@torch.no_grad()
def fuse_tiled(model, lr: torch.Tensor, expert_imgs: Dict[str, torch.Tensor],
               expert_feats: Optional[Dict[str, torch.Tensor]] = None, grid=(1, 2), halo: int = TILE_HALO_LR,
               rank: int = 0, world: int = 1, assemble: bool = True, group=None) -> torch.Tensor:
    """``forward_with_precomputed`` of ONE (batch of) image(s) computed tile by tile; rank r takes tiles r, r+world, ...

    Everything behind phase 2 has a finite receptive field, so a tile is the model applied to a window = core + halo
    (clipped at the image border, where the zero / clamp boundary rules then apply exactly as in the whole image).  The
    nine frequency bands are NOT local (global FFT mask; DWT resize ratio depends on H, W): every rank computes them on
    the whole 3-channel LR image (31 FLOP per HR pixel) and crops.  Halos are re-read from the inputs, never exchanged.
    With ``assemble`` (and world > 1) the cores are summed into the full image with one all-reduce (x + 0 is exact);
    without it the full-size tensor holds this rank's cores and zeros elsewhere.  ``rank`` / ``world`` are positions inside
    ``group`` (a torch.distributed process group of the ranks sharing this image; default: all ranks)."""
    eng_model = model
    if eng_model.training:
        raise RuntimeError("fuse_tiled is an inference path: call model.eval() first")
    from .fusion import EXPERT_ORDER
    from .pipeline import FusionEngine
    if eng_model._engine is None:
        eng_model._engine = FusionEngine(eng_model)
    eng = eng_model._engine
    B, _, H, W = lr.shape
    tiles = tile_grid(H, W, int(grid[0]), int(grid[1]))
    imgs = [expert_imgs[k] for k in EXPERT_ORDER if k in expert_imgs]
    feats = {k: expert_feats[k] for k in EXPERT_ORDER if k in expert_feats} if expert_feats is not None else {}
    bands = eng.frequency_bands(lr)
    out = torch.zeros(B, 3, 4 * H, 4 * W, device=lr.device, dtype=torch.float32)
    for t in range(rank, len(tiles), world):
        y0, y1, x0, x1 = tiles[t]
        wy0, wy1, wx0, wx1 = max(0, y0 - halo), min(H, y1 + halo), max(0, x0 - halo), min(W, x1 + halo)
        sr, _ = eng.forward(lr[:, :, wy0:wy1, wx0:wx1],
                            [im[:, :, 4 * wy0:4 * wy1, 4 * wx0:4 * wx1] for im in imgs],
                            {k: f[:, :, wy0:wy1, wx0:wx1] for k, f in feats.items()},
                            4 * (wy1 - wy0), 4 * (wx1 - wx0), False,
                            bands=bands[:, :, :, wy0:wy1, wx0:wx1].contiguous())
        oy, ox = 4 * (y0 - wy0), 4 * (x0 - wx0)
        out[:, :, 4 * y0:4 * y1, 4 * x0:4 * x1] = sr[:, :, oy:oy + 4 * (y1 - y0), ox:ox + 4 * (x1 - x0)]
    if assemble and world > 1:
        import torch.distributed as dist
        dist.all_reduce(out, group=group)
    return out
