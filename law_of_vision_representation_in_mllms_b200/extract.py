"""Feature extraction driver: host images -> device towers -> host features, data-parallel over ranks.

Replaces the role of llava/feature/extract.py:193-231 (DDP + DistributedSampler + per-image forward at batch 1)
and C_score/extract_feature.py:54-130 (batch-1 tower forward, `[1,C,h,w]` layout):

  * `shard_range(n, rank, world)`     contiguous DP shard per rank — extraction needs no collective;
                                      the only synchronisation is the caller's final barrier (extract.py:259).
  * `FeatureExtractor.run(...)`       chunked, double-buffered pipeline on three CUDA streams: H2D of chunk i+1 and
                                      D2H of chunk i-1 overlap the tower forward of chunk i (pinned host memory).
  * `to_c_score_layout(f, h, w)`      `[1,P,C] -> [1,C,h,w]` exactly as extract_feature.py:82 (`permute(0,2,1).view`).
"""
import torch

from . import towers as T


def shard_range(n, rank, world):
    """Contiguous slice [lo, hi) of n items for `rank` of `world` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def to_c_score_layout(features, h, w):
    """extract_feature.py:82/86/90: features [1,P,C] -> [1,C,h,w] with P = h*w in row-major patch order."""
    b, p, c = features.shape
    if p != h * w:
        raise RuntimeError(f"shape '[{b}, {c}, {h}, {w}]' is invalid for input of size {features.numel()}")
    return features.permute(0, 2, 1).reshape(b, c, h, w)


class FeatureExtractor:
    """Runs one tower (or a dual-tower ModuleList) over host-resident pixel tensors."""

    def __init__(self, vision_tower, chunk=128, out_dtype=None):
        self.tower = vision_tower
        self.towers = list(vision_tower) if isinstance(vision_tower, torch.nn.ModuleList) else [vision_tower]
        self.device = self.towers[0].device
        self.chunk = chunk
        self.out_dtype = out_dtype
        self._copy_in = torch.cuda.Stream(self.device)
        self._copy_out = torch.cuda.Stream(self.device)
        self._bufs = None

    def _alloc(self, pixel_shapes, pixel_dtype, out_shape, out_dtype):
        key = (tuple(pixel_shapes), pixel_dtype, tuple(out_shape), out_dtype)
        if self._bufs is not None and self._bufs["key"] == key:
            return self._bufs
        dev = self.device
        self._bufs = dict(
            key=key,
            px=[[torch.empty(s, dtype=pixel_dtype, device=dev) for s in pixel_shapes] for _ in range(2)],
            out=[torch.empty(out_shape, dtype=out_dtype, device=dev) for _ in range(2)],
            in_ready=[torch.cuda.Event() for _ in range(2)], fwd_done=[torch.cuda.Event() for _ in range(2)],
            out_free=[torch.cuda.Event() for _ in range(2)], in_free=[torch.cuda.Event() for _ in range(2)])
        return self._bufs

    @torch.no_grad()
    def run(self, host_pixels, host_out=None):
        """host_pixels: pinned CPU tensor [N,3,R,R] (single tower) or a list of them (one per tower).
        Returns a pinned CPU tensor [N,P,sum(hidden)] in the pixel dtype (the reference returns features in the
        caller's dtype: clip_encoder.py:45,49)."""
        pix = host_pixels if isinstance(host_pixels, (list, tuple)) else [host_pixels]
        if len(pix) != len(self.towers):
            raise ValueError(f"{len(self.towers)} tower(s) need {len(self.towers)} pixel tensor(s)")
        n = pix[0].shape[0]
        toks = {t.out_tokens(int(p.shape[-1])) for t, p in zip(self.towers, pix)}
        if len(toks) != 1:
            raise RuntimeError(f"towers disagree on token count: {sorted(toks)}")
        P, Dtot = toks.pop(), sum(t.hidden_size for t in self.towers)
        out_dtype = self.out_dtype or pix[0].dtype
        if host_out is None:
            host_out = torch.empty(n, P, Dtot, dtype=out_dtype, pin_memory=True)
        c = min(self.chunk, n)
        b = self._alloc([(c,) + tuple(p.shape[1:]) for p in pix], pix[0].dtype, (c, P, Dtot), out_dtype)
        main = torch.cuda.current_stream(self.device)
        chunks = [(s, min(n, s + c)) for s in range(0, n, c)]

        def stage_in(i):
            s, e = chunks[i]
            k = i & 1
            with torch.cuda.stream(self._copy_in):
                self._copy_in.wait_event(b["in_free"][k])
                for d, p in zip(b["px"][k], pix):
                    d[: e - s].copy_(p[s:e], non_blocking=True)
                b["in_ready"][k].record(self._copy_in)

        for k in range(2):
            b["in_free"][k].record(main)
            b["out_free"][k].record(main)
        stage_in(0)
        for i, (s, e) in enumerate(chunks):
            k = i & 1
            if i + 1 < len(chunks):
                stage_in(i + 1)
            main.wait_event(b["in_ready"][k])
            main.wait_event(b["out_free"][k])
            off = 0
            for t, d in zip(self.towers, b["px"][k]):
                t.forward_into(d[: e - s], b["out"][k][: e - s], off)
                off += t.hidden_size
            b["in_free"][k].record(main)
            b["fwd_done"][k].record(main)
            with torch.cuda.stream(self._copy_out):
                self._copy_out.wait_event(b["fwd_done"][k])
                host_out[s:e].copy_(b["out"][k][: e - s], non_blocking=True)
                b["out_free"][k].record(self._copy_out)
        main.wait_stream(self._copy_out)
        return host_out


    @torch.no_grad()
    def run_images(self, images, host_out=None, pixel_dtype=torch.bfloat16, pad_square=False):
        """Raw images (PIL images or HWC uint8 arrays, any sizes) -> pinned CPU features [N,P,sum(hidden)].

        The reference preprocesses on the CPU (DataLoader workers + HF processors, llava/feature/extract.py:176-190;
        `process_images`, llava/mm_utils.py:78-95) and copies float pixels to the GPU.  Here the 8-bit images are packed and
        copied as they are (half the bytes of bf16 pixels) and every tower's own processor configuration runs on the
        device (`lovr_preprocess`, bit-identical pixel values); packing chunk i+1 on the host overlaps the device work
        of chunk i because nothing in the loop synchronises."""
        from . import preprocess as PP
        n = len(images)
        if n == 0:
            raise ValueError("no images")
        procs = [t.image_processor for t in self.towers]
        sizes = [p.crop_size["height"] for p in procs]
        toks = {t.out_tokens(r) for t, r in zip(self.towers, sizes)}
        if len(toks) != 1:
            raise RuntimeError(f"towers disagree on token count: {sorted(toks)}")
        P, Dtot = toks.pop(), sum(t.hidden_size for t in self.towers)
        out_dtype = self.out_dtype or pixel_dtype
        if host_out is None:
            host_out = torch.empty(n, P, Dtot, dtype=out_dtype, pin_memory=True)
        c = min(self.chunk, n)
        b = self._alloc([(c, 3, r, r) for r in sizes], pixel_dtype, (c, P, Dtot), out_dtype)
        main = torch.cuda.current_stream(self.device)
        for k in range(2):
            b["out_free"][k].record(main)
        for i, s in enumerate(range(0, n, c)):
            e, k = min(n, s + c), i & 1
            host_raw, offsets, hw = PP.pack_images(images[s:e])
            with torch.cuda.stream(self._copy_in):
                dev_raw = host_raw.to(self.device, non_blocking=True)
                b["in_ready"][k].record(self._copy_in)
            dev_raw.record_stream(main)
            main.wait_event(b["in_ready"][k])
            main.wait_event(b["out_free"][k])
            off = 0
            for t, p, d in zip(self.towers, procs, b["px"][k]):
                PP.run_packed(dev_raw, offsets, hw, p._cfg(pad_square), out=d[: e - s])
                t.forward_into(d[: e - s], b["out"][k][: e - s], off)
                off += t.hidden_size
            b["fwd_done"][k].record(main)
            with torch.cuda.stream(self._copy_out):
                self._copy_out.wait_event(b["fwd_done"][k])
                host_out[s:e].copy_(b["out"][k][: e - s], non_blocking=True)
                b["out_free"][k].record(self._copy_out)
        main.wait_stream(self._copy_out)
        return host_out


def build_extractor(selector, select_layer=-2, select_feature="patch", chunk=128, **tower_args):
    """`--vision_tower` string -> towers (llava_arch.py grammar) -> FeatureExtractor."""
    import types
    cfg = types.SimpleNamespace(mm_vision_tower=selector, mm_vision_select_layer=select_layer,
                                mm_vision_select_feature=select_feature, **tower_args)
    return FeatureExtractor(T.build_towers_from_selector(cfg), chunk=chunk)
