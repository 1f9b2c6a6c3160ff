"""Feature extraction driver: host images -> device towers -> host features, data-parallel over ranks.

Replaces the role of llava/feature/extract.py:193-231 (DDP + DistributedSampler + per-image forward at batch 1)
and C_score/extract_feature.py:54-130 (batch-1 tower forward, `[1,C,h,w]` layout):

  * `shard_range(n, rank, world)`     contiguous DP shard per rank — extraction needs no collective;
                                      the only synchronisation is the caller's final barrier (extract.py:259).
  * `FeatureExtractor.run(...)`       chunked, double-buffered pipeline on three CUDA streams: H2D of chunk i+1 and
                                      D2H of chunk i-1 overlap the tower forward of chunk i (pinned host memory).
  * `to_c_score_layout(f, h, w)`      `[1,P,C] -> [1,C,h,w]` exactly as extract_feature.py:82 (`permute(0,2,1).view`).
  * `extract_folder(...)`             the dataset-level driver: walk a directory, shard it over ranks, skip images whose
                                      `.pt` exists, write per-image files in the reference's two layouts, final barrier.
"""
import os

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


def chunk_schedule(n, c):
    """[lo, hi) slices of n images for the copy/compute pipeline: full chunks of c (the size the GEMMs run best at) with HALF
    chunks at both ends — the first host-to-device copy and the last device-to-host copy have nothing to hide behind, so
    they are made half as long (512 images at c = 256: 128 + 256 + 128; +1.8 % end to end)."""
    edge = max(c // 2, 1)
    if n <= edge:
        sizes = [n]
    elif n <= c:
        sizes = [n // 2, n - n // 2]
    else:
        mid = n - 2 * edge
        sizes = [edge] + [c] * (mid // c) + ([mid % c] if mid % c else []) + [edge]
    out, lo = [], 0
    for sz in sizes:
        if sz > 0:
            out.append((lo, lo + sz))
            lo += sz
    return out


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
        self._raw_host = [None, None]      # pinned staging for packed raw images (run_images), reused across calls
        self._raw_done = [None, None]      # event after the H2D copy that last read the staging buffer

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
        chunks = chunk_schedule(n, c)

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
        self._copy_out.synchronize()      # the D2H copies are asynchronous: the host may only read host_out after this
        return host_out


    @torch.no_grad()
    def run_images(self, images, host_out=None, pixel_dtype=torch.bfloat16, pad_square=False, pre_cfgs=None):
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
        # pre_cfgs: one lovr_preprocess configuration per tower instead of the towers' own HF processors (the C-score
        # walker passes the `resize + (x/255 - .5)*2` map of C_score/extract_feature.py:65-67)
        cfgs = list(pre_cfgs) if pre_cfgs is not None else [t.image_processor._cfg(pad_square) for t in self.towers]
        sizes = [int(c.crop_size) for c in cfgs]
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
            if self._raw_done[k] is not None:
                self._raw_done[k].synchronize()          # the copy that last read this staging buffer has finished
            def grow(nbytes, k=k):
                self._raw_host[k] = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, pin_memory=True)
                return self._raw_host[k]
            host_raw, offsets, hw = PP.pack_images(images[s:e], self._raw_host[k], grow)
            with torch.cuda.stream(self._copy_in):
                dev_raw = host_raw.to(self.device, non_blocking=True)
                b["in_ready"][k].record(self._copy_in)
                self._raw_done[k] = torch.cuda.Event()
                self._raw_done[k].record(self._copy_in)
            dev_raw.record_stream(main)
            main.wait_event(b["in_ready"][k])
            main.wait_event(b["out_free"][k])
            off = 0
            for t, cfg, d in zip(self.towers, cfgs, b["px"][k]):
                PP.run_packed(dev_raw, offsets, hw, cfg, out=d[: e - s])
                t.forward_into(d[: e - s], b["out"][k][: e - s], off)
                off += t.hidden_size
            b["fwd_done"][k].record(main)
            with torch.cuda.stream(self._copy_out):
                self._copy_out.wait_event(b["fwd_done"][k])
                host_out[s:e].copy_(b["out"][k][: e - s], non_blocking=True)
                b["out_free"][k].record(self._copy_out)
        main.wait_stream(self._copy_out)
        self._copy_out.synchronize()      # the D2H copies are asynchronous: the host may only read host_out after this
        return host_out


IMAGE_EXTENSIONS = (".jpg", ".jpeg", ".png")      # C_score/extract_feature.py:113


def list_images(input_dir):
    """The walk of C_score/extract_feature.py:110-118 made deterministic (directories and files in sorted order, so every
    rank sees the same list): -> [(class_name = basename of the containing directory, image stem, path)]."""
    items = []
    for root, dirs, files in os.walk(input_dir):
        dirs.sort()
        for f in sorted(files):
            if f.endswith(IMAGE_EXTENSIONS):
                items.append((os.path.basename(root), os.path.splitext(f)[0], os.path.join(root, f)))
    return items


def _load_rgb(path):
    from PIL import Image
    with Image.open(path) as im:
        return im.convert("RGB")


def extract_folder(extractor, input_dir, output_dir, layout="c_score", model_tag="dino336", img_size=None, rank=None,
                   world=None, batch=256, loader=_load_rgb, barrier=True):
    """Dataset-level data-parallel extraction with the reference's observable contract — files.

    Replaces llava/feature/extract.py:193-231,259 (`inference()`: DistributedSampler shard per rank, per-image `.pt`,
    skip-if-exists at :225, final `dist.barrier()` at :259) and C_score/extract_feature.py:110-130 (`process_images`: walk,
    `<output>/<class>/<image>_<tag>.pt`), with the tower forward batched instead of batch-1.

      layout "c_score": pixels = `img.resize((S,S))`, `(x/255 - .5)*2` (extract_feature.py:65-67, on the device, fp32 like
                        the reference), file = fp32 `[1, C, h, w]` = `f.permute(0,2,1).view(1,C,h,w)` (:82-90) at
                        `<output_dir>/<class>/<stem>_<model_tag>.pt` (what pck_train.py:33-37 reads)
      layout "tokens":  pixels = the tower's own HF processor in bf16 (extract.py:228), file = bf16 `[P, D]`
                        (`out_list[0].squeeze().cpu()`, :230-231) at `<output_dir>/<parent dir>/<stem>.pt` (:222-231)

    Sharding: the sorted image list is cut into contiguous shards (`shard_range`), one per rank; no data-path collective.
    Images whose output file exists are dropped BEFORE anything is loaded or launched, so a finished job re-runs with zero
    kernel launches.  Files are written through a temporary name + rename: a killed job leaves no truncated `.pt` behind
    (which skip-if-exists would otherwise trust forever).  Returns the counts."""
    import torch.distributed as dist
    from . import _lib, preprocess as PP
    if layout not in ("c_score", "tokens"):
        raise ValueError(f"layout must be 'c_score' or 'tokens', got {layout!r}")
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    items = list_images(input_dir)
    lo, hi = shard_range(len(items), rank, world)

    def out_path(cls, stem, path):
        if layout == "c_score":
            return os.path.join(output_dir, cls, f"{stem}_{model_tag}.pt")
        return os.path.join(output_dir, cls, os.path.basename(path).split(".")[0] + ".pt")      # extract.py:224
    todo = [(c, s, p) for c, s, p in items[lo:hi] if not os.path.exists(out_path(c, s, p))]
    stats = dict(total=len(items), shard=hi - lo, skipped=(hi - lo) - len(todo), written=0, launches=0)
    pre_cfgs, pixel_dtype = None, torch.bfloat16
    if layout == "c_score":
        if len(extractor.towers) != 1:
            raise ValueError("the C-score feature files hold one tower each (pck_train_two.py loads two files)")
        tower = extractor.towers[0]
        size = int(img_size or tower._default_image_size())
        cfg = _lib.PreprocessConfig()
        cfg.resize_mode, cfg.resize_size, cfg.crop_size = _lib.RESIZE_EXACT, size, size
        cfg.pad_square, cfg.pixel_map = 0, _lib.PIX_CSCORE
        for i in range(3):
            cfg.mean[i], cfg.std[i] = 0.0, 1.0
        pre_cfgs, pixel_dtype = [cfg], torch.float32
    launches0 = _lib.launch_count() if todo else 0
    host_buf = None      # one pinned output buffer for the whole job (every file below is a clone of its rows)
    for s0 in range(0, len(todo), batch):
        part = todo[s0:s0 + batch]
        kw = {} if host_buf is None else dict(host_out=host_buf[: len(part)])
        feats = extractor.run_images([loader(p) for _, _, p in part], pixel_dtype=pixel_dtype, pre_cfgs=pre_cfgs, **kw)
        if host_buf is None and len(part) == batch:
            host_buf = feats
        g = int(round(feats.shape[1] ** 0.5))
        for (cls, stem, src), f in zip(part, feats):
            path = out_path(cls, stem, src)
            os.makedirs(os.path.dirname(path), exist_ok=True)
            t = to_c_score_layout(f[None], g, g).clone() if layout == "c_score" else f.clone()
            tmp = f"{path}.tmp{os.getpid()}"
            torch.save(t, tmp)
            os.replace(tmp, path)
            stats["written"] += 1
    if todo:
        stats["launches"] = _lib.launch_count() - launches0
    if barrier and dist.is_initialized() and dist.get_world_size() > 1:
        dist.barrier()      # extract.py:259: nobody leaves before every rank's files are on disk
    return stats


def build_extractor(selector, select_layer=-2, select_feature="patch", chunk=128, **tower_args):
    """`--vision_tower` string -> towers (llava_arch.py grammar) -> FeatureExtractor."""
    import types
    cfg = types.SimpleNamespace(mm_vision_tower=selector, mm_vision_select_layer=select_layer,
                                mm_vision_select_feature=select_feature, **tower_args)
    return FeatureExtractor(T.build_towers_from_selector(cfg), chunk=chunk)
