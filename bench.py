#!/usr/bin/env python
"""Headline benchmark (BASELINE.json): images/sec of CLIP ViT-L/14-336 patch-feature extraction on B200, with the score
kernels (A score, linear CKA, C score) timed beside it.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU code path on the host cores

A step = one tower forward over one batch of `--batch` synthetic 336px images (configs[1]: "CLIP ViT-L/14-336, 10k
synthetic 336px images, 1xB200 feature extract"; 10k images = 10000/batch steps).  Weights: seeded HF-init-style random
weights of the real architecture (no checkpoints offline).  `value` = whole-job images/s with inputs already resident in
HBM; `e2e` = the same through the public extractor (pinned host pixels -> H2D -> towers -> D2H features, host-synchronised
every step).  The headline runs with fp16 tensor-core operands — the mode that meets north_star's 1e-3 against the
reference's fp32 path (tests/test_gpu_tower.py; `parity` in the line re-measures it in this process) — and the same
measurements with bf16 operands are carried in the `bf16` block.  Under torchrun every rank runs its own shard (weak
scaling, no data-path collective for extraction); the score paths do ONE all-reduce each; timing is CUDA
events between barriers, max over ranks.  One JSON line on rank 0.
"""
import argparse
import contextlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

TOWER = "openai/clip-vit-large-patch14-336"
DUAL = "openai/clip-vit-large-patch14-336.facebook/dinov2-large"
R, T_PATCH, T_TOK, D = 336, 576, 577, 1024
# SURVEY.md §8d / BASELINE.md §4: algorithmic work per image, layers 1-23 + patch embed, T=577
FLOPS_PER_IMAGE = 366.03e9
METRIC = "images/sec ViT-L/14-336 extract"
WORKLOAD = "CLIP ViT-L/14-336, 10k synthetic 336px images, 1xB200 feature extract (configs[1])"
C3_PAIRS, C3_BLOCK = 10000, 250          # configs[2]; data generated in blocks seeded by the GLOBAL block index
C4_IMAGES_PER_RANK = 6250                # configs[3]: 50k dual-tower images over 8 GPUs
C5_PAIRS, C5_CHUNK, C5_P, C5_K, C5_CATS = 12234, 256, 60, 12, 18     # configs[4]: SPair-71k test size, P=60


def bench_config(world, batch, chunk):
    """The `config` object both arms print (the reference arm runs a bounded sample of the same workload)."""
    return dict(workload=WORKLOAD, tower=TOWER, images_per_step_per_gpu=batch, native_chunk=chunk,
                weights="seeded HF-init-style random weights (weights.synthetic_state_dict style=hf_init)",
                l2="inputs larger than L2: each step reads a different %d MB pixel batch" % (batch * 3 * R * R * 2 // 2**20),
                parallelism=f"dp{world} (images sharded, no data-path collective)")


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=float(d.get("bf16_tflops_sustained", d.get("bf16_tflops"))), tflops_burst=float(d.get("bf16_tflops", 0)),
                    hbm_gbs=float(d.get("hbm_gbs", 6500.0)), source="measured (MEASURED_PEAKS.json; sustained bf16 for kernels inside the "
                    "step, burst bf16 for kernels timed alone, copy bandwidth for HBM)")
    return dict(tflops=1400.0, tflops_burst=1600.0, hbm_gbs=6500.0, source="fallback (B200_PROFILING.md)")


def ncu_traffic():
    """Per-launch DRAM bytes of the dominant kernel from the committed ncu --set full capture of this round, if there is
    one (profiles/r02_traffic.json, written by scripts/ncu_traffic.py from the .ncu-rep); otherwise null."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if os.path.exists(p):
        return json.load(open(p))
    return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.idx)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        sm = sorted(int(r[1]) for r in self.rows if len(r) >= 8 and r[1].isdigit())
        reasons = set()
        for r in self.rows:
            if len(r) >= 8:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        mx = [int(r[2]) for r in self.rows if len(r) >= 8 and r[2].isdigit()]
        pw = [float(r[3]) for r in self.rows if len(r) >= 8 and r[3].replace(".", "").isdigit()]
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    power_w_max=max(pw) if pw else None, samples=len(sm), reasons=sorted(reasons))


# ===================================================================================================== CPU reference arm
_CPU = {}


def _calibrated_threads(run_probe):
    """The GPU boxes report 128 logical CPUs, but a 128-thread torch run of this model is ~50x slower than a 16-thread one
    (29.7 s vs 0.6 s per image measured): the count is calibrated over {all, 1/2, 1/4, 32, 16, 8} threads."""
    import torch
    cores = os.cpu_count() or 1
    best = (float("inf"), cores)
    for thr in sorted({cores, max(1, cores // 2), max(1, cores // 4), 32, 16, 8} & set(range(1, cores + 1)), reverse=True):
        torch.set_num_threads(thr)
        run_probe()
        t0 = time.perf_counter()
        run_probe()
        dt = time.perf_counter() - t0
        if dt < best[0]:
            best = (dt, thr)
    torch.set_num_threads(best[1])
    return best[1], cores


def _cpu_setup():
    """The CPU arm, built once.  kind "reference": the reference's OWN wrapper class (clip_encoder.py `CLIPVisionTower`, from
    oracle/_ref — see oracle/build_ref.py) over HF `CLIPVisionModel`, fp32, its stock forward path; only `from_pretrained` is
    replaced by a constructor with the seeded weights.  kind "port": oracle/vit_oracle.py when oracle/_ref is absent."""
    if _CPU:
        return _CPU
    import torch
    from law_of_vision_representation_in_mllms_b200.specs import TOWER_SPECS, tiny_spec
    from law_of_vision_representation_in_mllms_b200.weights import canonical_weights, synthetic_images, synthetic_state_dict
    spec = TOWER_SPECS[TOWER]
    sd = synthetic_state_dict(spec, seed=0, style="hf_init")
    kind, fwd = "port", None
    try:
        from oracle import ref_loader
        if ref_loader.encoder_dir() is not None:
            tower = ref_loader.reference_tower(spec, sd, -2, "patch")
            fwd = lambda imgs: tower(imgs)                                   # noqa: E731
            kind = "reference"
    except Exception as ex:      # transformers / the copied files unusable here: fall back to the restatement
        sys.stderr.write(f"bench: reference wrapper unavailable ({str(ex)[:120]}); timing the oracle port\n")
    if fwd is None:
        from oracle import vit_oracle
        cw = canonical_weights(spec, sd)
        fwd = lambda imgs: vit_oracle.tower_forward(cw, spec, imgs)          # noqa: E731
    probe = synthetic_images(1, R, seed=99)
    # thread calibration on a 2-layer model of the same widths (seconds, not minutes)
    spec2 = tiny_spec(spec, 2)
    from oracle import vit_oracle
    cw2 = canonical_weights(spec2, synthetic_state_dict(spec2, seed=0, style="hf_init"))
    with torch.no_grad():
        threads, cores = _calibrated_threads(lambda: vit_oracle.tower_forward(cw2, spec2, probe, select_layer=2))
    _CPU.update(fwd=fwd, kind=kind, threads=threads, cores=cores)
    return _CPU


def cpu_reference_leg(n_images):
    """-> (images/s, seconds, threads, kind) of the CPU arm over a bounded sample of the headline workload."""
    import torch
    from law_of_vision_representation_in_mllms_b200.weights import synthetic_images
    c = _cpu_setup()
    imgs = synthetic_images(n_images, R, seed=1234)
    t0 = time.perf_counter()
    with torch.no_grad():
        out = c["fwd"](imgs)
    dt = time.perf_counter() - t0
    assert tuple(out.shape) == (n_images, T_PATCH, D)
    return n_images / dt, dt, c["threads"], c["kind"]


def cpu_score_leg(threads, n_pairs=12):
    """A_score/compute.py's per-pair arithmetic as the reference executes it (broadcast F.cosine_similarity over [x,1,D] x
    [1,y,D], `a_pair_literal`) on a bounded sample of C3-shaped fp32 pairs; -> pairs/s and the extrapolated C3 seconds."""
    import torch
    from oracle import a_score_oracle
    g = torch.Generator().manual_seed(11)
    o = torch.randn(n_pairs, T_PATCH, D, generator=g)
    r = 0.5 * o + torch.randn(n_pairs, T_PATCH, D, generator=g)
    a_score_oracle.a_pair_literal(o[0], r[0])
    t0 = time.perf_counter()
    for i in range(n_pairs):
        a_score_oracle.a_pair_literal(o[i], r[i])
    dt = time.perf_counter() - t0
    return dict(kind="port", cores=threads, sample=f"{n_pairs} fp32 pairs of 576 x 576 x 1024, literal compute.py arithmetic, {dt:.1f} s",
                a_score_pairs_per_s=n_pairs / dt, a_score_c3_seconds_extrapolated=C3_PAIRS * dt / n_pairs)


def cpu_sample_text(per_step, thr, cores, kind):
    what = ("the reference's own CLIPVisionTower wrapper (oracle/_ref/clip_encoder.py) over HF CLIPVisionModel"
            if kind == "reference" else "fp32 oracle restatement of the HF CLIP forward")
    return (f"{per_step} synthetic 336px images per step, fp32, torch CPU, {what}, {thr} threads (calibrated best of "
            f"{cores} logical CPUs)")


def run_reference(args, rank, world):
    if rank != 0:
        return
    per_step = args.cpu_images
    rates = []
    for _ in range(args.warmup):
        cpu_reference_leg(1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r, _, thr, kind = cpu_reference_leg(per_step)
        rates.append(r)
    wall = time.perf_counter() - t0
    v = sum(rates) / len(rates)
    line = dict(impl="reference", metric=METRIC, value=v, unit="images/s", n_gpus=args.gpus, steps=args.steps,
                warmup=args.warmup, ms_per_step=1e3 * wall / args.steps, higher_is_better=True, scaling="weak",
                vs_baseline=None, dtype="f32", data="synthetic", config=bench_config(args.gpus, args.batch, args.chunk),
                cpu_baseline=dict(value=v, unit="images/s", cores=thr, kind=kind,
                                  sample=cpu_sample_text(per_step, thr, _CPU["cores"], kind)),
                e2e=dict(value=v, unit="images/s", h2d_bytes_per_step=0, d2h_bytes_per_step=0))
    print(json.dumps(line), flush=True)


@contextlib.contextmanager
def _stdout_to_stderr():
    """Point file descriptor 1 at stderr for the duration of the block (C-level prints included); no-op if that fails."""
    saved = None
    try:
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
    except OSError:
        saved = None
    try:
        yield
    finally:
        if saved is not None:
            try:
                sys.stdout.flush()
                os.dup2(saved, 1)
                os.close(saved)
            except OSError:
                pass


# ===================================================================================================== GPU arm
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=512, help="images per step per GPU")
    ap.add_argument("--chunk", type=int, default=256, help="images per native tower call (256: +2 % over 128, measured)")
    ap.add_argument("--cpu-images", type=int, default=32, help="images per CPU-baseline sample / reference-arm step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-a-score", "--no-scores", dest="no_scores", action="store_true",
                    help="headline only: skip the score kernels, the job-size rows and the other operand dtype")
    ap.add_argument("--operand-dtype", default="fp16", choices=["bf16", "fp16"],
                    help="16-bit tensor-core operand format of the headline (fp32 accumulate either way)")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import numpy as np
    import torch
    import torch.distributed as dist
    from law_of_vision_representation_in_mllms_b200 import _lib, a_score as A, c_score as CS, cka as CK, distributed as DD, extract as X

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback); use --impl reference for the CPU leg")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL can print its version banner on stdout from C (seen at 4 ranks): stdout is reserved for the one JSON line, so
        # communicator creation and the first collective run with fd 1 pointed at stderr.
        with _stdout_to_stderr():
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return t.item()
        return ms

    def device_ms(fn, reps=1):
        """fn() `reps` times between CUDA events on the current stream, barrier + synchronize on both sides; max over ranks."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = None
        for _ in range(reps):
            out = fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1) / reps), out

    def median_ms(fn, n=5):
        runs = []
        out = None
        for _ in range(n):
            ms, out = device_ms(fn)
            runs.append(ms)
        return sorted(runs)[n // 2], runs, out

    B = args.batch
    peaks = measured_peaks()
    tdt = {"bf16": torch.bfloat16, "fp16": torch.float16}
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    n_in = 3      # several distinct batches so every step reads pixels that are not L2-resident (B*677 KB >> L2 anyway)
    px32 = [torch.randn(B, 3, R, R, generator=g, device=dev) for _ in range(n_in)]

    def measure_mode(od, with_e2e=True):
        """Headline measurement for one operand dtype: resident images/s with per-kernel-class live timing, and e2e."""
        ex = X.build_extractor(TOWER, chunk=args.chunk, lovr_max_batch=args.chunk, lovr_device=dev,
                               lovr_state_dict=None, lovr_seed=0, lovr_operand_dtype=od)
        tower = ex.towers[0]
        io = tdt[od]
        dev_px = [p.to(io) for p in px32]
        dev_out = torch.empty(B, T_PATCH, D, dtype=io, device=dev)
        host_px = [p.cpu().pin_memory() for p in dev_px[:2]]
        host_out = torch.empty(B, T_PATCH, D, dtype=io).pin_memory()

        def step_resident(i):
            tower.forward_into(dev_px[i % n_in], dev_out, 0)

        def step_e2e(i):
            ex.run(host_px[i % 2], host_out)

        def timed(step_fn, steps, warmup, profile=False):
            for i in range(warmup):
                step_fn(i)
            barrier()
            sampler = ClockSampler(local_rank) if (rank == 0 and profile) else None
            if sampler:
                sampler.start()
            launches0 = _lib.launch_count()
            if profile:
                _lib.prof_read_kinds()
                _lib.prof_enable(True)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                step_fn(warmup + i)
            e1.record()
            barrier()
            ms = e0.elapsed_time(e1)
            prof = None
            if profile:
                _lib.prof_enable(False)
                prof = _lib.prof_read_kinds()
            clocks = sampler.stop() if sampler else None
            launches = _lib.launch_count() - launches0
            return max_over_ranks(ms), launches, prof, clocks

        ms, launches, prof, clocks = timed(step_resident, args.steps, max(args.warmup, 3), profile=True)
        res = dict(ex=ex, tower=tower, dev_px=dev_px, host_out=host_out, ms=ms, launches=launches, prof=prof, clocks=clocks,
                   value=world * B * args.steps / (ms / 1e3), io=io)
        if with_e2e:
            ms_e2e, _, _, _ = timed(step_e2e, args.steps, max(args.warmup, 3))
            res.update(ms_e2e=ms_e2e, e2e_value=world * B * args.steps / (ms_e2e / 1e3))
        return res

    def kernel_rooflines(m):
        """Per-kernel-class entries from the live CUDA-event timing inside the timed region (lovr_prof_read_kinds)."""
        out = []
        g_ms, g_fl, g_n = m["prof"]["gemm"]
        a_ms, a_fl, a_n = m["prof"]["attention"]
        l_ms, l_by, l_n = m["prof"]["layernorm"]
        if g_ms > 0:
            ach = g_fl / (g_ms / 1e3) / 1e12
            out.append(dict(kernel="gemm2_tcgen05_kernel / gemm_tcgen05_kernel (patch-embed, qkv, out-proj, fc1, fc2)", bound="tensor",
                            achieved=ach, peak=peaks["tflops"], unit="TFLOP/s", frac=ach / peaks["tflops"], launches=g_n,
                            share_of_step=g_ms / m["ms"], avg_launch_ms=g_ms / g_n))
        if a_ms > 0:
            ach = a_fl / (a_ms / 1e3) / 1e12
            out.append(dict(kernel="attention_tcgen05 (persistent, softmax fused)", bound="tensor", achieved=ach, peak=peaks["tflops"],
                            unit="TFLOP/s", frac=ach / peaks["tflops"], launches=a_n, share_of_step=a_ms / m["ms"],
                            avg_launch_ms=a_ms / a_n, note="the binding unit is MUFU (exp2), ~900 TFLOP/s at head_dim 64"))
        if l_ms > 0:
            ach = l_by / (l_ms / 1e3) / 1e9
            out.append(dict(kernel="layernorm_bf16_kernel", bound="hbm", achieved=ach, peak=peaks["hbm_gbs"], unit="GB/s",
                            frac=ach / peaks["hbm_gbs"], launches=l_n, share_of_step=l_ms / m["ms"], avg_launch_ms=l_ms / l_n))
        return out

    head = measure_mode(args.operand_dtype)
    other_od = "bf16" if args.operand_dtype == "fp16" else "fp16"
    other = None if args.no_scores else measure_mode(other_od)

    # ---- parity of both operand modes, re-measured here: the full-depth CLIP-336 fixture the reference wrapper produced
    parity = None
    if not args.no_scores:
        try:
            parity = parity_block(dev)
        except Exception as ex:      # a reported extra, never a reason to lose the bench line
            parity = dict(error=str(ex)[:200])

    tower, dev_px, io_dtype = head["tower"], head["dev_px"], head["io"]
    host_out = head["host_out"]
    ex = head["ex"]
    scores = None
    if not args.no_scores:
        scores = {}
        # ================================================================== A score at the C3 shape (10k x 576 x 1024 bf16)
        # Same data at every N: blocks of 250 pairs seeded by the GLOBAL block index; rank r holds a contiguous run of blocks.
        nblk = C3_PAIRS // C3_BLOCK
        blo, bhi = X.shard_range(nblk, rank, world)
        o_parts, r_parts = [], []
        for b in range(blo, bhi):
            gb = torch.Generator(device=dev).manual_seed(11 + b)
            ob = torch.randn(C3_BLOCK, T_PATCH, D, generator=gb, device=dev)
            rb = 0.5 * ob + torch.randn(C3_BLOCK, T_PATCH, D, generator=gb, device=dev)
            o_parts.append(ob.to(torch.bfloat16))
            r_parts.append(rb.to(torch.bfloat16))
            del ob, rb
        o, r_ = torch.cat(o_parts), torch.cat(r_parts)
        del o_parts, r_parts

        def a_job():
            s = A.a_score_pairs(o, r_).double().sum().reshape(1)
            if world > 1:
                dist.all_reduce(s)                   # the single collective of the score path
            return s
        a_job()                                      # same shapes as the timed call (workspace, NCCL buffers)
        a_ms, a_runs, s = median_ms(a_job)
        a_flops = 2.0 * C3_PAIRS * T_PATCH * T_PATCH * D
        a_bytes = 2.0 * C3_PAIRS * T_PATCH * D * 2
        # ridge roofline: the job needs max(bytes / HBM, flops / tensor) seconds
        a_min_s = max(a_bytes / (peaks["hbm_gbs"] * 1e9), a_flops / (peaks["tflops_burst"] * 1e12)) / world
        scores["a_score"] = dict(config="A-score max-cos over 10k x 576 x 1024 bf16 pairs (configs[2]), sharded by pair index, one all-reduce",
                                 seconds=a_ms / 1e3, mean_a=s.item() / C3_PAIRS, runs_ms=[round(t, 3) for t in a_runs],
                                 roofline=dict(bound="hbm/tensor ridge (288 FLOP/B)", achieved_gbs=a_bytes / (a_ms / 1e3) / 1e9 / world,
                                               achieved_tflops=a_flops / (a_ms / 1e3) / 1e12 / world, peak_gbs=peaks["hbm_gbs"],
                                               peak_tflops=peaks["tflops_burst"], min_seconds=a_min_s, frac=a_min_s / (a_ms / 1e3)))
        # ================================================================== linear CKA over the same rows: per-shard sums, ONE all-reduce
        def cka_job():
            gram, nrows, dx, dy = CK.cka_gram(o, r_)
            return DD.cka_reduce(gram, dx, dy)       # one all-reduce of one buffer (Gram + column sums + row count)
        cka_job()                                    # warm-up at the REAL size (the first 25 MB all-reduce sets up its channels)
        c_ms, c_runs, cka_val = median_ms(cka_job)
        c_flops = 2.0 * C3_PAIRS * T_PATCH * (D * 2 * D + D * D)
        scores["linear_cka"] = dict(config="linear CKA over 5.76M rows x (1024 | 1024) bf16 (configs[2] rows), rows sharded, ONE all-reduce "
                                           "of the fp64 partial-sum buffer (~25 MB)", seconds=c_ms / 1e3, value=cka_val,
                                    runs_ms=[round(t, 3) for t in c_runs], rows=C3_PAIRS * T_PATCH,
                                    roofline=dict(bound="tensor", achieved=c_flops / (c_ms / 1e3) / 1e12 / world, peak=peaks["tflops_burst"],
                                                  unit="TFLOP/s", frac=c_flops / (c_ms / 1e3) / 1e12 / world / peaks["tflops_burst"],
                                                  flops="3 x 2 x rows x D^2 = 36.2 TFLOP (SURVEY.md §8d)"))
        del o, r_
        torch.cuda.empty_cache()
        # ================================================================== hardware multi-GPU parity: sharded == unsharded
        try:
            scores["multi_gpu_parity"] = multi_gpu_parity(dev, rank, world)
        except Exception as ex_:
            scores["multi_gpu_parity"] = dict(error=str(ex_)[:200])
        # ================================================================== C score, configs[4] at job size
        try:
            scores["c_score"] = c_score_job(dev, rank, world, peaks, device_ms, max_over_ranks)
        except Exception as ex_:
            scores["c_score"] = dict(error=str(ex_)[:200])
        torch.cuda.empty_cache()
        # ================================================================== GPU image preprocessing
        from law_of_vision_representation_in_mllms_b200 import preprocess as PP
        rng = np.random.default_rng(rank)
        raw = [rng.integers(0, 256, (480, 640, 3), dtype=np.uint8) for _ in range(8)] * (B // 8)   # one step's worth
        host_raw, offs, hw = PP.pack_images(raw)
        dev_raw = host_raw.to(dev)
        pcfg = tower.image_processor._cfg()
        pws = torch.empty(PP.workspace_bytes(pcfg, hw), dtype=torch.uint8, device=dev)
        pout = torch.empty(len(raw), 3, R, R, dtype=io_dtype, device=dev)
        for _ in range(3):
            PP.run_packed(dev_raw, offs, hw, pcfg, out=pout, ws=pws)
        pre_ms, _ = device_ms(lambda: PP.run_packed(dev_raw, offs, hw, pcfg, out=pout, ws=pws), reps=5)
        pre_gbs = (host_raw.numel() + pout.numel() * 2) / pre_ms / 1e6
        scores["preprocess"] = dict(
            config=f"{B} x 480x640 8-bit RGB -> CLIP-336 pixels on the device (Pillow-exact bicubic, centre crop, HF normalise), per GPU",
            images_per_s=len(raw) / pre_ms * 1e3,
            roofline=dict(bound="hbm", achieved=pre_gbs, peak=peaks["hbm_gbs"], unit="GB/s", frac=pre_gbs / peaks["hbm_gbs"]))
        del dev_raw, pws, pout
        # the same end-to-end path as `e2e`, but from raw 8-bit images: host packing + H2D of bytes + GPU preprocessing +
        # tower + D2H (extract.FeatureExtractor.run_images), i.e. what replaces the reference's CPU DataLoader preprocessing
        ex.run_images(raw, host_out, pixel_dtype=io_dtype)
        t_raw = time.perf_counter()
        raw_ms, _ = device_ms(lambda: ex.run_images(raw, host_out, pixel_dtype=io_dtype), reps=3)
        scores["e2e_from_raw_images"] = dict(
            api=f"extract.FeatureExtractor.run_images ({B} x 480x640 8-bit RGB on the host -> features on the host)",
            images_per_s=world * len(raw) / raw_ms * 1e3, h2d_bytes_per_step=int(host_raw.numel()),
            host_wall_ms_per_step=(time.perf_counter() - t_raw) * 1e3 / 3)
        # ================================================================== configs[3]: dual tower at job size (6250 images per rank)
        ex2 = X.build_extractor(DUAL, chunk=args.chunk, lovr_max_batch=args.chunk, lovr_device=dev, lovr_seed=0, lovr_image_size=R,
                                lovr_operand_dtype=args.operand_dtype)
        n2 = args.chunk
        out2 = torch.empty(n2, T_PATCH, 2 * D, dtype=io_dtype, device=dev)

        def dual_job():
            done, i = 0, 0
            while done < C4_IMAGES_PER_RANK:
                n = min(n2, C4_IMAGES_PER_RANK - done)
                off = 0
                for tw in ex2.towers:                 # the pixel tensors of the two towers differ only in mean/std: same synthetic randn
                    tw.forward_into(dev_px[i % n_in][:n], out2[:n], off)
                    off += tw.hidden_size
                done += n
                i += 1
        off = 0
        for tw in ex2.towers:
            tw.forward_into(dev_px[0][:n2], out2, off)
            off += tw.hidden_size
        dual_ms, _ = device_ms(dual_job)
        scores["dual_tower"] = dict(
            config=f"clip336.dinov2-large @336 extract + fused concat -> [N,576,2048], {C4_IMAGES_PER_RANK} images per rank "
                   f"({world * C4_IMAGES_PER_RANK} images; configs[3] is the N=8 row: 50k images)",
            images=world * C4_IMAGES_PER_RANK, seconds=dual_ms / 1e3, images_per_s=world * C4_IMAGES_PER_RANK / dual_ms * 1e3,
            tflops_per_gpu=C4_IMAGES_PER_RANK * 2 * FLOPS_PER_IMAGE / (dual_ms / 1e3) / 1e12)
        del ex2, out2

    if rank == 0:
        kr = kernel_rooflines(head)
        gemm = kr[0]
        traffic = ncu_traffic()
        roofline = dict(bound="tensor", kernel=gemm["kernel"], achieved=gemm["achieved"], peak=gemm["peak"], unit="TFLOP/s",
                        frac=gemm["frac"], peak_source=peaks["source"], launches=gemm["launches"],
                        gemm_share_of_step=gemm["share_of_step"],
                        traffic=(traffic or {}).get("gemm_dram_bytes_per_launch"),
                        traffic_note=((traffic or {}).get("note") or "no ncu --set full capture of this round is committed yet: null rather than a constant"),
                        whole_step_tflops=head["value"] / world * FLOPS_PER_IMAGE / 1e12,
                        whole_step_frac=head["value"] / world * FLOPS_PER_IMAGE / 1e12 / peaks["tflops"],
                        kernels=kr)
        line = dict(
            metric=METRIC, value=head["value"], unit="images/s", n_gpus=world, steps=args.steps, warmup=max(args.warmup, 3),
            ms_per_step=head["ms"] / args.steps, higher_is_better=True, scaling="weak", vs_baseline=None, dtype=args.operand_dtype,
            data="synthetic", config=bench_config(world, B, args.chunk), gpu_launches=head["launches"], clocks=head["clocks"],
            e2e=dict(value=head["e2e_value"], unit="images/s", ms_per_step=head["ms_e2e"] / args.steps,
                     h2d_bytes_per_step=B * 3 * R * R * 2, d2h_bytes_per_step=B * T_PATCH * D * 2,
                     api="extract.FeatureExtractor.run (pinned host pixels -> device -> pinned host features, host-synchronised)"),
            roofline=roofline)
        if other is not None:
            okr = kernel_rooflines(other)
            line[other_od] = dict(value=other["value"], unit="images/s", ms_per_step=other["ms"] / args.steps,
                                  e2e=other["e2e_value"], gemm_tflops=okr[0]["achieved"], gemm_frac=okr[0]["frac"],
                                  whole_step_tflops=other["value"] / world * FLOPS_PER_IMAGE / 1e12, clocks=other["clocks"],
                                  note=f"the same measurement with {other_od} operands, same process, same inputs")
        if parity is not None:
            line["parity"] = parity
        if scores:
            line["scores"] = scores
        if not args.no_cpu_baseline and world == 1:      # reported on rank 0 at N=1 only
            cpu_reference_leg(1)
            v, dt, thr, kind = cpu_reference_leg(args.cpu_images)
            if scores is not None:
                try:      # the score scripts' arithmetic on the same host cores, bounded sample, reported beside the GPU seconds
                    scores["cpu_port"] = cpu_score_leg(thr)
                except Exception as ex_:
                    scores["cpu_port"] = dict(error=str(ex_)[:200])
            line["cpu_baseline"] = dict(value=v, unit="images/s", cores=thr, kind=kind,
                                        sample=cpu_sample_text(args.cpu_images, thr, _CPU["cores"], kind) + f", {dt:.1f} s")
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def parity_block(dev):
    """Both operand modes against the committed fixture tests/golden/vit_clip336_full_hf.pt (produced by the reference's own
    CLIPVisionTower wrapper in fp32, oracle/make_golden.py): the tower the headline times, at full depth, same weight recipe."""
    import types

    import torch
    from law_of_vision_representation_in_mllms_b200 import specs as S, towers
    from law_of_vision_representation_in_mllms_b200.weights import synthetic_images, synthetic_state_dict
    name = "clip336_full_hf"
    rec = torch.load(os.path.join(ROOT, "tests", "golden", f"vit_{name}.pt"))
    seed = 100 + sum(map(ord, name)) % 1000            # oracle/make_golden.py vit_case_inputs
    spec = S.TOWER_SPECS[TOWER]
    sd = synthetic_state_dict(spec, seed=seed, style="hf_init")
    imgs = synthetic_images(1, R, seed=seed + 1)
    out = dict(fixture=f"tests/golden/vit_{name}.pt (reference wrapper, fp32, 23 layers)", tolerance=1e-3,
               reference_own_bf16_path=rec.get("ref_bf16_rel"), reference_own_fp16_path=rec.get("ref_fp16_rel"))
    for od in ("fp16", "bf16"):
        a = types.SimpleNamespace(mm_vision_select_layer=-2, mm_vision_select_feature="patch", lovr_state_dict=sd,
                                  lovr_max_batch=1, lovr_device=dev, lovr_operand_dtype=od)
        t = towers.CLIPVisionTower(TOWER, a)
        got = t(imgs.to(dev))[:, rec["rows"]].float().cpu()
        out[od + "_rel_err"] = ((got - rec["values"]).norm() / rec["values"].norm()).item()
        del t
    out["fp16_within_tolerance"] = out["fp16_rel_err"] < 1e-3
    return out


def multi_gpu_parity(dev, rank, world):
    """Same seeded data on every rank and every N: each rank computes the UNSHARDED A score / C score / linear CKA locally and
    the SHARDED one (its contiguous shard + the path's single all-reduce over NCCL); they must agree (1e-9 / 1e-9 / 1e-5).
    At N = 1 the 'sharded' side splits the data into three uneven shards on the one GPU."""
    import torch
    import torch.distributed as dist
    from law_of_vision_representation_in_mllms_b200 import a_score as A, c_score as CS, cka as CK, distributed as DD, extract as X
    g = torch.Generator(device=dev).manual_seed(4242)
    n = 256
    o = torch.randn(n, T_PATCH, D, generator=g, device=dev).to(torch.bfloat16)
    r = (0.5 * o.float() + torch.randn(n, T_PATCH, D, generator=g, device=dev)).to(torch.bfloat16)
    shards = [X.shard_range(n, rank, world)] if world > 1 else [(0, 100), (100, 101), (101, n)]
    # A score
    whole_a = A.a_score_pairs(o, r).double().sum().item() / n
    part = torch.zeros(2, dtype=torch.float64, device=dev)
    for lo, hi in shards:
        part += torch.stack([A.a_score_pairs(o[lo:hi], r[lo:hi]).double().sum(), torch.tensor(float(hi - lo), dtype=torch.float64, device=dev)])
    DD.all_reduce_sum(part)
    shard_a = (part[0] / part[1]).item()
    # linear CKA
    whole_c = CK.linear_cka(o, r)
    gram = None
    for lo, hi in shards:
        gram, _, dx, dy = CK.cka_gram(o[lo:hi], r[lo:hi], gram)
    shard_c = DD.cka_reduce(gram, dx, dy)
    rows = gram[-1].item()
    del o, r, gram
    # C score: 48 P=60 pairs, 12 keypoints, 6 pseudo-categories
    npair, P, K, ncat = 48, 60, 12, 6
    d1 = torch.randn(npair, P * P, D, generator=g, device=dev).to(torch.bfloat16)
    d2 = (0.6 * torch.roll(d1.float(), 3, dims=1) + 0.8 * torch.randn(npair, P * P, D, generator=g, device=dev)).to(torch.bfloat16)
    gc = torch.Generator().manual_seed(4243)
    k1 = torch.cat([torch.rand(npair, K, 2, generator=gc) * 839.0, torch.ones(npair, K, 1)], -1)
    k2 = torch.cat([torch.rand(npair, K, 2, generator=gc) * 839.0, (torch.rand(npair, K, 1, generator=gc) > 0.2).float()], -1)
    thr = 200.0 + 600.0 * torch.rand(npair, generator=gc)
    cats = (torch.arange(npair) % ncat).tolist()
    sums_whole, _ = CS.c_score(d1, d2, k1, k2, thr, cats, ncat)
    whole_pck = CS.weighted_c_score(sums_whole)
    cshards = [X.shard_range(npair, rank, world)] if world > 1 else [(0, 20), (20, 21), (21, npair)]
    sums = torch.zeros(ncat, 4, dtype=torch.float64, device=dev)
    for lo, hi in cshards:
        s_, _ = CS.c_score(d1[lo:hi], d2[lo:hi], k1[lo:hi], k2[lo:hi], thr[lo:hi], cats[lo:hi], ncat)
        sums += s_
    shard_pck = DD.c_score_reduce(sums)
    res = dict(world=world, data="same seed on every rank and every N; A/CKA: 256 pairs x 576 x 1024 bf16; C: 48 P=60 pairs, 12 keypoints",
               a_score=dict(unsharded=whole_a, sharded=shard_a, abs_diff=abs(whole_a - shard_a), tolerance=1e-9),
               linear_cka=dict(unsharded=whole_c, sharded=shard_c, abs_diff=abs(whole_c - shard_c), tolerance=1e-5, rows_reduced=rows),
               c_score=dict(unsharded=whole_pck[0], sharded=shard_pck[0], abs_diff=abs(whole_pck[0] - shard_pck[0]), tolerance=1e-9),
               collective="nccl all_reduce(sum), one per score" if world > 1 else "single process (three uneven shards on one GPU)")
    res["ok"] = bool(res["a_score"]["abs_diff"] < 1e-9 and res["linear_cka"]["abs_diff"] < 1e-5 and res["c_score"]["abs_diff"] < 1e-9
                     and rows == n * T_PATCH)
    if world > 1:      # every rank must have seen the same thing
        flag = torch.tensor([1.0 if res["ok"] else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        res["ok"] = bool(flag.item() > 0)
    return res


def c_score_job(dev, rank, world, peaks, device_ms, max_over_ranks):
    """configs[4] at job size: 12 234 SPair-71k-shaped pairs (P = 60, D = 1024, bf16 descriptors, 12 keypoints, 18 pseudo-
    categories) in chunks of 256 seeded by the GLOBAL chunk index; chunks sharded contiguously over ranks; per chunk the
    keypoint path (what pck_train.py consumes: similarity rows at the keypoints -> windowed soft-argmax -> PCK -> category
    sums) and the dense P^2 x P^2 flow (the whole matrix the reference materialises) are timed on the device; descriptor
    generation stands in for loading feature files and is not timed.  One all-reduce of the [18, 4] category sums."""
    import torch
    from law_of_vision_representation_in_mllms_b200 import c_score as CS, distributed as DD, extract as X
    P, K = C5_P, C5_K
    nchunks = (C5_PAIRS + C5_CHUNK - 1) // C5_CHUNK
    clo, chi = X.shard_range(nchunks, rank, world)
    sums = torch.zeros(C5_CATS, 4, dtype=torch.float64, device=dev)
    kp_ms = dense_ms = 0.0
    pairs = 0
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    for c in range(clo, chi):
        n = min(C5_CHUNK, C5_PAIRS - c * C5_CHUNK)
        g = torch.Generator(device=dev).manual_seed(2000 + c)
        d1 = torch.randn(n, P * P, D, generator=g, device=dev).to(torch.bfloat16)
        d2 = (0.6 * torch.roll(d1.float(), 61, dims=1) + 0.8 * torch.randn(n, P * P, D, generator=g, device=dev)).to(torch.bfloat16)
        gc = torch.Generator().manual_seed(2000 + c)
        k1 = torch.cat([torch.rand(n, K, 2, generator=gc) * 839.0, torch.ones(n, K, 1)], -1)
        k2 = torch.cat([torch.rand(n, K, 2, generator=gc) * 839.0, (torch.rand(n, K, 1, generator=gc) > 0.2).float()], -1)
        thr = 200.0 + 600.0 * torch.rand(n, generator=gc)
        cats = ((torch.arange(n) + c * C5_CHUNK) % C5_CATS).tolist()
        if c == clo:      # warm-up at the real shapes
            CS.c_score(d1, d2, k1, k2, thr, cats, C5_CATS)
            CS.get_flow(d1, d2)
        torch.cuda.synchronize()
        e0.record()
        s_, _ = CS.c_score(d1, d2, k1, k2, thr, cats, C5_CATS)
        sums += s_
        e1.record()
        CS.get_flow(d1, d2)
        e2.record()
        torch.cuda.synchronize()
        kp_ms += e0.elapsed_time(e1)
        dense_ms += e1.elapsed_time(e2)
        pairs += n
        del d1, d2
    red_ms, pck = device_ms(lambda: DD.c_score_reduce(sums))
    kp_ms, dense_ms = max_over_ranks(kp_ms), max_over_ranks(dense_ms)
    kp_bytes = (P * P + K) * D * 2.0 * C5_PAIRS        # the target descriptors read once + the K source rows (bf16)
    dense_flops = 2.0 * P ** 4 * D * C5_PAIRS
    kp_gbs = kp_bytes / (kp_ms / 1e3) / 1e9 / world
    dense_tf = dense_flops / (dense_ms / 1e3) / 1e12 / world
    return dict(config=f"C-score, {C5_PAIRS} SPair-71k-shaped pairs (P=60, D=1024, bf16, K=12, 18 categories; configs[4]), chunks of "
                       f"{C5_CHUNK} sharded over {world} rank(s), one all-reduce of the category sums",
                pairs=C5_PAIRS, pairs_this_rank=pairs, c_score_pck010=pck[0], pck005=pck[1], pck001=pck[2],
                keypoint_path_seconds=(kp_ms + red_ms) / 1e3, keypoint_path_us_per_pair=1e3 * kp_ms * world / C5_PAIRS,
                dense_flow_seconds=dense_ms / 1e3, dense_flow_us_per_pair=1e3 * dense_ms * world / C5_PAIRS, reduce_ms=red_ms,
                roofline_keypoint=dict(bound="hbm", achieved=kp_gbs, peak=peaks["hbm_gbs"], unit="GB/s", frac=kp_gbs / peaks["hbm_gbs"],
                                       bytes="(P^2 + K) x D x 2 B = 7.4 MB per pair: only the K keypoint rows of the similarity are consumed"),
                roofline_dense=dict(bound="tensor", achieved=dense_tf, peak=peaks["tflops_burst"], unit="TFLOP/s",
                                    frac=dense_tf / peaks["tflops_burst"], flops="2 x P^4 x D = 26.5 GFLOP per pair"))


if __name__ == "__main__":
    main()
