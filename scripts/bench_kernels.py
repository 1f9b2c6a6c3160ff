"""Per-kernel timings on the device (CUDA events, L2-sized inputs): GEMM shapes of one ViT-L layer,
attention, LayerNorm.  Prints one line per kernel; used to fill profiles/ and DESIGN.md."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import gpu_util as G  # noqa: E402
from law_of_vision_representation_in_mllms_b200 import _lib  # noqa: E402


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(iters)]
    for a, b in evs:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2]


def main():
    n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    T, D, I, H = 577, 1024, 4096, 16
    M = n_img * T
    dev = "cuda"
    res = []
    A = torch.randn(M, D, device=dev).bfloat16()
    Hbuf = torch.randn(M, I, device=dev).bfloat16()
    x = torch.randn(M, D, device=dev)
    bias = {n: torch.randn(n, device=dev) for n in (D, 3 * D, I)}
    W = {k: (torch.randn(*s, device=dev) * 0.02).bfloat16() for k, s in
         dict(qkv=(3 * D, D), o=(D, D), fc1=(I, D), fc2=(D, I)).items()}
    cases = [
        ("gemm_qkv", lambda: G.gemm(A, W["qkv"], _lib.EPI_BF16, 0, bias[3 * D], C_inout=qkv), 2 * M * D * 3 * D),
        ("gemm_out_resid", lambda: G.gemm(A, W["o"], _lib.EPI_RESID_F32, 0, bias[D], C_inout=x), 2 * M * D * D),
        ("gemm_fc1_quickgelu", lambda: G.gemm(A, W["fc1"], _lib.EPI_BF16, 1, bias[I], C_inout=Hbuf), 2 * M * D * I),
        ("gemm_fc1_gelu_erf", lambda: G.gemm(A, W["fc1"], _lib.EPI_BF16, 2, bias[I], C_inout=Hbuf), 2 * M * D * I),
        ("gemm_fc2_resid", lambda: G.gemm(Hbuf, W["fc2"], _lib.EPI_RESID_F32, 0, bias[D], C_inout=x), 2 * M * D * I),
    ]
    qkv = torch.randn(M, 3 * D, device=dev).bfloat16()
    for name, fn, flops in cases:
        ms = timeit(fn)
        res.append(dict(kernel=name, ms=ms, tflops=flops / ms / 1e9))
    ms = timeit(lambda: torch.matmul(A, W["fc1"].T))
    res.append(dict(kernel="torch_matmul_fc1_shape(cuBLAS)", ms=ms, tflops=2 * M * D * I / ms / 1e9))
    for impl, nm in ((_lib.ATTN_TCGEN05_V3, "attention3"), (_lib.ATTN_TCGEN05_V4, "attention4 persistent (default)")):
        ms = timeit(lambda: G.attention(qkv, n_img, T, H, impl))
        res.append(dict(kernel=nm, ms=ms, tflops=4 * n_img * H * T * T * 64 / ms / 1e9))
    g, b = torch.ones(D, device=dev), torch.zeros(D, device=dev)
    ms = timeit(lambda: G.layernorm(x, g, b, 1e-5))
    res.append(dict(kernel="layernorm", ms=ms, gbs=M * D * 6 / ms / 1e6))
    res.append(preprocess_timing(256))
    for r in res:
        print(json.dumps(r))


def preprocess_timing(n):
    """lovr_preprocess on n device-resident 480x640 RGB images -> [n,3,336,336] bf16 (CLIP-336 processor):
    algorithmic bytes = source bytes read + output bytes written."""
    import numpy as np
    from law_of_vision_representation_in_mllms_b200 import preprocess as PP
    rng = np.random.default_rng(0)
    imgs = [rng.integers(0, 256, (480, 640, 3), dtype=np.uint8) for _ in range(8)] * (n // 8)
    host, offsets, hw = PP.pack_images(imgs)
    dev = host.cuda()
    proc = PP.B200ImageProcessor({"shortest_edge": 336}, 336)
    cfg = proc._cfg()
    ws = torch.empty(PP.workspace_bytes(cfg, hw), dtype=torch.uint8, device="cuda")
    out = torch.empty(n, 3, 336, 336, dtype=torch.bfloat16, device="cuda")
    ms = timeit(lambda: PP.run_packed(dev, offsets, hw, cfg, out=out, ws=ws))
    nbytes = host.numel() + out.numel() * 2
    return dict(kernel="preprocess_clip336_from_480x640", ms=ms, images_per_s=n / ms * 1e3, gbs=nbytes / ms / 1e6)


if __name__ == "__main__":
    main()
