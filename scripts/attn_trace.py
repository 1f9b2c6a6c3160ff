"""Dump a clock64 timeline of CTA 0 / pipeline 0 of the attention kernel (LOVR_ATTN_DEBUG=8) — diagnostics only."""
import ctypes, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import gpu_util as G
n, T, H = 64, 577, 16
qkv = torch.randn(n * T, 3 * H * 64, device="cuda").bfloat16()
G.attention(qkv, n, T, H, 3); torch.cuda.synchronize()
os.environ["LOVR_ATTN_DEBUG"] = "8"; os.environ["LOVR_ATTN_TRACE_PTR"] = "/tmp/trace_ptr.txt"
G.attention(qkv, n, T, H, 3); torch.cuda.synchronize()
ptr = int(open("/tmp/trace_ptr.txt").read())
buf = (ctypes.c_longlong * 8192)()
cudart = ctypes.CDLL("libcudart.so")
rc = cudart.cudaMemcpy(buf, ctypes.c_void_p(ptr), 8192 * 8, 2)
v = list(buf)
t0 = v[4000]
print("rc", rc, "kernel start->mma ready", v[4001] - t0, "k_full", v[4002] - t0, "end", v[4003] - t0)
for k in range(2):
    print(f"--- tile k={k} softmax warp4 lane0: unit: wait_s_full(start,end) ld_done exp_done p_empty_done arrived")
    for u in range(10):
        e = [v[(k * 16 + u) * 8 + i] - t0 for i in range(6)]
        print(u, e, "| deltas", [e[i + 1] - e[i] for i in range(5)])
    print("o_full wait", v[(k * 16 + 15) * 8 + 6] - t0, v[(k * 16 + 15) * 8 + 7] - t0)
    print(f"--- tile k={k} mma issuer: unit: qk(wait_start, issued) pv(wait_start, wait_end, committed)")
    for u in range(10):
        e = [v[2000 + (k * 16 + u) * 6 + i] - t0 for i in range(5)]
        print(u, e)
