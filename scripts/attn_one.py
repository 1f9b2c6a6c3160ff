import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import gpu_util as G
impl = int(sys.argv[1]) if len(sys.argv) > 1 else 3
n, T, H = 64, 577, 16
qkv = torch.randn(n * T, 3 * H * 64, device="cuda").bfloat16()
for _ in range(3):
    G.attention(qkv, n, T, H, impl)
torch.cuda.synchronize()
