"""`mm_projector` on the device: the reference saves A-score tensors AFTER the projector
(llava/model/llava_arch.py:262-263,475-476), so scoring reference-shaped features needs it.

Mirrors llava/model/multimodal_projector/builder.py:34-59 `build_vision_projector(config)`:
  'linear'          -> Linear(mm_hidden_size, hidden_size)
  'mlp{N}x_gelu'    -> Linear, (GELU, Linear) x (N-1)          (LLaVA-1.5 default: mlp2x_gelu, 1024|2048 -> 4096 -> 4096)
  'identity'        -> pass-through
  'perceiver{N}x'   -> out of scope here (NotImplementedError); anything else -> ValueError like the reference.
Every Linear(+GELU) is one launch of the tcgen05 GEMM with the bias/erf-GELU epilogue fused (lovr_gemm_bf16).
Weights: the reference modules' own state_dict — `nn.Sequential` keys ('0.weight', '0.bias', '2.weight', ...) for mlpNx_gelu,
bare `nn.Linear` keys ('weight', 'bias') for 'linear' — or seeded synthetic ones.
"""
import re

import torch
import torch.nn as nn

from . import _lib

ACT_NONE, ACT_GELU_ERF = 0, 2


def _gemm(A, W, bias, act, out_dtype):
    """A [M,K] 16-bit, W [N,K] same dtype -> [M,N] (16-bit with fused activation, or fp32)."""
    M, K = A.shape
    N = W.shape[0]
    epi = _lib.EPI_F32 if out_dtype == torch.float32 else _lib.EPI_BF16
    C = torch.empty(M, N, device=A.device, dtype=out_dtype)
    with torch.cuda.device(A.device):
        _lib.check(_lib.lib().lovr_gemm_bf16(_lib.ptr(A), A.stride(0), _lib.ptr(W), W.stride(0), _lib.ptr(C), C.stride(0),
                                             M, N, K, epi, act, _lib.torch_dtype_code(A.dtype), _lib.ptr(bias), None,
                                             _lib.stream_ptr(A.device)), "lovr_gemm_bf16")
    return C


class IdentityMap(nn.Module):
    def forward(self, x, *args, **kwargs):
        return x

    @property
    def config(self):
        return {"mm_projector_type": "identity"}


class B200MlpProjector(nn.Module):
    def __init__(self, dims, state_dict=None, device="cuda", operand_dtype=torch.bfloat16, seed=0):
        super().__init__()
        self.dims = list(dims)                     # [in, hidden, hidden, ...]
        self.operand_dtype = operand_dtype
        self._w, self._b = [], []
        g = torch.Generator().manual_seed(seed)
        for i in range(len(self.dims) - 1):
            key = str(2 * i)                       # nn.Sequential indices: 0, 2, 4 (GELU modules in between)
            if state_dict is not None:
                if f"{key}.weight" not in state_dict and len(self.dims) == 2 and "weight" in state_dict:
                    key = ""                       # 'linear' is a bare nn.Linear in the reference (builder.py:37-38): keys 'weight' / 'bias'
                pre = f"{key}." if key else ""
                w, b = state_dict[pre + "weight"], state_dict[pre + "bias"]
            else:
                w = torch.randn(self.dims[i + 1], self.dims[i], generator=g) * (self.dims[i] ** -0.5)
                b = torch.randn(self.dims[i + 1], generator=g) * 0.02
            if tuple(w.shape) != (self.dims[i + 1], self.dims[i]):
                raise ValueError(f"projector weight {key}: expected {(self.dims[i + 1], self.dims[i])}, got {tuple(w.shape)}")
            self._w.append(w.to(device, operand_dtype).contiguous())
            self._b.append(b.to(device, torch.float32).contiguous())
        if any(d % 64 for d in self.dims[:-1]) or any(d % 256 for d in self.dims[1:]):
            raise _lib.LovrError(f"projector widths {self.dims}: inputs must be multiples of 64, outputs of 256")

    @torch.no_grad()
    def forward(self, x):
        shape = x.shape
        out_dtype = x.dtype if x.dtype in (torch.float32, torch.bfloat16, torch.float16) else torch.float32
        h = x.reshape(-1, shape[-1])
        if h.dtype != self.operand_dtype:
            h = h.to(self.operand_dtype)           # dtype cast only (operands of the tensor-core GEMM are 16-bit)
        h = h.contiguous()
        n = len(self._w)
        for i in range(n):
            last = i == n - 1
            odt = self.operand_dtype
            if last:
                odt = torch.float32 if out_dtype == torch.float32 else self.operand_dtype
            h = _gemm(h, self._w[i], self._b[i], ACT_NONE if last else ACT_GELU_ERF, odt)
        if h.dtype != out_dtype:
            h = h.to(out_dtype)
        return h.reshape(*shape[:-1], self.dims[-1])


def build_vision_projector(config, delay_load=False, **kwargs):
    projector_type = getattr(config, "mm_projector_type", "linear")
    od = getattr(config, "lovr_operand_dtype", "bf16")
    extra = dict(state_dict=getattr(config, "lovr_projector_state_dict", None),
                 device=getattr(config, "lovr_device", "cuda"),
                 operand_dtype={"bf16": torch.bfloat16, "fp16": torch.float16, "f16": torch.float16}.get(od, od))
    if projector_type == "linear":
        return B200MlpProjector([config.mm_hidden_size, config.hidden_size], **extra)
    m = re.match(r"^mlp(\d+)x_gelu$", projector_type)
    if m:
        depth = int(m.group(1))
        return B200MlpProjector([config.mm_hidden_size] + [config.hidden_size] * depth, **extra)
    if projector_type == "identity":
        return IdentityMap()
    if re.match(r"^perceiver(\d+)x$", projector_type):
        raise NotImplementedError("perceiver projectors are outside the scoring hot path (SURVEY.md §2 row 7)")
    raise ValueError(f"Unknown projector type: {projector_type}")
