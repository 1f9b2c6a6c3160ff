"""Parity of each CUDA building block, called through the C ABI, against plain fp32 PyTorch on the same
bf16-representable inputs.  Runs on the B200 box only."""
import pytest
import torch
import torch.nn.functional as F

from law_of_vision_representation_in_mllms_b200 import _lib
import gpu_util as G

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _fp32_reference_math():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.cuda.synchronize()


def _rand(shape, seed, scale=1.0, dtype=torch.bfloat16):
    g = torch.Generator(device="cuda").manual_seed(seed)
    return (torch.randn(*shape, generator=g, device="cuda") * scale).to(dtype)


@pytest.mark.parametrize("M,N,K", [(128, 256, 64), (128, 256, 1024), (200, 1024, 640), (577 * 3, 3072, 1024),
                                   (1154, 768, 3072), (65, 256, 128), (4096 + 37, 4096, 1024)])
def test_gemm_f32_epilogue(M, N, K):
    A, W = _rand((M, K), 1), _rand((N, K), 2, 0.05)
    bias = _rand((N,), 3, 1.0, torch.float32)
    got = G.gemm(A, W, _lib.EPI_F32, bias=bias)
    want = A.float() @ W.float().T + bias
    assert G.rel_err(got, want) < 1e-5          # fp32 accumulation order only
    assert (got - want).abs().max() < 1e-4 * want.abs().max()


@pytest.mark.parametrize("act", [0, 1, 2, 3])
def test_gemm_bf16_epilogue_activations(act):
    M, N, K = 577 * 2 + 5, 1024, 1024
    A, W = _rand((M, K), 4), _rand((N, K), 5, 0.05)
    bias = _rand((N,), 6, 0.5, torch.float32)
    got = G.gemm(A, W, _lib.EPI_BF16, act=act, bias=bias)
    z = A.float() @ W.float().T + bias
    want = [z, z * torch.sigmoid(1.702 * z), F.gelu(z), F.gelu(z, approximate="tanh")][act]
    assert got.dtype == torch.bfloat16
    # bf16 output rounding: 2^-9 relative per element
    assert G.rel_err(got, want) < 3e-3
    assert G.rel_err(got, want.to(torch.bfloat16)) < 1.5e-3


@pytest.mark.parametrize("use_ls", [False, True])
def test_gemm_residual_epilogue(use_ls):
    M, N, K = 1300, 1024, 4096
    A, W = _rand((M, K), 7), _rand((N, K), 8, 0.02)
    bias = _rand((N,), 9, 0.5, torch.float32)
    ls = _rand((N,), 10, 1.0, torch.float32) if use_ls else None
    x0 = _rand((M, N), 11, 3.0, torch.float32)
    x = x0.clone()
    G.gemm(A, W, _lib.EPI_RESID_F32, bias=bias, colscale=ls, C_inout=x)
    delta = A.float() @ W.float().T + bias
    want = x0 + (delta * ls if use_ls else delta)
    assert G.rel_err(x, want) < 1e-5


@pytest.mark.parametrize("act", [0, 1])
def test_gemm_fp16_operands(act):
    M, N, K = 700, 1024, 1024
    A, W = _rand((M, K), 21, 1.0, torch.float16), _rand((N, K), 22, 0.05, torch.float16)
    bias = _rand((N,), 23, 0.5, torch.float32)
    got = G.gemm(A, W, _lib.EPI_BF16, act=act, bias=bias)
    z = A.float() @ W.float().T + bias
    want = [z, z * torch.sigmoid(1.702 * z)][act]
    assert got.dtype == torch.float16
    assert G.rel_err(got, want) < 4e-4          # fp16 output rounding: 2^-12 relative
    x = _rand((M, N), 24, 2.0, torch.float32)
    x0 = x.clone()
    G.gemm(A, W, _lib.EPI_RESID_F32, bias=bias, C_inout=x)
    assert G.rel_err(x, x0 + z) < 1e-5


def test_attention_fp16_operands():
    n, T, H = 2, 577, 16
    qkv = _rand((n * T, 3 * H * 64), 31, 1.0, torch.float16)
    want = _torch_attention(qkv, n, T, H)
    for impl in (_lib.ATTN_TCGEN05_V3, _lib.ATTN_TCGEN05_V4, _lib.ATTN_SIMT_CHECK):
        got = G.attention(qkv, n, T, H, impl)
        assert got.dtype == torch.float16
        assert G.rel_err(got, want) < 6e-4


def test_gemm_rejects_bad_shapes():
    A, W = _rand((128, 100), 1), _rand((256, 100), 2)
    with pytest.raises(_lib.LovrError):
        G.gemm(A, W, _lib.EPI_F32)


@pytest.mark.parametrize("D,eps", [(1024, 1e-5), (768, 1e-6)])
def test_layernorm(D, eps):
    x = _rand((1000, D), 12, 2.0, torch.float32) + 0.7
    g, b = _rand((D,), 13, 0.2, torch.float32) + 1, _rand((D,), 14, 0.2, torch.float32)
    got = G.layernorm(x, g, b, eps)
    want = F.layer_norm(x, (D,), g, b, eps)
    assert G.rel_err(got, want.to(torch.bfloat16)) < 1e-3
    assert G.rel_err(got, want) < 3e-3


@pytest.mark.parametrize("R,P,kpad,dtype", [(224, 14, 640, torch.float32), (336, 14, 640, torch.bfloat16),
                                             (224, 16, 768, torch.float16)])
def test_im2col_matches_unfold(R, P, kpad, dtype):
    px = _rand((3, 3, R, R), 15, 1.0, dtype)
    cols = G.im2col(px, P, kpad)
    want = F.unfold(px.float(), kernel_size=P, stride=P).transpose(1, 2).reshape(-1, 3 * P * P)
    assert torch.equal(cols[:, : 3 * P * P].float(), want.to(torch.bfloat16).float())
    assert cols[:, 3 * P * P:].abs().max().item() == 0 if kpad > 3 * P * P else True
    if dtype != torch.float32:      # pixels already in the operand format: the word-copy kernel, exact
        same = G.im2col(px, P, kpad, out_dtype=dtype)
        assert torch.equal(same[:, : 3 * P * P].float(), want.to(dtype).float())
        assert kpad == 3 * P * P or same[:, 3 * P * P:].abs().max().item() == 0


def _torch_attention(qkv, n, T, H):
    D = H * 64
    q, k, v = qkv.float().view(n, T, 3, H, 64).permute(2, 0, 3, 1, 4)
    p = torch.softmax(q @ k.transpose(-1, -2) * 0.125, dim=-1)
    return (p @ v).transpose(1, 2).reshape(n * T, D)


@pytest.mark.parametrize("impl", [_lib.ATTN_SIMT_CHECK, _lib.ATTN_TCGEN05_V3, _lib.ATTN_TCGEN05_V4])
@pytest.mark.parametrize("n,T,H,scale", [(2, 577, 16, 1.0), (3, 257, 16, 2.0), (2, 196, 12, 1.0), (1, 50, 16, 1.0),
                                          (1, 128, 16, 3.0), (1, 640, 16, 1.0)])
def test_attention(impl, n, T, H, scale):
    qkv = _rand((n * T, 3 * H * 64), 16 + T, scale)
    got = G.attention(qkv, n, T, H, impl)
    want = _torch_attention(qkv, n, T, H)
    assert torch.isfinite(got.float()).all()
    # P is rounded to bf16 before P·V on the tensor-core path; output is bf16
    assert G.rel_err(got, want) < (3e-3 if impl == _lib.ATTN_SIMT_CHECK else 4e-3)


def test_attention_tcgen05_matches_simt_check():
    n, T, H = 4, 577, 16
    qkv = _rand((n * T, 3 * H * 64), 99, 1.5)
    a = G.attention(qkv, n, T, H, _lib.ATTN_AUTO)
    b = G.attention(qkv, n, T, H, _lib.ATTN_SIMT_CHECK)
    c = G.attention(qkv, n, T, H, _lib.ATTN_TCGEN05_V3)
    assert G.rel_err(a, b.float()) < 4e-3
    assert G.rel_err(c, b.float()) < 4e-3
    with pytest.raises(_lib.LovrError, match="not in this build"):      # the removed generations fail loudly
        G.attention(qkv, n, T, H, 0)


@pytest.mark.parametrize("n,T,H", [(40, 577, 16), (70, 257, 16), (100, 196, 12), (19, 640, 16), (10, 300, 16),
                                   (23, 129, 16), (31, 256, 12), (13, 384, 16), (11, 513, 16), (29, 200, 16), (10, 128, 16)])
def test_persistent_attention_many_items(n, T, H):
    """attention4: more (image, head) items than SMs, so every CTA rolls several items through its three pipelines; the
    per-tile arithmetic is attention3's, so the two must agree bit for bit, and both with the fp32 reference."""
    qkv = _rand((n * T, 3 * H * 64), 7 + T, 1.5)
    got = G.attention(qkv, n, T, H, _lib.ATTN_TCGEN05_V4)
    # same tiles, same products; the two kernels only differ in which row maximum the exponentials are taken against
    assert G.rel_err(got, G.attention(qkv, n, T, H, _lib.ATTN_TCGEN05_V3).float()) < 2e-3
    assert G.rel_err(got, _torch_attention(qkv, n, T, H)) < 4e-3
    for _ in range(3):     # no state leaks between launches
        assert torch.equal(G.attention(qkv, n, T, H, _lib.ATTN_TCGEN05_V4), got)


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float16])
@pytest.mark.parametrize("T", [577, 257, 200])
def test_attention_growing_logits_redo_path(dtype, T):
    """Keys whose scores dwarf the first 64-key unit's maximum (by far more than 2^15 / 2^60 in the exp2 domain): the
    persistent kernel takes its reference maximum from the first unit only and must detect the overflow through the unit's
    row sum, rescale and redo the unit.  Also the opposite case (first unit dominates everything after it)."""
    n, H = 3, 16
    g = torch.Generator().manual_seed(5 + T)
    qkv = torch.randn(n, T, 3, H, 64, generator=g)
    grow = torch.ones(T)
    grow[64:] = 12.0
    grow[T // 2:] = 40.0                       # logits of later keys: std ~ 8 * 40 = 320 -> 2^(320 * 0.18) per sigma
    for scale in (grow, grow.flip(0)):
        x = qkv.clone()
        x[:, :, 1] *= scale[None, :, None, None]
        x = x.reshape(n * T, 3 * H * 64).to(dtype).cuda()
        want = _torch_attention(x, n, T, H)
        for impl in (_lib.ATTN_TCGEN05_V4, _lib.ATTN_TCGEN05_V3):
            got = G.attention(x, n, T, H, impl)
            assert torch.isfinite(got.float()).all()
            assert G.rel_err(got, want) < (4e-3 if dtype == torch.bfloat16 else 1e-3), (impl, T)


@pytest.mark.parametrize("op_dtype,tol", [(torch.bfloat16, 6e-3), (torch.float16, 8e-4)])
@pytest.mark.parametrize("M,D,row_mean", [(3 * 577, 1024, 0.0), (100, 768, 0.0), (700, 1024, 40.0)])
def test_layernorm_folded_into_gemm_epilogues(op_dtype, tol, M, D, row_mean):
    """LN(x) W^T + b = rstd (x^ W'^T - mu cw) + bw (gemm.cuh): the residual epilogue emits x^ and per-128-column statistics
    slots, lovr_ln_fold_weights folds gamma / beta into the next GEMM's weights, the 16-bit epilogue applies mu / rstd.
    Against plain torch fp32: x + GEMM, then LayerNorm, then Linear + activation.  `row_mean` = 40 puts a DC offset of 40
    standard deviations on every row (the shifted sums keep the variance from cancelling)."""
    g = torch.Generator().manual_seed(M + D)
    K1, N2 = 512, 768
    A = (torch.randn(M, K1, generator=g) * 0.5).to(op_dtype)
    W1 = (torch.randn(D, K1, generator=g) * 0.05).to(op_dtype)
    b1 = torch.randn(D, generator=g) * 0.1
    ls = 0.5 + torch.rand(D, generator=g)
    x0 = torch.randn(M, D, generator=g) + row_mean
    x0[:, 7] += 25.0                                                    # an outlier channel
    want_x = x0 + (A.float() @ W1.float().T + b1) * ls
    x = x0.clone().cuda()
    xhat, stats = G.gemm_resid_lnfold(A.cuda(), W1.cuda(), x, b1.cuda(), ls.cuda())
    assert G.rel_err(x.cpu(), want_x) < 1e-5
    assert torch.equal(xhat.cpu(), x.cpu().to(op_dtype))                # x^ is exactly the rounded residual stream
    st = stats.cpu().double()                                           # slots -> row mean / variance (pairwise merge)
    seg_mean = st[..., 0] + st[..., 1] / 128
    seg_m2 = st[..., 2] - st[..., 1] ** 2 / 128
    mean = seg_mean.mean(1)
    var = (seg_m2.sum(1) + 128 * ((seg_mean - mean[:, None]) ** 2).sum(1)) / D
    assert (mean - want_x.double().mean(1)).abs().max() < 1e-4 * (1 + abs(row_mean))
    assert ((var - want_x.double().var(1, unbiased=False)) / want_x.double().var(1, unbiased=False)).abs().max() < 1e-4
    # consumer
    gamma, beta = 0.5 + torch.rand(D, generator=g), 0.2 * torch.randn(D, generator=g)
    W2 = torch.randn(N2, D, generator=g) * 0.04
    b2 = torch.randn(N2, generator=g) * 0.1
    W16, cw, bw = G.ln_fold_weights(W2.cuda(), gamma.cuda(), beta.cuda(), b2.cuda(), op_dtype)
    assert torch.equal(W16.cpu(), (W2 * gamma).to(op_dtype))
    assert torch.allclose(cw.cpu(), W16.float().cpu().sum(1), rtol=1e-5, atol=1e-5)
    assert torch.allclose(bw.cpu(), W2 @ beta + b2, rtol=1e-5, atol=1e-5)
    for act, fn in ((0, lambda t: t), (1, lambda t: t * torch.sigmoid(1.702 * t))):
        got = G.gemm_consume_lnfold(xhat, stats, W16, cw, bw, 1e-5, act).float().cpu()
        want = fn(torch.nn.functional.layer_norm(want_x, (D,), gamma, beta, 1e-5) @ W2.T + b2)
        rel = G.rel_err(got, want)
        # x^ is rounded relative to |x|, not |x - mu|: a row whose mean is 40 standard deviations loses log2(40) bits of
        # (x - mu) — the documented limit of the fold (real residual streams have |mu| < sigma); the statistics stay exact
        assert rel < tol * (12 if row_mean else 1), (act, rel)
