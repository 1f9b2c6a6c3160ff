"""Linear CKA on the device against the fp64 oracle restatement (builder-defined metric; parity unpinned against the
reference, which has no CKA).  Tolerance 1e-5 (north_star) on bf16-representable inputs."""
import pytest
import torch

from law_of_vision_representation_in_mllms_b200 import cka as K
from oracle import a_score_oracle as ao

pytestmark = pytest.mark.gpu


def _data(R, Dx, Dy, seed, shift=0.0):
    g = torch.Generator().manual_seed(seed)
    X = torch.randn(R, Dx, generator=g)
    M = torch.randn(Dx, Dy, generator=g) / Dx ** 0.5
    Y = 0.6 * X @ M + 0.8 * torch.randn(R, Dy, generator=g)
    X, Y = X + shift, Y - 2 * shift
    return X.bfloat16(), Y.bfloat16()


@pytest.mark.parametrize("R,Dx,Dy,shift", [(5000, 256, 512, 0.0), (40000, 1024, 1024, 0.0), (20000, 512, 256, 0.5),
                                            (3 * 576, 768, 1024, 0.0)])
def test_linear_cka_matches_fp64_oracle(R, Dx, Dy, shift):
    X, Y = _data(R, Dx, Dy, seed=R + Dx, shift=shift)
    want = ao.linear_cka(X.float(), Y.float())
    got = K.linear_cka(X.cuda(), Y.cuda())
    assert abs(got - want) < 1e-5, (got, want)


def test_cka_shards_add_up_and_identity():
    X, Y = _data(30000, 256, 256, seed=7)
    full = K.linear_cka(X.cuda(), Y.cuda())
    gram, n1, Dx, Dy = K.cka_gram(X[:11000].cuda(), Y[:11000].cuda())
    gram, n2, _, _ = K.cka_gram(X[11000:].cuda(), Y[11000:].cuda(), gram)        # second shard adds into the same sums
    assert abs(K.cka_from_gram(gram, n1 + n2, Dx, Dy) - full) < 2e-6      # different split boundaries, fp32 partials
    assert gram[-1].item() == 30000                                            # the row count rides in the buffer ...
    assert abs(K.cka_from_gram(gram, None, Dx, Dy) - K.cka_from_gram(gram, n1 + n2, Dx, Dy)) < 1e-12   # ... and finalize can use it (fp64 atomics: order-dependent last bits)
    want = ao.cka_gram_buffer(X.float(), Y.float())                             # buffer layout = the oracle's restatement
    assert tuple(gram.shape) == tuple(want.shape)
    assert ((gram.cpu() - want).abs().max() / want.abs().max()).item() < 1e-5
    assert abs(K.linear_cka(X.cuda(), X.cuda()) - 1.0) < 1e-6
    feats = X.reshape(50, 600, 256)                                             # [N, P, D] layout is flattened
    assert abs(K.linear_cka(feats.cuda(), Y.reshape(50, 600, 256).cuda()) - full) < 1e-12
