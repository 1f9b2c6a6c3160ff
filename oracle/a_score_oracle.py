"""ORACLE (test infrastructure — never imported by the product path).

CPU restatement of the reference's A score (max-cosine matching).
Every function cites the reference lines it follows.  Pinned by tests/golden/a_score_*.pt,
which oracle/make_golden.py produces by exec'ing /root/reference/A_score/compute.py itself.
"""
import torch
import torch.nn.functional as F


def normalize_feat(feat, epsilon=1e-10):
    """A_score/compute.py:12-15 — x / (||x||_2 + 1e-10), in the input dtype."""
    norms = torch.linalg.norm(feat, dim=-1, keepdim=True)
    return feat / (norms + epsilon)


def a_pair_literal(other, ref):
    """A_score/compute.py:54-72 for one (other, reference) pair, literally: normalise, broadcast
    cosine_similarity over [x,1,D] x [1,y,D], max over reference tokens, mean over `other` tokens.
    Materialises the [x,y,D] product like the reference — small shapes only."""
    o = normalize_feat(other).unsqueeze(1)
    r = normalize_feat(ref).unsqueeze(0)
    sim = F.cosine_similarity(o, r, dim=-1)
    return sim.max(dim=1).values.mean().item()


def a_pair(other, ref):
    """Same quantity in GEMM form (agrees with a_pair_literal to ~1e-8 on fp32):
    cosine_similarity clamps each norm at eps=1e-8 after the 1e-10 normalisation."""
    o = normalize_feat(other.double())
    r = normalize_feat(ref.double())
    on = torch.linalg.norm(o, dim=-1).clamp_min(1e-8)
    rn = torch.linalg.norm(r, dim=-1).clamp_min(1e-8)
    sim = (o @ r.T) / (on[:, None] * rn[None, :])
    return sim.max(dim=1).values.mean().item()


def a_score(others, refs336, refs224, pair_fn=a_pair):
    """A_score/compute.py:41-81 — mean over images per reference, then (avg336+avg224)/2."""
    s336 = [pair_fn(o, r) for o, r in zip(others, refs336)]
    s224 = [pair_fn(o, r) for o, r in zip(others, refs224)]
    a336 = sum(s336) / len(s336)
    a224 = sum(s224) / len(s224)
    return (a336 + a224) / 2, s336, s224


def linear_cka(X, Y):
    """Builder-defined metric (named by BASELINE.json north_star; ABSENT from the reference —
    parity unpinned).  Linear CKA over all rows, fp64:
        ||Yc^T Xc||_F^2 / (||Xc^T Xc||_F ||Yc^T Yc||_F),  Xc = X - mean_rows(X)."""
    X = X.reshape(-1, X.shape[-1]).double()
    Y = Y.reshape(-1, Y.shape[-1]).double()
    Xc = X - X.mean(0, keepdim=True)
    Yc = Y - Y.mean(0, keepdim=True)
    hsic = (Yc.T @ Xc).pow(2).sum()
    return (hsic / ((Xc.T @ Xc).pow(2).sum().sqrt() * (Yc.T @ Yc).pow(2).sum().sqrt())).item()


def cka_gram_buffer(X, Y):
    """The per-shard partial-sum buffer of the multi-GPU CKA path, restated in fp64 (layout documented at
    include/lovr_b200.h `lovr_cka_accumulate`): [X^T[X|Y] (Dx x (Dx+Dy)) | Y^T Y (Dy x Dy) | column sums (Dx+Dy) | row count].
    Sums over ranks of these buffers are sufficient statistics of linear CKA."""
    X = X.reshape(-1, X.shape[-1]).double()
    Y = Y.reshape(-1, Y.shape[-1]).double()
    Z = torch.cat([X, Y], dim=1)
    return torch.cat([(X.T @ Z).flatten(), (Y.T @ Y).flatten(), Z.sum(0), torch.tensor([float(X.shape[0])], dtype=torch.float64)])


def cka_from_gram_buffer(buf, Dx, Dy):
    """Centring applied after the reduction: Gc = G - s s^T / n."""
    Dz = Dx + Dy
    na, nb = Dx * Dz, Dy * Dy
    Ga, Gb = buf[:na].reshape(Dx, Dz), buf[na:na + nb].reshape(Dy, Dy)
    s, n = buf[na + nb:na + nb + Dz], buf[na + nb + Dz]
    Gac = Ga - torch.outer(s[:Dx], s) / n
    Gbc = Gb - torch.outer(s[Dx:], s[Dx:]) / n
    xx, xy, yy = Gac[:, :Dx].pow(2).sum(), Gac[:, Dx:].pow(2).sum(), Gbc.pow(2).sum()
    return (xy / (xx.sqrt() * yy.sqrt())).item()
