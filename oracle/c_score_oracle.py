"""ORACLE (test infrastructure — never imported by the product path).

CPU restatement of the reference's C score (dense cosine correspondence -> windowed soft-argmax
-> PCK).  Pinned by tests/golden/c_score_*.pt, produced by
oracle/make_golden.py calling /root/reference/C_score/utils/utils_correspondence.py itself.
`normalize_feats` and the PCK thresholding live in C_score/pck_train.py, which is not importable
offline (ODISE/fvcore imports) — they are restated from the cited lines.
"""
import numpy as np
import torch

ANNO_SIZE = 840


def normalize_feats(feat, epsilon=1e-10):
    """C_score/pck_train.py:24-29 — feat [1,P*P,C]: x / (||x|| + 1e-10)."""
    norms = torch.linalg.norm(feat, dim=-1)[:, :, None]
    return feat / (norms + epsilon)


def normalize_two(feat1, feat2, epsilon=1e-10):
    """C_score/pck_train_two.py:24-35 — per-model normalise, concat channels, renormalise."""
    f = torch.cat([normalize_feats(feat1, epsilon), normalize_feats(feat2, epsilon)], dim=-1)
    return normalize_feats(f, epsilon)


def kpts_to_patch_idx(kps, num_patches, anno_size=ANNO_SIZE):
    """utils_correspondence.py:384-389 — P*int(P/ANNO*y) + int(P/ANNO*x); kps [K,>=2] (x,y,...)."""
    y, x = kps[:, 1].numpy(), kps[:, 0].numpy()
    yp = (num_patches / anno_size * y).astype(np.int32)
    xp = (num_patches / anno_size * x).astype(np.int32)
    return num_patches * yp + xp


def get_flow(sim, window, P, beta=0.02):
    """utils_correspondence.py:297-337 with soft_argmax :234-256 and
    unnormalise_and_convert_mapping_to_flow :258-277.  sim [P*P, P*P] (rows = source).
    Returns [P*P, 2] = (x, y) in patch units, fp32 like the reference."""
    sim = sim.float()
    R = sim.shape[0]
    if window > 0:
        idx = torch.argmax(sim, dim=-1)
        mx, my = idx % P, idx // P
        off = torch.arange(-window, window + 1)
        ox, oy = torch.meshgrid(off, off, indexing="ij")
        ox, oy = ox.flatten(), oy.flatten()
        wx = (mx[:, None] + ox[None, :]).clamp(0, P - 1)
        wy = (my[:, None] + oy[None, :]).clamp(0, P - 1)
        mask = torch.zeros(R, P, P)
        mask[torch.arange(R)[:, None], wy, wx] = 1
        sim = sim * mask.view(R, P * P)            # masked entries become 0, NOT -inf
    M = sim.max(dim=1, keepdim=True).values
    e = torch.exp((sim - M) / beta)
    p = e / e.sum(dim=1, keepdim=True)
    lin = torch.tensor(np.linspace(-1, 1, P)).float()
    p = p.view(R, P, P)                            # [row, ty, tx]
    gx = (p.sum(dim=1) * lin[None, :]).sum(dim=1)
    gy = (p.sum(dim=2) * lin[None, :]).sum(dim=1)
    fx = (gx + 1) * (P - 1) / 2.0
    fy = (gy + 1) * (P - 1) / 2.0
    return torch.stack([fx, fy], dim=1)


def keypoint_transformation(desc1, desc2, patch_idx, P, soft=True, window=5, anno_size=ANNO_SIZE):
    """utils_correspondence.py:345-382.  desc [1,P*P,C] already normalised; returns [K,2] (x,y) px."""
    sim = torch.matmul(desc1, desc2.permute(0, 2, 1))[0]
    stride = anno_size / P
    patch_idx = torch.as_tensor(patch_idx, dtype=torch.long)
    if soft:
        flow = get_flow(sim, window, P)[patch_idx]
        ny, nx = flow[:, 1].clamp(0, P - 1), flow[:, 0].clamp(0, P - 1)
    else:
        nn_ = torch.max(sim[patch_idx], dim=-1).indices
        ny, nx = (nn_ // P).float(), (nn_ % P).float()
    half = stride // 2                              # python float floor-div, e.g. 14.0//2 = 7.0
    return torch.stack([nx * stride + half, ny * stride + half]).permute(1, 0)


def pck_per_image(pred, gt, vis, thr, alphas=(0.1, 0.05, 0.01)):
    """C_score/pck_train.py:150-163 — fraction of visible kps with ||pred-gt|| < alpha*thr."""
    err = (gt[vis] - pred[vis]).norm(dim=-1)
    return [float((err < a * thr).float().mean()) for a in alphas]


def weighted_pck(per_category_means, weights):
    """C_score/utils/logger.py:61-65 — np.average over categories weighted by #pairs."""
    return float(np.average(per_category_means, weights=weights))


# ---- adaptive flip (pck_train.py:111-126) ------------------------------------------------------------------------
def mutual_nn_distance(desc1, desc2):
    """utils_correspondence.py:55-73 get_distance_mutual_nn on normalised descriptors [1,P*P,C], restated in fp64
    (direct differences; the reference's torch.cdist takes the fp32 matmul form for P*P > 25).
    -> (mean distance over mutual nearest neighbours, their count)."""
    a, b = desc1[0].double(), desc2[0].double()
    d = torch.cdist(a, b, compute_mode="donot_use_mm_for_euclid_dist")
    nn12 = d.argmin(dim=1)
    nn21 = d.argmin(dim=0)
    mutual = nn21[nn12] == torch.arange(d.shape[0])
    return float(d.min(dim=1).values[mutual].mean()), int(mutual.sum())


def permute_indices(flip_list, vis=None):
    """utils/utils_geoware.py:151-189 — left/right keypoint swaps; a group is swapped only if all its members are visible."""
    flat = [i for item in flip_list for i in (item if isinstance(item, list) else [item])]
    indices = list(range(max(flat) + 1))
    for item in flip_list:
        if isinstance(item, list) and (vis is None or all(vis[i] for i in item)):
            for i in range(len(item)):
                indices[item[i]] = item[(i + 1) % len(item)]
    return indices


def flip_keypoints(kps, img_size, permute_list=None):
    """utils/utils_geoware.py:199-204."""
    out = kps.detach().clone()
    out[:, 0] = img_size - out[:, 0]
    if permute_list is not None:
        out = out[permute_list]
    return out


def optimized_kps_1_to_2(kps_1_to_2, kps_1_to_2_flip, img1_kps, img2_kps, flip_dist, original_dist, vis, permute_list,
                         anno_size=ANNO_SIZE):
    """utils/utils_geoware.py:269-280 — take the flipped prediction for the keypoints that stay mutually visible after
    the flip when the flipped source is closer (mutual-NN distance) to the target."""
    mutual = img1_kps * vis.unsqueeze(-1).float()
    flipped = flip_keypoints(mutual, anno_size, permute_indices(permute_list, None))
    vis_flip = flipped[:, 2] * img2_kps[:, 2] * img1_kps[:, 2] > 0
    out = kps_1_to_2.clone()
    if flip_dist < original_dist:
        out[vis_flip] = kps_1_to_2_flip[vis_flip]
    return out
