"""C score on the device (reference: C_score/pck_train.py + C_score/utils/utils_correspondence.py).

Reference call chain per image pair (pck_train.py:96-163):
    desc = torch.load(<img>_<MODEL>.pt) [1,C,h,w] -> [1,P*P,C]          prepare_feature_paths_and_load :31-44
    desc = normalize_feats(desc)                                           :24-29
    idx  = kpts_to_patch_idx(args, img1_kps, P)                            utils_correspondence.py:384-389
    kps  = calculate_keypoint_transformation(args, d1, d2, idx, P)         :345-382  (sim -> get_flow -> index -> pixels)
    pck  = mean(||kps - gt|| < alpha * bbox)  alpha in {.1,.05,.01}        pck_train.py:150-163
    C    = np.average(per-category mean PCK@0.10, weights=#pairs)          utils/logger.py:61-65

Here `keypoint_transformation` (only the K keypoint rows, plain fp32 — what the callers consume) and `get_flow`
(the dense P^2 x P^2 path on tcgen05) take UN-normalised descriptors; normalisation is fused.
"""
import ctypes as C
from collections import defaultdict

import numpy as np
import torch

from . import _lib

ANNO_SIZE = 840
_DENSE_PAIRS_PER_CALL = 8


def desc_from_feature_map(fmap):
    """pck_train.py:38-39 with the identity aggregation net: [1,C,h,w] -> [1, h*w, C]."""
    return fmap.reshape(1, 1, fmap.shape[1], -1).permute(0, 1, 3, 2)[0]


def concat_two_models(desc_a, desc_b):
    """pck_train_two.py:24-35 — normalise each model's descriptors and concatenate the channels (`lovr_c_score_concat_two`);
    the renormalisation of the concatenation is fused into the similarity kernels.  desc [B,P*P,C1], [B,P*P,C2] on the
    device, same dtype (fp32 or bf16) -> [B,P*P,C1+C2]."""
    if desc_a.shape[:-1] != desc_b.shape[:-1]:
        raise ValueError(f"descriptor grids differ: {tuple(desc_a.shape)} vs {tuple(desc_b.shape)}")
    if not desc_a.is_cuda:
        raise _lib.LovrError("C-score kernels need CUDA tensors: this package has no CPU path")
    if desc_a.dtype not in (torch.float32, torch.bfloat16):
        desc_a = desc_a.float()
    a, b = desc_a.contiguous(), desc_b.to(desc_a.dtype).contiguous()
    out = torch.empty(*a.shape[:-1], a.shape[-1] + b.shape[-1], dtype=a.dtype, device=a.device)
    rows = a.numel() // a.shape[-1]
    with torch.cuda.device(a.device):
        _lib.check(_lib.lib().lovr_c_score_concat_two(_lib.ptr(a), _lib.ptr(b), _lib.torch_dtype_code(a.dtype), rows,
                                                      a.shape[-1], b.shape[-1], _lib.ptr(out), _lib.stream_ptr(a.device)),
                   "lovr_c_score_concat_two")
    return out


def kpts_to_patch_idx(kps, num_patches, anno_size=ANNO_SIZE):
    """utils_correspondence.py:384-389 (host integer arithmetic, numpy like the reference)."""
    k = kps.detach().cpu()
    y, x = k[..., 1].numpy(), k[..., 0].numpy()
    yp = (num_patches / anno_size * y).astype(np.int32)
    xp = (num_patches / anno_size * x).astype(np.int32)
    return num_patches * yp + xp


def _prep(d1, d2, mult):
    if d1.dim() == 2:
        d1, d2 = d1[None], d2[None]
    if d1.shape != d2.shape or d1.dim() != 3:
        raise ValueError(f"expected two [B,P*P,D] tensors, got {tuple(d1.shape)} and {tuple(d2.shape)}")
    if not d1.is_cuda:
        raise _lib.LovrError("C-score kernels need CUDA tensors: this package has no CPU path")
    if d1.dtype not in (torch.float32, torch.bfloat16):
        d1, d2 = d1.float(), d2.float()
    d2 = d2.to(d1.dtype)
    D = d1.shape[-1]
    if D % mult:
        pad = mult - D % mult                   # zero channels change neither dot products nor norms
        d1, d2 = torch.nn.functional.pad(d1, (0, pad)), torch.nn.functional.pad(d2, (0, pad))
    P = int(round(d1.shape[1] ** 0.5))
    if P * P != d1.shape[1]:
        raise ValueError(f"{d1.shape[1]} patches is not a square grid")
    return d1.contiguous(), d2.contiguous(), P


def _workspace(nbytes, device):
    ws = torch.empty(nbytes + 1024, dtype=torch.uint8, device=device)
    return ws, C.c_void_p((ws.data_ptr() + 1023) // 1024 * 1024)


def get_flow(desc1, desc2, window=5, soft=True, return_argmax=False):
    """Dense correspondence for every source patch: == utils_correspondence.get_flow(sim, window, P).reshape(-1,2)
    per pair, with sim = n(desc1) @ n(desc2)^T.  -> flow [B,P*P,2] (x,y in patch units) (, argmax [B,P*P])."""
    d1, d2, P = _prep(desc1, desc2, 64)
    B, PP, D = d1.shape
    L = _lib.lib()
    dt = _lib.torch_dtype_code(d1.dtype)
    flow = torch.empty(B, PP, 2, dtype=torch.float32, device=d1.device)
    amax = torch.empty(B, PP, dtype=torch.int32, device=d1.device)
    step = min(B, _DENSE_PAIRS_PER_CALL)
    nbytes = L.lovr_c_score_workspace_bytes(dt, step, P, D, 0)
    ws, base = _workspace(nbytes, d1.device)
    with torch.cuda.device(d1.device):
        st = _lib.stream_ptr(d1.device)
        for s in range(0, B, step):
            e = min(B, s + step)
            _lib.check(L.lovr_c_score_flow(_lib.ptr(d1[s:e]), _lib.ptr(d2[s:e]), dt, e - s, P, D, window, int(soft),
                                           _lib.ptr(flow[s:e]), _lib.ptr(amax[s:e]), base, nbytes, st),
                       "lovr_c_score_flow")
    return (flow, amax) if return_argmax else flow


def keypoint_transformation(desc1, desc2, patch_idx, soft=True, window=5, anno_size=ANNO_SIZE):
    """== calculate_keypoint_transformation(args(SOFT_EVAL=soft, SOFT_EVAL_WINDOW=window, ANNO_SIZE), n(d1), n(d2),
    patch_idx, P) for a batch of pairs: desc [B,P*P,D], patch_idx [B,K] int -> [B,K,2] (x,y) pixels."""
    d1, d2, P = _prep(desc1, desc2, 32)
    B, PP, D = d1.shape
    idx_host = np.asarray(patch_idx.detach().cpu() if isinstance(patch_idx, torch.Tensor) else patch_idx).reshape(B, -1)
    if idx_host.size and (int(idx_host.min()) < 0 or int(idx_host.max()) >= PP):      # on the host: no device round trip
        raise IndexError(f"index out of bounds for dimension 0 with size {PP}")
    idx = torch.as_tensor(idx_host.astype(np.int32)).to(d1.device).contiguous()
    K = idx.shape[1]
    L = _lib.lib()
    dt = _lib.torch_dtype_code(d1.dtype)
    out = torch.empty(B, K, 2, dtype=torch.float32, device=d1.device)
    nbytes = L.lovr_c_score_workspace_bytes(dt, B, P, D, K)
    ws, base = _workspace(nbytes, d1.device)
    with torch.cuda.device(d1.device):
        _lib.check(L.lovr_c_score_keypoints(_lib.ptr(d1), _lib.ptr(d2), dt, B, P, D, _lib.ptr(idx), K, window, int(soft),
                                            float(anno_size), _lib.ptr(out), base, nbytes, _lib.stream_ptr(d1.device)),
                   "lovr_c_score_keypoints")
    return out


def pck_per_pair(pred, gt, visible, thresholds):
    """pck_train.py:150-163 for a batch: pred/gt [B,K,2] (x,y), visible [B,K] bool, thresholds [B] (bbox size, or
    ANNO_SIZE when the dataset has none) -> [B,4] = PCK@.10, @.05, @.01, #visible."""
    B, K, _ = pred.shape
    dev = pred.device
    out = torch.empty(B, 4, dtype=torch.float32, device=dev)
    # every device tensor whose raw pointer goes into the call is held in a local until after the launch: a temporary freed
    # between two argument evaluations can be handed to the NEXT host-to-device copy by the caching allocator, which then
    # overwrites it before the kernel reads it (seen on hardware: thresholds landing in gt's block changed a PCK)
    vis = visible.to(dev, torch.uint8).contiguous()
    pred_f = pred.float().contiguous()
    gt_f = gt.to(dev).float().contiguous()
    thr_f = thresholds.to(dev).float().contiguous()
    with torch.cuda.device(dev):
        _lib.check(_lib.lib().lovr_pck(_lib.ptr(pred_f), _lib.ptr(gt_f), _lib.ptr(vis), _lib.ptr(thr_f), B, K,
                                       _lib.ptr(out), _lib.stream_ptr(dev)), "lovr_pck")
    return out


def category_sums(pck, categories, n_categories):
    """Per-rank partial sums for the data-parallel reduction: [n_categories, 4] =
    sum PCK@.10, @.05, @.01 over this rank's pairs and #pairs.  One all-reduce(sum) of this tensor, then
    `weighted_c_score`."""
    sums = torch.zeros(n_categories, 4, dtype=torch.float64, device=pck.device)
    cat = torch.as_tensor(categories, device=pck.device, dtype=torch.long)
    sums[:, :3].index_add_(0, cat, pck[:, :3].double())
    sums[:, 3].index_add_(0, cat, torch.ones(len(cat), dtype=torch.float64, device=pck.device))
    return sums


def weighted_c_score(sums):
    """compute_pck's per-category image means (pck_train.py:204-209) then np.average(..., weights=#pairs)
    (utils/logger.py:61-65).  -> (PCK@0.10, @0.05, @0.01); the first one is the paper's C score (`corres`)."""
    s = sums.detach().cpu().numpy()
    n = s[:, 3]
    keep = n > 0
    means = s[keep, :3] / n[keep, None]
    return tuple(float(np.average(means[:, i], weights=n[keep])) for i in range(3))


def c_score(desc1, desc2, kps1, kps2, thresholds, categories, n_categories, soft=True, window=5, anno_size=ANNO_SIZE):
    """End to end for a batch of pairs on this rank: descriptors [B,P*P,D], kps [B,K,3] (x,y,visible).
    Returns (category partial sums [n_categories,4], predicted kps [B,K,2])."""
    P = int(round(desc1.shape[1] ** 0.5))
    idx = kpts_to_patch_idx(kps1.reshape(-1, 3), P, anno_size).reshape(kps1.shape[0], -1)
    pred = keypoint_transformation(desc1, desc2, idx, soft, window, anno_size)
    vis = (kps1[..., 2] * kps2[..., 2]) > 0                      # pck_train.py:131
    pck = pck_per_pair(pred, kps2[..., :2], vis, thresholds)
    return category_sums(pck, categories, n_categories), pred


# ---- adaptive flip (pck_train.py:111-126, ADAPT_FLIP + MUTUAL_NN) ---------------------------------------------------
def get_distance_mutual_nn(desc1, desc2, return_count=False):
    """== utils_correspondence.get_distance_mutual_nn(n(desc1), n(desc2)) per pair (:55-73): the mean descriptor distance
    over mutually-nearest patch pairs.  desc [B,P*P,D] UN-normalised (normalisation fused) -> [B] fp32 (NaN where no mutual
    pair exists, like torch's empty mean)."""
    d1, d2, P = _prep(desc1, desc2, 64)
    B, PP, D = d1.shape
    L = _lib.lib()
    dt = _lib.torch_dtype_code(d1.dtype)
    out = torch.empty(B, 2, dtype=torch.float32, device=d1.device)
    step = min(B, _DENSE_PAIRS_PER_CALL)
    nbytes = L.lovr_c_score_mutual_nn_workspace_bytes(dt, step, P, D)
    ws, base = _workspace(nbytes, d1.device)
    with torch.cuda.device(d1.device):
        st = _lib.stream_ptr(d1.device)
        for s in range(0, B, step):
            e = min(B, s + step)
            _lib.check(L.lovr_c_score_mutual_nn(_lib.ptr(d1[s:e]), _lib.ptr(d2[s:e]), dt, e - s, P, D, _lib.ptr(out[s:e]),
                                                base, nbytes, st), "lovr_c_score_mutual_nn")
    return (out[:, 0], out[:, 1].long()) if return_count else out[:, 0]


def permute_indices(flip_list, vis=None):
    """utils/utils_geoware.py:151-189: index permutation that swaps left/right keypoint groups (`[4, 5]` -> 4<->5, longer
    groups rotate); with `vis`, a group is only swapped when every member is visible."""
    members = [g if isinstance(g, list) else [g] for g in flip_list]
    perm = list(range(max(i for g in members for i in g) + 1))
    for g in members:
        if len(g) > 1 and (vis is None or all(bool(vis[i]) for i in g)):
            for pos, i in enumerate(g):
                perm[i] = g[(pos + 1) % len(g)]
    return perm


def flip_keypoints(keypoints, img_size, permute_list=None):
    """utils/utils_geoware.py:199-204: mirror x, then reorder."""
    out = keypoints.detach().clone()
    out[:, 0] = img_size - out[:, 0]
    return out if permute_list is None else out[permute_list]


def optimized_kps_1_to_2(kps_1_to_2, kps_1_to_2_flip, img1_kps, img2_kps, flip_dist, original_dist, vis, permute_list,
                         anno_size=ANNO_SIZE):
    """utils/utils_geoware.py:269-280: if the mirrored source is closer to the target, take its predictions for the
    keypoints that stay mutually visible after the flip."""
    masked = img1_kps * vis.unsqueeze(-1).to(img1_kps.dtype)
    flipped = flip_keypoints(masked, anno_size, permute_indices(permute_list, None))
    vis_flip = (flipped[:, 2] * img2_kps[:, 2] * img1_kps[:, 2] > 0).to(kps_1_to_2.device)
    if float(flip_dist) < float(original_dist):
        kps_1_to_2 = kps_1_to_2.clone()
        kps_1_to_2[vis_flip] = kps_1_to_2_flip[vis_flip]
    return kps_1_to_2


def keypoint_transformation_adaptive_flip(desc1, desc1_flip, desc2, kps1, kps2, permute_list, soft=True, window=5,
                                          anno_size=ANNO_SIZE):
    """pck_train.py:100-126 for ONE pair with ADAPT_FLIP and MUTUAL_NN: desc* [1,P*P,D] un-normalised, kps [K,3].
    -> (kps_1_to_2 [K,2], original_dist, flip_dist).  Both keypoint transformations run as one batch of two pairs and
    both distances as one batch of two dense similarities."""
    P = int(round(desc1.shape[-2] ** 0.5))
    k1, k2 = kps1.detach().cpu(), kps2.detach().cpu()
    vis = k1[:, 2] * k2[:, 2] > 0
    k1_flip = flip_keypoints(k1, anno_size, permute_indices(permute_list, vis))
    idx = np.stack([kpts_to_patch_idx(k1, P, anno_size), kpts_to_patch_idx(k1_flip, P, anno_size)])
    src = torch.cat([desc1.reshape(1, P * P, -1), desc1_flip.reshape(1, P * P, -1)])
    tgt = desc2.reshape(1, P * P, -1).expand(2, -1, -1)
    pred = keypoint_transformation(src, tgt, idx, soft, window, anno_size)
    dist = get_distance_mutual_nn(src, tgt).cpu()
    out = optimized_kps_1_to_2(pred[0], pred[1], k1, k2, dist[1], dist[0], vis, permute_list, anno_size)
    return out, float(dist[0]), float(dist[1])
