"""Recalibration path on the GPU: descriptor matching and RANSAC homography.

Host mirror of ``StitcherBase.matchKeypoints`` (reference
PostScripts/Stitcher/StitcherClass.py:405-448) on top of
``mcs_match_hamming_top2`` and ``mcs_ransac_homography`` (include/mcs.h).
"""
import ctypes

import numpy as np
import torch

from . import _cabi

RANSAC_MAX_ITERS = 2000      # cv2.findHomography default maxIters
RANSAC_SEED = 0x5EED         # fixed seed: the estimate is deterministic call to call, like cv2's
REFINE_ITERS = 10            # cv2 refines the inlier fit with 10 LM iterations


def _device():
    if not torch.cuda.is_available():
        raise _cabi.McsError("no CUDA device: the recalibration path has no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr() if t is not None else 0)


# ---------------------------------------------------------------------------
def match_top2_batch(q, t, nq=None, nt=None, ratio=0.75):
    """Brute-force 2-NN + ratio test for a batch of pairs: Hamming distance for uint8 (binary, e.g. ORB)
    descriptors, L2 for float32 ones (the reference's SIFT branch, StitcherClass.py:380-386, :423-424).

    ``q``: CUDA tensor [B, NQ, D] (query = featuresA), ``t``: [B, NT, D]
    (train = featuresB); ``nq``/``nt``: optional int32 CUDA tensors [B] with the
    real counts.  Returns ``(idx2 [B,NQ,2] int32, dist2 [B,NQ,2] int32 | float32, keep
    [B,NQ] uint8)`` on the device."""
    lib = _cabi.load()
    if not q.is_cuda or not t.is_cuda or q.dtype != t.dtype or q.dtype not in (torch.uint8, torch.float32):
        raise TypeError("descriptors must be uint8 or float32 CUDA tensors of one dtype")
    q = q.contiguous()
    t = t.contiguous()
    B, NQ, D = q.shape
    NT = t.shape[1]
    binary = q.dtype == torch.uint8
    idx2 = torch.empty((B, NQ, 2), dtype=torch.int32, device=q.device)
    dist2 = torch.empty((B, NQ, 2), dtype=torch.int32 if binary else torch.float32, device=q.device)
    keep = torch.empty((B, NQ), dtype=torch.uint8, device=q.device)
    with torch.cuda.device(q.device):
        stream = torch.cuda.current_stream().cuda_stream
        fn = lib.mcs_match_hamming_top2 if binary else lib.mcs_match_l2_top2
        _cabi.check(fn(_ptr(q), _ptr(nq), NQ, _ptr(t), _ptr(nt), NT, D, float(ratio), _ptr(idx2), _ptr(dist2),
                       _ptr(keep), B, ctypes.c_void_p(stream)),
                    "mcs_match_hamming_top2" if binary else "mcs_match_l2_top2")
    return idx2, dist2, keep


def ransac_batch(ptsA, ptsB, samples, reproj_thresh, n=None):
    """Score 4-point hypotheses.  ``ptsA``/``ptsB``: float32 CUDA [B, N, 2];
    ``samples``: int32 CUDA [B, K, 4].  Returns ``(inlier_counts [B,K], H_k
    [B,K,9] float64, best_idx [B], best_mask [B,N] uint8)``."""
    lib = _cabi.load()
    ptsA = ptsA.contiguous().float()
    ptsB = ptsB.contiguous().float()
    samples = samples.contiguous().int()
    B, N, _ = ptsA.shape
    K = samples.shape[1]
    dev = ptsA.device
    counts = torch.empty((B, K), dtype=torch.int32, device=dev)
    H_k = torch.empty((B, K, 9), dtype=torch.float64, device=dev)
    best = torch.empty((B,), dtype=torch.int32, device=dev)
    mask = torch.empty((B, N), dtype=torch.uint8, device=dev)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream().cuda_stream
        _cabi.check(lib.mcs_ransac_homography(_ptr(ptsA), _ptr(ptsB), _ptr(n), N, _ptr(samples), K,
                                              float(reproj_thresh), _ptr(counts), _ptr(H_k), _ptr(best),
                                              _ptr(mask), B, ctypes.c_void_p(stream)),
                    "mcs_ransac_homography")
    return counts, H_k, best, mask


_SAMPLE_CACHE = {}


def draw_samples(n_points, k, seed=RANSAC_SEED):
    """``k`` minimal samples of 4 distinct point indices (host-seeded, so deterministic per ``(n_points, k,
    seed)`` and cached: a rig re-calibrates with about the same number of matches every time)."""
    key = (int(n_points), int(k), int(seed))
    hit = _SAMPLE_CACHE.get(key)
    if hit is None:
        if len(_SAMPLE_CACHE) > 64:
            _SAMPLE_CACHE.clear()
        hit = _SAMPLE_CACHE[key] = _draw_samples(*key)
    return hit


def _draw_samples(n_points, k, seed):
    rng = np.random.default_rng(seed)
    if n_points < 4:
        return np.zeros((k, 4), dtype=np.int32)
    # argsort of uniform noise per row = 4 distinct indices, vectorised
    if n_points <= 64:
        return np.argsort(rng.random((k, n_points)), axis=1)[:, :4].astype(np.int32)
    s = rng.integers(0, n_points, size=(k, 4), dtype=np.int64)
    for _ in range(8):  # redraw the (rare) rows with a repeated index
        srt = np.sort(s, axis=1)
        bad = (srt[:, 1:] == srt[:, :-1]).any(axis=1)
        if not bad.any():
            break
        s[bad] = rng.integers(0, n_points, size=(int(bad.sum()), 4), dtype=np.int64)
    return s.astype(np.int32)


# ---------------------------------------------------------------------------
def _refit(ptsA, ptsB, H0, status):
    """Host end of ``cv2.findHomography``: least-squares fit on the winner's inliers + LM refinement, in the
    library (``mcs_refit_homography``: normalised DLT + ``REFINE_ITERS`` Levenberg-Marquardt iterations)."""
    return _cabi.refit_homography(ptsA, ptsB, status, H0, REFINE_ITERS)


def find_homography_ransac_batch(pairs, reproj_thresh, max_iters=RANSAC_MAX_ITERS, seed=RANSAC_SEED):
    """``cv2.findHomography(ptsA, ptsB, cv2.RANSAC, reproj_thresh)`` for several point-set pairs at once: ONE
    hypothesis-scoring launch over all pairs (padded to the longest), one device->host copy, then the refit of
    every pair on the host.  ``pairs``: list of ``(ptsA, ptsB)`` float32 N_i x 2.  Returns a list of
    ``(H 3x3 float64 | None, status N_i x 1 uint8 | None)``."""
    dev = _device()
    pts = [(np.ascontiguousarray(a, dtype=np.float32).reshape(-1, 2), np.ascontiguousarray(b, dtype=np.float32).reshape(-1, 2))
           for a, b in pairs]
    out = [(None, None)] * len(pts)
    todo = [i for i, (a, _) in enumerate(pts) if len(a) >= 4]
    if not todo:
        return out
    n_max = max(len(pts[i][0]) for i in todo)
    B = len(todo)
    hA = np.zeros((B, n_max, 2), np.float32)
    hB = np.zeros((B, n_max, 2), np.float32)
    hS = np.zeros((B, max_iters, 4), np.int32)
    hN = np.zeros((B,), np.int32)
    for j, i in enumerate(todo):
        a, b = pts[i]
        hA[j, :len(a)] = a
        hB[j, :len(b)] = b
        hN[j] = len(a)
        hS[j] = draw_samples(len(a), max_iters, seed)
    a_d, b_d = torch.from_numpy(hA).to(dev), torch.from_numpy(hB).to(dev)
    counts, H_k, best, mask = ransac_batch(a_d, b_d, torch.from_numpy(hS).to(dev), reproj_thresh,
                                           n=torch.from_numpy(hN).to(dev))
    best_h = best.cpu().numpy()                      # one synchronisation for the whole batch
    mask_h = mask.cpu().numpy()
    H_best = H_k[torch.arange(B, device=dev), best.clamp(min=0).long()].cpu().numpy()
    for j, i in enumerate(todo):
        if best_h[j] < 0:
            continue
        a, b = pts[i]
        status = mask_h[j, :len(a)].reshape(-1, 1).copy()
        out[i] = (_refit(a, b, H_best[j].reshape(3, 3), status), status)
    return out


def find_homography_ransac(ptsA, ptsB, reproj_thresh, max_iters=RANSAC_MAX_ITERS, seed=RANSAC_SEED):
    """GPU RANSAC (hypothesis scoring) + host refit on the winner's inliers.
    Returns ``(H 3x3 float64 | None, status N x 1 uint8 | None)`` like
    ``cv2.findHomography(ptsA, ptsB, cv2.RANSAC, reproj_thresh)``."""
    return find_homography_ransac_batch([(ptsA, ptsB)], reproj_thresh, max_iters, seed)[0]


def match_keypoints_batch(items, ratio=0.75, reprojThresh=4.0):
    """``StitcherBase.matchKeypoints`` (reference :405-448) for several image pairs at once - BASELINE config 4's
    "4 x 1080p pairs": ONE matching launch over all pairs (descriptor sets padded to the longest), one
    device->host copy of the survivors, ONE RANSAC scoring launch, host refits.

    ``items``: list of ``(kpsA, kpsB, featuresA, featuresB)``, all features of one dtype (uint8 -> Hamming,
    float32 -> L2).  Returns a list of ``(H, matches, status)``."""
    dev = _device()
    feats = [(np.ascontiguousarray(fa), np.ascontiguousarray(fb)) for (_, _, fa, fb) in items]
    kinds = {f.dtype for pair in feats for f in pair}
    if len(kinds) != 1 or next(iter(kinds)) not in (np.dtype(np.uint8), np.dtype(np.float32)):
        raise TypeError("descriptors must all be uint8 (Hamming) or all float32 (L2), got %s" % sorted(map(str, kinds)))
    dt = next(iter(kinds))
    B = len(items)
    D = feats[0][0].shape[1]
    nq_max = max(len(fa) for fa, _ in feats)
    nt_max = max(len(fb) for _, fb in feats)
    hq = np.zeros((B, nq_max, D), dt)
    ht = np.zeros((B, max(nt_max, 1), D), dt)
    nq = np.zeros((B,), np.int32)
    nt = np.zeros((B,), np.int32)
    for j, (fa, fb) in enumerate(feats):
        hq[j, :len(fa)] = fa
        ht[j, :len(fb)] = fb
        nq[j], nt[j] = len(fa), len(fb)
    idx2, _dist2, keep = match_top2_batch(torch.from_numpy(hq).to(dev), torch.from_numpy(ht).to(dev),
                                          torch.from_numpy(nq).to(dev), torch.from_numpy(nt).to(dev), ratio=ratio)
    packed = torch.where(keep.bool(), idx2[:, :, 0], torch.full_like(idx2[:, :, 0], -1)).cpu().numpy()   # one copy
    all_matches, point_pairs = [], []
    for j, (kpsA, kpsB, _, _) in enumerate(items):
        query = np.nonzero(packed[j, :nq[j]] >= 0)[0]
        train = packed[j, query]
        matches = list(zip(train.tolist(), query.tolist()))     # (trainIdx, queryIdx), reference :428-433
        all_matches.append(matches)
        if len(matches) > 4:
            point_pairs.append((np.asarray(kpsA, dtype=np.float32)[query], np.asarray(kpsB, dtype=np.float32)[train]))
        else:
            point_pairs.append((np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32)))
    fits = find_homography_ransac_batch(point_pairs, reprojThresh)
    return [(fits[j][0], all_matches[j], fits[j][1]) for j in range(B)]


def match_keypoints(kpsA, kpsB, featuresA, featuresB, ratio=0.75, reprojThresh=4.0):
    """``StitcherBase.matchKeypoints`` (reference :405-448): returns ``(H, matches, status)``.  Binary (uint8)
    descriptors are matched by Hamming distance, float32 ones (the reference's own SIFT branch) by L2."""
    return match_keypoints_batch([(kpsA, kpsB, featuresA, featuresB)], ratio, reprojThresh)[0]
