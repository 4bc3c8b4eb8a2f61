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
    """Brute-force Hamming 2-NN + ratio test for a batch of pairs.

    ``q``: uint8 CUDA tensor [B, NQ, D] (query = featuresA), ``t``: [B, NT, D]
    (train = featuresB); ``nq``/``nt``: optional int32 CUDA tensors [B] with the
    real counts.  Returns ``(idx2 [B,NQ,2] int32, dist2 [B,NQ,2] int32, keep
    [B,NQ] uint8)`` on the device."""
    lib = _cabi.load()
    if q.dtype != torch.uint8 or t.dtype != torch.uint8 or not q.is_cuda or not t.is_cuda:
        raise TypeError("descriptors must be uint8 CUDA tensors")
    q = q.contiguous()
    t = t.contiguous()
    B, NQ, D = q.shape
    NT = t.shape[1]
    idx2 = torch.empty((B, NQ, 2), dtype=torch.int32, device=q.device)
    dist2 = torch.empty((B, NQ, 2), dtype=torch.int32, device=q.device)
    keep = torch.empty((B, NQ), dtype=torch.uint8, device=q.device)
    with torch.cuda.device(q.device):
        stream = torch.cuda.current_stream().cuda_stream
        _cabi.check(lib.mcs_match_hamming_top2(_ptr(q), _ptr(nq), NQ, _ptr(t), _ptr(nt), NT, D,
                                               float(ratio), _ptr(idx2), _ptr(dist2), _ptr(keep),
                                               B, ctypes.c_void_p(stream)),
                    "mcs_match_hamming_top2")
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


def draw_samples(n_points, k, seed=RANSAC_SEED):
    """``k`` minimal samples of 4 distinct point indices (host-seeded)."""
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
def _normalise(p):
    c = p.mean(axis=0)
    d = np.abs(p - c).mean(axis=0)
    d = np.where(d > 1e-12, d, 1.0)
    s = 1.0 / d
    T = np.array([[s[0], 0, -c[0] * s[0]], [0, s[1], -c[1] * s[1]], [0, 0, 1.0]])
    return (p - c) * s, T


def fit_homography_dlt(a, b):
    """Least-squares homography a -> b (normalised DLT, the structure of
    OpenCV's HomographyEstimatorCallback::runKernel)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    an, Ta = _normalise(a)
    bn, Tb = _normalise(b)
    n = len(a)
    A = np.zeros((2 * n, 9))
    A[0::2, 0:2] = an
    A[0::2, 2] = 1
    A[0::2, 6:8] = -bn[:, 0:1] * an
    A[0::2, 8] = -bn[:, 0]
    A[1::2, 3:5] = an
    A[1::2, 5] = 1
    A[1::2, 6:8] = -bn[:, 1:2] * an
    A[1::2, 8] = -bn[:, 1]
    _, _, vt = np.linalg.svd(A.T @ A)
    Hn = vt[-1].reshape(3, 3)
    H = np.linalg.inv(Tb) @ Hn @ Ta
    return H / H[2, 2]


def refine_homography(H, a, b, iters=REFINE_ITERS):
    """Levenberg-Marquardt on the reprojection error over 8 parameters, the
    role of OpenCV's HomographyRefineCallback."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    h = (H / H[2, 2]).ravel()[:8].copy()

    def residual(h8):
        w = h8[6] * a[:, 0] + h8[7] * a[:, 1] + 1.0
        x = (h8[0] * a[:, 0] + h8[1] * a[:, 1] + h8[2]) / w
        y = (h8[3] * a[:, 0] + h8[4] * a[:, 1] + h8[5]) / w
        return x, y, w

    lam = 1e-3
    x, y, w = residual(h)
    err = np.concatenate([x - b[:, 0], y - b[:, 1]])
    cost = float(err @ err)
    for _ in range(iters):
        n = len(a)
        J = np.zeros((2 * n, 8))
        iw = 1.0 / w
        J[:n, 0] = a[:, 0] * iw
        J[:n, 1] = a[:, 1] * iw
        J[:n, 2] = iw
        J[:n, 6] = -a[:, 0] * x * iw
        J[:n, 7] = -a[:, 1] * x * iw
        J[n:, 3] = a[:, 0] * iw
        J[n:, 4] = a[:, 1] * iw
        J[n:, 5] = iw
        J[n:, 6] = -a[:, 0] * y * iw
        J[n:, 7] = -a[:, 1] * y * iw
        JtJ = J.T @ J
        g = J.T @ err
        improved = False
        for _try in range(6):
            try:
                step = np.linalg.solve(JtJ + lam * np.diag(np.diag(JtJ)), -g)
            except np.linalg.LinAlgError:
                lam *= 10
                continue
            x2, y2, w2 = residual(h + step)
            err2 = np.concatenate([x2 - b[:, 0], y2 - b[:, 1]])
            cost2 = float(err2 @ err2)
            if cost2 < cost:
                h, x, y, w, err, cost = h + step, x2, y2, w2, err2, cost2
                lam = max(lam * 0.1, 1e-12)
                improved = True
                break
            lam *= 10
        if not improved:
            break
    return np.append(h, 1.0).reshape(3, 3)


def find_homography_ransac(ptsA, ptsB, reproj_thresh, max_iters=RANSAC_MAX_ITERS, seed=RANSAC_SEED):
    """GPU RANSAC (hypothesis scoring) + host refit on the winner's inliers.
    Returns ``(H 3x3 float64 | None, status N x 1 uint8 | None)`` like
    ``cv2.findHomography(ptsA, ptsB, cv2.RANSAC, reproj_thresh)``."""
    dev = _device()
    ptsA = np.ascontiguousarray(ptsA, dtype=np.float32).reshape(-1, 2)
    ptsB = np.ascontiguousarray(ptsB, dtype=np.float32).reshape(-1, 2)
    n = len(ptsA)
    if n < 4:
        return None, None
    samples = draw_samples(n, max_iters, seed)
    a = torch.from_numpy(ptsA).to(dev)[None]
    b = torch.from_numpy(ptsB).to(dev)[None]
    s = torch.from_numpy(samples).to(dev)[None]
    counts, H_k, best, mask = ransac_batch(a, b, s, reproj_thresh)
    best_i = int(best[0].item())
    if best_i < 0:
        return None, None
    status = mask[0].cpu().numpy().reshape(-1, 1)
    inl = status.ravel().astype(bool)
    H = H_k[0, best_i].cpu().numpy().reshape(3, 3)
    if inl.sum() >= 4:
        if inl.sum() > 4:
            H = fit_homography_dlt(ptsA[inl], ptsB[inl])
        H = refine_homography(H, ptsA[inl], ptsB[inl])
    return H, status


def match_keypoints(kpsA, kpsB, featuresA, featuresB, ratio=0.75, reprojThresh=4.0):
    """``StitcherBase.matchKeypoints`` (reference :405-448): returns
    ``(H, matches, status)``."""
    featuresA = np.ascontiguousarray(featuresA)
    featuresB = np.ascontiguousarray(featuresB)
    if featuresA.dtype != np.uint8 or featuresB.dtype != np.uint8:
        raise NotImplementedError(
            "only binary (uint8, e.g. ORB) descriptors are matched on the GPU; float descriptors "
            "(the reference's SIFT/L2 branch) are outside BASELINE.json's recalibration workload")
    dev = _device()
    q = torch.from_numpy(featuresA).to(dev)[None]
    t = torch.from_numpy(featuresB).to(dev)[None]
    idx2, _dist2, keep = match_top2_batch(q, t, ratio=ratio)
    keep_h = keep[0].cpu().numpy().astype(bool)
    idx_h = idx2[0, :, 0].cpu().numpy()
    query = np.nonzero(keep_h)[0]
    matches = [(int(idx_h[i]), int(i)) for i in query]
    H = None
    status = None
    if len(matches) > 4:
        ptsA = np.float32([kpsA[i] for (_, i) in matches])
        ptsB = np.float32([kpsB[i] for (i, _) in matches])
        H, status = find_homography_ransac(ptsA, ptsB, reprojThresh)
    return H, matches, status
