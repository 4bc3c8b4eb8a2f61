"""GPU: the recalibration kernels (Hamming top-2 matcher, RANSAC scoring) through the C ABI
against cv2 / the numpy oracle.  Match indices and the ratio-test survivors must be
bit-exact; the homography is compared through reprojection, as SURVEY.md section 8(c) states."""
import cv2
import numpy as np
import pytest
import torch

from multicamera_stitching_b200 import StitcherBase, Stitcher, recalib, synthetic
from oracle import match_model, stitcher_ref

pytestmark = pytest.mark.gpu


def random_desc(rng, n, nbytes=32):
    return rng.integers(0, 256, size=(n, nbytes), dtype=np.uint8)


def gpu_match(fa_list, fb_list, ratio=0.75, device="cuda"):
    """Batch of pairs with ragged counts -> per-pair (idx, dist, keep) numpy arrays."""
    B = len(fa_list)
    nbytes = fa_list[0].shape[1]
    nq_max = max(1, max(len(f) for f in fa_list))
    nt_max = max(1, max(len(f) for f in fb_list))
    q = torch.zeros((B, nq_max, nbytes), dtype=torch.uint8)
    t = torch.zeros((B, nt_max, nbytes), dtype=torch.uint8)
    for i in range(B):
        q[i, :len(fa_list[i])] = torch.from_numpy(fa_list[i])
        t[i, :len(fb_list[i])] = torch.from_numpy(fb_list[i])
    nq = torch.tensor([len(f) for f in fa_list], dtype=torch.int32, device=device)
    nt = torch.tensor([len(f) for f in fb_list], dtype=torch.int32, device=device)
    idx2, dist2, keep = recalib.match_top2_batch(q.to(device), t.to(device), nq, nt, ratio)
    torch.cuda.synchronize()
    out = []
    for i in range(B):
        n = len(fa_list[i])
        out.append((idx2[i, :n].cpu().numpy(), dist2[i, :n].cpu().numpy(), keep[i, :n].cpu().numpy().astype(bool)))
    return out


@pytest.mark.parametrize("nq,nt,nbytes", [(2000, 2000, 32), (333, 4100, 32), (100, 37, 64), (50, 50, 16), (5, 1, 32)])
def test_matcher_equals_bfmatcher(cuda_device, nq, nt, nbytes):
    rng = np.random.default_rng(nq + 7 * nt)
    fb = random_desc(rng, nt, nbytes)
    fa = random_desc(rng, nq, nbytes)
    n_copy = min(nq, nt) // 2
    fa[:n_copy] = fb[rng.permutation(nt)[:n_copy]]           # half the queries have a true partner ...
    fa[:n_copy, :4] ^= random_desc(rng, n_copy, 4) & 0x11    # ... a few bits away
    (idx, dist, keep), = gpu_match([fa], [fb])
    ridx, rdist, rkeep, _ = match_model.match(fa, fb, 0.75)
    assert np.array_equal(idx, ridx)
    assert np.array_equal(dist, rdist)
    assert np.array_equal(keep, rkeep)
    # and against cv2 directly, exactly as the reference's loop reads it
    raw = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(fa, fb, 2)
    ref = [(m[0].trainIdx, m[0].queryIdx) for m in raw if len(m) == 2 and m[0].distance < m[1].distance * 0.75]
    got = [(int(idx[i, 0]), i) for i in np.nonzero(keep)[0]]
    assert got == ref


def test_matcher_ties_and_low_entropy(cuda_device):
    rng = np.random.default_rng(3)
    fb = random_desc(rng, 300)
    fb[250] = fb[10]
    fb[251] = fb[10]
    fb[299] = fb[0]
    fa = np.concatenate([fb[[10, 0, 299, 250]], random_desc(rng, 60) & 0x03])
    fb2 = np.concatenate([fb, random_desc(rng, 100) & 0x03])
    (idx, dist, keep), = gpu_match([fa], [fb2])
    ridx, rdist, rkeep, _ = match_model.match(fa, fb2)
    assert np.array_equal(idx, ridx) and np.array_equal(dist, rdist) and np.array_equal(keep, rkeep)
    assert idx[0].tolist() == [10, 250] and idx[1].tolist() == [0, 299]


def test_matcher_ragged_batch_and_empty_sets(cuda_device):
    rng = np.random.default_rng(9)
    sizes = [(700, 900), (1, 5), (0, 40), (64, 0), (300, 1), (2000, 1500)]
    fa_list = [random_desc(rng, a) for a, _ in sizes]
    fb_list = [random_desc(rng, b) for _, b in sizes]
    res = gpu_match(fa_list, fb_list, ratio=0.8)
    for (idx, dist, keep), fa, fb in zip(res, fa_list, fb_list):
        ridx, rdist, rkeep, _ = match_model.match(fa, fb, 0.8)
        assert np.array_equal(idx, ridx) and np.array_equal(dist, rdist) and np.array_equal(keep, rkeep)


@pytest.mark.parametrize("ratio", [0.5, 0.75, 0.95, 1.0])
def test_matcher_ratio_values(cuda_device, ratio):
    rng = np.random.default_rng(21)
    fb = random_desc(rng, 800)
    fa = fb[rng.permutation(800)[:500]].copy()
    fa ^= random_desc(rng, 500) & random_desc(rng, 500) & random_desc(rng, 500)
    (idx, dist, keep), = gpu_match([fa], [fb], ratio)
    _, _, rkeep, _ = match_model.match(fa, fb, ratio)
    assert np.array_equal(keep, rkeep)


# ---------------------------------------------------------------------------
def synthetic_correspondences(n, outlier_frac, seed, w=1920, h=1080, sigma=0.5):
    rng = np.random.default_rng(seed)
    H = np.array([[0.97, 0.02, 740.3], [-0.012, 0.985, 8.1], [1.2e-5, -6e-6, 1.0]])
    a = np.stack([rng.uniform(0, w, n), rng.uniform(0, h, n)], axis=1)
    p = np.concatenate([a, np.ones((n, 1))], axis=1) @ H.T
    b = p[:, :2] / p[:, 2:3] + rng.normal(0, sigma, (n, 2))
    n_out = int(outlier_frac * n)
    out = rng.permutation(n)[:n_out]
    b[out] = np.stack([rng.uniform(0, 2 * w, n_out), rng.uniform(0, h, n_out)], axis=1)
    inl = np.ones(n, bool)
    inl[out] = False
    return a.astype(np.float32), b.astype(np.float32), H, inl


def test_ransac_hypotheses_against_the_four_point_oracle(cuda_device):
    a, b, H_true, _ = synthetic_correspondences(600, 0.3, seed=1)
    k = 256
    samples = recalib.draw_samples(len(a), k, seed=123)
    samples[5] = [0, 0, 1, 2]                                   # repeated index: degenerate
    ta = torch.from_numpy(a).cuda()[None]
    tb = torch.from_numpy(b).cuda()[None]
    ts = torch.from_numpy(samples).cuda()[None]
    counts, H_k, best, mask = recalib.ransac_batch(ta, tb, ts, 3.0)
    torch.cuda.synchronize()
    counts = counts[0].cpu().numpy()
    H_k = H_k[0].cpu().numpy().reshape(k, 3, 3)
    assert counts[5] == -1
    checked = 0
    for i in range(k):
        a4, b4 = a[samples[i]], b[samples[i]]
        if len(set(samples[i].tolist())) == 4:
            assert (counts[i] >= 0) == match_model.sample_is_valid(a4, b4), i   # OpenCV's checkSubset rule
        if counts[i] < 0:
            continue
        Hi = match_model.homography_from_4(a4, b4)
        assert Hi is not None
        # same map: compare where the frame corners land
        if abs(np.linalg.det(Hi)) < 1e-6:
            continue
        assert match_model.corner_error(H_k[i], Hi, 1920, 1080) < 1e-3 * max(1.0, np.abs(Hi).max()), i
        err = match_model.reprojection_errors_sq(H_k[i].astype(np.float64), a, b)
        sure_in = int((err <= 9.0 * (1 - 1e-3)).sum())
        sure_max = int((err <= 9.0 * (1 + 1e-3)).sum())
        assert sure_in <= counts[i] <= sure_max, (i, counts[i], sure_in, sure_max)
        checked += 1
    assert checked > 50
    # winner = arg max count, lowest index on ties; its mask is the inlier set of that hypothesis
    valid = np.where(counts >= 0, counts, -1)
    assert int(best[0]) == int(np.argmax(valid))
    m = mask[0].cpu().numpy().astype(bool)
    assert m.sum() == counts[int(best[0])]


# tolerance in pixels at the frame corners: 0.5 px at the match counts of BASELINE.json's config 4
# (SURVEY.md section 8c), wider where a dozen noisy (sigma 0.5 px) points cannot do better - cv2's
# own estimate is that far from the planted homography there
@pytest.mark.parametrize("n,outliers,seed,tol", [(800, 0.33, 0, 0.5), (2000, 0.5, 1, 0.5), (60, 0.2, 2, 1.5),
                                                 (12, 0.0, 3, 2.5)])
def test_find_homography_agrees_with_cv2(cuda_device, n, outliers, seed, tol):
    a, b, H_true, inl_true = synthetic_correspondences(n, outliers, seed)
    H, status = recalib.find_homography_ransac(a, b, 3.0)
    Hc, sc = cv2.findHomography(a, b, cv2.RANSAC, 3.0)
    assert H is not None and status.shape == (n, 1) and status.dtype == np.uint8
    # both land within half a pixel of each other and of the ground truth at the frame corners
    assert match_model.corner_error(H, Hc, 1920, 1080) < tol
    assert match_model.corner_error(H, H_true, 1920, 1080) < tol
    agree = (status.ravel().astype(bool) == sc.ravel().astype(bool)).mean()
    assert agree >= 0.97, agree
    assert (status.ravel().astype(bool) & ~inl_true).sum() <= max(2, 0.02 * n)   # almost no outlier accepted


def test_find_homography_degenerate_inputs(cuda_device):
    a = np.float32([[0, 0], [1, 1], [2, 2], [3, 3], [4, 4], [5, 5]])
    H, status = recalib.find_homography_ransac(a, a + 1, 3.0)
    assert H is None and status is None                      # every sample collinear
    H, status = recalib.find_homography_ransac(a[:3], a[:3], 3.0)
    assert H is None and status is None                      # fewer than 4 points


def test_match_keypoints_signature_and_result(cuda_device):
    imageB, imageA, H_true = synthetic.make_pair(540, 960, seed=5)
    sb = StitcherBase()
    kpsA, fa = sb.detectAndDescribe(imageA)
    kpsB, fb = sb.detectAndDescribe(imageB)
    H, matches, status = sb.matchKeypoints(kpsA, kpsB, fa, fb, ratio=0.75, reprojThresh=3.0)
    Hr, matches_ref, status_ref = stitcher_ref.match_keypoints(kpsA, kpsB, fa, fb, 0.75, 3.0)
    assert matches == matches_ref                            # (trainIdx, queryIdx), bit-exact
    assert status.shape == status_ref.shape
    assert match_model.corner_error(H, Hr, 960, 540) < 1.0
    assert match_model.corner_error(H, H_true, 960, 540) < 2.5


def test_calibrate_then_stitch_matches_the_cv2_chain_on_the_same_state(cuda_device):
    imageB, imageA, H_true = synthetic.make_pair(360, 640, seed=8)
    images = {"CAM1": imageB, "CAM2": imageA}
    st = Stitcher(images)
    st.calibrate_stitcher(images, save=False)
    sb = st.stitchers[0]
    assert sb.cachedAH is not None
    # the recovered homography (before the canvas translation) is the planted one
    T = np.array([[1, 0, -sb.Bpts[0][0]], [0, 1, -sb.Bpts[0][1]], [0, 0, 1.0]])
    assert match_model.corner_error(T @ sb.cachedAH, H_true, 640, 360) < 2.5
    got = st.stitch(images)
    state = stitcher_ref.new_state()
    for key in state:
        if hasattr(sb, key):
            state[key] = getattr(sb, key)
    ref = stitcher_ref.stitch_chain([state], list(st.img_labels), images)
    assert got.shape == ref.shape and np.array_equal(got, ref)


# ---------------------------------------------------------------------------
# float descriptors / L2: the reference's own matcher branch (StitcherClass.py:380-386, :423-433)
def sift_like(rng, n, dim=128):
    """Integer-valued float32 descriptors in 0..255, the form OpenCV's SIFT emits."""
    return np.minimum(255, rng.gamma(0.6, 40.0, size=(n, dim))).astype(np.uint8).astype(np.float32)


def bf_l2(fa, fb, ratio):
    raw = cv2.BFMatcher(cv2.NORM_L2).knnMatch(fa, fb, 2)
    idx = -np.ones((len(fa), 2), np.int32)
    dist = -np.ones((len(fa), 2), np.float32)
    for i, m in enumerate(raw):
        for j, mm in enumerate(m[:2]):
            idx[i, j], dist[i, j] = mm.trainIdx, mm.distance
    keep = np.array([len(m) == 2 and m[0].distance < m[1].distance * ratio for m in raw], dtype=bool)
    return idx, dist, keep


@pytest.mark.parametrize("nq,nt,dim", [(2000, 2000, 128), (333, 1100, 128), (100, 37, 64), (5, 1, 128), (40, 300, 36)])
def test_l2_matcher_equals_bfmatcher_on_sift_like_descriptors(cuda_device, nq, nt, dim):
    rng = np.random.default_rng(nq + 3 * nt)
    fb = sift_like(rng, nt, dim)
    fa = sift_like(rng, nq, dim)
    n_copy = min(nq, nt) // 2
    fa[:n_copy] = np.clip(fb[rng.permutation(nt)[:n_copy]] + rng.integers(-6, 7, (n_copy, dim)), 0, 255).astype(np.float32)
    if nt >= 3 and nq >= 2:            # exact ties: equal distance, the lower train index comes first
        fb[nt - 1] = fb[0]
        fa[nq - 1] = fb[0]
    q = torch.from_numpy(fa).to(cuda_device)[None]
    t = torch.from_numpy(fb).to(cuda_device)[None]
    idx2, dist2, keep = recalib.match_top2_batch(q, t, ratio=0.75)
    torch.cuda.synchronize()
    ridx, rdist, rkeep = bf_l2(fa, fb, 0.75)
    assert np.array_equal(idx2[0].cpu().numpy(), ridx)
    assert np.array_equal(dist2[0].cpu().numpy(), rdist)      # sums of squares are exact in float32: bit-identical
    assert np.array_equal(keep[0].cpu().numpy().astype(bool), rkeep)


def test_l2_matcher_general_floats_and_ragged_batch(cuda_device):
    """Arbitrary float descriptors: the float32 sum may round differently from OpenCV's SIMD order in the last
    bit, so distances are compared to 1e-5 relative and indices wherever the two best are not that close."""
    rng = np.random.default_rng(5)
    sizes = [(400, 500), (1, 3), (0, 10), (30, 0), (257, 129)]
    fa_list = [rng.normal(0, 1, (a, 128)).astype(np.float32) for a, _ in sizes]
    fb_list = [rng.normal(0, 1, (b, 128)).astype(np.float32) for _, b in sizes]
    nq_max, nt_max = max(a for a, _ in sizes), max(b for _, b in sizes)
    q = torch.zeros((len(sizes), nq_max, 128))
    t = torch.zeros((len(sizes), nt_max, 128))
    for i, (fa, fb) in enumerate(zip(fa_list, fb_list)):
        q[i, :len(fa)] = torch.from_numpy(fa)
        t[i, :len(fb)] = torch.from_numpy(fb)
    nq = torch.tensor([a for a, _ in sizes], dtype=torch.int32, device=cuda_device)
    nt = torch.tensor([b for _, b in sizes], dtype=torch.int32, device=cuda_device)
    idx2, dist2, keep = recalib.match_top2_batch(q.to(cuda_device), t.to(cuda_device), nq, nt, ratio=0.9)
    torch.cuda.synchronize()
    for i, (fa, fb) in enumerate(zip(fa_list, fb_list)):
        n = len(fa)
        gi, gd = idx2[i, :n].cpu().numpy(), dist2[i, :n].cpu().numpy()
        if n == 0:
            continue
        if len(fb) == 0:
            assert (gi == -1).all()
            continue
        ridx, rdist, _ = bf_l2(fa, fb, 0.9)
        assert np.allclose(gd, rdist, rtol=1e-5, atol=0)
        clear = np.ones(n, bool) if len(fb) < 3 else (rdist[:, 1] - rdist[:, 0]) > 1e-4 * rdist[:, 1]
        assert np.array_equal(gi[clear, 0], ridx[clear, 0])


def test_match_keypoints_float_descriptors_equal_the_cv2_loop(cuda_device):
    """``StitcherBase.matchKeypoints`` on float (SIFT-like) descriptors - the branch the reference itself runs:
    the match list equals cv2's loop, the homography agrees through reprojection."""
    rng = np.random.default_rng(12)
    nA, nB = 600, 700
    fb = sift_like(rng, nB)
    perm = rng.permutation(nB)[:nA]
    fa = np.clip(fb[perm] + rng.integers(-5, 6, (nA, 128)), 0, 255).astype(np.float32)
    fa[450:] = sift_like(rng, 150)
    H_true = np.array([[0.97, 0.02, 310.0], [-0.015, 1.01, 7.5], [1e-5, -2e-5, 1.0]])
    kpsA = rng.uniform(0, 1900, (nA, 2)).astype(np.float32)
    proj = np.c_[kpsA, np.ones(nA)] @ H_true.T
    kpsB = rng.uniform(0, 1900, (nB, 2)).astype(np.float32)
    kpsB[perm[:450]] = (proj[:450, :2] / proj[:450, 2:]).astype(np.float32) + rng.normal(0, 0.3, (450, 2)).astype(np.float32)
    H, matches, status = StitcherBase().matchKeypoints(kpsA, kpsB, fa, fb, ratio=0.75, reprojThresh=3.0)
    Hc, mc, sc = stitcher_ref.match_keypoints(kpsA, kpsB, fa, fb, ratio=0.75, reprojThresh=3.0)
    assert matches == mc and len(matches) > 400
    corners = np.float64([[0, 0, 1], [1920, 0, 1], [1920, 1080, 1], [0, 1080, 1]])
    pa, pb = corners @ H.T, corners @ Hc.T
    assert np.abs(pa[:, :2] / pa[:, 2:] - pb[:, :2] / pb[:, 2:]).max() < 0.5
    assert (status.ravel() == sc.ravel()).mean() > 0.98


def test_match_keypoints_batch_equals_single_calls(cuda_device):
    """The batched form (config 4's four pairs in one matching and one RANSAC launch) returns, pair by pair,
    what the single-pair calls return."""
    items = []
    for k in range(4):
        imageB, imageA, _ = synthetic.make_pair(270, 480, seed=k)
        sb = StitcherBase()
        sb.nfeatures = 500
        kA, fA = sb.detectAndDescribe(imageA)
        kB, fB = sb.detectAndDescribe(imageB)
        items.append((kA, kB, fA, fB))
    batch = recalib.match_keypoints_batch(items, ratio=0.75, reprojThresh=3.0)
    for item, (H, matches, status) in zip(items, batch):
        H1, m1, s1 = recalib.match_keypoints(*item, ratio=0.75, reprojThresh=3.0)
        assert matches == m1
        assert (H is None) == (H1 is None)
        if H is not None:
            assert np.allclose(H, H1) and np.array_equal(status, s1)
