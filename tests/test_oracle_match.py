"""CPU: the numpy restatement of the matching path (oracle/match_model.py) against
cv2.BFMatcher itself, driven as StitcherClass.py:423-433 drives it."""
import cv2
import numpy as np
import pytest

from oracle import match_model, stitcher_ref


def cv2_knn(fa, fb):
    raw = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(fa, fb, 2)
    idx = -np.ones((len(fa), 2), np.int32)
    dist = -np.ones((len(fa), 2), np.int32)
    for i, m in enumerate(raw):
        for j, mm in enumerate(m[:2]):
            idx[i, j] = mm.trainIdx
            dist[i, j] = int(mm.distance)
    return idx, dist, raw


def random_desc(rng, n, nbytes=32):
    return rng.integers(0, 256, size=(n, nbytes), dtype=np.uint8)


@pytest.mark.parametrize("nq,nt,nbytes", [(300, 400, 32), (64, 1000, 32), (100, 37, 64), (50, 50, 16)])
def test_knn_top2_equals_cv2(nq, nt, nbytes):
    rng = np.random.default_rng(nq * 1000 + nt)
    fa, fb = random_desc(rng, nq, nbytes), random_desc(rng, nt, nbytes)
    idx, dist = match_model.knn_top2(fa, fb)
    cidx, cdist, _ = cv2_knn(fa, fb)
    assert np.array_equal(idx, cidx)
    assert np.array_equal(dist, cdist)


def test_ties_go_to_the_lower_train_index():
    rng = np.random.default_rng(7)
    fb = random_desc(rng, 200)
    fb[150] = fb[20]          # exact duplicates: distance ties for every query
    fb[151] = fb[20]
    fb[90] = fb[3]
    fa = fb[[20, 3, 77, 150]].copy()
    fa[2, 0] ^= 0x81          # near 77
    idx, dist = match_model.knn_top2(fa, fb)
    cidx, cdist, _ = cv2_knn(fa, fb)
    assert np.array_equal(idx, cidx) and np.array_equal(dist, cdist)
    assert idx[0].tolist() == [20, 150] and dist[0].tolist() == [0, 0]
    assert idx[1].tolist() == [3, 90]
    # low-entropy descriptors: many coincident distances
    fa2 = (random_desc(rng, 120) & 0x03)
    fb2 = (random_desc(rng, 130) & 0x03)
    i2, d2 = match_model.knn_top2(fa2, fb2)
    c2, cd2, _ = cv2_knn(fa2, fb2)
    assert np.array_equal(i2, c2) and np.array_equal(d2, cd2)


def test_small_train_sets():
    rng = np.random.default_rng(11)
    fa = random_desc(rng, 10)
    one = random_desc(rng, 1)
    idx, dist = match_model.knn_top2(fa, one)
    cidx, cdist, raw = cv2_knn(fa, one)
    assert np.array_equal(idx, cidx) and np.array_equal(dist, cdist)
    assert all(len(m) == 1 for m in raw)
    keep, matches = match_model.ratio_test(idx, dist)
    assert not keep.any() and matches == []          # `len(m) == 2` fails (reference :431)
    idx0, dist0 = match_model.knn_top2(fa, np.zeros((0, 32), np.uint8))
    assert (idx0 == -1).all() and (dist0 == -1).all()


@pytest.mark.parametrize("ratio", [0.75, 0.6, 0.9, 1.0])
def test_ratio_loop_equals_reference_loop(ratio):
    rng = np.random.default_rng(5)
    fb = random_desc(rng, 500)
    fa = fb[rng.permutation(500)[:300]].copy()
    flips = rng.integers(0, 256, size=fa.shape, dtype=np.uint8) & rng.integers(0, 256, size=fa.shape, dtype=np.uint8) \
        & rng.integers(0, 256, size=fa.shape, dtype=np.uint8)
    fa ^= flips                                        # ~12 % of the bits flipped
    _, matches_ref, _ = stitcher_ref.match_keypoints(np.zeros((300, 2), np.float32), np.zeros((500, 2), np.float32),
                                                     fa, fb, ratio=ratio)[0:3]
    idx, dist, keep, matches = match_model.match(fa, fb, ratio)
    assert matches == matches_ref
    # integer form used nowhere in the product but worth pinning: 4*d0 < 3*d1 for ratio 0.75
    if ratio == 0.75:
        assert np.array_equal(keep, (idx[:, 1] >= 0) & (4 * dist[:, 0] < 3 * dist[:, 1]))


def test_orb_descriptors_of_a_synthetic_pair():
    from multicamera_stitching_b200 import synthetic
    imageB, imageA, H_true = synthetic.make_pair(360, 640, seed=3)
    orb = cv2.ORB_create(nfeatures=500)
    ka, fa = orb.detectAndCompute(imageA, None)
    kb, fb = orb.detectAndCompute(imageB, None)
    ka = np.float32([k.pt for k in ka])
    kb = np.float32([k.pt for k in kb])
    H, matches_ref, status = stitcher_ref.match_keypoints(ka, kb, fa, fb, 0.75, 3.0)
    _, _, _, matches = match_model.match(fa, fb, 0.75)
    assert matches == matches_ref and len(matches) > 50
    assert match_model.corner_error(H, H_true, 640, 360) < 3.0


def test_four_point_solver_equals_cv2():
    rng = np.random.default_rng(2)
    for _ in range(20):
        a = np.float32([[10, 20], [600, 35], [580, 400], [25, 380]]) + rng.normal(0, 5, (4, 2)).astype(np.float32)
        b = a + rng.normal(0, 20, (4, 2)).astype(np.float32)
        H = match_model.homography_from_4(a, b)
        Hc = cv2.getPerspectiveTransform(a, b)
        assert np.allclose(H, Hc, rtol=1e-6, atol=1e-8)
        assert match_model.reprojection_errors_sq(H, a, b).max() < 1e-12
    assert match_model.homography_from_4([[0, 0], [1, 1], [2, 2], [3, 3]], [[0, 0], [1, 0], [1, 1], [0, 1]]) is None
