"""The host end of ``cv2.findHomography(..., RANSAC)`` (StitcherClass.py:443-444): the library's inlier refit
(``mcs_refit_homography``: normalised DLT + Levenberg-Marquardt) against its numpy model (oracle/refit_model.py)
and against OpenCV's own least-squares ``findHomography`` on the same points.  No GPU involved.

Tolerances: library vs model 1e-9 relative on H (same arithmetic, different summation order of the 9 x 9 / 8 x 8
systems); library vs cv2 0.02 px on the reprojected frame corners (cv2 differs in its normalisation and LM
details; both are least-squares fits of the same inliers)."""
import cv2
import numpy as np
import pytest

from multicamera_stitching_b200 import _cabi
from oracle import refit_model

CORNERS = np.array([[0, 0], [1919, 0], [1919, 1079], [0, 1079]], dtype=np.float64)


def _project(H, p):
    q = np.c_[p, np.ones(len(p))] @ np.asarray(H).T
    return q[:, :2] / q[:, 2:]


def _case(seed, n=600, noise=0.5, outliers=0.3):
    rng = np.random.default_rng(seed)
    H = np.array([[1 + rng.normal(0, 0.03), rng.normal(0, 0.03), rng.uniform(-400, 400)],
                  [rng.normal(0, 0.03), 1 + rng.normal(0, 0.03), rng.uniform(-60, 60)],
                  [rng.normal(0, 2e-5), rng.normal(0, 2e-5), 1.0]])
    a = rng.uniform(0, [1920, 1080], (n, 2)).astype(np.float32)
    b = (_project(H, a) + rng.normal(0, noise, (n, 2))).astype(np.float32)
    mask = (rng.random(n) >= outliers).astype(np.uint8)
    bad = mask == 0
    b[bad] += rng.uniform(-300, 300, (int(bad.sum()), 2)).astype(np.float32)
    H0 = H * (1 + rng.normal(0, 1e-3, (3, 3)))
    H0 /= H0[2, 2]
    return H, H0, a, b, mask


@pytest.mark.parametrize("seed", range(12))
def test_refit_equals_model_and_agrees_with_cv2(seed):
    H, H0, a, b, mask = _case(seed, n=[5, 6, 9, 40, 600, 2000][seed % 6] + 8)
    got = _cabi.refit_homography(a, b, mask, H0)
    want = refit_model.refit(a, b, H0, mask)
    assert got[2, 2] == 1.0
    assert np.allclose(got, want, rtol=1e-9, atol=1e-12), np.abs(got - want).max()
    inl = mask.astype(bool)
    if inl.sum() >= 12:
        Hcv, _ = cv2.findHomography(a[inl], b[inl], 0)
        assert np.abs(_project(got, CORNERS) - _project(Hcv, CORNERS)).max() < 0.02


def test_refit_noise_free_points_recover_the_homography():
    H, H0, a, b, mask = _case(3, n=300, noise=0.0, outliers=0.0)
    b = _project(H, a.astype(np.float64)).astype(np.float32)
    got = _cabi.refit_homography(a, b, None, H0)
    assert np.abs(_project(got, CORNERS) - _project(H, CORNERS)).max() < 2e-3      # float32 points


def test_refit_small_inlier_sets():
    H, H0, a, b, mask = _case(5, n=50)
    few = np.zeros(50, np.uint8)
    few[:3] = 1
    assert np.array_equal(_cabi.refit_homography(a, b, few, H0), H0)                # < 4 inliers: the hypothesis itself
    few[3] = 1                                                                      # exactly 4: LM from the hypothesis
    got, want = _cabi.refit_homography(a, b, few, H0), refit_model.refit(a, b, H0, few)
    assert np.allclose(got, want, rtol=1e-7, atol=1e-10)
    assert np.abs(_project(got, a[:4].astype(np.float64)) - b[:4]).max() < 1e-3     # four points are fitted exactly
    assert np.array_equal(_cabi.refit_homography(a[:0], b[:0], None, H0), H0)       # no points at all
    got0 = _cabi.refit_homography(a, b, mask, H0, lm_iters=0)                       # DLT only
    assert np.allclose(got0, refit_model.fit_homography_dlt(a[mask.astype(bool)], b[mask.astype(bool)]), rtol=1e-9, atol=1e-12)


def test_refit_degenerate_points_return_finite_or_the_hypothesis():
    H0 = np.eye(3)
    a = np.zeros((20, 2), np.float32)
    a[:, 0] = np.arange(20)                      # collinear
    got = _cabi.refit_homography(a, a.copy(), None, H0)
    assert got.shape == (3, 3)                   # never raises; the caller validates H like the reference does
    same = np.ones((20, 2), np.float32)          # one repeated point
    assert _cabi.refit_homography(same, same, None, H0).shape == (3, 3)


def test_refit_argument_errors():
    with pytest.raises(ValueError):
        _cabi.refit_homography(np.zeros((5, 2)), np.zeros((4, 2)), None, np.eye(3))
    with pytest.raises(ValueError):
        _cabi.refit_homography(np.zeros((5, 2)), np.zeros((5, 2)), np.zeros(3), np.eye(3))
    lib = _cabi.load()
    assert lib.mcs_refit_homography(None, None, None, 3, None, 10, None) == -1
    assert b"NULL" in lib.mcs_last_error()
    assert lib.mcs_refit_homography(None, None, None, -1, None, 10, None) == -1
