"""CPU: the pre-warp path's host logic and oracle.

* ``prewarp.undistort_maps`` builds the map pair exactly the way ``cv::undistort``
  does: remapping through it with cv2 reproduces ``cv2.undistort`` bit for bit;
* ``oracle/prewarp_ref.remap_fixed_point`` (the readable specification of a
  REMAP layer) equals ``cv2.remap`` on those maps and on random fixed-point maps
  with taps outside the source."""
import cv2
import numpy as np
import pytest

from multicamera_stitching_b200 import prewarp
from oracle import prewarp_ref


def camera(h, w, strength=1.0):
    f = 0.8 * w
    mtx = np.array([[f, 0, w / 2 + 3.3], [0, f * 1.01, h / 2 - 2.1], [0, 0, 1]])
    dist = np.array([-0.32, 0.12, 0.001, -0.0007, -0.02]) * strength
    return mtx, dist


@pytest.mark.parametrize("h,w,c", [(360, 640, 3), (720, 1280, 3), (240, 320, 1), (97, 131, 4), (50, 5000, 3)])
def test_striped_maps_reproduce_cv2_undistort(h, w, c):
    rng = np.random.default_rng(h + w)
    img = rng.integers(0, 256, size=(h, w, c) if c > 1 else (h, w), dtype=np.uint8)
    mtx, dist = camera(h, w)
    xy, frac = prewarp.undistort_maps(mtx, dist, (w, h))
    assert xy.dtype == np.int16 and xy.shape == (h, w, 2) and frac.dtype == np.uint16 and frac.shape == (h, w)
    ref = cv2.undistort(img, mtx, dist)
    assert np.array_equal(cv2.remap(img, xy, frac, cv2.INTER_LINEAR), ref)
    assert np.array_equal(prewarp_ref.remap_fixed_point(img, xy, frac), ref)


def test_undistort_maps_options():
    h, w = 120, 200
    mtx, dist = camera(h, w)
    img = np.random.default_rng(3).integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    new = mtx.copy()
    new[0, 0] *= 0.7
    new[1, 1] *= 0.7
    xy, frac = prewarp.undistort_maps(mtx, dist, (w, h), newCameraMatrix=new)
    assert np.array_equal(cv2.remap(img, xy, frac, cv2.INTER_LINEAR), cv2.undistort(img, mtx, dist, None, new))
    xy, frac = prewarp.undistort_maps(mtx, None, (w, h))     # no distortion: identity map
    assert np.array_equal(cv2.remap(img, xy, frac, cv2.INTER_LINEAR), img)


def test_remap_model_equals_cv2_on_random_maps():
    rng = np.random.default_rng(11)
    for c in (1, 3, 4):
        src = rng.integers(0, 256, size=(60, 80, c) if c > 1 else (60, 80), dtype=np.uint8)
        xy = np.stack([rng.integers(-5, 86, size=(70, 90)), rng.integers(-5, 66, size=(70, 90))], axis=-1).astype(np.int16)
        frac = rng.integers(0, 1024, size=(70, 90)).astype(np.uint16)
        assert np.array_equal(prewarp_ref.remap_fixed_point(src, xy, frac), cv2.remap(src, xy, frac, cv2.INTER_LINEAR))


def test_prewarp_sequence_is_the_callers():
    h, w = 180, 320
    mtx, dist = camera(h, w)
    img = np.random.default_rng(5).integers(0, 256, size=(h, w, 3), dtype=np.uint8)
    M = cv2.getPerspectiveTransform(np.float32([(60, 90), (260, 90), (310, 170), (10, 170)]),
                                    np.float32([(0, 0), (300, 0), (300, 200), (0, 200)]))
    ic, ec = {"mtx": mtx, "dist": dist}, {"M": M, "dst_size": (300, 200)}
    out = prewarp_ref.prewarp(img, ic, ec)
    assert out.shape == (200, 300, 3)
    assert np.array_equal(out, cv2.warpPerspective(cv2.undistort(img, mtx, dist), M, (300, 200)))
    assert prewarp_ref.prewarp(img, {"mtx": None, "dist": None}, ec) is img
    assert np.array_equal(prewarp_ref.prewarp(img, ic, {"M": None, "dst_size": None}), cv2.undistort(img, mtx, dist))
    # no device is needed to pass a frame through an uncalibrated PreWarp
    assert prewarp.PreWarp()(img) is img
