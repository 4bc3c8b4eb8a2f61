"""CPU: pins the oracle's integer warp model to cv2.warpPerspective itself
(the reference's call, StitcherClass.py:239) - bit-exact, every case."""
import cv2
import numpy as np
import pytest

from oracle import warp_model

HOMS = {
    "near_identity": [[0.97, 0.02, 100.3], [-0.015, 0.99, 40.7], [2e-5, -1e-5, 1.0]],
    "integer_shift": [[1, 0, 37.0], [0, 1, 12.0], [0, 0, 1]],
    "rotate_scale": [[0.6, -0.5, 200.0], [0.5, 0.6, -30.0], [0, 0, 1]],
    "perspective": [[1.1, 0.1, -20.0], [0.05, 0.9, 15.0], [4e-4, -3e-4, 1.0]],
    "downscale": [[0.31, 0.0, 5.5], [0.0, 0.29, 7.25], [0, 0, 1]],
    "upscale": [[3.7, 0.2, -50.0], [0.1, 4.1, -80.0], [1e-4, 0, 1]],
    "w_crosses_zero": [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [-0.01, 0.0, 1.0]],
}


@pytest.mark.parametrize("name", sorted(HOMS))
@pytest.mark.parametrize("shape,dsize", [
    ((240, 320, 3), (500, 300)),
    ((100, 50), (63, 40)),          # grayscale, destination narrower than one 64-column block
    ((77, 129, 3), (200, 33)),
    ((64, 64, 4), (130, 130)),
])
def test_model_equals_cv2(name, shape, dsize):
    rng = np.random.default_rng(hash(name) % 1000)
    src = rng.integers(0, 256, size=shape, dtype=np.uint8)
    M = np.array(HOMS[name], dtype=np.float64)
    ref = cv2.warpPerspective(src, M, dsize)
    got = warp_model.warp_perspective_u8(src, M, dsize)
    assert np.array_equal(got, ref)


def test_window_evaluation_keeps_block_phase():
    rng = np.random.default_rng(3)
    src = rng.integers(0, 256, size=(200, 300, 3), dtype=np.uint8)
    M = np.array(HOMS["perspective"])
    ref = cv2.warpPerspective(src, M, (400, 250))
    got = warp_model.warp_perspective_u8(src, M, (400, 250), x_range=(70, 333), y_range=(11, 200))
    assert np.array_equal(got, ref[11:200, 70:333])


def test_invert3x3_is_cv2_invert():
    rng = np.random.default_rng(1)
    for _ in range(3000):
        M = np.array([[rng.uniform(.5, 1.5), rng.uniform(-.2, .2), rng.uniform(-3000, 3000)],
                      [rng.uniform(-.2, .2), rng.uniform(.5, 1.5), rng.uniform(-2000, 2000)],
                      [rng.uniform(-1e-4, 1e-4), rng.uniform(-1e-4, 1e-4), rng.uniform(.9, 1.1)]])
        ok, inv = cv2.invert(M)
        assert ok != 0
        assert np.array_equal(inv, warp_model.invert3x3(M))
    assert warp_model.invert3x3([[1, 2, 3], [2, 4, 6], [0, 0, 1]]) is None


def test_singular_matrix_matches_cv2():
    rng = np.random.default_rng(2)
    src = rng.integers(0, 256, size=(40, 60, 3), dtype=np.uint8)
    M = np.array([[1, 2, 3.0], [2, 4, 6.0], [0, 0, 1]])
    assert np.array_equal(warp_model.warp_perspective_u8(src, M, (50, 30)), cv2.warpPerspective(src, M, (50, 30)))


def test_float32_matrix_is_promoted_like_cv2():
    rng = np.random.default_rng(4)
    src = rng.integers(0, 256, size=(90, 120, 3), dtype=np.uint8)
    M32 = np.array(HOMS["near_identity"], dtype=np.float32)
    assert np.array_equal(warp_model.warp_perspective_u8(src, M32, (200, 150)),
                          cv2.warpPerspective(src, M32, (200, 150)))
