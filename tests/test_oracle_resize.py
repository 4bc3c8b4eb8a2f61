"""CPU: the integer model of ``cv2.resize(INTER_LINEAR)`` for uint8 frames
(oracle/resize_model.py) against cv2 itself - the pin of the shape fix-up path
(StitcherClass.py:226-233) - and the host logic that decides where the fix-up
strikes in a chain (``StitcherClass._segments``)."""
import cv2
import numpy as np
import pytest

from helpers import synthetic_chain
from multicamera_stitching_b200 import StitcherClass
from oracle.resize_model import resize_linear_u8

CASES = [
    ((720, 1280, 3), (1920, 1080)),   # upscale to the calibrated 1080p
    ((1080, 1920, 3), (1280, 720)),   # 1.5 x decimation
    ((1280, 720), (1920, 1080)),      # MediaPlayer: 2-D transposed frame (view.py:408-409)
    ((480, 640, 3), (320, 240)),      # exact 2 x 2 decimation -> OpenCV's area kernel
    ((200, 300, 3), (150, 50)),       # 2 x in one axis only stays linear
    ((37, 53, 3), (101, 77)),
    ((100, 100, 4), (33, 17)),
    ((64, 64, 3), (64, 64)),          # identity
    ((5, 7, 3), (1, 1)),
    ((1, 1, 3), (9, 5)),
    ((2, 2, 1), (5, 5)),
    ((1080, 1920, 3), (1919, 1079)),  # scale just above 1
]


def _same(got, ref):
    if ref.ndim == 2 and got.ndim == 3:   # cv2 drops a trailing channel axis of 1
        got = got[:, :, 0]
    assert got.shape == ref.shape
    assert np.array_equal(got, ref)


@pytest.mark.parametrize("shape,dsize", CASES)
def test_model_equals_cv2(shape, dsize):
    rng = np.random.default_rng(hash((shape, dsize)) % (2 ** 32))
    img = rng.integers(0, 256, size=shape, dtype=np.uint8)
    _same(resize_linear_u8(img, dsize), cv2.resize(img, dsize, interpolation=cv2.INTER_LINEAR))


def test_model_equals_cv2_random_sizes():
    rng = np.random.default_rng(7)
    for _ in range(60):
        c = int(rng.choice([0, 1, 3, 4]))
        shape = (int(rng.integers(1, 200)), int(rng.integers(1, 200))) + ((c,) if c else ())
        dsize = (int(rng.integers(1, 300)), int(rng.integers(1, 300)))
        img = rng.integers(0, 256, size=shape, dtype=np.uint8)
        _same(resize_linear_u8(img, dsize), cv2.resize(img, dsize, interpolation=cv2.INTER_LINEAR))


def test_segments_regular_chain_is_one_segment():
    st, _, labels, images = synthetic_chain(4, 90, 160, 3)
    shapes = [images[l].shape for l in labels]
    segs = StitcherClass._segments(st.stitchers, shapes)
    assert segs == [{"a": 0, "b": 3, "base_hw": None, "resize": {}, "n": 3}]


def test_segments_camera_frames_are_resized_in_place():
    st, _, labels, images = synthetic_chain(4, 90, 160, 3)
    shapes = [images[l].shape for l in labels]
    shapes[0] = (45, 80, 3)
    shapes[2] = (100, 100, 3)
    segs = StitcherClass._segments(st.stitchers, shapes)
    assert segs == [{"a": 0, "b": 3, "base_hw": (90, 160), "resize": {2: (90, 160)}, "n": 3}]
    # a channel-only difference (2-D frames against a 3-channel calibration) warns but resizes nothing
    logged = []

    class Log(object):
        def debugger(self, level, msg, log_type="info"):
            logged.append((log_type, msg))

    segs = StitcherClass._segments(st.stitchers, [s[:2] for s in [images[l].shape for l in labels]], Log())
    assert segs == [{"a": 0, "b": 3, "base_hw": None, "resize": {}, "n": 3}]
    assert len(logged) == 6 and all(t == "warn" for t, _ in logged)
    assert "ImageB size should be (90, 160, 3), Image will be resized" in logged[0][1]


def test_segments_cut_where_a_composited_canvas_must_be_resized():
    st, _, labels, images = synthetic_chain(4, 90, 160, 3)
    shapes = [images[l].shape for l in labels]
    # stage 1 was calibrated against a canvas of another size (a mixed set of saved calibrations)
    h, w = st.stitchers[1].BimgSize[:2]
    st.stitchers[1].BimgSize = (h + 6, w - 10, 3)
    segs = StitcherClass._segments(st.stitchers, shapes)
    assert [(s["a"], s["b"], s["base_hw"]) for s in segs] == [(0, 1, None), (1, 3, (h + 6, w - 10))]
