"""GPU parity of the shape fix-up path: the
``mcs_resize_linear_u8`` kernel against ``cv2.resize(INTER_LINEAR)`` - bit-exact -
and ``Stitcher.stitch`` on frames whose shape differs from the calibrated one
against the reference's chain with its own cv2 resizes
(oracle/stitcher_ref.stitch_pair = StitcherClass.py:226-233)."""
import cv2
import numpy as np
import pytest
import torch

from helpers import compare_u8, synthetic_chain
from multicamera_stitching_b200 import synthetic
from multicamera_stitching_b200.engine import CompositeEngine
from oracle import stitcher_ref
from oracle.resize_model import resize_linear_u8

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape,dsize", [
    ((720, 1280, 3), (1920, 1080)),
    ((1080, 1920, 3), (1280, 720)),
    ((1280, 720), (1920, 1080)),
    ((480, 640, 3), (320, 240)),      # exact 2 x 2 decimation (area kernel)
    ((200, 300, 3), (150, 50)),
    ((37, 53, 3), (101, 77)),
    ((100, 100, 4), (33, 17)),
    ((64, 64, 3), (64, 64)),
    ((5, 7, 3), (1, 1)),
    ((1, 1, 3), (9, 5)),
    ((2160, 3840, 3), (1920, 1080)),  # 4K -> 1080p (area kernel)
    ((1080, 1920, 3), (3840, 2160)),
    ((600, 1000, 3), (100, 60)),      # 10 x decimation
    ((300, 403), (200, 100)),         # one-channel rows that are not a multiple of 4 bytes
    ((64, 50, 3), (300, 300)),        # upscale across several tiles from a narrow source
])
def test_resize_kernel_equals_cv2(cuda_device, shape, dsize):
    rng = np.random.default_rng(hash((shape, dsize)) % (2 ** 32))
    img = rng.integers(0, 256, size=shape, dtype=np.uint8)
    ref = cv2.resize(img, dsize, interpolation=cv2.INTER_LINEAR)
    got = CompositeEngine().resize(torch.from_numpy(img).to(cuda_device), (dsize[1], dsize[0])).cpu().numpy()
    assert got.shape == ref.shape and np.array_equal(got, ref)
    if img.size <= 1 << 20:
        assert np.array_equal(got, resize_linear_u8(img, dsize))


def test_resize_kernel_batched_and_strided(cuda_device):
    rng = np.random.default_rng(5)
    frames = rng.integers(0, 256, size=(5, 123, 211, 3), dtype=np.uint8)
    dev = torch.from_numpy(frames).to(cuda_device)
    eng = CompositeEngine()
    out = eng.resize(dev, (77, 300), batched=True).cpu().numpy()
    for f in range(5):
        assert np.array_equal(out[f], cv2.resize(frames[f], (300, 77), interpolation=cv2.INTER_LINEAR))
    # a window of a larger tensor: row pitch and frame stride larger than the image
    win = dev[1:4, 10:100, 20:150]
    out = eng.resize(win, (45, 260), batched=True).cpu().numpy()
    for f in range(3):
        ref = cv2.resize(np.ascontiguousarray(frames[1 + f, 10:100, 20:150]), (260, 45), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(out[f], ref)


def test_stitch_resizes_mismatched_frames_like_the_reference(cuda_device):
    st, states, labels, images = synthetic_chain(4, 180, 320, 3, kind="noise")
    rng = np.random.default_rng(1)
    images = dict(images)
    images[labels[0]] = rng.integers(0, 256, size=(90, 200, 3), dtype=np.uint8)     # camera 0 (imageB of stage 0)
    images[labels[2]] = rng.integers(0, 256, size=(360, 640, 3), dtype=np.uint8)    # exact 2 x 2 decimation
    images[labels[3]] = rng.integers(0, 256, size=(200, 300, 3), dtype=np.uint8)
    ref = stitcher_ref.stitch_chain(states, labels, images)
    got = st.stitch(images)
    assert compare_u8(got, ref) == (0, 1.0)
    dev = {l: torch.from_numpy(images[l]).to(cuda_device) for l in labels}
    assert compare_u8(st.stitch(dev).cpu().numpy(), ref) == (0, 1.0)
    # the pair API
    pair = (images[labels[0]], images[labels[1]])
    assert compare_u8(st.stitchers[0].stitch(pair), stitcher_ref.stitch_pair(states[0], pair)) == (0, 1.0)


def test_stitch_2d_frames_against_a_3_channel_calibration(cuda_device):
    """MediaPlayer/view.py:408-409: calibrated on BGR frames, fed ``[:, :, 0].T`` planes."""
    st, states, labels, images = synthetic_chain(3, 180, 320, 3, kind="noise")
    planes = {l: np.ascontiguousarray(images[l][:, :, 0].T) for l in labels}       # 320 x 180, 2-D
    ref = stitcher_ref.stitch_chain(states, labels, planes)
    got = st.stitch(planes)
    assert got.ndim == 2 and compare_u8(got, ref) == (0, 1.0)


def test_stitch_resizes_a_composited_canvas(cuda_device):
    st, states, labels, images = synthetic_chain(4, 120, 200, 3, kind="noise")
    # stage 1 calibrated against a canvas of another size: the reference resizes the canvas
    # stitched so far before it goes on (StitcherClass.py:226-229)
    h, w = st.stitchers[1].BimgSize[:2]
    shapeB = (h + 6, w - 10, 3)
    shapes = [images[l].shape for l in labels]
    for k in (1, 2):
        H = synthetic.make_homography(k, 120, 200, shapeB[1])
        st.stitchers[k].set_homography(H, shapeA=shapes[k + 1], shapeB=shapeB, xoffset=0, yoffset=0)
        states[k] = stitcher_ref.new_state(sid=str(k))
        stitcher_ref.geometry_from_homography(states[k], H, shapes[k + 1], shapeB, 0, 0)
        shapeB = st.stitchers[k].result_shape()
    ref = stitcher_ref.stitch_chain(states, labels, images)
    got = st.stitch(images)
    assert compare_u8(got, ref) == (0, 1.0)
    batch = {l: torch.from_numpy(np.stack([images[l], images[l][::-1].copy()])).to(cuda_device) for l in labels}
    outb = st.stitch_batch(batch).cpu().numpy()
    assert compare_u8(outb[0], ref) == (0, 1.0)
    flipped = {l: images[l][::-1].copy() for l in labels}
    assert compare_u8(outb[1], stitcher_ref.stitch_chain(states, labels, flipped)) == (0, 1.0)
