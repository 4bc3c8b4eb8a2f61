"""GPU parity of the per-camera pre-warp through the C ABI
(``mcs_plan_create_maps`` REMAP layers + ``mcs_stitch_u8``): bit-exact against
``cv2.undistort`` / ``cv2.warpPerspective`` called the way the reference's callers call them
(oracle/prewarp_ref.py = MediaPlayer/view.py:378-388, video_mapping_node.py:155-158)."""
import cv2
import numpy as np
import pytest
import torch

from multicamera_stitching_b200 import Utils, prewarp
from multicamera_stitching_b200.engine import CompiledPlan
from multicamera_stitching_b200.plan import LAYER_REMAP, FlatPlan, Layer
from oracle import prewarp_ref
from test_oracle_prewarp import camera

pytestmark = pytest.mark.gpu


def _img(h, w, c, seed=0):
    return np.random.default_rng(seed).integers(0, 256, size=(h, w, c) if c > 1 else (h, w), dtype=np.uint8)


@pytest.mark.parametrize("h,w,c", [(360, 640, 3), (1080, 1920, 3), (240, 320, 1), (96, 132, 4), (97, 131, 3)])
def test_undistort_equals_cv2(cuda_device, h, w, c):
    img = _img(h, w, c, h)
    mtx, dist = camera(h, w)
    ref = cv2.undistort(img, mtx, dist)
    got = prewarp.undistort(img, mtx, dist)
    assert isinstance(got, np.ndarray) and got.shape == ref.shape and np.array_equal(got, ref)
    dev = prewarp.undistort(torch.from_numpy(img).to(cuda_device), mtx, dist)
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), ref)


@pytest.mark.parametrize("variant", [1, 2])
def test_remap_layer_both_kernel_variants(cuda_device, variant):
    h, w = 360, 640
    img = _img(h, w, 3, 1)
    mtx, dist = camera(h, w, strength=1.5)
    maps = prewarp.undistort_maps(mtx, dist, (w, h))
    flat = FlatPlan([Layer(0, LAYER_REMAP, None, 0, 0, (0, 0, w, h), (h, w), maps)], w, h, 3, 3)
    plan = CompiledPlan(flat, cuda_device)
    assert plan.handle.tiled_status() == ""
    plan.handle.force_variant(variant)
    out = plan.run([torch.from_numpy(img).to(cuda_device)])
    assert plan.handle.last_variant() == variant
    ref = cv2.undistort(img, mtx, dist)
    assert np.array_equal(out.cpu().numpy(), ref)
    assert np.array_equal(prewarp_ref.remap_fixed_point(img, *maps), ref)
    # owned pixels of the layer = output pixels with a tap inside the source
    xy = maps[0].astype(np.int64)
    inside = ((xy[..., 0] >= -1) & (xy[..., 0] < w) & (xy[..., 1] >= -1) & (xy[..., 1] < h)).sum()
    assert plan.owned_pixels() == [int(inside)]


def test_remap_layer_random_map_with_taps_outside(cuda_device):
    rng = np.random.default_rng(2)
    src = rng.integers(0, 256, size=(60, 80, 3), dtype=np.uint8)
    ys, xs = np.mgrid[0:70, 0:92]
    xy = np.stack([xs - 6 + rng.integers(-1, 2, size=xs.shape), ys - 5 + rng.integers(-1, 2, size=xs.shape)],
                  axis=-1).astype(np.int16)
    frac = rng.integers(0, 1024, size=xs.shape).astype(np.uint16)
    ref = cv2.remap(src, xy, frac, cv2.INTER_LINEAR)
    flat = FlatPlan([Layer(0, LAYER_REMAP, None, 0, 0, (0, 0, 92, 70), (60, 80), (xy, frac))], 92, 70, 3, 3)
    for variant in (1, 2):
        plan = CompiledPlan(flat, cuda_device)
        plan.handle.force_variant(variant)
        out = plan.run([torch.from_numpy(src).to(cuda_device)])
        assert np.array_equal(out.cpu().numpy(), ref)


def test_warp_perspective_equals_cv2(cuda_device):
    h, w = 360, 640
    img = _img(h, w, 3, 3)
    M, _ = Utils.CalculateProjectionMatrix([(120, 180), (520, 180), (620, 340), (20, 340)],
                                           [(0, 0), (300, 0), (300, 200), (0, 200)])
    ref = cv2.warpPerspective(src=img, M=M, dsize=(300, 200))
    assert np.array_equal(prewarp.warpPerspective(img, M, (300, 200)), ref)


def test_prewarp_sequence_and_batch(cuda_device):
    h, w = 360, 640
    mtx, dist = camera(h, w)
    M, _ = Utils.CalculateProjectionMatrix([(120, 180), (520, 180), (620, 340), (20, 340)],
                                           [(0, 0), (300, 0), (300, 200), (0, 200)])
    ic, ec = {"mtx": mtx, "dist": dist}, {"M": M, "dst_size": (300, 200)}
    pw = prewarp.PreWarp(ic, ec)
    frames = np.stack([_img(h, w, 3, 10 + f) for f in range(4)])
    refs = [prewarp_ref.prewarp(frames[f], ic, ec) for f in range(4)]
    assert np.array_equal(pw(frames[0]), refs[0])
    out = pw(torch.from_numpy(frames).to(cuda_device), batched=True)
    assert out.shape == (4, 200, 300, 3)
    for f in range(4):
        assert np.array_equal(out[f].cpu().numpy(), refs[f])
    # undistortion only (the node: video_mapping_node.py:155-158)
    only = prewarp.PreWarp(ic)
    assert np.array_equal(only(frames[1]), cv2.undistort(frames[1], mtx, dist))
