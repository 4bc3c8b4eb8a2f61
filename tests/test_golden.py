"""Golden fixtures (tests/golden, written by scripts/make_golden.py).

CPU: the oracle and the product's host-side helpers reproduce the fixtures - this pins the
oracle (the reference itself ships no vectors) and guards against a cv2
upgrade silently moving the target.  GPU: the kernels reproduce them through the C ABI.
Inputs are regenerated from the generator's seeds; nothing here reads /root/reference."""
import importlib.util
import os

import numpy as np
import pytest

from oracle import match_model, stitcher_ref, warp_model

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = os.path.join(HERE, "golden")


def _gen():
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(os.path.dirname(HERE), "scripts",
                                                                             "make_golden.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load(name):
    return np.load(os.path.join(GOLD, name))


# ---- CPU ---------------------------------------------------------------------------------------
def test_utils_match_the_reference_module_outputs():
    from multicamera_stitching_b200 import (CalculateProjectionMatrix, get_projection_point_dst,
                                            get_projection_point_src)
    g = load("utils_reference.npz")
    for i in range(len(g["src"])):
        M, INVM = CalculateProjectionMatrix(g["src"][i], g["dst"][i])
        assert np.array_equal(M, g["M"][i]) and np.array_equal(INVM, g["INVM"][i])
        fwd = [get_projection_point_dst(tuple(p), M) for p in g["pts"]]
        back = [get_projection_point_src(tuple(p), INVM) for p in g["pts"]]
        assert np.array_equal(np.array(fwd), g["fwd"][i])
        assert np.array_equal(np.array(back), g["back"][i])
        assert stitcher_ref.projection_point_dst(tuple(g["pts"][0]), M) == fwd[0]


def test_warp_model_reproduces_cv2_golden():
    mg = _gen()
    g = load("warp_cv2.npz")
    src = mg.warp_source()
    for name, H in mg.WARP_HOMS.items():
        out = warp_model.warp_perspective_u8(src, np.array(H, dtype=np.float64), (140, 90))
        assert np.array_equal(out, g[name]), name


def test_chain_oracle_reproduces_golden_panorama_and_geometry():
    from helpers import synthetic_chain
    g = load("chain_cv2.npz")
    st, states, labels, images = synthetic_chain(3, 96, 128, 3, kind="noise", xoffset=3, yoffset=5)
    assert np.array_equal(stitcher_ref.stitch_chain(states, labels, images), g["pano"])
    for k, (s, sb) in enumerate(zip(states, st.stitchers)):
        for obj in (s, {f: getattr(sb, f) for f in ("cachedAH", "Bpts", "ABSize", "x_limits", "y_limits")}):
            assert np.array_equal(np.asarray(obj["cachedAH"]), g["cachedAH_%d" % k])
            assert np.array_equal(np.asarray(obj["Bpts"]), g["Bpts_%d" % k])
            assert np.array_equal(np.asarray(obj["ABSize"]), g["ABSize_%d" % k])
            assert np.array_equal(np.asarray([obj["x_limits"], obj["y_limits"]]), g["limits_%d" % k])


@pytest.mark.parametrize("case", ["c3", "c4_super", "c6"])
def test_chain_oracle_reproduces_the_reference_classes_golden(case):
    """chain_ref.npz was produced by the reference's own Stitcher / StitcherBase (scripts/make_golden.py through
    oracle/build_ref.py): the oracle restatement and the product's host-side state reproduce it."""
    from helpers import synthetic_chain
    mg = _gen()
    g = load("chain_ref.npz")
    n, h, w, super_mode, kind = mg.CHAIN_REF_CASES[case]
    st, states, labels, images = synthetic_chain(n, h, w, 3, super_mode=super_mode, kind=kind)
    pano = stitcher_ref.stitch_chain(states, labels, images)
    assert pano.shape == g[case + "_pano"].shape and np.array_equal(pano, g[case + "_pano"])
    for k, (s, sb) in enumerate(zip(states, st.stitchers)):
        for obj in (s, {f: getattr(sb, f) for f in ("cachedAH", "Bpts", "ABSize", "x_limits", "y_limits")}):
            assert np.array_equal(np.asarray(obj["cachedAH"]), g["%s_cachedAH_%d" % (case, k)])
            assert np.array_equal(np.asarray(obj["Bpts"]), g["%s_Bpts_%d" % (case, k)])
            assert np.array_equal(np.asarray(obj["ABSize"]), g["%s_ABSize_%d" % (case, k)])
            assert np.array_equal(np.asarray([obj["x_limits"], obj["y_limits"]]), g["%s_limits_%d" % (case, k)])


def test_match_model_reproduces_bfmatcher_golden():
    g = load("match_cv2.npz")
    idx, dist, keep, matches = match_model.match(g["fa"], g["fb"], 0.75)
    assert np.array_equal(idx, g["idx"]) and np.array_equal(dist, g["dist"])
    assert np.array_equal(np.array(matches, dtype=np.int32).reshape(-1, 2), g["matches"])
    mg = _gen()
    fa, fb = mg.match_descriptors()
    assert np.array_equal(fa, g["fa"]) and np.array_equal(fb, g["fb"])   # the seeds still give these sets


def test_resize_model_reproduces_cv2_golden():
    from oracle.resize_model import resize_linear_u8
    mg = _gen()
    g = load("resize_cv2.npz")
    src = mg.resize_source()
    for name, dsize in mg.RESIZE_CASES.items():
        assert np.array_equal(resize_linear_u8(src, dsize), g[name]), name
    assert np.array_equal(resize_linear_u8(np.ascontiguousarray(src[:, :, 0].T), (60, 82)), g["plane"])


def test_prewarp_oracle_reproduces_cv2_golden():
    from multicamera_stitching_b200 import CalculateProjectionMatrix, prewarp
    from oracle import prewarp_ref
    mg = _gen()
    g = load("prewarp_cv2.npz")
    src = mg.prewarp_source()
    mtx, dist = np.array(mg.PREWARP_CAMERA["mtx"]), np.array(mg.PREWARP_CAMERA["dist"])
    M, _ = CalculateProjectionMatrix(src_pts=mg.PREWARP_QUAD[0], dst_pts=mg.PREWARP_QUAD[1])
    assert np.array_equal(M, g["M"])                      # our Utils mirror = the reference's own module
    maps = prewarp.undistort_maps(mtx, dist, (src.shape[1], src.shape[0]))
    und = prewarp_ref.remap_fixed_point(src, *maps)
    assert np.array_equal(und, g["undistorted"])
    assert np.array_equal(warp_model.warp_perspective_u8(und, M, (120, 80)), g["birdseye"])
    ic, ec = {"mtx": mtx, "dist": dist}, {"M": M, "dst_size": (120, 80)}
    assert np.array_equal(prewarp_ref.prewarp(src, ic, ec), g["birdseye"])


# ---- GPU ---------------------------------------------------------------------------------------
@pytest.mark.gpu
def test_gpu_resize_equals_golden(cuda_device):
    import torch
    from multicamera_stitching_b200.engine import CompositeEngine
    mg = _gen()
    g = load("resize_cv2.npz")
    src = torch.from_numpy(mg.resize_source()).to(cuda_device)
    eng = CompositeEngine()
    for name, dsize in mg.RESIZE_CASES.items():
        assert np.array_equal(eng.resize(src, (dsize[1], dsize[0])).cpu().numpy(), g[name]), name
    plane = src[:, :, 0].T.contiguous()
    assert np.array_equal(eng.resize(plane, (82, 60)).cpu().numpy(), g["plane"])


@pytest.mark.gpu
def test_gpu_prewarp_equals_golden(cuda_device):
    from multicamera_stitching_b200 import prewarp
    mg = _gen()
    g = load("prewarp_cv2.npz")
    src = mg.prewarp_source()
    pw = prewarp.PreWarp({"mtx": np.array(mg.PREWARP_CAMERA["mtx"]), "dist": np.array(mg.PREWARP_CAMERA["dist"])},
                         {"M": g["M"], "dst_size": (120, 80)})
    assert np.array_equal(pw.undistort(src), g["undistorted"])
    assert np.array_equal(pw(src), g["birdseye"])



@pytest.mark.gpu
def test_gpu_chain_equals_golden_panorama(cuda_device):
    from helpers import synthetic_chain
    g = load("chain_cv2.npz")
    st, states, labels, images = synthetic_chain(3, 96, 128, 3, kind="noise", xoffset=3, yoffset=5)
    got = st.stitch(images)
    assert got.shape == g["pano"].shape and np.array_equal(got, g["pano"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["c3", "c4_super", "c6"])
def test_gpu_chain_equals_the_reference_classes_golden(cuda_device, case):
    """The CUDA path against panoramas made by the reference's own classes (chain_ref.npz)."""
    from helpers import synthetic_chain
    mg = _gen()
    g = load("chain_ref.npz")
    n, h, w, super_mode, kind = mg.CHAIN_REF_CASES[case]
    st, states, labels, images = synthetic_chain(n, h, w, 3, super_mode=super_mode, kind=kind)
    got = st.stitch(images)
    assert got.shape == g[case + "_pano"].shape and np.array_equal(got, g[case + "_pano"])


@pytest.mark.gpu
def test_gpu_warp_equals_golden(cuda_device):
    """A single warpPerspective through the product path: a 2-camera chain whose first camera
    is an empty-looking 1 x 1 paste does not exist in the reference, so the plan is built
    directly (one WARP layer covering the whole canvas)."""
    import torch
    from multicamera_stitching_b200 import _cabi
    mg = _gen()
    g = load("warp_cv2.npz")
    src = mg.warp_source()
    dev_src = torch.from_numpy(src).cuda()
    for name, H in mg.WARP_HOMS.items():
        plan = _cabi.Plan([_cabi.MCS_LAYER_WARP], [src.shape[:2]], [np.array(H, dtype=np.float64)], [(0, 0)],
                          [(0, 0, 140, 90)], 140, 90, 3)
        out = torch.empty((90, 140, 3), dtype=torch.uint8, device="cuda")
        plan.stitch([dev_src.data_ptr()], [src.shape[1] * 3], [0], 1, out.data_ptr(), 140 * 3, 0,
                    torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), g[name]), name


@pytest.mark.gpu
def test_gpu_matcher_equals_golden(cuda_device):
    import torch
    from multicamera_stitching_b200 import recalib
    g = load("match_cv2.npz")
    q = torch.from_numpy(g["fa"]).cuda()[None]
    t = torch.from_numpy(g["fb"]).cuda()[None]
    idx2, dist2, keep = recalib.match_top2_batch(q, t, ratio=0.75)
    torch.cuda.synchronize()
    assert np.array_equal(idx2[0].cpu().numpy(), g["idx"])
    assert np.array_equal(dist2[0].cpu().numpy(), g["dist"])
    k = keep[0].cpu().numpy().astype(bool)
    got = np.stack([idx2[0, :, 0].cpu().numpy()[k], np.nonzero(k)[0]], axis=1).astype(np.int32)
    assert np.array_equal(got, g["matches"])
