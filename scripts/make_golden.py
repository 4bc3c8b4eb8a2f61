#!/usr/bin/env python
"""Generates the fixtures under tests/golden/ (run in the authoring container, where
/root/reference and cv2 are present; the fixtures travel, this script's inputs do not).

  python scripts/make_golden.py [/root/reference]

  utils_reference.npz   outputs of the REFERENCE'S OWN PostScripts/Calibration_Utils/Utils.py
                        (imported from the reference tree, it runs under Python 3):
                        get_projection_point_dst / _src, CalculateProjectionMatrix
  warp_cv2.npz          cv2.warpPerspective(INTER_LINEAR, BORDER_CONSTANT 0) called as
                        StitcherClass.py:239 calls it, on a seeded noise image
  chain_ref.npz         panoramas and stage states of the REFERENCE'S OWN Stitcher / StitcherBase classes
                        (StitcherClass.py made importable by oracle/build_ref.py): 3-, 4- (super mode) and 6-camera
                        chains calibrated by calibrate_stitcher on injected stage homographies
  chain_cv2.npz         the 3-camera warp+paste chain of StitcherClass.py:131-136 / :239-241
                        (oracle/stitcher_ref.py driving cv2), with its geometry (:293-351)
  match_cv2.npz         cv2.BFMatcher(NORM_HAMMING).knnMatch(k=2) + the ratio loop of :428-433
                        on seeded descriptors with planted partners and exact ties
  resize_cv2.npz        cv2.resize(INTER_LINEAR) as StitcherClass.py:229 / :233 call it (up, down, exact 2 x 2
                        decimation, one axis halved, identity, a transposed one-channel plane)
  prewarp_cv2.npz       cv2.undistort + cv2.warpPerspective as the callers run them in front of the stitcher
                        (view.py:378-388), extrinsic matrix from the reference's own CalculateProjectionMatrix
All inputs are regenerated from seeds by the tests; only the expected outputs (and the
descriptor sets) are stored.
"""
import importlib.util
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
OUT = os.path.join(ROOT, "tests", "golden")


def load_reference_utils(ref_root):
    path = os.path.join(ref_root, "PostScripts", "Calibration_Utils", "Utils.py")
    spec = importlib.util.spec_from_file_location("reference_Utils", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def utils_cases():
    rng = np.random.default_rng(42)
    quads = []
    for _ in range(12):
        w, h = int(rng.integers(200, 2000)), int(rng.integers(200, 1200))
        src = np.float64([[0, 0], [w, 0], [w, h], [0, h]]) + rng.uniform(-30, 30, (4, 2))
        dst = src * rng.uniform(0.6, 1.4) + rng.uniform(-120, 400, (1, 2)) + rng.uniform(-25, 25, (4, 2))
        quads.append((src, dst))
    pts = np.concatenate([rng.uniform(-50, 2100, (40, 2)), np.ones((40, 1))], axis=1)
    return quads, pts


def golden_utils(ref_root):
    U = load_reference_utils(ref_root)
    quads, pts = utils_cases()
    Ms, INVMs, fwd, back = [], [], [], []
    for src, dst in quads:
        M, INVM = U.CalculateProjectionMatrix(src, dst)
        Ms.append(M)
        INVMs.append(INVM)
        fwd.append([U.get_projection_point_dst(tuple(p), M) for p in pts])
        back.append([U.get_projection_point_src(tuple(p), INVM) for p in pts])
    np.savez_compressed(os.path.join(OUT, "utils_reference.npz"),
                        src=np.array([q[0] for q in quads]), dst=np.array([q[1] for q in quads]), pts=pts,
                        M=np.array(Ms), INVM=np.array(INVMs), fwd=np.array(fwd, dtype=np.int64),
                        back=np.array(back, dtype=np.int64))


WARP_HOMS = {
    "near_identity": [[0.98, 0.015, 13.2], [-0.01, 1.003, 4.7], [1.1e-5, -4e-6, 1.0]],
    "rotate_scale": [[0.61, -0.52, 60.0], [0.48, 0.66, -5.0], [2e-4, 1e-4, 1.0]],
    "strong_perspective": [[1.2, 0.1, -20.0], [0.05, 1.3, -10.0], [1.5e-3, 8e-4, 1.0]],
}


def warp_source():
    return np.random.default_rng(1234).integers(0, 256, size=(61, 83, 3), dtype=np.uint8)


def golden_warp():
    src = warp_source()
    out = {}
    for name, H in WARP_HOMS.items():
        out[name] = cv2.warpPerspective(src, np.array(H, dtype=np.float64), (140, 90))
    np.savez_compressed(os.path.join(OUT, "warp_cv2.npz"), **out)


def golden_chain():
    from helpers import synthetic_chain
    from oracle import stitcher_ref
    st, states, labels, images = synthetic_chain(3, 96, 128, 3, kind="noise", xoffset=3, yoffset=5)
    pano = stitcher_ref.stitch_chain(states, labels, images)
    geo = {}
    for k, s in enumerate(states):
        geo["cachedAH_%d" % k] = np.asarray(s["cachedAH"], dtype=np.float64)
        geo["Bpts_%d" % k] = np.asarray(s["Bpts"], dtype=np.int64)
        geo["ABSize_%d" % k] = np.asarray(s["ABSize"], dtype=np.int64)
        geo["limits_%d" % k] = np.asarray([s["x_limits"], s["y_limits"]], dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "chain_cv2.npz"), pano=pano, **geo)


CHAIN_REF_CASES = {
    # name: (cameras, height, width, super_mode, frame kind, xoffset, yoffset)  -- offsets of calibrate_stitcher are 0
    "c3": (3, 96, 128, False, "noise"),
    "c4_super": (4, 90, 160, True, "noise"),
    "c6": (6, 68, 120, False, "smooth"),
}


def chain_ref_homographies(st, n, h, w):
    """The stage homographies ``helpers.synthetic_chain`` uses (they depend on the running canvas width)."""
    from multicamera_stitching_b200 import synthetic
    homs, cw = [], w
    for k in range(n - 1):
        homs.append(synthetic.make_homography(k, h, w, cw))
        cw = st.stitchers[k].result_shape()[1]
    return homs


def golden_chain_ref():
    """Panoramas and stage states produced by the REFERENCE'S OWN Stitcher / StitcherBase classes
    (oracle/build_ref.py makes StitcherClass.py importable): calibrate_stitcher (:77-112) runs its real
    canvas geometry (:293-351) on injected stage homographies, stitch (:114-136, :211-256) makes the panorama."""
    from helpers import synthetic_chain
    from oracle import build_ref
    ref = build_ref.load()
    if ref is None:
        raise SystemExit("make_golden.py: the reference tree is needed for chain_ref.npz")
    out = {}
    for name, (n, h, w, super_mode, kind) in CHAIN_REF_CASES.items():
        st, states, labels, images = synthetic_chain(n, h, w, 3, super_mode=super_mode, kind=kind)
        rs = ref.Stitcher(images, super_mode=super_mode)
        for sb, H in zip(rs.stitchers, chain_ref_homographies(st, n, h, w)):
            sb.detectAndDescribe = lambda image: (np.zeros((1, 2), np.float32), None)
            sb.matchKeypoints = (lambda Hk: (lambda **kw: (np.array(Hk, dtype=np.float64), [(0, 0)] * 5,
                                                          np.ones((5, 1), np.uint8))))(H)
        rs.calibrate_stitcher(images, save=False)
        out[name + "_pano"] = rs.stitch(images)
        for k, sb in enumerate(rs.stitchers):
            out["%s_cachedAH_%d" % (name, k)] = np.asarray(sb.cachedAH, dtype=np.float64)
            out["%s_Bpts_%d" % (name, k)] = np.asarray(sb.Bpts, dtype=np.int64)
            out["%s_ABSize_%d" % (name, k)] = np.asarray(sb.ABSize, dtype=np.int64)
            out["%s_limits_%d" % (name, k)] = np.asarray([sb.x_limits, sb.y_limits], dtype=np.int64)
    np.savez_compressed(os.path.join(OUT, "chain_ref.npz"), **out)


def match_descriptors():
    rng = np.random.default_rng(77)
    fb = rng.integers(0, 256, size=(350, 32), dtype=np.uint8)
    fa = rng.integers(0, 256, size=(300, 32), dtype=np.uint8)
    fa[:150] = fb[rng.permutation(350)[:150]]
    fa[:150, :3] ^= rng.integers(0, 256, size=(150, 3), dtype=np.uint8) & 0x21
    fb[340] = fb[7]      # exact ties
    fb[341] = fb[7]
    fa[299] = fb[7]
    return fa, fb


def golden_match():
    fa, fb = match_descriptors()
    raw = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(fa, fb, 2)
    idx = np.array([[m[0].trainIdx, m[1].trainIdx] for m in raw], dtype=np.int32)
    dist = np.array([[int(m[0].distance), int(m[1].distance)] for m in raw], dtype=np.int32)
    matches = np.array([(m[0].trainIdx, m[0].queryIdx) for m in raw
                        if len(m) == 2 and m[0].distance < m[1].distance * 0.75], dtype=np.int32)
    np.savez_compressed(os.path.join(OUT, "match_cv2.npz"), fa=fa, fb=fb, idx=idx, dist=dist, matches=matches)


RESIZE_CASES = {"up": (101, 77), "down": (40, 23), "half_area": (41, 30), "identity": (82, 60), "one_axis": (41, 90)}


def resize_source():
    # 60 x 82 so that (41, 30) is an exact 2 x 2 decimation (OpenCV's area kernel)
    return np.random.default_rng(4321).integers(0, 256, size=(60, 82, 3), dtype=np.uint8)


def golden_resize():
    """cv2.resize(INTER_LINEAR) as StitcherClass.py:229 / :233 call it."""
    src = resize_source()
    out = {}
    for name, dsize in RESIZE_CASES.items():
        out[name] = cv2.resize(src, dsize, interpolation=cv2.INTER_LINEAR)
    out["plane"] = cv2.resize(np.ascontiguousarray(src[:, :, 0].T), (60, 82), interpolation=cv2.INTER_LINEAR)
    np.savez_compressed(os.path.join(OUT, "resize_cv2.npz"), **out)


PREWARP_CAMERA = dict(mtx=[[130.0, 0.0, 84.3], [0.0, 131.5, 47.9], [0.0, 0.0, 1.0]],
                      dist=[-0.32, 0.12, 0.001, -0.0007, -0.02])
PREWARP_QUAD = ([(40, 45), (130, 45), (155, 85), (10, 85)], [(0, 0), (120, 0), (120, 80), (0, 80)])


def prewarp_source():
    return np.random.default_rng(2468).integers(0, 256, size=(96, 168, 3), dtype=np.uint8)


def golden_prewarp(ref_root):
    """cv2.undistort + cv2.warpPerspective as the callers run them (view.py:378-388), with the
    extrinsic matrix from the REFERENCE'S OWN Utils.CalculateProjectionMatrix (Extrinsic.py:96-99)."""
    U = load_reference_utils(ref_root)
    src = prewarp_source()
    mtx, dist = np.array(PREWARP_CAMERA["mtx"]), np.array(PREWARP_CAMERA["dist"])
    M, _ = U.CalculateProjectionMatrix(src_pts=PREWARP_QUAD[0], dst_pts=PREWARP_QUAD[1])
    und = cv2.undistort(src=src, cameraMatrix=mtx, distCoeffs=dist)
    bird = cv2.warpPerspective(src=und, M=M, dsize=(120, 80))
    np.savez_compressed(os.path.join(OUT, "prewarp_cv2.npz"), undistorted=und, birdseye=bird, M=M)


def recorded_csv_text():
    """A small, well-formed data.csv in the capture node's format (two captures, three cameras)."""
    rows = [["capture_id", "timestamp", "camera_label", "image_file"]]
    for cap, n in ((0, 4), (1, 3)):
        for i in range(n):
            ts = 1565270000000 + 1000 * cap + 33 * i
            for cam in ("C", "LL", "RR"):
                rows.append([cap, ts, cam, "ab%02d-%d_%s.jpg" % (cap, ts, cam)])
    return "\n".join(",".join(str(v) for v in r) for r in rows) + "\n"


def golden_recorded(ref_root):
    """The reference's own reader (MediaPlayer/model.py) on that file."""
    import importlib.util
    import json
    import tempfile
    spec = importlib.util.spec_from_file_location("ref_model", os.path.join(ref_root, "MediaPlayer", "model.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    text = recorded_csv_text()
    with open(os.path.join(OUT, "recorded_data.csv"), "w") as f:
        f.write(text)
    with tempfile.TemporaryDirectory() as d:
        with open(os.path.join(d, "data.csv"), "w") as f:
            f.write(text)
        r = mod.data_reader()
        r.load_data(d)
    with open(os.path.join(OUT, "recorded_reference.json"), "w") as f:
        json.dump({"camera_labels": r.camera_labels, "timestamps": r.timestamps, "images": r.images,
                   "line_count": r.line_count}, f, indent=1)


if __name__ == "__main__":
    ref_root = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
    os.makedirs(OUT, exist_ok=True)
    golden_utils(ref_root)
    golden_warp()
    golden_chain()
    golden_chain_ref()
    golden_match()
    golden_resize()
    golden_prewarp(ref_root)
    golden_recorded(ref_root)
    with open(os.path.join(OUT, "README.md"), "w") as f:
        f.write("Fixtures written by scripts/make_golden.py (cv2 %s, numpy %s).\n"
                "utils_reference.npz comes from the reference's own Calibration_Utils/Utils.py;\n"
                "chain_ref.npz from the reference's own StitcherClass.py (Stitcher.calibrate_stitcher + stitch, made\n"
                "importable by oracle/build_ref.py) on injected stage homographies;\n"
                "recorded_reference.json is what the reference's MediaPlayer/model.py data_reader parses from\n"
                "recorded_data.csv; the others come from cv2 driven as PostScripts/Stitcher/StitcherClass.py\n"
                "drives it.\n"
                % (cv2.__version__, np.__version__))
    for n in sorted(os.listdir(OUT)):
        print(n, os.path.getsize(os.path.join(OUT, n)))
