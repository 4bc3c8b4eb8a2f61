"""CPU: the Python surface mirrors the reference's Stitcher / StitcherBase
(names, defaults, label scheme, persistence, degraded-mode returns)."""
import inspect
import os
import pickle

import numpy as np
import pytest

from helpers import synthetic_chain
from multicamera_stitching_b200 import Stitcher, StitcherBase, synthetic
from multicamera_stitching_b200 import Utils
from oracle import stitcher_ref


def test_labels_follow_reference_scheme():
    images = synthetic.make_frames(4, 36, 64)
    st = Stitcher(images)
    assert list(st.img_labels) == ["CAM1", "CAM2", "CAM3", "CAM4"]
    assert st.stitcher_labels == ["(CAM1&CAM2)", "((CAM1&CAM2)&CAM3)", "(((CAM1&CAM2)&CAM3)&CAM4)"]
    assert st.stitcher_labels == stitcher_ref.stitcher_labels(stitcher_ref.sorted_labels(images))
    assert [s.sid for s in st.stitchers] == st.stitcher_labels
    assert all(isinstance(s, StitcherBase) and s.cachedAH is None for s in st.stitchers)


def test_signatures_match_reference():
    def params(f):
        return [(p.name, p.default) for p in inspect.signature(f).parameters.values() if p.name != "self"]
    E = inspect.Parameter.empty
    assert params(Stitcher.__init__) == [("images_dic", E), ("super_mode", False)]
    assert params(Stitcher.stitch) == [("images_dic", E), ("draw_descriptors", False)]
    assert params(Stitcher.calibrate_stitcher) == [("images_dic", E), ("save", True), ("save_path", "")]
    assert params(Stitcher.save_stitcher) == [("save_path", E)]
    assert params(Stitcher.load_stitcher) == [("load_path", E)]
    assert params(StitcherBase.__init__) == [("sid", None), ("super_mode", False)]
    assert params(StitcherBase.stitch) == [("images", E), ("draw_descriptors", False)]
    assert params(StitcherBase.calibrate) == [("images", E), ("ratio", 0.75), ("reprojThresh", 4.0),
                                              ("xoffset", 10), ("yoffset", 10)]
    assert params(StitcherBase.matchKeypoints) == [("kpsA", E), ("kpsB", E), ("featuresA", E), ("featuresB", E),
                                                   ("ratio", 0.75), ("reprojThresh", 4.0)]
    assert params(StitcherBase.detectAndDescribe) == [("image", E)]
    assert params(StitcherBase.draw_descriptors) == [("img_src", E)]
    for name in ("params_to_list", "params_to_array", "reset", "__str__"):
        assert hasattr(StitcherBase, name)


def test_degraded_returns_need_no_gpu():
    images = synthetic.make_frames(3, 36, 64)
    st = Stitcher(images)
    # fewer images than labels -> the last label's image comes back untouched (reference :126-128)
    fewer = {k: images[k] for k in ("CAM2", "CAM3")}
    assert st.stitch(fewer) is images["CAM3"]
    # uncalibrated chain -> the first image passes through every stage (reference :255-256)
    assert st.stitch(images) is images["CAM1"]
    assert st.stitchers[0].stitch((images["CAM1"], images["CAM2"])) is images["CAM1"]
    assert str(st.stitchers[0]) == "Stitcher:(CAM1&CAM2)| Matches:0| StitcherSize:None"


def test_save_load_round_trip(tmp_path):
    st, states, labels, images = synthetic_chain(3, 72, 128, 3)
    path = str(tmp_path / "Stitcher_config.pkl")
    st.save_stitcher(path)
    assert os.path.isfile(path)
    # matrices are arrays again after saving (reference :146-148)
    assert isinstance(st.stitchers[0].cachedAH, np.ndarray)
    fresh = Stitcher(images)
    loaded = fresh.load_stitcher(path)          # callers rebind the return value
    assert loaded is not fresh
    for a, b in zip(loaded.stitchers, st.stitchers):
        assert np.array_equal(a.cachedAH, b.cachedAH) and a.ABSize == b.ABSize and a.Bpts == b.Bpts
        assert isinstance(a.cachedAH, np.ndarray)
    # missing file -> warning, the same (uncalibrated) object comes back
    assert fresh.load_stitcher(str(tmp_path / "nope.pkl")) is fresh
    # the engine (GPU handles) is never pickled
    st.__dict__["_engine"] = object()
    assert "_engine" not in pickle.loads(pickle.dumps(st)).__dict__


def test_pickle_written_under_reference_module_name_loads(tmp_path):
    """Robots hold pickles whose classes live in module ``StitcherClass``."""
    st, _, _, images = synthetic_chain(2, 72, 128, 3)
    for s in st.stitchers:
        s.params_to_list()
    data = pickle.dumps(st, 2).replace(b"multicamera_stitching_b200.StitcherClass", b"StitcherClass")
    for s in st.stitchers:
        s.params_to_array()
    p = tmp_path / "legacy.pkl"
    p.write_bytes(data)
    loaded = Stitcher(images).load_stitcher(str(p))
    assert np.array_equal(loaded.stitchers[0].cachedAH, st.stitchers[0].cachedAH)


def test_python2_pickle_fixture_loads():
    """tests/golden/stitcher_py2.pkl holds the opcode stream CPython 2.7 + numpy write for the reference's
    save_stitcher (protocol 2, str opcodes, an OBJ and a NEWOBJ instance, numpy arrays with raw byte-string
    states; scripts/make_py2_pickle.py assembles it by hand).  It must load into this module's classes with
    str labels, float64 matrices and the class-level defaults the reference's objects never had."""
    import pickletools
    path = os.path.join(os.path.dirname(__file__), "golden", "stitcher_py2.pkl")
    ops = {op.name for op, _, _ in pickletools.genops(open(path, "rb").read())}
    assert "SHORT_BINSTRING" in ops and "OBJ" in ops and "NEWOBJ" in ops and not ops & {"BINUNICODE", "SHORT_BINUNICODE"}
    st, _, labels, images = synthetic_chain(3, 72, 128, 3)
    loaded = Stitcher(images).load_stitcher(path)
    assert isinstance(loaded, Stitcher) and [str(l) for l in loaded.img_labels] == list(labels)
    assert all(isinstance(l, str) for l in loaded.img_labels.tolist())
    assert loaded.stitcher_labels == st.stitcher_labels and loaded.feather_log2 == 0
    for a, b in zip(loaded.stitchers, st.stitchers):
        assert isinstance(a, StitcherBase) and isinstance(a.cachedAH, np.ndarray) and a.cachedAH.dtype == np.float64
        for f in ("cachedAH", "cachedAINVH", "cachedBH", "cachedBINVH"):
            assert np.array_equal(getattr(a, f), getattr(b, f))
        assert tuple(a.ABSize) == tuple(b.ABSize) and [tuple(p) for p in a.Bpts] == [tuple(p) for p in b.Bpts]
        assert tuple(a.BimgSize) == tuple(b.BimgSize) and tuple(a.AimgSize) == tuple(b.AimgSize)
        assert a.x_limits == b.x_limits and a.y_limits == b.y_limits and a.sid == b.sid
        assert a.descriptor == "ORB" and a.nfeatures == 2000       # defaults a reference pickle lacks
    # the loaded chain flattens to the same plan as the one it was written from
    from multicamera_stitching_b200.plan import stage_signature
    assert [stage_signature(s) for s in loaded.stitchers] == [stage_signature(s) for s in st.stitchers]


def test_reset_clears_every_field():
    st, _, _, _ = synthetic_chain(2, 72, 128, 3)
    s = st.stitchers[0]
    assert s.cachedAH is not None
    s.reset()
    for f in ("cachedBH", "cachedBINVH", "Bpts", "cachedAH", "cachedAINVH", "Apts", "matches", "status",
              "ABSize", "x_limits", "y_limits", "AimgSize", "BimgSize"):
        assert getattr(s, f) is None


def test_utils_projection_helpers():
    M, INVM = Utils.CalculateProjectionMatrix([(0, 0), (100, 0), (100, 50), (0, 50)],
                                              [(10, 5), (120, 9), (115, 70), (4, 66)])
    assert M.shape == (3, 3) and np.allclose(M @ INVM, np.eye(3), atol=1e-9)
    assert Utils.get_projection_point_dst((0, 0, 1), M) == [10, 5]
    assert Utils.get_projection_point_dst((100, 50, 1), M) in ([115, 70], [114, 69], [115, 69], [114, 70])
    # int() truncation toward zero, not rounding (Utils.py:33-35)
    T = np.array([[1, 0, -0.9], [0, 1, 2.9], [0, 0, 1.0]])
    assert Utils.get_projection_point_dst((0, 0, 1), T) == [0, 2]
    assert Utils.get_projection_point_src((0, 0, 1), T) == [0, 2]
    assert stitcher_ref.projection_point_dst((0, 0, 1), T) == [0, 2]
