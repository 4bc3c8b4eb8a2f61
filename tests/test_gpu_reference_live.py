"""The CUDA path against the REFERENCE'S OWN ``Stitcher`` class, run live on the GPU box.

``oracle/_ref`` (the reference's ``StitcherClass.py`` made importable by ``oracle/build_ref.py``, with the stand-in
logging module and a byte copy of the reference's ``Utils.py``) is a build output that travels with the repository
snapshot; the reference tree itself does not.  Where that output is present, the reference class is calibrated by
its own ``calibrate_stitcher`` with the stage homographies of the synthetic rigs and its ``stitch(images_dic)`` is
held against the product's, bit for bit.  (Where it is absent the committed fixture tests/golden/chain_ref.npz,
made by the same classes, carries the pin: tests/test_golden.py.)"""
import numpy as np
import pytest

from multicamera_stitching_b200 import synthetic
from oracle import build_ref

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    mod = build_ref.load()
    if mod is None:
        pytest.skip("oracle/_ref was not built (python oracle/build_ref.py where the reference tree is present)")
    return mod


@pytest.mark.parametrize("n,h,w,super_mode,kind", [
    (3, 720, 1280, False, "smooth"),     # BASELINE config 1
    (3, 720, 1280, False, "noise"),
    (6, 270, 480, False, "noise"),
    (4, 180, 320, True, "noise"),        # super mode: the reference returns the cropped view
    (8, 135, 240, False, "smooth"),
])
def test_cuda_chain_equals_the_reference_class(cuda_device, ref, n, h, w, super_mode, kind):
    st, homographies, labels, images = synthetic.synthetic_stitcher(n, h, w, 3, super_mode=super_mode, kind=kind)
    rs = build_ref.calibrated_stitcher(ref, images, homographies, super_mode=super_mode)
    for frame_index in (0, 1):
        frames = synthetic.make_frames(n, h, w, 3, frame_index=frame_index, kind=kind)
        want = rs.stitch(frames)
        got = st.stitch(frames)
        assert got.shape == want.shape and got.dtype == want.dtype
        assert np.array_equal(got, want)


def test_cuda_pair_equals_the_reference_class(cuda_device, ref):
    st, homographies, labels, images = synthetic.synthetic_stitcher(2, 200, 300, 3, kind="noise")
    rs = build_ref.calibrated_stitcher(ref, images, homographies)
    pair = (images[labels[0]], images[labels[1]])
    assert np.array_equal(st.stitchers[0].stitch(pair), rs.stitchers[0].stitch(images=pair))
