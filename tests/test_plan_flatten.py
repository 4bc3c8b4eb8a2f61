"""CPU: the one-pass layer plan (multicamera_stitching_b200.plan) evaluated by
the oracle's integer model must reproduce the reference's sequential chain
(cv2 warp + paste, N-1 stages) bit for bit - this checks the host logic that
feeds mcs_plan_create without needing a GPU."""
import numpy as np
import pytest

from helpers import compare_u8, synthetic_chain
from multicamera_stitching_b200 import Stitcher, synthetic
from multicamera_stitching_b200.plan import PlanUnsupported, flatten_chain, plan_tables
from oracle import composite_model, stitcher_ref


@pytest.mark.parametrize("n,h,w,c,super_mode,off,points", [
    (3, 180, 320, 3, False, 0, True),
    (3, 180, 320, 3, True, 0, False),
    (4, 120, 160, 1, False, 7, False),
    (6, 135, 240, 3, False, 0, False),
    (8, 90, 160, 3, False, 3, True),
    (2, 97, 131, 4, True, 5, False),
])
def test_flat_plan_equals_sequential_chain(n, h, w, c, super_mode, off, points):
    st, states, labels, images = synthetic_chain(n, h, w, c, super_mode=super_mode, kind="noise",
                                                 xoffset=off, yoffset=off, use_points_first=points)
    ref = stitcher_ref.stitch_chain(states, labels, images)
    flat = flatten_chain(st.stitchers, [images[l].shape for l in labels])
    got, owned = composite_model.composite(flat.layers, flat.out_w, flat.out_h, [images[l] for l in labels])
    assert compare_u8(got, ref) == (0, 1.0)
    assert (flat.out_h, flat.out_w) == ref.shape[:2]
    assert len(owned) == n and all(v >= 0 for v in owned)


def test_product_geometry_equals_oracle_geometry():
    st, states, labels, images = synthetic_chain(5, 144, 256, 3, xoffset=4, yoffset=9)
    for a, b in zip(st.stitchers, states):
        assert np.array_equal(a.cachedAH, b["cachedAH"])
        assert np.array_equal(a.cachedBH, b["cachedBH"])
        assert a.Bpts == b["Bpts"] and a.Apts == b["Apts"]
        assert a.ABSize == b["ABSize"]
        assert a.x_limits == b["x_limits"] and a.y_limits == b["y_limits"]
        assert tuple(a.AimgSize) == tuple(b["AimgSize"]) and tuple(a.BimgSize) == tuple(b["BimgSize"])


def test_uncalibrated_stages_pass_through():
    st, states, labels, images = synthetic_chain(4, 90, 160, 3, kind="noise")
    shapes = [images[l].shape for l in labels]
    # nothing calibrated -> no plan
    blank = Stitcher(images)
    assert flatten_chain(blank.stitchers, shapes) is None
    # middle stage uncalibrated: the canvas passes through it; the following stage was calibrated
    # against a different canvas shape, which the one-pass plan must refuse
    st.stitchers[1].reset()
    with pytest.raises(PlanUnsupported):
        flatten_chain(st.stitchers, shapes)
    # last stage uncalibrated: result is the canvas of the first two stages
    st, states, labels, images = synthetic_chain(4, 90, 160, 3, kind="noise")
    st.stitchers[2].reset()
    states[2] = stitcher_ref.new_state()
    flat = flatten_chain(st.stitchers, shapes)
    got, _ = composite_model.composite(flat.layers, flat.out_w, flat.out_h, [images[l] for l in labels])
    assert compare_u8(got, stitcher_ref.stitch_chain(states, labels, images)) == (0, 1.0)


def test_shape_mismatch_is_refused_not_silently_wrong():
    st, states, labels, images = synthetic_chain(3, 90, 160, 3)
    shapes = [images[l].shape for l in labels]
    shapes[2] = (91, 160, 3)
    with pytest.raises(PlanUnsupported):
        flatten_chain(st.stitchers, shapes)


def test_plan_tables_layout():
    st, states, labels, images = synthetic_chain(3, 90, 160, 3, xoffset=2, yoffset=3)
    flat = flatten_chain(st.stitchers, [images[l].shape for l in labels])
    kind, src_hw, fwd, origin, rect = plan_tables(flat)
    assert kind.tolist() == [0, 1, 1] and kind.dtype == np.int32
    assert src_hw.shape == (3, 2) and fwd.shape == (3, 9) and origin.shape == (3, 2) and rect.shape == (3, 4)
    assert np.array_equal(fwd[1].reshape(3, 3), st.stitchers[0].cachedAH)
    # rectangles are nested, innermost first, the last one is the whole panorama
    for i in range(2):
        assert rect[i][0] >= rect[i + 1][0] and rect[i][2] <= rect[i + 1][2]
    assert rect[2].tolist() == [0, 0, flat.out_w, flat.out_h]


def test_synthetic_frames_are_deterministic():
    a = synthetic.make_frame(72, 128, 3, 1, 2)
    b = synthetic.make_frame(72, 128, 3, 1, 2)
    assert np.array_equal(a, b) and a.dtype == np.uint8 and a.flags["C_CONTIGUOUS"]
    assert not np.array_equal(a, synthetic.make_frame(72, 128, 3, 2, 2))
    assert synthetic.make_frame(72, 128, 1, 0, 0, "noise").shape == (72, 128)
