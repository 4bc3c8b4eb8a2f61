"""Shared builders for the tests: synthetic chains in both the product classes
and the oracle's dict states."""
import numpy as np

from multicamera_stitching_b200 import Stitcher, synthetic
from oracle import stitcher_ref


def synthetic_chain(n_cams, h, w, channels=3, super_mode=False, kind="smooth", frame_index=0,
                    xoffset=0, yoffset=0, use_points_first=False):
    """Returns (stitcher, oracle_states, labels, images_dic)."""
    images = synthetic.make_frames(n_cams, h, w, channels, frame_index, kind)
    st = Stitcher(images, super_mode=super_mode)
    labels = list(st.img_labels)
    shapes = [images[l].shape for l in labels]
    # homographies depend on the running canvas width -> calibrate stage by stage
    states = []
    shapeB = tuple(shapes[0])
    for k in range(n_cams - 1):
        cw = shapeB[1]
        if use_points_first and k == 0:
            H = synthetic.homography_from_points(h, w, cw)
        else:
            H = synthetic.make_homography(k, h, w, cw)
        st.stitchers[k].set_homography(H, shapeA=shapes[k + 1], shapeB=shapeB, xoffset=xoffset, yoffset=yoffset)
        ost = stitcher_ref.new_state(sid=str(k), super_mode=super_mode)
        stitcher_ref.geometry_from_homography(ost, H, shapes[k + 1], shapeB, xoffset, yoffset)
        states.append(ost)
        shapeB = st.stitchers[k].result_shape()
    return st, states, labels, images


def compare_u8(got, ref):
    """(max abs diff, fraction exactly equal)."""
    got = np.asarray(got)
    ref = np.asarray(ref)
    assert got.shape == ref.shape, (got.shape, ref.shape)
    d = np.abs(got.astype(np.int16) - ref.astype(np.int16))
    return int(d.max()) if d.size else 0, float((d == 0).mean()) if d.size else 1.0
