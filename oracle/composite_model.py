"""ORACLE (test infrastructure, never imported by the product package).

One-pass CPU model of the fused compositing kernel: consumes the *same layer
table* the CUDA plan is built from (``multicamera_stitching_b200.plan``) and
evaluates it with the integer warp model of ``oracle/warp_model.py``.  It lets
the CPU-only test-suite check the plan flattening against the reference's
sequential chain (``oracle/stitcher_ref.stitch_chain`` = StitcherClass.py:
114-136, :211-256) without a GPU, and gives the GPU tests a second,
transparent checker beside cv2 itself.
"""
import numpy as np

from . import warp_model

LAYER_COPY = 0
LAYER_WARP = 1


def composite(layers, out_w, out_h, frames, channels=None):
    """``layers``: iterable of objects with fields ``cam kind H ox oy rect
    src_hw`` ordered innermost first; ``frames``: list of camera frames indexed
    by ``layer.cam``.  Returns ``(panorama, owned)`` where ``owned[i]`` is the
    number of output pixels taken from layer ``i`` with at least one tap inside
    its source."""
    f0 = np.asarray(frames[layers[0].cam])
    two_d = f0.ndim == 2
    C = 1 if two_d else f0.shape[2]
    out = np.zeros((out_h, out_w, C), dtype=np.uint8)
    taken = np.zeros((out_h, out_w), dtype=bool)
    owned = []
    for l in layers:
        x0, y0, x1, y1 = [int(v) for v in l.rect]
        x0, y0, x1, y1 = max(0, x0), max(0, y0), min(out_w, x1), min(out_h, y1)
        if x1 <= x0 or y1 <= y0:
            owned.append(0)
            continue
        free = ~taken[y0:y1, x0:x1]
        src = np.asarray(frames[l.cam])
        src3 = src[:, :, None] if src.ndim == 2 else src
        if l.kind == LAYER_COPY:
            vals = src3[y0 - l.oy:y1 - l.oy, x0 - l.ox:x1 - l.ox]
            touched = np.ones(free.shape, dtype=bool)
        else:
            Mi = warp_model.invert3x3(l.H)
            if Mi is None:
                Mi = np.zeros((3, 3))
            X, Y = warp_model.fixed_point_coords(Mi, np.arange(x0 - l.ox, x1 - l.ox),
                                                 np.arange(y0 - l.oy, y1 - l.oy))
            vals, touched = warp_model.sample_fixed_point(src3, X, Y)
        win = out[y0:y1, x0:x1]
        win[free] = vals[free]
        owned.append(int((free & touched).sum()))
        taken[y0:y1, x0:x1] = True
    return (out[:, :, 0] if two_d else out), owned
