"""Flatten a calibrated stitcher chain into the layer table of one compositing
plan (host logic, no GPU needed).

The reference composites with N-1 sequential stages, each of which warps the
next camera into a fresh canvas and pastes the running canvas over it
(StitcherClass.py:131-136, :237-251).  Because every paste is an axis-aligned
rectangle overwrite, the final value of an output pixel is decided by the
*innermost* pasted rectangle that contains it.  ``flatten_chain`` walks the
stage states once and produces, for every camera, its rectangle and canvas
origin in final-panorama coordinates; ``mcs_plan_create`` (include/mcs.h) turns
that table into the plan the fused kernel executes.
"""
from collections import namedtuple

import numpy as np

LAYER_COPY = 0
LAYER_WARP = 1
LAYER_REMAP = 2   # coordinates from a fixed-point map pair (prewarp.py), not from a homography

# ``map``: for LAYER_REMAP the (int16 H x W x 2, uint16 H x W) pair cv2.convertMaps produces
# ``paste``: the layer's rectangle as the next stage pasted it (never cut by later super-mode crops,
# only shifted with them) - what the feather blend measures its distances against; None = ``rect``
Layer = namedtuple("Layer", "cam kind H ox oy rect src_hw map paste", defaults=(None, None))
FlatPlan = namedtuple("FlatPlan", "layers out_w out_h channels ndim")


class PlanUnsupported(Exception):
    """The chain needs something the fused one-pass plan cannot express."""


def _field(st, name):
    return st[name] if isinstance(st, dict) else getattr(st, name)


def _numpy_window(start, stop, size):
    """Start/stop of ``a[start:stop]`` on an axis of length ``size`` with
    Python's slice semantics (negative indices wrap, out-of-range clamps)."""
    lo, hi, _ = slice(start, stop).indices(size)
    return lo, max(lo, hi)


def stage_signature(st):
    """Hashable snapshot of everything of a stage that shapes the plan."""
    H = _field(st, "cachedAH")
    if H is None:
        return None
    Bpts = _field(st, "Bpts")
    xl, yl = _field(st, "x_limits"), _field(st, "y_limits")
    return (np.asarray(H, dtype=np.float64).tobytes(),
            int(Bpts[0][0]), int(Bpts[0][1]),
            tuple(int(v) for v in _field(st, "ABSize")),
            tuple(_field(st, "AimgSize")), tuple(_field(st, "BimgSize")),
            bool(_field(st, "super_mode")),
            tuple(xl) if xl is not None else None, tuple(yl) if yl is not None else None)


def flatten_chain(stages, cam_shapes):
    """``stages``: the N-1 ``StitcherBase`` objects (or dicts with the same
    field names) in chain order.  ``cam_shapes``: ``ndarray.shape`` of the N
    camera frames in label order.  Returns a :class:`FlatPlan` whose layers are
    ordered innermost first, or ``None`` when no stage is calibrated (the
    reference then hands the first frame straight through,
    StitcherClass.py:255-256)."""
    shape0 = tuple(int(v) for v in cam_shapes[0])
    if len(shape0) not in (2, 3):
        raise PlanUnsupported("frames must be HxW or HxWxC, got shape %r" % (shape0,))
    ndim = len(shape0)
    channels = 1 if ndim == 2 else shape0[2]
    tail = shape0[2:]
    layers = [Layer(0, LAYER_COPY, None, 0, 0, (0, 0, shape0[1], shape0[0]), (shape0[0], shape0[1]))]
    cur_shape = shape0
    any_calibrated = False

    for s, st in enumerate(stages):
        H = _field(st, "cachedAH")
        if H is None:
            continue  # uncalibrated stage: the running canvas passes through
        any_calibrated = True
        shapeA = tuple(int(v) for v in cam_shapes[s + 1])
        # Only the geometry has to agree here.  The reference compares whole shape tuples and
        # resizes on any difference (StitcherClass.py:226-233); a frame that differs from the
        # calibrated one in its channel layout only is "resized" to its own size, i.e. unchanged.
        # Frames of another height / width are resized by the caller before planning
        # (StitcherClass._composite), which also splits the chain where a composited canvas
        # would have to be resized.
        if cur_shape[:2] != tuple(_field(st, "BimgSize"))[:2]:
            raise PlanUnsupported(
                "stage %d: running canvas has shape %r but was calibrated for %r (resizing a "
                "composited canvas is not expressible in one pass)" % (s, cur_shape, tuple(_field(st, "BimgSize"))))
        if shapeA[:2] != tuple(_field(st, "AimgSize"))[:2]:
            raise PlanUnsupported("stage %d: frame shape %r != calibrated %r"
                                  % (s, shapeA, tuple(_field(st, "AimgSize"))))
        if shapeA[2:] != tail:
            raise PlanUnsupported("stage %d: channel layout %r differs from %r" % (s, shapeA[2:], tail))
        Bpts = _field(st, "Bpts")
        bx, by = int(Bpts[0][0]), int(Bpts[0][1])
        W, Hh = int(_field(st, "ABSize")[0]), int(_field(st, "ABSize")[1])
        hB, wB = cur_shape[0], cur_shape[1]
        # dst[by:by+hB, bx:bx+wB] = imageB  (numpy raises if the window is not hB x wB)
        y0, y1 = _numpy_window(by, by + hB, Hh)
        x0, x1 = _numpy_window(bx, bx + wB, W)
        if (y1 - y0, x1 - x0) != (hB, wB):
            raise ValueError("could not broadcast input array from shape %r into shape %r"
                             % ((hB, wB) + tail, (y1 - y0, x1 - x0) + tail))
        def shifted(r, dx, dy):
            return None if r is None else (r[0] + dx, r[1] + dy, r[2] + dx, r[3] + dy)
        layers = [l._replace(ox=l.ox + x0, oy=l.oy + y0, rect=shifted(l.rect, x0, y0), paste=shifted(l.paste, x0, y0))
                  for l in layers]
        # the outermost layer so far is the canvas this stage pastes: remember the rectangle it is pasted as
        layers[-1] = layers[-1]._replace(paste=layers[-1].rect)
        layers.append(Layer(s + 1, LAYER_WARP, np.array(H, dtype=np.float64).reshape(3, 3), 0, 0,
                            (0, 0, W, Hh), (shapeA[0], shapeA[1])))
        cw, ch = W, Hh
        if _field(st, "super_mode"):
            yl, xl = _field(st, "y_limits"), _field(st, "x_limits")
            cy0, cy1 = _numpy_window(yl[0], yl[1], Hh)
            cx0, cx1 = _numpy_window(xl[0], xl[1], W)
            cw, ch = cx1 - cx0, cy1 - cy0
            clipped = []
            for l in layers:
                r = (max(0, l.rect[0] - cx0), max(0, l.rect[1] - cy0),
                     min(cw, l.rect[2] - cx0), min(ch, l.rect[3] - cy0))
                r = (r[0], r[1], max(r[0], r[2]), max(r[1], r[3]))
                clipped.append(l._replace(ox=l.ox - cx0, oy=l.oy - cy0, rect=r,
                                          paste=None if l.paste is None else
                                          (l.paste[0] - cx0, l.paste[1] - cy0, l.paste[2] - cx0, l.paste[3] - cy0)))
            layers = clipped
        cur_shape = (ch, cw) + tail

    if not any_calibrated:
        return None
    return FlatPlan(layers, cur_shape[1], cur_shape[0], channels, ndim)


def plan_tables(flat):
    """Arrays in the layout ``mcs_plan_create`` takes."""
    n = len(flat.layers)
    kind = np.array([l.kind for l in flat.layers], dtype=np.int32)
    src_hw = np.array([l.src_hw for l in flat.layers], dtype=np.int32).reshape(n, 2)
    origin = np.array([(l.ox, l.oy) for l in flat.layers], dtype=np.int32).reshape(n, 2)
    rect = np.array([l.rect for l in flat.layers], dtype=np.int32).reshape(n, 4)
    fwd = np.zeros((n, 9), dtype=np.float64)
    for i, l in enumerate(flat.layers):
        fwd[i] = np.eye(3).ravel() if l.H is None else np.asarray(l.H, dtype=np.float64).ravel()
    return kind, src_hw, fwd, origin, rect


def plan_paste(flat):
    """``[n, 4]`` int32 paste rectangles for ``mcs_plan_set_blend`` (a layer that was never pasted - the last
    one - and layers whose rectangle was not cut report their visible rectangle)."""
    return np.array([l.rect if l.paste is None else l.paste for l in flat.layers], dtype=np.int32).reshape(-1, 4)


def plan_maps(flat):
    """Per-layer map pairs for ``mcs_plan_create_maps`` (None when no layer has one)."""
    maps = [l.map for l in flat.layers]
    return maps if any(m is not None for m in maps) else None
