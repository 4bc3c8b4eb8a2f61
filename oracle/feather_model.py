"""ORACLE (test infrastructure, never imported by the product package).

Feather blend mode of the compositing chain - SURVEY.md section 8 row f1.  The
reference has NO feather blend: its "blend" is the rectangle overwrite
``dst[By:By+hB, Bx:Bx+wB] = imageB`` (PostScripts/Stitcher/StitcherClass.py:
240-241).  This mode is an extension behind the same ``stitch`` call, so its
parity is **unpinned** by construction; the definition below is the
specification, chosen so that it reduces bit-exactly to the reference's
overwrite when the feather is one pixel wide (``feather_log2 == 0``).

Definition, per stage (same walk as ``stitcher_ref.stitch_pair``):

  warped  = cv2.warpPerspective(imageA, cachedAH, ABSize)            (:239)
  touched = pixels of ``warped`` with at least one bilinear tap inside imageA
            (integer part of the 1/32-px coordinates of oracle/warp_model.py)
  inside the pasted rectangle of imageB, with F = 2**feather_log2 and
    a = min(F, 1 + distance in pixels to the nearest edge of the rectangle),
  the result is   (a*imageB + (F - a)*warped + F/2) >> feather_log2   where
  ``touched`` and a < F, and imageB elsewhere; outside the rectangle it is
  ``warped``.  Every stage rounds to uint8 like the reference's chain does.

Weight maps generalise the ramp: ``weights[k]`` (one optional ``hB x wB`` integer array per
stage, values 0 .. F, larger values count as F) replaces ``a`` of stage k; the ramp is the
instance ``a = min(F, 1 + distance)``.  A map that is F everywhere is the reference's overwrite;
with ``feather_log2 == 0`` a {0, 1} map is a mask (0 shows the warped camera where it touches its
source).  The super-mode crop (:248-251) is applied after the blended paste, exactly where
``stitcher_ref.stitch_pair`` applies it after the overwrite: distances and maps refer to the
pasted rectangle, whatever later crops cut away.
"""
import numpy as np
import cv2

from . import warp_model


def touched_mask(M, src_hw, dsize):
    """True where cv2.warpPerspective(src, M, dsize) reads at least one tap inside src."""
    W, H = dsize
    Mi = warp_model.invert3x3(np.asarray(M, dtype=np.float64))
    X, Y = warp_model.fixed_point_coords(Mi, np.arange(W), np.arange(H))
    sx = np.clip(X >> 5, -32768, 32767)
    sy = np.clip(Y >> 5, -32768, 32767)
    h, w = src_hw
    okx = ((sx >= 0) & (sx < w)) | ((sx + 1 >= 0) & (sx + 1 < w))
    oky = ((sy >= 0) & (sy < h)) | ((sy + 1 >= 0) & (sy + 1 < h))
    return okx & oky


def ramp_weights(shape_hw, feather_log2):
    """The distance ramp as a weight map: ``min(F, 1 + distance to the nearest edge)``."""
    hB, wB = shape_hw
    F = 1 << feather_log2
    yy, xx = np.mgrid[0:hB, 0:wB]
    a = np.minimum(np.minimum(xx, wB - 1 - xx), np.minimum(yy, hB - 1 - yy)) + 1
    return np.minimum(a, F).astype(np.int32)


def feather_pair(st, images, feather_log2, weights=None):
    imageB, imageA = images
    if st["cachedAH"] is None:
        return imageB
    F = 1 << feather_log2
    W, H = int(st["ABSize"][0]), int(st["ABSize"][1])
    bx, by = int(st["Bpts"][0][0]), int(st["Bpts"][0][1])
    hB, wB = imageB.shape[:2]
    warped = cv2.warpPerspective(src=imageA, M=st["cachedAH"], dsize=(W, H))
    touched = touched_mask(st["cachedAH"], imageA.shape[:2], (W, H))
    if weights is None:
        a = ramp_weights((hB, wB), feather_log2)
    else:
        a = np.minimum(np.asarray(weights).astype(np.int32), F)
        if a.shape != (hB, wB):
            raise ValueError("weight map of shape %r for a %d x %d paste" % (a.shape, hB, wB))
    outer = warped[by:by + hB, bx:bx + wB].astype(np.int32)
    inner = imageB.astype(np.int32)
    t = touched[by:by + hB, bx:bx + wB] & (a < F)
    if imageB.ndim == 3:
        a3, t3 = a[..., None], t[..., None]
    else:
        a3, t3 = a, t
    blend = (a3 * inner + (F - a3) * outer + (F >> 1)) >> feather_log2
    dst = warped.copy()
    dst[by:by + hB, bx:bx + wB] = np.where(t3, blend, inner).astype(np.uint8)
    if st["super_mode"]:   # StitcherClass.py:248-251
        dst = dst[st["y_limits"][0]:st["y_limits"][1], st["x_limits"][0]:st["x_limits"][1]]
    return dst


def feather_chain(states, img_labels, images_dic, feather_log2, weights=None):
    """The chain of ``stitcher_ref.stitch_chain`` with the blended paste; ``weights`` = one optional
    map per stage."""
    dst = None
    for i in range(len(img_labels) - 1):
        pair = (images_dic[img_labels[0]] if i == 0 else dst, images_dic[img_labels[i + 1]])
        dst = feather_pair(states[i], pair, feather_log2, None if weights is None else weights[i])
    return dst if dst is not None else images_dic[img_labels[-1]]
