"""ORACLE (test infrastructure, never imported by the product package).

CPU restatement of the per-camera pre-warp both callers of the reference apply
in front of the stitcher, driving OpenCV exactly
the way they do, plus a transparent integer model of ``cv2.remap`` with
fixed-point maps so the GPU REMAP layers can be checked against a readable
specification.

Where the arithmetic lives: OpenCV (third party, unpinned by the reference;
opencv-python 4.13.0 here).  The reference has no tests or golden vectors for
this path: the model is pinned by bit-equality with
``cv2.remap`` / ``cv2.undistort`` in ``tests/test_oracle_prewarp.py``.
"""
import cv2
import numpy as np

from . import warp_model


def prewarp(image, intrinsic_calibration, extrinsic_calibration):
    """MediaPlayer/view.py:378-388 (the node applies the first half,
    video_mapping_node.py:155-158): undistort when the intrinsic matrix is
    known, then the extrinsic bird's-eye projection when ``M`` is known."""
    if intrinsic_calibration["mtx"] is not None:
        image = cv2.undistort(src=image, cameraMatrix=intrinsic_calibration["mtx"],
                              distCoeffs=intrinsic_calibration["dist"])
        if extrinsic_calibration["M"] is not None:
            image = cv2.warpPerspective(src=image, M=extrinsic_calibration["M"],
                                        dsize=extrinsic_calibration["dst_size"])
    return image


def remap_fixed_point(src, xy, frac):
    """Model of ``cv2.remap(src, xy, frac, cv2.INTER_LINEAR)`` (BORDER_CONSTANT 0)
    for uint8 ``src`` with a CV_16SC2 ``xy`` map and a CV_16UC1 ``frac`` map:
    ``frac = (fy << 5) | fx`` indexes the 32 x 32 bilinear table whose weights
    are ``32 * {(32-fy)(32-fx), (32-fy)fx, fy(32-fx), fy*fx}`` (sum 32768) and
    ``out = (sum w*p + 16384) >> 15`` - the interpolation of
    ``warp_model.sample_fixed_point`` at ``X = 32*x + fx, Y = 32*y + fy``."""
    xy = np.asarray(xy).astype(np.int64)
    frac = np.zeros(xy.shape[:2], dtype=np.int64) if frac is None else np.asarray(frac).astype(np.int64)
    X = xy[..., 0] * 32 + (frac & 31)
    Y = xy[..., 1] * 32 + ((frac >> 5) & 31)
    out, _ = warp_model.sample_fixed_point(np.asarray(src), X, Y)
    return out if np.asarray(src).ndim == 3 else out[:, :, 0]
