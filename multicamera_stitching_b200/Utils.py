"""Calibration helper math used by the stitching path.

Same names, argument meaning and return conventions as the three functions of
the reference's ``PostScripts/Calibration_Utils/Utils.py`` that feed the hot
path:

* ``get_projection_point_dst``  - Utils.py:23-37
* ``get_projection_point_src``  - Utils.py:39-54
* ``CalculateProjectionMatrix`` - Utils.py:111-129

The drawing / GUI helpers of that file are out of scope.
"""
import numpy as np
import cv2


def _project_truncated(M, pt):
    # Homogeneous projection followed by Python ``int()`` truncation (toward
    # zero).  The truncation is part of the reference's canvas geometry
    # (Utils.py:33-35) and must not be "fixed" into a rounding.
    v = np.matmul(M, pt)
    v = v / v[2]
    return [int(v[0]), int(v[1])]


def get_projection_point_dst(pt_src, M):
    """Project ``pt_src`` (x, y, 1) from the original view into the surface
    projection space with ``M``; returns ``[int(x), int(y)]``."""
    return _project_truncated(M, pt_src)


def get_projection_point_src(coords_dst, INVM):
    """Project ``coords_dst`` (x, y, 1) back into the original view with the
    inverse matrix ``INVM``; returns ``[int(x), int(y)]``."""
    return _project_truncated(INVM, coords_dst)


def CalculateProjectionMatrix(src_pts, dst_pts):
    """Four-point projection matrix and its inverse: ``(M, INVM)``.

    ``M`` comes from ``cv2.getPerspectiveTransform`` on float32 points, the
    inverse from ``numpy.linalg.inv`` - as in the reference.
    """
    src = np.array(src_pts, dtype=np.float32)
    dst = np.array(dst_pts, dtype=np.float32)
    M = cv2.getPerspectiveTransform(src=src, dst=dst)
    try:
        INVM = np.linalg.inv(M)
    except IOError:
        return None, None
    return M, INVM
