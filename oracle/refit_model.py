"""Numpy model of the inlier refit behind ``cv2.findHomography(..., RANSAC)`` - normalised DLT on the inliers
(the structure of OpenCV's HomographyEstimatorCallback::runKernel) + Levenberg-Marquardt on the reprojection error
(HomographyRefineCallback) - as ``StitcherBase.matchKeypoints`` reaches it (StitcherClass.py:443-444).

Test infrastructure only: the checker of the library's ``mcs_refit_homography`` (csrc/mcs_refit.cu), which follows
it statement by statement.  Pinned by ``tests/test_refit.py`` against ``cv2.findHomography`` itself (least-squares
method, same points) within the reprojection tolerance written there; H is not expected bit-exact with OpenCV.
"""
import numpy as np

REFINE_ITERS = 10            # cv2 refines the inlier fit with 10 LM iterations


def _normalise(p):
    c = p.mean(axis=0)
    d = np.abs(p - c).mean(axis=0)
    d = np.where(d > 1e-12, d, 1.0)
    s = 1.0 / d
    T = np.array([[s[0], 0, -c[0] * s[0]], [0, s[1], -c[1] * s[1]], [0, 0, 1.0]])
    return (p - c) * s, T


def fit_homography_dlt(a, b):
    """Least-squares homography a -> b (normalised DLT, the structure of
    OpenCV's HomographyEstimatorCallback::runKernel)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    an, Ta = _normalise(a)
    bn, Tb = _normalise(b)
    n = len(a)
    A = np.zeros((2 * n, 9))
    A[0::2, 0:2] = an
    A[0::2, 2] = 1
    A[0::2, 6:8] = -bn[:, 0:1] * an
    A[0::2, 8] = -bn[:, 0]
    A[1::2, 3:5] = an
    A[1::2, 5] = 1
    A[1::2, 6:8] = -bn[:, 1:2] * an
    A[1::2, 8] = -bn[:, 1]
    _, _, vt = np.linalg.svd(A.T @ A)
    Hn = vt[-1].reshape(3, 3)
    H = np.linalg.inv(Tb) @ Hn @ Ta
    return H / H[2, 2]


def refine_homography(H, a, b, iters=REFINE_ITERS):
    """Levenberg-Marquardt on the reprojection error over 8 parameters, the
    role of OpenCV's HomographyRefineCallback.  The Jacobian rows of a point are
    (x, y, 1, 0, 0, 0, -x u, -y u) / w and (0, 0, 0, x, y, 1, -x v, -y v) / w: they are held as one
    2n x 8 array that is rebuilt in place per iteration."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    h = (H / H[2, 2]).ravel()[:8].copy()
    n = len(a)
    ax, ay, bx, by = a[:, 0], a[:, 1], b[:, 0], b[:, 1]
    J = np.zeros((2 * n, 8))
    err = np.empty(2 * n)

    def residual(h8, out):
        w = h8[6] * ax + h8[7] * ay + 1.0
        x = (h8[0] * ax + h8[1] * ay + h8[2]) / w
        y = (h8[3] * ax + h8[4] * ay + h8[5]) / w
        out[:n] = x - bx
        out[n:] = y - by
        return x, y, w

    lam = 1e-3
    x, y, w = residual(h, err)
    cost = float(err @ err)
    err2 = np.empty(2 * n)
    for _ in range(iters):
        iw = 1.0 / w
        xw, yw = ax * iw, ay * iw
        J[:n, 0] = xw
        J[:n, 1] = yw
        J[:n, 2] = iw
        J[:n, 6] = -xw * x
        J[:n, 7] = -yw * x
        J[n:, 3] = xw
        J[n:, 4] = yw
        J[n:, 5] = iw
        J[n:, 6] = -xw * y
        J[n:, 7] = -yw * y
        JtJ = J.T @ J
        g = J.T @ err
        dg = np.diag(np.diag(JtJ))
        improved = False
        for _try in range(6):
            try:
                step = np.linalg.solve(JtJ + lam * dg, -g)
            except np.linalg.LinAlgError:
                lam *= 10
                continue
            x2, y2, w2 = residual(h + step, err2)
            cost2 = float(err2 @ err2)
            if cost2 < cost:
                converged = cost - cost2 <= 1e-9 * cost     # the fit has stopped moving
                h, x, y, w, cost = h + step, x2, y2, w2, cost2
                err, err2 = err2, err
                lam = max(lam * 0.1, 1e-12)
                improved = not converged
                break
            lam *= 10
        if not improved or cost < 1e-18:
            break
    return np.append(h, 1.0).reshape(3, 3)


def refit(ptsA, ptsB, H0, status, iters=REFINE_ITERS):
    """What the product does with a RANSAC winner ``H0`` and its inlier mask ``status``."""
    inl = np.asarray(status).ravel().astype(bool)
    H = np.asarray(H0, dtype=np.float64).reshape(3, 3)
    if inl.sum() >= 4:
        if inl.sum() > 4:
            H = fit_homography_dlt(ptsA[inl], ptsB[inl])
        H = refine_homography(H, ptsA[inl], ptsB[inl], iters)
    return H
