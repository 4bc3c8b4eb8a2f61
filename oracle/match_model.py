"""ORACLE (test infrastructure, never imported by the product package).

Transparent numpy restatement of the descriptor-matching half of
``StitcherBase.matchKeypoints`` (reference PostScripts/Stitcher/
StitcherClass.py:405-448) for binary descriptors:

  :423-424  ``matcher.knnMatch(featuresA, featuresB, 2)``  - brute force, query =
            featuresA, train = featuresB, Hamming norm for uint8 (ORB) descriptors
  :428-433  keep ``m`` iff ``len(m) == 2 and m[0].distance < m[1].distance * ratio``
            -> ``(m[0].trainIdx, m[0].queryIdx)``

The arithmetic lives in OpenCV's BFMatcher (third party, unpinned by the
reference; opencv-python 4.13.0 here).  Its published behaviour, restated:
distance = popcount(a XOR b) summed over the descriptor bytes; the k best train
descriptors per query in ascending distance, ties resolved toward the LOWER
train index (a stable selection); fewer than k train descriptors -> a shorter
list.  ``tests/test_oracle_match.py`` pins this restatement against
``cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch`` itself, including forced ties.
"""
import numpy as np

_POPCOUNT = np.array([bin(i).count("1") for i in range(256)], dtype=np.uint8)


def hamming_matrix(featuresA, featuresB):
    """All-pairs Hamming distances, int32 [len(A), len(B)]."""
    a = np.ascontiguousarray(featuresA, dtype=np.uint8)
    b = np.ascontiguousarray(featuresB, dtype=np.uint8)
    out = np.zeros((len(a), len(b)), dtype=np.int32)
    for lo in range(0, len(a), 256):   # bounded temporaries
        x = a[lo:lo + 256, None, :] ^ b[None, :, :]
        out[lo:lo + 256] = _POPCOUNT[x].sum(axis=2, dtype=np.int32)
    return out


def knn_top2(featuresA, featuresB):
    """``knnMatch(k=2)`` as arrays: ``idx`` [N,2] int32 train indices and ``dist``
    [N,2] int32 distances, best first, -1 where the train set is too small."""
    n, m = len(featuresA), len(featuresB)
    idx = -np.ones((n, 2), dtype=np.int32)
    dist = -np.ones((n, 2), dtype=np.int32)
    if n == 0 or m == 0:
        return idx, dist
    d = hamming_matrix(featuresA, featuresB)
    order = np.argsort(d, axis=1, kind="stable")[:, :2]   # stable: lower train index wins ties
    k = order.shape[1]
    idx[:, :k] = order
    dist[:, :k] = np.take_along_axis(d, order, axis=1)
    return idx, dist


def ratio_test(idx, dist, ratio=0.75):
    """The loop of StitcherClass.py:428-433.  The comparison is evaluated like the
    reference's Python expression: float distances, ``d1 * ratio`` in float64."""
    keep = np.zeros(len(idx), dtype=bool)
    matches = []
    for i in range(len(idx)):
        if idx[i, 1] >= 0 and float(dist[i, 0]) < float(dist[i, 1]) * ratio:
            keep[i] = True
            matches.append((int(idx[i, 0]), i))   # (trainIdx, queryIdx)
    return keep, matches


def match(featuresA, featuresB, ratio=0.75):
    idx, dist = knn_top2(featuresA, featuresB)
    keep, matches = ratio_test(idx, dist, ratio)
    return idx, dist, keep, matches


# ---------------------------------------------------------------------------
# RANSAC scoring, restated for the per-hypothesis outputs of mcs_ransac_homography
def homography_from_4(a, b):
    """Exact homography through 4 correspondences a -> b (float64, h22 = 1):
    the 8x8 linear system of the minimal solver OpenCV's RANSAC runs per sample
    (cv::findHomography -> HomographyEstimatorCallback::runKernel on 4 points).
    Returns None for a singular system."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    A = np.zeros((8, 8))
    rhs = np.zeros(8)
    for i in range(4):
        x, y = a[i]
        u, v = b[i]
        A[2 * i] = [x, y, 1, 0, 0, 0, -u * x, -u * y]
        A[2 * i + 1] = [0, 0, 0, x, y, 1, -v * x, -v * y]
        rhs[2 * i] = u
        rhs[2 * i + 1] = v
    try:
        h = np.linalg.solve(A, rhs)
    except np.linalg.LinAlgError:
        return None
    if not np.all(np.isfinite(h)):
        return None
    return np.append(h, 1.0).reshape(3, 3)


def sample_is_valid(a, b):
    """The sample filter of cv::findHomography's RANSAC (HomographyEstimatorCallback::
    checkSubset): for every cyclic triple of the 4 correspondences the orientation
    (sign of the cross product) must be the same, and non-zero, in both images."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)

    def cross(p, i, j, k):
        return (p[j, 0] - p[i, 0]) * (p[k, 1] - p[i, 1]) - (p[j, 1] - p[i, 1]) * (p[k, 0] - p[i, 0])

    for i in range(4):
        j, k = (i + 1) % 4, (i + 2) % 4
        if not cross(a, i, j, k) * cross(b, i, j, k) > 0.0:
            return False
    return True


def reprojection_errors_sq(H, a, b):
    """Squared forward reprojection error of every correspondence (the inlier
    criterion of cv::findHomography's RANSAC: err <= thresh^2)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    w = H[2, 0] * a[:, 0] + H[2, 1] * a[:, 1] + H[2, 2]
    x = (H[0, 0] * a[:, 0] + H[0, 1] * a[:, 1] + H[0, 2]) / w
    y = (H[1, 0] * a[:, 0] + H[1, 1] * a[:, 1] + H[1, 2]) / w
    return (x - b[:, 0]) ** 2 + (y - b[:, 1]) ** 2


def corner_error(H1, H2, w, h):
    """Max distance between the images of the frame corners under two homographies."""
    c = np.array([[0, 0, 1], [w, 0, 1], [w, h, 1], [0, h, 1]], dtype=np.float64).T
    p1 = H1 @ c
    p2 = H2 @ c
    p1 = p1[:2] / p1[2]
    p2 = p2[:2] / p2[2]
    return float(np.sqrt(((p1 - p2) ** 2).sum(axis=0)).max())
