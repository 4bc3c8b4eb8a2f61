"""ORACLE (test infrastructure, never imported by the product package).

CPU restatement of the reference's compositing and recalibration path, driving
OpenCV exactly the way ``PostScripts/Stitcher/StitcherClass.py`` does, function
by function, each citing the lines it follows.  The arithmetic lives in OpenCV
(third party, unpinned by the reference; opencv-python 4.13.0 here).

PINNED against the reference itself: the reference ships no tests or golden
vectors, but its ``StitcherClass.py`` runs here once
``oracle/build_ref.py`` has applied a four-line mechanical patch list (Python 2
indentation and dict idioms, the OpenCV-version switch of the feature
detector).  ``tests/test_oracle_ref_pin.py`` holds every function below against
the reference method it restates - state fields, panoramas, match lists, bit
for bit - on the seeded inputs of ``multicamera_stitching_b200.synthetic``, and
``tests/golden/chain_ref.npz`` (``scripts/make_golden.py``) carries panoramas
and stage states made by the reference's own classes to machines without the
reference tree.

State is carried in plain dicts with the reference's field names
(StitcherClass.py:190-209).
"""
import numpy as np
import cv2


# --------------------------------------------------------------------------
# Calibration_Utils/Utils.py:23-37
def projection_point_dst(pt_src, M):
    v = np.matmul(M, pt_src)
    v = v / v[2]
    return [int(v[0]), int(v[1])]


# --------------------------------------------------------------------------
def new_state(sid=None, super_mode=False):
    """StitcherClass.py:182-209 (fields) / :507-525 (reset)."""
    return dict(sid=sid, super_mode=super_mode, cachedBH=None, cachedBINVH=None,
                Bpts=None, BimgSize=None, cachedAH=None, cachedAINVH=None,
                Apts=None, AimgSize=None, matches=None, status=None,
                ABSize=None, x_limits=None, y_limits=None)


def geometry_from_homography(st, H, shapeA, shapeB, xoffset=10, yoffset=10):
    """Canvas geometry of ``StitcherBase.calibrate`` once ``cachedAH`` is
    known: StitcherClass.py:273-281 and :293-351.  ``H`` maps image A into
    image B's frame; it is translated in place exactly as the reference does.
    """
    xoffset = abs(xoffset)
    yoffset = abs(yoffset)
    st["BimgSize"] = tuple(shapeB)
    st["AimgSize"] = tuple(shapeA)
    hA, wA = shapeA[0], shapeA[1]
    hB, wB = shapeB[0], shapeB[1]
    H = np.array(H, dtype=np.float64)
    cornersA = [(0, 0), (wA, 0), (wA, hA), (0, hA)]
    Apts = [projection_point_dst((p[0], p[1], 1), H) for p in cornersA]
    Bpts = [(0, 0), (wB, 0), (wB, hB), (0, hB)]
    allpts = np.concatenate((Apts, Bpts), axis=0)
    x_min = min(p[0] for p in allpts)
    y_min = min(p[1] for p in allpts)
    st["cachedBH"] = np.float32([[1, 0, x_min + xoffset], [0, 1, y_min + yoffset], [0, 0, 1]])
    st["cachedBINVH"] = np.linalg.inv(st["cachedBH"])
    H[0][2] += -x_min + xoffset
    H[1][2] += -y_min + yoffset
    st["cachedAH"] = H
    st["cachedAINVH"] = np.linalg.inv(H)
    xoff = -x_min + xoffset
    yoff = -y_min + yoffset
    st["Bpts"] = [(xoff, yoff), (xoff + wB, yoff), (xoff + wB, hB + yoff), (xoff, hB + yoff)]
    st["Apts"] = [projection_point_dst((p[0], p[1], 1), H) for p in cornersA]
    pts = np.concatenate((st["Apts"], st["Bpts"]), axis=0)
    xs = [p[0] for p in pts]
    ys = [p[1] for p in pts]
    st["ABSize"] = (int(abs(max(xs)) + xoffset), int(abs(max(ys)) + yoffset))
    st["x_limits"] = [max([v for v in xs if v < st["ABSize"][0] * 0.5]),
                      min([v for v in xs if v > st["ABSize"][0] * 0.5])]
    st["y_limits"] = [max([v for v in ys if v < st["ABSize"][1] * 0.5]),
                      min([v for v in ys if v > st["ABSize"][1] * 0.5])]
    return st


def stitch_pair(st, images):
    """``StitcherBase.stitch`` - StitcherClass.py:211-256 (without the
    optional debug overlay of :244-245)."""
    imageB, imageA = images
    if st["cachedAH"] is None:
        return imageB
    if imageB.shape != st["BimgSize"]:
        imageB = cv2.resize(imageB, (st["BimgSize"][1], st["BimgSize"][0]),
                            interpolation=cv2.INTER_LINEAR)
    if imageA.shape != st["AimgSize"]:
        imageA = cv2.resize(imageA, (st["AimgSize"][1], st["AimgSize"][0]),
                            interpolation=cv2.INTER_LINEAR)
    bx = int(st["Bpts"][0][0])
    by = int(st["Bpts"][0][1])
    dst = cv2.warpPerspective(src=imageA, M=st["cachedAH"],
                              dsize=(st["ABSize"][0], st["ABSize"][1]))
    dst[by:by + imageB.shape[0], bx:bx + imageB.shape[1]] = imageB
    if st["super_mode"]:
        dst = dst[st["y_limits"][0]:st["y_limits"][1], st["x_limits"][0]:st["x_limits"][1]]
    return dst


def sorted_labels(images_dic):
    """StitcherClass.py:61 under Python 3."""
    return list(np.sort(list(images_dic.keys())))


def stitcher_labels(img_labels):
    """StitcherClass.py:64-71."""
    out = []
    for i in range(len(img_labels) - 1):
        left = img_labels[i] if i == 0 else out[-1]
        out.append("({}&{})".format(left, img_labels[i + 1]))
    return out


def stitch_chain(states, img_labels, images_dic):
    """``Stitcher.stitch`` - StitcherClass.py:114-136."""
    if len(images_dic) < len(img_labels):
        return images_dic[img_labels[-1]]
    dst = None
    for i in range(len(img_labels) - 1):
        if i == 0:
            pair = (images_dic[img_labels[0]], images_dic[img_labels[1]])
        else:
            pair = (dst, images_dic[img_labels[i + 1]])
        dst = stitch_pair(states[i], pair)
    return dst if dst is not None else images_dic[img_labels[-1]]


def calibrate_chain_from_homographies(img_shapes, homographies, super_mode=False,
                                      xoffset=0, yoffset=0):
    """``Stitcher.calibrate_stitcher`` (StitcherClass.py:96-104) with the
    feature matching replaced by given homographies: stage ``i`` receives
    ``homographies[i]`` (camera i+1 -> running canvas) and the *shape of the
    stitched canvas so far* as its image B, exactly like :99-104 where the next
    pair calibrates against ``img_result``."""
    states = []
    shapeB = tuple(img_shapes[0])
    for i, H in enumerate(homographies):
        st = new_state(sid=str(i), super_mode=super_mode)
        geometry_from_homography(st, H, tuple(img_shapes[i + 1]), shapeB, xoffset, yoffset)
        states.append(st)
        w, h = st["ABSize"]
        if super_mode:
            ys = slice(st["y_limits"][0], st["y_limits"][1]).indices(h)
            xs = slice(st["x_limits"][0], st["x_limits"][1]).indices(w)
            h = max(0, ys[1] - ys[0])
            w = max(0, xs[1] - xs[0])
        shapeB = (h, w) + tuple(img_shapes[i + 1][2:])
    return states


# --------------------------------------------------------------------------
def match_keypoints(kpsA, kpsB, featuresA, featuresB, ratio=0.75, reprojThresh=4.0,
                    norm=None):
    """``StitcherBase.matchKeypoints`` - StitcherClass.py:405-448.

    The reference builds ``DescriptorMatcher_create("BruteForce")`` (L2, for
    its float SIFT descriptors).  BASELINE.json's config 4 runs the same
    routine on ORB descriptors, for which the brute-force matcher is the
    Hamming one; ``norm`` defaults to Hamming for uint8 features and L2
    otherwise."""
    if norm is None:
        norm = cv2.NORM_HAMMING if featuresA.dtype == np.uint8 else cv2.NORM_L2
    matcher = cv2.BFMatcher(norm)
    raw = matcher.knnMatch(featuresA, featuresB, 2)
    matches = []
    H = None
    status = None
    for m in raw:
        if len(m) == 2 and m[0].distance < m[1].distance * ratio:
            matches.append((m[0].trainIdx, m[0].queryIdx))
    if len(matches) > 4:
        ptsA = np.float32([kpsA[i] for (_, i) in matches])
        ptsB = np.float32([kpsB[i] for (i, _) in matches])
        H, status = cv2.findHomography(srcPoints=ptsA, dstPoints=ptsB, method=cv2.RANSAC,
                                       ransacReprojThreshold=reprojThresh)
    return H, matches, status


def knn_top2(featuresA, featuresB, norm=None):
    """Raw ``knnMatch(k=2)`` result as arrays (idx Nx2, dist Nx2), -1 padded."""
    if norm is None:
        norm = cv2.NORM_HAMMING if featuresA.dtype == np.uint8 else cv2.NORM_L2
    raw = cv2.BFMatcher(norm).knnMatch(featuresA, featuresB, 2)
    idx = -np.ones((len(raw), 2), dtype=np.int32)
    dist = -np.ones((len(raw), 2), dtype=np.float32)
    for i, m in enumerate(raw):
        for j, mm in enumerate(m[:2]):
            idx[i, j] = mm.trainIdx
            dist[i, j] = mm.distance
    return idx, dist
