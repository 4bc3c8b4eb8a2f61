"""Per-camera pre-warp in front of the stitcher.

Both callers of the reference prepare every camera frame with the same two
OpenCV calls before it reaches the stitcher (video_mapping_node.py:155-163,
MediaPlayer/view.py:378-388; calibration tools: Intrinsic.py:234-235,
Extrinsic.py:98-99)::

    image = cv2.undistort(src=image, cameraMatrix=intrinsic["mtx"], distCoeffs=intrinsic["dist"])
    image = cv2.warpPerspective(src=image, M=extrinsic["M"], dsize=extrinsic["dst_size"])

Here both run on the GPU through the compositing kernels of ``libmcs_b200.so``
- a pre-warp is a one-layer plan: a REMAP layer driven by the fixed-point
undistortion map (``mcs_plan_create_maps``) or a WARP layer driven by ``M`` -
so they share the TMA-staged resampling kernel, its frame batching and its
bit-exactness with OpenCV.  :func:`undistort` and :func:`warpPerspective` are
drop-ins for the two cv2 calls (same argument names; numpy in -> numpy out,
CUDA tensor in -> CUDA tensor out), :class:`PreWarp` bundles the sequence for
one camera from the reference's calibration dictionaries.

What stays on the host is calibration-time work, done once per camera: the
undistortion map is built with ``cv2.initUndistortRectifyMap`` exactly the way
``cv::undistort`` builds it internally (row stripes, CV_16SC2 + CV_16UC1).
There is no CPU fallback for the per-frame resampling.
"""
import cv2
import numpy as np

from .plan import LAYER_REMAP, LAYER_WARP, FlatPlan, Layer


def _is_tensor(x):
    return type(x).__module__.startswith("torch") and hasattr(x, "is_cuda")


def undistort_maps(cameraMatrix, distCoeffs, size, newCameraMatrix=None):
    """The fixed-point map pair ``cv2.undistort`` resamples through for images of
    ``size = (width, height)``: ``(xy int16 H x W x 2, frac uint16 H x W)``.

    ``cv::undistort`` does not build one map for the image: it walks it in
    stripes of ``max(1, 4096 // width)`` rows, shifts the principal point of the
    new camera matrix by the stripe's first row and calls
    ``initUndistortRectifyMap(A, dist, I, Ar, (width, rows), CV_16SC2)`` per
    stripe.  The same construction is used here so that every map entry is the
    one cv2 would have used."""
    w, h = int(size[0]), int(size[1])
    A = np.array(cameraMatrix, dtype=np.float64).reshape(3, 3)
    Ar = A.copy() if newCameraMatrix is None else np.array(newCameraMatrix, dtype=np.float64).reshape(3, 3)
    if distCoeffs is None or np.size(distCoeffs) == 0:
        dist = np.zeros((5, 1), dtype=np.float64)
    else:
        dist = np.array(distCoeffs, dtype=np.float64)
    eye = np.eye(3, dtype=np.float64)
    stripe = min(max(1, (1 << 12) // max(w, 1)), h)
    xy = np.empty((h, w, 2), dtype=np.int16)
    frac = np.empty((h, w), dtype=np.uint16)
    v0 = Ar[1, 2]
    for y in range(0, h, stripe):
        rows = min(stripe, h - y)
        Ar[1, 2] = v0 - y
        m1, m2 = cv2.initUndistortRectifyMap(A, dist, eye, Ar, (w, rows), cv2.CV_16SC2)
        xy[y:y + rows] = m1
        frac[y:y + rows] = m2
    return xy, frac


class _OnePlan(object):
    """A one-layer plan per (operation, frame shape, device) with the host<->device plumbing."""

    def __init__(self):
        self._plans = {}
        self._engine = None

    def _eng(self):
        if self._engine is None:
            from .engine import CompositeEngine
            self._engine = CompositeEngine()
        return self._engine

    def plan(self, key, shape, device, make_layer, out_wh):
        from .engine import CompiledPlan
        full = (key, tuple(shape), str(device))
        p = self._plans.pop(full, None)
        if p is not None:
            self._plans[full] = p                        # most recently used last
        else:
            if len(shape) not in (2, 3):
                raise ValueError("frames must be H x W or H x W x C, got shape %r" % (tuple(shape),))
            channels = 1 if len(shape) == 2 else int(shape[2])
            flat = FlatPlan([make_layer()], int(out_wh[0]), int(out_wh[1]), channels, len(shape))
            while len(self._plans) >= 16:                # evict the least recently used, one at a time
                self._plans.pop(next(iter(self._plans)))
            p = CompiledPlan(flat, device)
            self._plans[full] = p
        return p

    def run(self, key, src, make_layer, out_wh, batched=False, out=None):
        import torch
        eng = self._eng()
        if _is_tensor(src):
            if not src.is_cuda:
                raise TypeError("frames must be numpy arrays or uint8 CUDA tensors")
            shape = tuple(int(v) for v in (src.shape[1:] if batched else src.shape))
            plan = self.plan(key, shape, src.device, make_layer, out_wh)
            with torch.cuda.device(src.device):
                return plan.run([src], out=out, n_frames=int(src.shape[0]) if batched else None)
        if batched:
            raise TypeError("batched pre-warps expect uint8 CUDA tensors [F, H, W(, C)]")
        src = np.asarray(src)
        if src.dtype != np.uint8:
            raise TypeError("the pre-warp kernels resample uint8 frames, got %s" % src.dtype)
        device = eng.device
        plan = self.plan(key, src.shape, device, make_layer, out_wh)
        with torch.cuda.device(device):
            res = plan.run([eng.upload(0, src, device)])
            host = torch.empty(res.shape, dtype=torch.uint8)
            host.copy_(res)
        return host.numpy()


_shared = _OnePlan()


def _mat_key(*arrays):
    return tuple(None if a is None else np.asarray(a, dtype=np.float64).tobytes() for a in arrays)


def _frame_hw(src, batched):
    s = src.shape[1:] if batched else src.shape
    return int(s[0]), int(s[1])


def undistort(src, cameraMatrix, distCoeffs, dst=None, newCameraMatrix=None, batched=False, _cache=None):
    """``cv2.undistort(src, cameraMatrix, distCoeffs[, dst[, newCameraMatrix]])`` for uint8 frames
    (video_mapping_node.py:157-158, view.py:380-381), bit-exact with cv2.  ``src``: numpy H x W[x C]
    (returns numpy) or a CUDA tensor (returns a CUDA tensor; ``batched=True`` for [F, H, W(, C)]).
    ``dst`` is accepted for signature compatibility and ignored unless it is a CUDA tensor of the
    result's shape, which is then filled."""
    h, w = _frame_hw(src, batched)
    key = ("undistort",) + _mat_key(cameraMatrix, distCoeffs, newCameraMatrix)

    def layer():
        maps = undistort_maps(cameraMatrix, distCoeffs, (w, h), newCameraMatrix)
        return Layer(0, LAYER_REMAP, None, 0, 0, (0, 0, w, h), (h, w), maps)

    out = dst if (_is_tensor(dst) and _is_tensor(src)) else None
    return (_cache or _shared).run(key, src, layer, (w, h), batched, out)


def warpPerspective(src, M, dsize, dst=None, batched=False, _cache=None):
    """``cv2.warpPerspective(src, M, dsize)`` with the default flags (INTER_LINEAR,
    BORDER_CONSTANT 0) for uint8 frames (view.py:387-388, Extrinsic.py:99), bit-exact with cv2."""
    h, w = _frame_hw(src, batched)
    dw, dh = int(dsize[0]), int(dsize[1])
    key = ("warp", dw, dh) + _mat_key(M)

    def layer():
        return Layer(0, LAYER_WARP, np.array(M, dtype=np.float64).reshape(3, 3), 0, 0, (0, 0, dw, dh), (h, w))

    out = dst if (_is_tensor(dst) and _is_tensor(src)) else None
    return (_cache or _shared).run(key, src, layer, (dw, dh), batched, out)


class PreWarp(object):
    """The pre-warp of one camera from the reference's calibration dictionaries:
    ``intrinsic_calibration`` as ``Intrinsic.load_intrinsic_calibration`` returns it (``mtx``,
    ``dist``; Intrinsic.py:166-207) and ``extrinsic_calibration`` as
    ``Extrinsic.load_extrinsic_calibration`` returns it (``M``, ``dst_size``;
    Extrinsic.py:207-241).  Calling the object applies what MediaPlayer/view.py:378-388 applies:
    the undistortion when the intrinsic matrix is known and then, when ``M`` is known, the
    bird's-eye projection.  With neither the frame passes through."""

    def __init__(self, intrinsic_calibration=None, extrinsic_calibration=None):
        self.intrinsic_calibration = intrinsic_calibration or {"mtx": None, "dist": None}
        self.extrinsic_calibration = extrinsic_calibration or {"M": None, "dst_size": None}
        self._cache = _OnePlan()

    def undistort(self, src, batched=False):
        ic = self.intrinsic_calibration
        return undistort(src, ic["mtx"], ic["dist"], batched=batched, _cache=self._cache)

    def warpPerspective(self, src, batched=False):
        ec = self.extrinsic_calibration
        return warpPerspective(src, ec["M"], ec["dst_size"], batched=batched, _cache=self._cache)

    def __call__(self, image, batched=False):
        if self.intrinsic_calibration.get("mtx") is None:
            return image
        two = self.extrinsic_calibration.get("M") is not None
        if two and not _is_tensor(image):
            # numpy in, numpy out: one upload, both stages on the device, one download
            import torch
            eng = self._cache._eng()
            device = eng.device
            with torch.cuda.device(device):
                res = self.warpPerspective(self.undistort(eng.upload(0, np.asarray(image), device)))
                host = torch.empty(res.shape, dtype=torch.uint8)
                host.copy_(res)
            return host.numpy()
        image = self.undistort(image, batched=batched)
        if two:
            image = self.warpPerspective(image, batched=batched)
        return image
