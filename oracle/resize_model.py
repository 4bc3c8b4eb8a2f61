"""ORACLE (test infrastructure, never imported by the product package).

Transparent integer model of ``cv2.resize(img, (W, H), interpolation=
cv2.INTER_LINEAR)`` for uint8 images, the shape fix-up the reference applies to
a frame whose shape differs from the calibrated one
(PostScripts/Stitcher/StitcherClass.py:226-233; the MediaPlayer always takes
this branch, MediaPlayer/view.py:408-409 hands 2-D transposed frames in).

Where the arithmetic lives: OpenCV (third party, not vendored in the reference,
no version pinned there); this model is pinned by asserting bit-equality with
``cv2.resize`` of opencv-python 4.13.0 in ``tests/test_oracle_resize.py``.  The
reference holds no golden vectors for this path.

Recipe (OpenCV ``resize.cpp``: ``resize_`` coefficient tables, ``HResizeLinear``
and ``VResizeLinear`` for 8-bit data):

  scale_x = 1 / (W_dst / W_src)   (float64, that association), scale_y likewise
  exact 2x2 decimation (scale_x == scale_y == 2) is served by the area kernel:
      out = (p00 + p01 + p10 + p11 + 2) >> 2
  otherwise, per destination column dx:
      fx = float32((dx + 0.5) * scale_x - 0.5); sx = floor(fx); fx -= sx   (float32)
      sx < 0        -> sx = 0, fx = 0
      sx >= W_src-1 -> sx = W_src-1, fx = 0 (the second tap is not read)
      a0 = rint(float32(1 - fx) * 2048), a1 = rint(fx * 2048)           (half to even)
  per destination row dy the same without the edge rule (b0, b1, sy); the two
  source rows are clip(sy, 0, H_src-1) and clip(sy+1, 0, H_src-1)
      h_r[x] = p[r][sx]*a0 + p[r][sx+1]*a1                                (int32)
      out    = (((b0 * (h_0 >> 4)) >> 16) + ((b1 * (h_1 >> 4)) >> 16) + 2) >> 2
"""
import numpy as np

COEF_SCALE = 2048


def _axis_tables(n_dst, n_src, clamp_edges):
    """(index, w0, w1) of one axis; ``clamp_edges`` = the horizontal rule."""
    inv_scale = np.float64(n_dst) / np.float64(n_src)
    scale = np.float64(1.0) / inv_scale
    d = np.arange(n_dst, dtype=np.float64)
    f = ((d + 0.5) * scale - 0.5).astype(np.float32)
    s = np.floor(f).astype(np.int32)
    f = (f - s.astype(np.float32)).astype(np.float32)
    if clamp_edges:
        lo = s < 0
        f[lo] = 0.0
        s[lo] = 0
        hi = s >= n_src - 1
        f[hi] = 0.0
        s[hi] = n_src - 1
    w0 = np.rint((np.float32(1.0) - f).astype(np.float32) * np.float32(COEF_SCALE)).astype(np.int32)
    w1 = np.rint(f * np.float32(COEF_SCALE)).astype(np.int32)
    return s, w0, w1, float(scale)


def resize_linear_u8(img, dsize):
    """``cv2.resize(img, dsize, interpolation=cv2.INTER_LINEAR)`` for a uint8
    H x W or H x W x C image; ``dsize = (W_dst, H_dst)``."""
    img = np.asarray(img)
    assert img.dtype == np.uint8 and img.ndim in (2, 3)
    wd, hd = int(dsize[0]), int(dsize[1])
    hs, ws = img.shape[:2]
    src = img.reshape(hs, ws, -1).astype(np.int32)
    sx, a0, a1, scale_x = _axis_tables(wd, ws, True)
    sy, b0, b1, scale_y = _axis_tables(hd, hs, False)
    if (hs, ws) == (hd, wd):
        return img.copy()
    if scale_x == 2.0 and scale_y == 2.0:
        out = (src[0:2 * hd:2, 0:2 * wd:2] + src[0:2 * hd:2, 1:2 * wd:2] +
               src[1:2 * hd:2, 0:2 * wd:2] + src[1:2 * hd:2, 1:2 * wd:2] + 2) >> 2
    else:
        sx1 = np.minimum(sx + 1, ws - 1)
        hrow = src[:, sx, :] * a0[None, :, None] + src[:, sx1, :] * a1[None, :, None]   # H_src x W_dst x C
        r0 = np.clip(sy, 0, hs - 1)
        r1 = np.clip(sy + 1, 0, hs - 1)
        h0 = hrow[r0] >> 4
        h1 = hrow[r1] >> 4
        out = (((b0[:, None, None] * h0) >> 16) + ((b1[:, None, None] * h1) >> 16) + 2) >> 2
    out = out.astype(np.uint8)
    return out.reshape((hd, wd) + img.shape[2:])
