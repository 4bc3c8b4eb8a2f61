"""Deterministic synthetic frames and camera geometry for the BASELINE.json
configurations (SURVEY.md section 8 row d).

Used by the tests, ``bench.py`` and ``__graft_entry__.smoke()`` so that the
GPU path, the CPU oracle and the CPU baseline all see the same inputs.
Nothing here touches the GPU.
"""
import numpy as np
import cv2

# name -> (n_cameras, height, width) of the BASELINE.json configs
CONFIGS = {
    "cfg1_3x720p": (3, 720, 1280),
    "cfg2_6x1080p": (6, 1080, 1920),
    "cfg3_8x2160p": (8, 2160, 3840),
    "cfg5_6x1080p_seq": (6, 1080, 1920),
}


def frame_seed(cam_index, frame_index):
    return int(cam_index) + 1000 * int(frame_index)


def make_frame(height, width, channels=3, cam_index=0, frame_index=0, kind="smooth"):
    """uint8, C-contiguous ``height x width x channels`` (``height x width``
    when ``channels == 1``) frame.

    ``smooth``: low-resolution uniform noise, Gaussian blur (sigma 2), bicubic
    upsample - natural-image-like statistics.  ``noise``: white noise, the
    stress case where one 1/32-px bucket flip moves a pixel by several levels.
    """
    rng = np.random.default_rng(frame_seed(cam_index, frame_index))
    if kind == "noise":
        img = rng.integers(0, 256, size=(height, width, channels), dtype=np.uint8)
    elif kind == "smooth":
        lh, lw = max(2, height // 4), max(2, width // 4)
        low = rng.integers(0, 256, size=(lh, lw, channels), dtype=np.uint8)
        low = cv2.GaussianBlur(low, (0, 0), 2.0)
        img = cv2.resize(low, (width, height), interpolation=cv2.INTER_CUBIC)
        img = img.reshape(height, width, channels)
    else:
        raise ValueError("unknown frame kind %r" % (kind,))
    img = np.ascontiguousarray(img)
    return img[:, :, 0].copy() if channels == 1 else img


def make_frames(n_cams, height, width, channels=3, frame_index=0, kind="smooth"):
    """``{"CAM1": frame, ...}`` - the ``images_dic`` the reference's callers
    build (video_mapping_node.py:101-103)."""
    return {"CAM%d" % (k + 1): make_frame(height, width, channels, k, frame_index, kind)
            for k in range(n_cams)}


def make_homography(stage, height, width, canvas_width, overlap=0.4):
    """Homography mapping camera ``stage + 1`` into the running canvas of the
    previous stage: ~40 % overlap with the canvas' right edge, scale just under
    1, small shear, a vertical tilt and a mild perspective term whose signs
    alternate from stage to stage."""
    sign = 1.0 if stage % 2 == 0 else -1.0
    scale = 0.95 + 0.01 * (stage % 5)
    shear = 0.02 * sign
    tilt = 18.0 * (height / 720.0) * sign
    persp = 2e-5 * (720.0 / height) * sign
    tx = canvas_width - overlap * width + 0.37 * (stage + 1)
    ty = tilt + 0.21 * stage
    return np.array([[scale, shear, tx],
                     [-0.6 * shear, scale + 0.004, ty],
                     [persp, -0.5 * persp, 1.0]], dtype=np.float64)


def homography_from_points(height, width, canvas_width, overlap=0.4):
    """Same idea, but through the 4-point route of
    ``Calibration_Utils.CalculateProjectionMatrix`` (config 1: "fixed
    homographies from Calibration_Utils")."""
    from .Utils import CalculateProjectionMatrix
    src = [(0, 0), (width, 0), (width, height), (0, height)]
    x0 = canvas_width - overlap * width
    dst = [(x0 + 3.0, 11.0), (x0 + 0.97 * width, 2.0),
           (x0 + 0.985 * width - 4.0, 0.98 * height + 9.0), (x0 - 2.0, 0.99 * height + 14.0)]
    M, _ = CalculateProjectionMatrix(src, dst)
    return M


def synthetic_stitcher(n_cams, h, w, channels=3, super_mode=False, kind="smooth", frame_index=0,
                       xoffset=0, yoffset=0, use_points_first=False):
    """A ``Stitcher`` calibrated on the synthetic camera geometry above, stage by stage (each stage's
    homography depends on the width of the canvas stitched so far, like ``calibrate_stitcher`` calibrates
    against ``img_result``, reference StitcherClass.py:96-104).

    Returns ``(stitcher, homographies, labels, images_dic)``."""
    from .StitcherClass import Stitcher
    images = make_frames(n_cams, h, w, channels, frame_index, kind)
    st = Stitcher(images, super_mode=super_mode)
    labels = list(st.img_labels)
    shapes = [images[l].shape for l in labels]
    homographies = []
    shapeB = tuple(shapes[0])
    for k in range(n_cams - 1):
        cw = shapeB[1]
        H = homography_from_points(h, w, cw) if (use_points_first and k == 0) else make_homography(k, h, w, cw)
        homographies.append(H)
        st.stitchers[k].set_homography(H, shapeA=shapes[k + 1], shapeB=shapeB, xoffset=xoffset, yoffset=yoffset)
        shapeB = st.stitchers[k].result_shape()
    return st, homographies, labels, images


# ---------------------------------------------------------------------------
# recalibration workload (BASELINE.json config 4): feature-rich image pairs
def make_textured(height, width, seed=0, channels=3):
    """A frame full of corners (random filled rectangles and discs, lightly
    blurred): ``cv2.ORB_create(2000)`` finds its full quota of key-points on it,
    which the smooth frames above do not offer."""
    rng = np.random.default_rng(10_000 + int(seed))
    img = np.full((height, width, 3), 96, np.uint8)
    n_shapes = max(50, height * width // 1400)
    for i in range(n_shapes):
        x, y = int(rng.integers(0, width)), int(rng.integers(0, height))
        r = int(rng.integers(4, 40))
        colour = tuple(int(v) for v in rng.integers(0, 255, 3))
        if i % 2:
            cv2.circle(img, (x, y), r, colour, -1)
        else:
            cv2.rectangle(img, (x, y), (x + r, y + int(rng.integers(4, 40))), colour, -1)
    img = cv2.GaussianBlur(img, (0, 0), 1.0)
    if channels == 1:
        return np.ascontiguousarray(img[:, :, 0])
    return np.ascontiguousarray(img)


def make_pair(height, width, seed=0, overlap=0.6, noise_sigma=2.0):
    """``(imageB, imageA, H_true)``: two views of one textured scene.  ``H_true``
    maps imageA pixel coordinates into imageB's frame (the homography
    ``StitcherBase.calibrate`` has to recover): imageA is the scene resampled
    through ``inv(H_true)`` plus sensor noise, imageB a crop of the scene."""
    scene_w = int(width * (2.0 - overlap)) + 64
    scene = make_textured(height + 64, scene_w, seed)
    rng = np.random.default_rng(20_000 + int(seed))
    imageB = scene[32:32 + height, 32:32 + width].copy()
    tx = (1.0 - overlap) * width
    H_true = np.array([[0.97, 0.015, tx + 1.3],
                       [-0.01, 0.985, 6.7],
                       [1.5e-5 * (720.0 / height), -0.8e-5 * (720.0 / height), 1.0]], dtype=np.float64)
    # scene coordinates = imageB coordinates + 32
    T = np.array([[1, 0, 32.0], [0, 1, 32.0], [0, 0, 1]], dtype=np.float64)
    A_to_scene = T @ H_true
    imageA = cv2.warpPerspective(scene, np.linalg.inv(A_to_scene), (width, height), flags=cv2.INTER_LINEAR)
    for im in (imageA, imageB):
        noise = rng.normal(0.0, noise_sigma, im.shape)
        np.clip(im.astype(np.float64) + noise, 0, 255, out=noise)
        im[...] = noise.astype(np.uint8)
    return np.ascontiguousarray(imageB), np.ascontiguousarray(imageA), H_true
