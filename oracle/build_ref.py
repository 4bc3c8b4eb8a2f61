"""ORACLE RECIPE (test infrastructure): make the reference's own ``StitcherClass.py`` importable here.

    python oracle/build_ref.py          # writes oracle/_ref/ (git-ignored), prints what it patched

``/root/reference/PostScripts/Stitcher/StitcherClass.py`` is Python 2 era code that Python 3 cannot even
compile (mixed space + tab indentation of ``__str__``, ``np.sort(dict.keys())``) and imports a logging module
the repository does not ship.  This recipe reads the file WHERE IT LIES, applies the mechanical patch list
below - nothing that touches the arithmetic of the stitch / geometry / match path - and writes the result to
``oracle/_ref/StitcherClass_ref.py`` next to a stand-in ``extended_rospylogs.py``.  The reference's
``Calibration_Utils/Utils.py`` is imported UNCHANGED from the reference tree (it runs under Python 3 as is); a byte
copy of it is placed in the build output for the GPU box, which has no reference tree (``prebuilt()``).
No reference source is copied into the repository: ``oracle/_ref/`` is a build output, like a ``.so``.

The patched module is what ``tests/test_oracle_ref_pin.py`` holds ``oracle/stitcher_ref.py`` against and what
``scripts/make_golden.py`` generates ``tests/golden/chain_ref.npz`` from.

Patch list (each must apply exactly as many times as stated, or the recipe fails loudly):
  1. ``:527``  `` \\tdef __str__`` (space + tab)          -> ``\\tdef __str__``           TabError otherwise
  2. ``:61``   ``np.sort(images_dic.keys())``            -> ``np.sort(list(images_dic.keys()))``   dict view in py3
  3. ``:376``  ``if is_cv3(or_better=False):`` in ``detectAndDescribe`` -> ``or_better=True``: under OpenCV 4 the
     strict test sends the code into the OpenCV 2.4 branch (``cv2.FeatureDetector_create``, gone since 3.0)
  4. ``:380``  ``cv2.xfeatures2d.SIFT_create()`` in ``detectAndDescribe`` -> ``cv2.SIFT_create()``: SIFT moved out of
     contrib in OpenCV 4.4; same detector, same descriptor
Patches 3 and 4 only concern feature DETECTION (CPU by contract); the chain check of
``calibrate_stitcher`` (:87-93, also ``is_cv3(or_better=False)``) is left alone - it is simply skipped on OpenCV 4.
"""
import os
import sys

REF_ROOT = os.environ.get("MCS_REFERENCE_ROOT", "/root/reference")
REF_FILE = os.path.join(REF_ROOT, "PostScripts", "Stitcher", "StitcherClass.py")
REF_UTILS_DIR = os.path.join(REF_ROOT, "PostScripts", "Calibration_Utils")
OUT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")

PATCHES = [
    # (description, old, new, expected count)
    ("space+tab indentation of __str__", "\n \tdef __str__(self):", "\n\tdef __str__(self):", 1),
    ("dict view passed to np.sort", "np.sort(images_dic.keys())", "np.sort(list(images_dic.keys()))", 1),
    ("strict OpenCV-3 test in detectAndDescribe",
     "\t\t# check to see if we are using OpenCV 3.X\n\t\tif is_cv3(or_better=False):",
     "\t\t# check to see if we are using OpenCV 3.X\n\t\tif is_cv3(or_better=True):", 1),
    ("SIFT constructor of detectAndDescribe",
     "\t\t\t\tdescriptor = cv2.xfeatures2d.SIFT_create()", "\t\t\t\tdescriptor = cv2.SIFT_create()", 1),
]

SHIM = '''"""Stand-in for the ``extended_rospylogs`` module the reference imports but does not ship (build output of
oracle/build_ref.py).  Messages are collected on the object so that tests can look at them."""
DEBUG_LEVEL_0, DEBUG_LEVEL_1, DEBUG_LEVEL_2, DEBUG_LEVEL_3, DEBUG_LEVEL_4 = range(5)


class Debugger(object):
    def debugger(self, level, msg, log_type="info"):
        self.__dict__.setdefault("_log", []).append((level, log_type, msg))


def update_debuggers(*args, **kwargs):
    pass


def loginfo_cond(*args, **kwargs):
    pass


def logerr_cond(*args, **kwargs):
    pass
'''


def available():
    return os.path.isfile(REF_FILE) and os.path.isfile(os.path.join(REF_UTILS_DIR, "Utils.py"))


def build(verbose=False):
    """Writes oracle/_ref/; returns the path of the patched module (None when the reference is absent)."""
    if not available():
        return None
    with open(REF_FILE, "r") as f:
        src = f.read()
    for what, old, new, count in PATCHES:
        n = src.count(old)
        if n != count:
            raise RuntimeError("oracle/build_ref.py: patch %r matched %d times, expected %d - the reference file is not "
                               "the one this recipe was written for" % (what, n, count))
        src = src.replace(old, new)
        if verbose:
            print("patched: %s" % what)
    os.makedirs(OUT_DIR, exist_ok=True)
    out = os.path.join(OUT_DIR, "StitcherClass_ref.py")
    header = ("# BUILD OUTPUT of oracle/build_ref.py from %s - do not edit, do not commit.\n" % REF_FILE)
    with open(out, "w") as f:
        f.write(header + src)
    with open(os.path.join(OUT_DIR, "extended_rospylogs.py"), "w") as f:
        f.write(SHIM)
    # The GPU box has no /root/reference: the helper module the class imports travels with the build output,
    # byte for byte as it lies in the reference tree (used only where that tree is absent, see load()).
    os.makedirs(os.path.join(OUT_DIR, "Calibration_Utils"), exist_ok=True)
    with open(os.path.join(REF_UTILS_DIR, "Utils.py"), "rb") as f, \
            open(os.path.join(OUT_DIR, "Calibration_Utils", "Utils.py"), "wb") as g:
        g.write(f.read())
    compile(src, out, "exec")   # fails here, loudly, if the patch list no longer makes it Python 3
    return out


def prebuilt():
    """The build output of an earlier run (it travels to the GPU box, the reference tree does not)."""
    out = os.path.join(OUT_DIR, "StitcherClass_ref.py")
    ok = os.path.isfile(out) and os.path.isfile(os.path.join(OUT_DIR, "extended_rospylogs.py")) and \
        os.path.isfile(os.path.join(OUT_DIR, "Calibration_Utils", "Utils.py"))
    return out if ok else None


def load():
    """Imports the patched reference module (building it first); None when the reference is absent."""
    path = build()
    utils_dir = REF_UTILS_DIR
    if path is None:
        path = prebuilt()
        utils_dir = os.path.join(OUT_DIR, "Calibration_Utils")
    if path is None:
        return None
    for p in (OUT_DIR, utils_dir):
        if p not in sys.path:
            sys.path.insert(0, p)
    import importlib
    return importlib.import_module("StitcherClass_ref")


def calibrated_stitcher(ref, images_dic, homographies, super_mode=False):
    """The reference's own ``Stitcher`` (``ref`` = the module from ``load()``), calibrated by its own
    ``calibrate_stitcher`` (StitcherClass.py:77-112) with the given stage homographies standing in for matched
    features: ``StitcherBase.calibrate`` (:258-354) runs its real geometry code, only the feature detection and
    matching in front of it are replaced.  Every stage calibrates against the stitched canvas so far."""
    import numpy as np
    rs = ref.Stitcher(images_dic, super_mode=super_mode)

    def fixed(H):
        H = np.array(H, dtype=np.float64)
        return lambda *a, **kw: (H.copy(), [(0, 0)] * 5, np.ones((5, 1), np.uint8))

    for sb, H in zip(rs.stitchers, homographies):
        sb.detectAndDescribe = lambda image: (np.zeros((1, 2), np.float32), None)
        sb.matchKeypoints = fixed(H)
    rs.calibrate_stitcher(images_dic, save=False)
    return rs


if __name__ == "__main__":
    p = build(verbose=True)
    print(p if p else "reference not found under %s" % REF_ROOT)
