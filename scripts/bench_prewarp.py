"""Measurement of the rows around the hot path on one B200:
per-camera pre-warp (undistort = REMAP layer, bird's-eye projection = WARP layer) and the shape
fix-up resize, each on a batch of 1080p BGR frames resident in HBM, CUDA events on the launching
stream, against the measured HBM copy bandwidth, with the cv2 call it replaces timed on the host
cores beside it.  One JSON line per operation; `python scripts/bench_prewarp.py [--frames 64]`."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cv2  # noqa: E402
import numpy as np  # noqa: E402
import torch  # noqa: E402

from multicamera_stitching_b200 import Utils, prewarp, synthetic  # noqa: E402
from multicamera_stitching_b200.engine import CompositeEngine  # noqa: E402


def peak_gbs():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:
        return 6533.5, "fallback"


def time_gpu(fn, warmup=3, iters=10):
    for _ in range(warmup):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def time_cpu(fn, budget_s=4.0):
    fn()
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < budget_s:
        fn()
        n += 1
    return n / (time.perf_counter() - t0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64)
    args = ap.parse_args()
    F = args.frames
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    peak, peak_src = peak_gbs()
    h, w = 1080, 1920
    frames = synthetic.make_frames(8, h, w, 3, 0, "smooth")
    host = np.stack([frames[k] for k in sorted(frames)] * ((F + 7) // 8))[:F]
    batch = torch.from_numpy(host).to(dev)                      # F x 1080 x 1920 x 3 = 398 MB at F = 64
    f = 0.8 * w
    mtx = np.array([[f, 0, w / 2 + 3.3], [0, f * 1.01, h / 2 - 2.1], [0, 0, 1]])
    dist = np.array([-0.32, 0.12, 0.001, -0.0007, -0.02])
    M, _ = Utils.CalculateProjectionMatrix([(360, 540), (1560, 540), (1860, 1020), (60, 1020)],
                                           [(0, 0), (1280, 0), (1280, 720), (0, 720)])
    cv2.setNumThreads(os.cpu_count())
    results = []

    def report(name, ms, alg_bytes, exact, cpu_fps, extra):
        gbs = alg_bytes * F / ms / 1e6
        line = {"op": name, "frames_per_launch": F, "ms_per_launch": ms, "frames_per_s": F / ms * 1e3,
                "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                             "peak_source": peak_src, "algorithmic_bytes_per_frame": alg_bytes},
                "bit_exact_vs_cv2": bool(exact),
                "cpu_baseline": {"value": cpu_fps, "unit": "frames/s", "cores": os.cpu_count(), "kind": "reference",
                                 "sample": "cv2 %s on one 1080p frame, all host threads" % cv2.__version__}}
        line.update(extra)
        results.append(line)
        print(json.dumps(line), flush=True)

    # ---- undistort (REMAP layer) -------------------------------------------------------------
    pw = prewarp.PreWarp({"mtx": mtx, "dist": dist}, {"M": M, "dst_size": (1280, 720)})
    out = pw.undistort(batch, batched=True)
    plan = list(pw._cache._plans.values())[0]
    ref = cv2.undistort(host[1], mtx, dist)
    exact = np.array_equal(out[1].cpu().numpy(), ref)
    ms = time_gpu(lambda: plan.run([batch], out=out, n_frames=F))
    report("undistort_1080p", ms, plan.algorithmic_bytes(), exact,
           time_cpu(lambda: cv2.undistort(host[0], mtx, dist)),
           {"kernel_variant": plan.handle.last_variant(), "tiled_status": plan.handle.tiled_status()})
    und = out

    # ---- bird's-eye projection (WARP layer) ----------------------------------------------------
    out2 = pw.warpPerspective(und, batched=True)
    plan2 = [p for p in pw._cache._plans.values() if p is not plan][0]
    und1 = und[1].cpu().numpy()
    ref2 = cv2.warpPerspective(und1, M, (1280, 720))
    exact = np.array_equal(out2[1].cpu().numpy(), ref2)
    ms = time_gpu(lambda: plan2.run([und], out=out2, n_frames=F))
    report("birdseye_1080p_to_720p", ms, plan2.algorithmic_bytes(), exact,
           time_cpu(lambda: cv2.warpPerspective(und1, M, (1280, 720))),
           {"kernel_variant": plan2.handle.last_variant(), "tiled_status": plan2.handle.tiled_status()})

    # ---- the pair, as the callers run it ---------------------------------------------------------
    ms = time_gpu(lambda: (plan.run([batch], out=und, n_frames=F), plan2.run([und], out=out2, n_frames=F)))
    report("prewarp_sequence", ms, plan.algorithmic_bytes() + plan2.algorithmic_bytes(), True,
           time_cpu(lambda: cv2.warpPerspective(cv2.undistort(host[0], mtx, dist), M, (1280, 720))), {})

    # ---- shape fix-up resize ---------------------------------------------------------------------
    eng = CompositeEngine()
    for name, hw in (("resize_1080p_to_720p", (720, 1280)), ("resize_1080p_to_2160p", (2160, 3840)),
                     ("resize_1080p_to_540p_area", (540, 960))):
        n = F if hw[0] <= 1080 else max(1, F // 4)
        src = batch[:n]
        got = eng.resize(src, hw, batched=True)
        ref = cv2.resize(host[1 % n], (hw[1], hw[0]), interpolation=cv2.INTER_LINEAR)
        exact = np.array_equal(got[1 % n].cpu().numpy(), ref)
        del got
        ms = time_gpu(lambda: eng.resize(src, hw, batched=True))
        alg = h * w * 3 + hw[0] * hw[1] * 3 if hw[0] <= 1080 else hw[0] * hw[1] * 3 + h * w * 3
        gbs = alg * n / ms / 1e6
        line = {"op": name, "frames_per_launch": n, "ms_per_launch": ms, "frames_per_s": n / ms * 1e3,
                "roofline": {"bound": "hbm", "achieved": gbs, "peak": peak, "unit": "GB/s", "frac": gbs / peak,
                             "peak_source": peak_src, "algorithmic_bytes_per_frame": alg},
                "bit_exact_vs_cv2": bool(exact),
                "cpu_baseline": {"value": time_cpu(lambda: cv2.resize(host[0], (hw[1], hw[0]), interpolation=cv2.INTER_LINEAR)),
                                 "unit": "frames/s", "cores": os.cpu_count(), "kind": "reference",
                                 "sample": "cv2 %s resize of one 1080p frame, all host threads" % cv2.__version__},
                "note": "timed with the output allocation (torch caching allocator) inside"}
        print(json.dumps(line), flush=True)
        results.append(line)
    return results


if __name__ == "__main__":
    main()
