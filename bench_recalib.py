#!/usr/bin/env python
"""Benchmark of the recalibration kernels (BASELINE.json config 4): ORB-sized brute-force Hamming
top-2 + ratio test and batched RANSAC homography scoring on 4 image pairs of 2000 keypoints.

  python bench_recalib.py [--steps K] [--warmup W] [--pairs 4] [--keypoints 2000]

Prints ONE JSON line: pairs/s of the GPU match + RANSAC kernels (inputs resident in HBM, CUDA
events), the popcount rate of the matcher against nothing but itself (it is latency/launch bound
at this size; the ncu capture in profiles/ gives the INT pipe utilisation), and the same work done
by OpenCV on this box's host cores (BFMatcher.knnMatch + ratio loop, findHomography(RANSAC)) - the
calls of StitcherBase.matchKeypoints (reference StitcherClass.py:423-444)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def make_pairs(n_pairs, n_kp, seed=4):
    """Descriptors with ~60 % true correspondences (a few bits flipped), the rest random; points
    related by a known homography with 0.5 px noise and gross outliers."""
    rng = np.random.default_rng(seed)
    fa, fb, pa, pb = [], [], [], []
    for p in range(n_pairs):
        b = rng.integers(0, 256, size=(n_kp, 32), dtype=np.uint8)
        a = rng.integers(0, 256, size=(n_kp, 32), dtype=np.uint8)
        n_true = int(0.6 * n_kp)
        perm = rng.permutation(n_kp)[:n_true]
        a[:n_true] = b[perm]
        a[:n_true, :4] ^= rng.integers(0, 256, size=(n_true, 4), dtype=np.uint8) & 0x11
        H = np.array([[0.97, 0.02, 40.0 + 5 * p], [-0.015, 0.98, 12.0], [1.5e-5, -1e-5, 1.0]])
        xa = rng.uniform([0, 0], [1920, 1080], size=(n_kp, 2))
        h = np.c_[xa, np.ones(n_kp)] @ H.T
        xb_true = h[:, :2] / h[:, 2:3] + rng.normal(0, 0.5, size=(n_kp, 2))
        xb = rng.uniform([0, 0], [1920, 1080], size=(n_kp, 2))
        xb[perm] = xb_true[:n_true]
        fa.append(a); fb.append(b); pa.append(xa.astype(np.float32)); pb.append(xb.astype(np.float32))
    return np.stack(fa), np.stack(fb), np.stack(pa), np.stack(pb)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--pairs", type=int, default=4)
    ap.add_argument("--keypoints", type=int, default=2000)
    ap.add_argument("--hypotheses", type=int, default=2000)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()

    import cv2
    import torch
    from multicamera_stitching_b200 import _cabi, recalib
    if not torch.cuda.is_available():
        raise SystemExit("bench_recalib.py: no CUDA device (there is no CPU fallback)")
    _cabi.load()
    dev = torch.device("cuda", 0)
    fa, fb, pa, pb = make_pairs(args.pairs, args.keypoints)
    q = torch.from_numpy(fa).to(dev)
    t = torch.from_numpy(fb).to(dev)

    # ---- parity spot check against cv2 (pair 0)
    idx2, dist2, keep = recalib.match_top2_batch(q, t, ratio=0.75)
    raw = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(fa[0], fb[0], 2)
    ref_idx = np.array([[m[0].trainIdx, m[1].trainIdx] for m in raw], dtype=np.int32)
    ref_keep = np.array([m[0].distance < m[1].distance * 0.75 for m in raw])
    assert np.array_equal(idx2[0].cpu().numpy(), ref_idx), "match indices differ from cv2"
    assert np.array_equal(keep[0].cpu().numpy().astype(bool), ref_keep), "ratio test differs from cv2"

    # matched point sets for RANSAC (padded to a common length)
    keep_h = keep.cpu().numpy().astype(bool)
    idx_h = idx2[:, :, 0].cpu().numpy()
    n_m = [int(k.sum()) for k in keep_h]
    n_max = max(n_m)
    A = np.zeros((args.pairs, n_max, 2), np.float32)
    B = np.zeros((args.pairs, n_max, 2), np.float32)
    for p in range(args.pairs):
        qi = np.nonzero(keep_h[p])[0]
        A[p, :len(qi)] = pa[p][qi]
        B[p, :len(qi)] = pb[p][idx_h[p][qi]]
    dA, dB = torch.from_numpy(A).to(dev), torch.from_numpy(B).to(dev)
    dn = torch.tensor(n_m, dtype=torch.int32, device=dev)
    samples = np.stack([recalib.draw_samples(n_m[p], args.hypotheses, seed=7 + p) for p in range(args.pairs)])
    ds = torch.from_numpy(samples).to(dev)

    def timed(fn):
        for _ in range(args.warmup):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _cabi.launch_count()
        e0.record()
        for _ in range(args.steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / args.steps, (_cabi.launch_count() - l0) // args.steps

    ms_match, l_match = timed(lambda: recalib.match_top2_batch(q, t, ratio=0.75))
    ms_ransac, l_ransac = timed(lambda: recalib.ransac_batch(dA, dB, ds, 3.0, n=dn))
    counts, H_k, best, mask = recalib.ransac_batch(dA, dB, ds, 3.0, n=dn)
    inl = [int(mask[p, :n_m[p]].sum().item()) for p in range(args.pairs)]

    popc = args.pairs * args.keypoints * args.keypoints * 8   # 32-bit popcounts per launch
    line = {
        "metric": "recalib_pairs_per_sec", "unit": "pairs/s", "higher_is_better": True, "n_gpus": 1,
        "value": args.pairs / ((ms_match + ms_ransac) * 1e-3),
        "steps": args.steps, "warmup": args.warmup, "dtype": "u8 popcount / f32 reprojection", "data": "synthetic",
        "config": {"workload": "cfg4_recalib", "pairs": args.pairs, "keypoints": args.keypoints,
                   "descriptor_bytes": 32, "ransac_hypotheses": args.hypotheses, "matches_per_pair": n_m,
                   "inliers_per_pair": inl},
        "match": {"ms_per_launch": ms_match, "launches": l_match, "popc32_per_s": popc / (ms_match * 1e-3),
                  "indices_equal_cv2": True},
        "ransac": {"ms_per_launch": ms_ransac, "launches": l_ransac,
                   "projections_per_s": args.hypotheses * sum(n_m) / (ms_ransac * 1e-3)},
        "gpu_launches": (l_match + l_ransac) * args.steps,
    }
    if not args.no_cpu:
        cv2.setNumThreads(os.cpu_count() or 1)
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            for p in range(args.pairs):
                raw = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(fa[p], fb[p], 2)
                [m for m in raw if len(m) == 2 and m[0].distance < m[1].distance * 0.75]
        t_match = (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()
        for _ in range(reps):
            for p in range(args.pairs):
                cv2.findHomography(A[p, :n_m[p]], B[p, :n_m[p]], cv2.RANSAC, 3.0)
        t_ransac = (time.perf_counter() - t0) / reps
        line["cpu_baseline"] = {"value": args.pairs / (t_match + t_ransac), "unit": "pairs/s",
                                "cores": cv2.getNumThreads(), "kind": "port",
                                "match_ms": 1e3 * t_match, "ransac_ms": 1e3 * t_ransac,
                                "sample": "%d x %d pairs: cv2 %s BFMatcher(NORM_HAMMING).knnMatch(k=2) + ratio "
                                          "loop, findHomography(RANSAC, 3.0) (StitcherClass.py:423-444)"
                                          % (reps, args.pairs, cv2.__version__)}
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
