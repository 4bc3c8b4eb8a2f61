"""Where the time of one batched `matchKeypoints` call goes (config 4: four 1080p pairs, ORB-2000)."""
import cProfile
import os
import pstats
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from multicamera_stitching_b200 import StitcherBase, recalib, synthetic  # noqa: E402


def main():
    items = []
    for k in range(4):
        imageB, imageA, _ = synthetic.make_pair(1080, 1920, seed=k)
        sb = StitcherBase()
        kA, fA = sb.detectAndDescribe(imageA)
        kB, fB = sb.detectAndDescribe(imageB)
        items.append((kA, kB, fA, fB))
    for _ in range(5):
        recalib.match_keypoints_batch(items, 0.75, 3.0)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(100):
        recalib.match_keypoints_batch(items, 0.75, 3.0)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / 100 * 1e3
    print("match_keypoints_batch, 4 pairs: %.3f ms per call = %.0f pairs/s" % (ms, 4e3 / ms))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(50):
        recalib.match_keypoints_batch(items, 0.75, 3.0)
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(14)


if __name__ == "__main__":
    main()
