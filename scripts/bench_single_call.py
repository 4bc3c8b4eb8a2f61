"""Latency of the reference's own call, one frame-set at a time (`Stitcher.stitch(images_dic)`,
StitcherClass.py:114-136) on config 2: numpy frames in / numpy panorama out (pageable host memory on
both sides, like a caller of the reference has), CUDA tensors in / CUDA tensor out, and the cv2 chain
on the host beside them.  One JSON line."""
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

from helpers import synthetic_chain  # noqa: E402
from oracle import stitcher_ref  # noqa: E402


def timed(fn, n):
    fn()
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def main():
    st, states, labels, images = synthetic_chain(6, 1080, 1920, 3, kind="smooth")
    dev = torch.device("cuda", 0)
    ref = stitcher_ref.stitch_chain(states, labels, images)
    got = st.stitch(images)
    exact = bool(np.array_equal(got, ref))
    ms_numpy = timed(lambda: st.stitch(images), 30)
    dimages = {l: torch.from_numpy(images[l]).to(dev) for l in labels}
    ms_cuda = timed(lambda: st.stitch(dimages), 200)
    cv2.setNumThreads(os.cpu_count())
    t0 = time.perf_counter()
    for _ in range(10):
        stitcher_ref.stitch_chain(states, labels, images)
    ms_cv2 = (time.perf_counter() - t0) / 10 * 1e3
    print(json.dumps({"op": "Stitcher.stitch(images_dic), one cfg2 frame-set per call", "bit_exact_vs_cv2": exact,
                      "ms_per_call_numpy_in_numpy_out": ms_numpy, "ms_per_call_cuda_in_cuda_out": ms_cuda,
                      "ms_per_call_cv2_chain": ms_cv2, "host_threads": os.cpu_count()}))


if __name__ == "__main__":
    main()
