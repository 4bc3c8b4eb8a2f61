"""Phases of one numpy-in / numpy-out `Stitcher.stitch` call after the staging change (config 2)."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from multicamera_stitching_b200 import engine as E  # noqa: E402
from multicamera_stitching_b200 import synthetic  # noqa: E402


def timed(fn, n=40):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
        torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3


def main():
    st, _, labels, images = synthetic.synthetic_stitcher(6, 1080, 1920, 3, kind="smooth")
    dev = torch.device("cuda", 0)
    eng = st._engine_()
    frames = [images[l] for l in labels]
    cams = list(range(len(frames)))
    for piece in (1 << 19, 1 << 20, 2 << 20, 4 << 20):
        eng.STAGE_PIECE_BYTES = piece
        for thr in ("1", "3", "4", "6", "8"):
            os.environ["MCS_UPLOAD_THREADS"] = thr
            print("piece %4d KB threads %s: whole call %.3f ms   upload_frames+sync %.3f ms" % (
                piece >> 10, thr, timed(lambda: st.stitch(images)),
                timed(lambda: eng.upload_frames(cams, frames, dev, eng.plan_for(st.stitchers, [f.shape for f in frames], dev).upload_bands()))))
    eng.STAGE_PIECE_BYTES = 2 << 20
    os.environ["MCS_UPLOAD_THREADS"] = "4"
    plan = eng.plan_for(st.stitchers, [f.shape for f in frames], dev)
    bands = plan.upload_bands()
    print("copies per camera:", [len(bands[c][2]) for c in cams])
    d = eng.upload_frames(cams, frames, dev, bands)
    dl = [d[c] for c in cams]
    pins = [eng._pinned_in[(c, tuple(frames[c].shape))][0] for c in cams]
    print("DMA from pinned, issue + sync   %.3f ms" % timed(lambda: [eng.upload(c, pins[c], dev, bands[c]) for c in cams]))
    t0 = time.perf_counter()
    for _ in range(20):
        [eng.upload(c, pins[c], dev, bands[c]) for c in cams]
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print("DMA from pinned, issue only     %.3f ms" % ((t1 - t0) / 20 * 1e3))
    res = plan.run(dl)
    print("plan.run + sync                 %.3f ms" % timed(lambda: plan.run(dl)))
    print("download (pool)                 %.3f ms" % timed(lambda: eng.download(res)))
    print("plan_for + bands                %.3f ms" % timed(lambda: eng.plan_for(st.stitchers, [f.shape for f in frames], dev).upload_bands()))


if __name__ == "__main__":
    main()
