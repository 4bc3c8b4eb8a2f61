"""Where the time of one `Stitcher.stitch(images_dic)` call with numpy frames goes (config 2, one frame-set per
call): the window uploads from pageable memory, the launch, the download into a fresh pageable array - and what
the alternatives cost (pinned result buffers, frames staged through pinned memory by host threads)."""
import os
import sys
import time
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from multicamera_stitching_b200 import synthetic  # noqa: E402


def timed(fn, n=30):
    fn()
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
    print("whole call                     %.3f ms" % timed(lambda: st.stitch(images)))
    eng = st._engine_()
    frames = [images[l] for l in labels]
    plan = eng.plan_for(st.stitchers, [f.shape for f in frames], dev)
    bands = plan.upload_bands()

    def up():
        return [eng.upload(c, frames[c], dev, bands[c]) for c in range(len(frames))]

    print("uploads (windows, pageable)    %.3f ms" % timed(up))
    print("uploads (whole, pageable)      %.3f ms" % timed(
        lambda: [eng.upload(c, frames[c], dev, None) for c in range(len(frames))]))
    d = up()
    res = plan.run(d)
    print("launch                         %.3f ms" % timed(lambda: plan.run(d)))

    def down_fresh():
        h = torch.empty(res.shape, dtype=torch.uint8)
        h.copy_(res)
        return h

    print("download, fresh pageable       %.3f ms" % timed(down_fresh))
    keep = torch.empty(res.shape, dtype=torch.uint8)
    print("download, reused pageable      %.3f ms" % timed(lambda: keep.copy_(res)))
    pin = torch.empty(res.shape, dtype=torch.uint8, pin_memory=True)

    def down_pin():
        pin.copy_(res, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    print("download, pinned               %.3f ms" % timed(down_pin))

    # frames staged through pinned memory by host threads, one camera per task, DMA as each lands
    pins = [torch.empty(f.shape, dtype=torch.uint8, pin_memory=True) for f in frames]
    pin_np = [p.numpy() for p in pins]
    for nthr in (1, 2, 3, 6):
        pool = ThreadPoolExecutor(nthr)

        def stage(c):
            row, h, copies = bands[c]
            f2 = frames[c].reshape(h, row)
            p2 = pin_np[c].reshape(h, row)
            for w in copies:
                p2[w["y0"]:w["y0"] + w["rows"], w["b0"]:w["b0"] + w["nbytes"]] = \
                    f2[w["y0"]:w["y0"] + w["rows"], w["b0"]:w["b0"] + w["nbytes"]]
            return c

        def up_staged():
            out = []
            for c in pool.map(stage, range(len(frames))):
                out.append(eng.upload(c, pins[c], dev, bands[c]))
            return out

        print("uploads staged by %d threads    %.3f ms" % (nthr, timed(up_staged)))
        pool.shutdown()
    whole_bytes = sum(f.nbytes for f in frames)
    win_bytes = sum(w["nbytes"] * w["rows"] for c in range(len(frames)) for w in bands[c][2])
    print("bytes: whole %.1f MB, windows %.1f MB, panorama %.1f MB" % (whole_bytes / 1e6, win_bytes / 1e6, res.numel() / 1e6))


if __name__ == "__main__":
    main()
