"""BASELINE.json config 5: a 10 000-frame synthetic 6-camera 1080p sequence sharded by frame range
across the GPUs of one box, end to end (pinned host frames in, pinned host panoramas out).

The sequence is not materialised (373 GB): every rank cycles a pinned ring of R distinct frame-sets,
frame f of the sequence living in slot f % R, and writes its panoramas into a pinned ring of the same
length (SURVEY.md section 8 d).  Rank r composites frames [r * N / G, (r + 1) * N / G) through
`sequence.SequencePipeline.run(..., ring=True)`; there is no data-path collective, torch.distributed
only carries the barrier and the max-over-ranks time.  One JSON line from rank 0.

  python scripts/bench_sequence.py [--frames 10000]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node G --master-addr 127.0.0.1 scripts/bench_sequence.py
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from helpers import synthetic_chain  # noqa: E402
from multicamera_stitching_b200 import synthetic  # noqa: E402
from multicamera_stitching_b200.sequence import SequencePipeline, pinned_like, shard_range  # noqa: E402
from multicamera_stitching_b200.shard import ShardContext, bind_to_gpu_numa  # noqa: E402
from oracle import stitcher_ref  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=10000)
    ap.add_argument("--ring", type=int, default=32)
    ap.add_argument("--chunk", type=int, default=16)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    bind_to_gpu_numa(local_rank)
    ctx = ShardContext.from_env(backend="nccl", device=device)

    st, states, labels, images = synthetic_chain(6, 1080, 1920, 3, kind="smooth")
    shapes = [images[l].shape for l in labels]
    pipe = SequencePipeline(st, shapes, device, chunk=args.chunk, depth=3)
    R = args.ring
    distinct = 8
    sets = [synthetic.make_frames(6, 1080, 1920, 3, frame_index=rank * 1000 + f, kind="smooth") for f in range(distinct)]
    host = {l: pinned_like((R,) + tuple(images[l].shape)) for l in labels}
    for l in labels:
        for f in range(R):
            host[l][f].copy_(torch.from_numpy(sets[f % distinct][l]))
    out = pinned_like((R,) + pipe.plan.out_shape())
    lo, hi = shard_range(args.frames, world, rank)

    pipe.run(host, out, lo, min(hi, lo + 3 * R), ring=True)          # warm-up, also the parity sample
    torch.cuda.synchronize()
    last = min(hi, lo + 3 * R) - 1
    ref = stitcher_ref.stitch_chain(states, labels, sets[(last % R) % distinct])
    exact = bool(np.array_equal(out[last % R].numpy(), ref))

    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    done = pipe.run(host, out, lo, hi, ring=True)
    e1.record()
    torch.cuda.synchronize()
    ctx.barrier()
    ms = ctx.max_over_ranks(e0.elapsed_time(e1))
    h2d, d2h = pipe.bytes_per_frame()
    ok = ctx.max_over_ranks(0.0 if exact else 1.0) == 0.0
    if rank == 0:
        print(json.dumps({"metric": "panoramas_per_sec", "workload": "cfg5: %d-frame 6 x 1080p sequence, frame-range shards" % args.frames,
                          "n_gpus": world, "value": args.frames / (ms * 1e-3), "unit": "panoramas/s", "seconds": ms * 1e-3,
                          "frames_rank0": done, "h2d_bytes_per_panorama": h2d, "d2h_bytes_per_panorama": d2h,
                          "host_to_device_GBps": h2d * args.frames / ms / 1e6, "device_to_host_GBps": d2h * args.frames / ms / 1e6,
                          "ring_frame_sets": R, "chunk": args.chunk, "bit_exact_sample_all_ranks": ok,
                          "collectives_on_data_path": 0}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
