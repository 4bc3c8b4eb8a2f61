#!/usr/bin/env python
"""Benchmark of the fused warp+paste compositing path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one pass of the hot path over one batch of synthetic frame-sets
(default: BASELINE.json config 2, 6 x 1080p cameras, a batch of 64 frame-sets
resident in HBM, ~2.4 GB of input per step so consecutive steps never see a
warm L2; the end-to-end leg moves 32 frame-sets per step).  Rank 0 prints ONE JSON line:

  value     panoramas/s, whole job, inputs already resident in HBM
  e2e       the same metric through the host-facing sequence API: pinned host
            frames -> H2D -> kernel -> D2H -> pinned host panoramas, every step
  roofline  algorithmic bytes per launch / CUDA-event launch duration against
            the measured HBM copy bandwidth (MEASURED_PEAKS.json)
  cpu_baseline  the reference's own OpenCV chain (oracle/stitcher_ref.py) timed
            on this box's host cores on a bounded sample of the same workload

``--impl reference`` times only that CPU chain (all host threads) and prints
the same line shape with ``"impl": "reference"``.

With N > 1 (torchrun, one rank per GPU) the sequence shards by frame range:
every rank composites its own batch, there is no data-path collective
(torch.distributed is used for the barrier and the max-over-ranks time only).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402

WORKLOADS = {
    # name: (n_cams, H, W, batch of frame-sets per step, e2e frame-sets per step)
    "cfg1_3x720p": (3, 720, 1280, 64, 64),
    "cfg2_6x1080p": (6, 1080, 1920, 64, 32),
    "cfg3_8x2160p": (8, 2160, 3840, 16, 8),
    "ns_8x1080p": (8, 1080, 1920, 48, 24),     # the geometry north_star's 70 % target names
}
FALLBACK_HBM_GBS = 6650.0


def build_chain(name):
    """Calibrated Stitcher + oracle states for a workload (host only)."""
    from helpers import synthetic_chain
    n, h, w, _, _ = WORKLOADS[name]
    return synthetic_chain(n, h, w, 3, kind="smooth")


def recorded_traffic(workload, batch):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (None if the
    capture was taken at another batch size or does not exist)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_dram_traffic.json")) as f:
            rec = json.load(f)[workload]
        if int(rec["panoramas_per_launch"]) != int(batch):
            return None
        return int(rec["dram_read_bytes"]) + int(rec["dram_write_bytes"])
    except Exception:
        return None


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ---------------------------------------------------------------------------
class ClockSampler(object):
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    QUERY = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        try:
            for line in open(self.path):
                parts = [p.strip() for p in line.split(",")]
                if len(parts) < 7:
                    continue
                try:
                    sm.append(float(parts[0]))
                    mx.append(float(parts[1]))
                except ValueError:
                    continue
                for nm, v in zip(names, parts[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            os.unlink(self.path)
        except Exception:
            pass
        self.sm, self.mx, self.reasons = sm, mx, reasons
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        return out

    @staticmethod
    def merged(*samplers):
        """One clocks record over the timed regions of several samplers."""
        sm, mx, reasons = [], [], set()
        for s in samplers:
            sm += getattr(s, "sm", [])
            mx += getattr(s, "mx", [])
            reasons |= getattr(s, "reasons", set())
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": sorted(reasons), "samples": len(sm)}
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx))
        return out


# ---------------------------------------------------------------------------
def time_cpu_chain(states, labels, frame_sets, budget_s, min_panos=4, threads=None):
    """Reference CPU path: the sequential cv2 chain of StitcherClass.py:131-136."""
    import cv2
    from oracle import stitcher_ref
    if threads is not None:
        cv2.setNumThreads(threads)
    for fs in frame_sets[:2]:
        stitcher_ref.stitch_chain(states, labels, fs)  # warm-up
    n = 0
    t0 = time.perf_counter()
    while True:
        stitcher_ref.stitch_chain(states, labels, frame_sets[n % len(frame_sets)])
        n += 1
        dt = time.perf_counter() - t0
        if (dt >= budget_s and n >= min_panos) or n >= 100000:
            break
    return n / dt, n, dt, cv2.getNumThreads()


def make_frame_sets(name, count):
    from multicamera_stitching_b200 import synthetic
    n, h, w, _, _ = WORKLOADS[name]
    return [synthetic.make_frames(n, h, w, 3, frame_index=f, kind="smooth") for f in range(count)]


# ---------------------------------------------------------------------------
def run_reference(args, rank, world):
    if rank != 0:
        return
    st, states, labels, images = build_chain(args.workload)
    frame_sets = make_frame_sets(args.workload, 4)
    import cv2
    from oracle import stitcher_ref
    cv2.setNumThreads(os.cpu_count() or 1)
    ref = stitcher_ref.stitch_chain(states, labels, frame_sets[0])
    out_h, out_w = ref.shape[:2]
    per_step = args.ref_panos_per_step
    for _ in range(args.warmup):
        for i in range(per_step):
            stitcher_ref.stitch_chain(states, labels, frame_sets[i % 4])
    t0 = time.perf_counter()
    for _ in range(args.steps):
        for i in range(per_step):
            stitcher_ref.stitch_chain(states, labels, frame_sets[i % 4])
    dt = time.perf_counter() - t0
    pps = args.steps * per_step / dt
    line = {
        "impl": "reference", "metric": "panoramas_per_sec", "value": pps, "unit": "panoramas/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "output_mp_per_s": pps * out_w * out_h / 1e6,
        "config": {"workload": args.workload, "cameras": WORKLOADS[args.workload][0],
                   "frame_hw": list(WORKLOADS[args.workload][1:3]), "panorama_wh": [out_w, out_h],
                   "panoramas_per_step": per_step},
        "cpu_baseline": {"value": pps, "unit": "panoramas/s", "cores": cv2.getNumThreads(), "kind": "port",
                         "sample": "%d steps x %d panoramas, cv2 %s chain (warpPerspective + paste, "
                                   "StitcherClass.py:131-136, :239-241), %d threads"
                                   % (args.steps, per_step, cv2.__version__, cv2.getNumThreads())},
        "e2e": {"value": pps, "unit": "panoramas/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------
def run_ours(args, rank, local_rank, world):
    import torch
    from multicamera_stitching_b200 import _cabi
    from multicamera_stitching_b200.sequence import SequencePipeline, pinned_like
    from multicamera_stitching_b200.shard import ShardContext, bind_to_gpu_numa

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (there is no CPU fallback for the GPU arm)")
    _cabi.load()
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    # pinned buffers of this rank on its GPU's socket ($MCS_BENCH_NO_NUMA=1 leaves the threads alone)
    numa_bound = False if os.environ.get("MCS_BENCH_NO_NUMA") else bind_to_gpu_numa(local_rank)
    ctx = ShardContext.from_env(backend="nccl", device=device)   # control path only: barrier + max time
    barrier = ctx.barrier
    max_over_ranks = ctx.max_over_ranks

    n_cams, H, W, batch, e2e_batch = WORKLOADS[args.workload]
    if args.batch:
        batch = e2e_batch = args.batch
    st, states, labels, images = build_chain(args.workload)
    st.feather_log2 = args.feather   # 0 = the reference's overwrite (the headline); n = feather over 2**n px
    shapes = [images[l].shape for l in labels]
    plan = st.plan(shapes, device)
    out_w, out_h = plan.out_w, plan.out_h
    algo_bytes = plan.algorithmic_bytes()

    # Device-resident ring of distinct frame-sets (cycled to fill the batch): frames of
    # rank r start at frame index r * batch so every rank works on its own frame range.
    distinct = min(batch, 8)
    ring = make_frame_sets_offset(args.workload, distinct, rank * batch)
    dev_frames = {}
    for l in labels:
        stack = np.stack([ring[f % distinct][l] for f in range(batch)])
        dev_frames[l] = torch.from_numpy(stack).to(device)
    # panorama rows padded to a 128-byte pitch in HBM: every 128-column cell row of the tiled kernel
    # then starts and ends on a 32-byte sector boundary (no partial-sector writes)
    out = plan.new_output(batch, pitch_align=args.pitch_align)
    in_bytes = sum(int(t.numel()) for t in dev_frames.values())

    # ---- parity spot-check of what is about to be timed (outside the timed region)
    st.stitch_batch(dev_frames, out=out)
    torch.cuda.synchronize()
    if rank == 0:
        from oracle import stitcher_ref
        if args.feather:
            from oracle import feather_model
            ref = feather_model.feather_chain(states, labels, ring[0], args.feather)
        else:
            ref = stitcher_ref.stitch_chain(states, labels, ring[0])
        got = out[0].cpu().numpy()
        d = np.abs(got.astype(np.int16) - ref.astype(np.int16))
        parity = {"max_abs_diff": int(d.max()), "exact_fraction": float((d == 0).mean())}
        if (parity["max_abs_diff"] > 1 or parity["exact_fraction"] < 0.999) and not os.environ.get("MCS_BENCH_ABLATION"):
            raise SystemExit("bench.py: GPU panorama differs from the cv2 chain: %r" % (parity,))
    else:
        parity = None

    # ---- kernel-resident timing ------------------------------------------------
    for _ in range(args.warmup):
        st.stitch_batch(dev_frames, out=out)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = _cabi.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        st.stitch_batch(dev_frames, out=out)
    e1.record()
    barrier()
    launches = _cabi.launch_count() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    pps = world * batch * args.steps / (ms_total * 1e-3)
    launch_ms = ms_total / max(launches, 1)
    # algorithmic bytes of a step over the time of a step (a step is one launch in the reference's
    # overwrite mode; the feather mode adds a second, small launch over the seam bands)
    achieved = algo_bytes * batch / (ms_step * 1e-3) / 1e9
    peak, peak_src = measured_peak()

    # ---- end to end through the host-facing sequence API -------------------------
    if args.no_e2e:   # kernel experiments only: the line then carries no end-to-end number
        ctx.close()
        if rank == 0:
            print(json.dumps({"metric": "panoramas_per_sec", "value": pps, "ms_per_step": ms_step,
                              "roofline": {"achieved": achieved, "frac": achieved / peak}, "parity": parity,
                              "gpu_launches": launches, "clocks": clocks, "e2e": None,
                              "ctas_per_sm": plan.handle.tiled_ctas_per_sm(),
                              "tiled": plan.handle.tiled_stats()}), flush=True)
        return
    pipe = SequencePipeline(st, shapes, device, chunk=args.chunk, depth=3, windows=not args.whole_frames)
    host_frames = {l: pinned_like((e2e_batch,) + tuple(images[l].shape)) for l in labels}
    for l in labels:
        for f in range(e2e_batch):
            host_frames[l][f].copy_(torch.from_numpy(ring[f % distinct][l]))
    host_out = pinned_like((e2e_batch,) + plan.out_shape())
    h2d_b, d2h_b = pipe.bytes_per_frame()
    for _ in range(max(1, min(args.warmup, 3))):
        pipe.run(host_frames, host_out)
    barrier()
    if rank == 0:
        got = host_out[0].numpy()
        d = np.abs(got.astype(np.int16) - ref.astype(np.int16))
        if int(d.max()) > 1:
            raise SystemExit("bench.py: e2e panorama differs from the cv2 chain")
    e2e_steps = max(3, min(args.steps, 10))
    sampler2 = ClockSampler(local_rank)
    if rank == 0:
        sampler2.start()
    # The steps are consecutive batches of one streaming pipeline: frame f of the timed sequence lives
    # in slot f % batch of the pinned input ring and its panorama lands in the same slot of the pinned
    # output ring, so every step uploads its inputs and downloads its panoramas, and the pipeline is
    # not drained between steps (it is at the end: run() returns after the last device->host copy).
    e0.record()
    pipe.run(host_frames, host_out, 0, e2e_steps * e2e_batch, ring=True)
    e1.record()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    if rank == 0:
        sampler2.stop()
        clocks = ClockSampler.merged(sampler, sampler2)   # both timed regions
    e2e_pps = world * e2e_batch * e2e_steps / (e2e_ms * 1e-3)

    # ---- CPU baseline (rank 0, N = 1 only): the reference's cv2 chain ----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        import cv2
        cpu_pps, n, dt, threads = time_cpu_chain(states, labels, ring[:4], args.cpu_budget,
                                                 threads=os.cpu_count() or 1)
        cpu = {"value": cpu_pps, "unit": "panoramas/s", "cores": threads, "kind": "port",
               "sample": "%d panoramas of %s in %.1f s: cv2 %s warpPerspective+paste chain "
                         "(oracle/stitcher_ref.py, StitcherClass.py:131-136), %d threads"
                         % (n, args.workload, dt, cv2.__version__, threads)}

    ctx.close()
    if rank != 0:
        return
    line = {
        "metric": "panoramas_per_sec", "value": pps, "unit": "panoramas/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic", "output_mp_per_s": pps * out_w * out_h / 1e6,
        "config": {"workload": args.workload, "cameras": n_cams, "frame_hw": [H, W],
                   "panorama_wh": [out_w, out_h], "panoramas_per_step_per_gpu": batch,
                   "sharding": "frame range per rank, no collective",
                   "l2": "inputs per step %.0f MB + outputs %.0f MB per GPU, larger than the 126 MB L2"
                         % (in_bytes / 1e6, out.numel() / 1e6),
                   "panorama_pitch_bytes": int(out.stride(1)), "kernel_variant": plan.handle.last_variant(), "feather_log2": args.feather,
                   "tiled_ctas_per_sm": plan.handle.tiled_ctas_per_sm()},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": recorded_traffic(args.workload, batch),
                     "traffic_source": "ncu dram__bytes_read+write per launch, profiles/r1_dram_traffic.json",
                     "algorithmic_bytes_per_launch": algo_bytes * batch, "launches_per_step": launches // max(args.steps, 1),
                     "peak_source": peak_src,
                     "algorithmic_bytes_per_panorama": algo_bytes,
                     "launch_ms": launch_ms, "panoramas_per_launch": batch},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_pps, "unit": "panoramas/s", "h2d_bytes_per_step": h2d_b * e2e_batch,
                "d2h_bytes_per_step": d2h_b * e2e_batch, "steps": e2e_steps,
                "ms_per_step": e2e_ms / e2e_steps, "api": "sequence.SequencePipeline.run(ring=True) (pinned host in/out, %d frame-sets per chunk)" % args.chunk,
                "numa_bound": bool(numa_bound)},
        "gpu_launches": launches,
        "clocks": clocks,
        "parity": parity,
    }
    print(json.dumps(line), flush=True)


def make_frame_sets_offset(name, count, first_frame):
    from multicamera_stitching_b200 import synthetic
    n, h, w, _, _ = WORKLOADS[name]
    return [synthetic.make_frames(n, h, w, 3, frame_index=first_frame + f, kind="smooth") for f in range(count)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2_6x1080p", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="frame-sets per step per GPU (0 = workload default)")
    ap.add_argument("--whole-frames", action="store_true",
                    help="e2e path uploads whole camera frames instead of the windows the panorama can see")
    ap.add_argument("--chunk", type=int, default=16, help="frame-sets per pipeline chunk of the e2e path")
    ap.add_argument("--cpu-budget", type=float, default=12.0, help="seconds of CPU baseline work")
    ap.add_argument("--pitch-align", type=int, default=128, help="row pitch alignment of the device-resident panoramas")
    ap.add_argument("--feather", type=int, default=0, help="feather blend over 2**n pixels (0 = reference overwrite)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="kernel experiments: skip the end-to-end leg")
    ap.add_argument("--ref-panos-per-step", type=int, default=4)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
