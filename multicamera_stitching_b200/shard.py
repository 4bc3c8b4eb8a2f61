"""Multi-GPU plumbing of the sequence path: one process per GPU, frame-range sharding.

The compositing path has no cross-GPU data dependency (every panorama needs only its own
frames and the replicated calibration; reference StitcherClass.py:114-136 keeps no state
between calls), so ``torch.distributed`` is used for exactly three things, none of them on
the data path: the barrier around a timed region, the max-over-ranks of a device time, and
gathering small per-rank summaries (counts, checksums) to rank 0.  Backend ``nccl`` on the
GPU box, ``gloo`` in the CPU tests (tests/test_sharding_gloo.py).
"""
import os

import torch
import torch.distributed as dist

from .sequence import shard_range


def bind_to_gpu_numa(device_index):
    """Pin the calling thread to the CPUs closest to GPU ``device_index`` (NVML's ideal affinity),
    so that the pinned host buffers it allocates next live on that GPU's NUMA node and every
    rank's host<->device traffic stays on its own socket.  Best effort: returns False when NVML
    is missing or refuses."""
    try:
        import pynvml
        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(device_index))
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return True
    except Exception:
        return False


class ShardContext(object):
    """Rank / world size of this process plus the three collectives the path needs."""

    def __init__(self, rank=0, world_size=1, device=None):
        self.rank = int(rank)
        self.world_size = int(world_size)
        self.device = device            # tensor device for the collectives (cuda for nccl, cpu for gloo)
        self.initialised_here = False

    @classmethod
    def from_env(cls, backend=None, device=None):
        """Read RANK / WORLD_SIZE / MASTER_* (torchrun's contract); initialise the default
        process group when WORLD_SIZE > 1."""
        rank = int(os.environ.get("RANK", "0"))
        world = int(os.environ.get("WORLD_SIZE", "1"))
        ctx = cls(rank, world, device)
        if world > 1 and not dist.is_initialized():
            if backend is None:
                backend = "nccl" if (device is not None and torch.device(device).type == "cuda") else "gloo"
            kwargs = {}
            if backend == "nccl" and device is not None:
                kwargs["device_id"] = torch.device(device)
            dist.init_process_group(backend, rank=rank, world_size=world, **kwargs)
            ctx.initialised_here = True
        return ctx

    def close(self):
        if self.initialised_here and dist.is_initialized():
            dist.destroy_process_group()
        self.initialised_here = False

    # -- sharding ----------------------------------------------------------------
    def frame_range(self, n_frames):
        """Contiguous ``[lo, hi)`` of this rank (see ``sequence.shard_range``)."""
        return shard_range(n_frames, self.world_size, self.rank)

    # -- collectives (control path only) --------------------------------------------
    def barrier(self):
        if self.world_size > 1:
            dist.barrier()
        if self.device is not None and torch.device(self.device).type == "cuda":
            torch.cuda.synchronize(self.device)

    def _reduce(self, value, op):
        if self.world_size == 1:
            return float(value)
        t = torch.tensor([float(value)], dtype=torch.float64, device=self.device or "cpu")
        dist.all_reduce(t, op=op)
        return float(t.item())

    def max_over_ranks(self, value):
        return self._reduce(value, dist.ReduceOp.MAX)

    def sum_over_ranks(self, value):
        return self._reduce(value, dist.ReduceOp.SUM)

    def gather_frame_summaries(self, lo, summaries):
        """Rank 0 receives every rank's per-frame summaries (a 1-D int64 tensor, one entry
        per frame of ``[lo, lo + len)``) ordered by frame index; other ranks get ``None``.
        Used for checksums / counts only - panoramas themselves never cross ranks."""
        summaries = torch.as_tensor(summaries, dtype=torch.int64).reshape(-1)
        if self.world_size == 1:
            return summaries.clone()
        dev = self.device or "cpu"
        meta = torch.tensor([int(lo), int(summaries.numel())], dtype=torch.int64, device=dev)
        metas = [torch.zeros_like(meta) for _ in range(self.world_size)]
        dist.all_gather(metas, meta)
        n_max = max(int(m[1]) for m in metas)
        padded = torch.zeros(max(n_max, 1), dtype=torch.int64, device=dev)
        padded[:summaries.numel()] = summaries.to(dev)
        parts = [torch.zeros_like(padded) for _ in range(self.world_size)]
        dist.all_gather(parts, padded)
        if self.rank != 0:
            return None
        total = sum(int(m[1]) for m in metas)
        out = torch.zeros(total, dtype=torch.int64)
        for m, p in zip(metas, parts):
            lo_r, n_r = int(m[0]), int(m[1])
            out[lo_r:lo_r + n_r] = p[:n_r].cpu()
        return out
