"""Recorded multi-camera sequences: the on-disk format of the reference's capture node
(``data.csv`` + ``data/`` image folder) read into batched frames for the sequence pipeline.

Format (reference TestTrackVision/.../data_capture_node.py:107-130, :171-179): ``<folder>/data.csv``
with the header ``capture_id,timestamp,camera_label,image_file`` and one row per saved image, rows
of one timestamp adjacent, captures numbered 0, 1, ...; the images are ``<folder>/data/<image_file>``
(MediaPlayer/view.py:409, :491), JPEG by default.  ``RecordedSequence`` indexes it the way the
reference's reader does (MediaPlayer/model.py:51-128: ``images[capture][camera][timestamp]``,
``timestamps[capture]``, ``camera_labels`` in order of first appearance, ``get_image``).

Decoding stays on the CPU (cv2.imread on a thread pool, straight into pinned host tensors); the
stitching of the decoded batches is ``sequence.SequencePipeline`` (H2D / kernel / D2H overlapped),
sharded by frame range across ranks with ``sequence.shard_range``.
"""
import csv
import os
from concurrent.futures import ThreadPoolExecutor


HEADER = ["capture_id", "timestamp", "camera_label", "image_file"]


class RecordedSequence(object):
    """Index of a capture folder."""

    def __init__(self, path=None):
        self.path = None
        self.timestamps = [[]]       # [capture] -> list of timestamp strings
        self.camera_labels = {}      # label -> camera index (order of first appearance)
        self.images = [[[]]]         # [capture][camera][timestamp index] -> image file name
        self.line_count = None
        self.header = None
        if path is not None:
            self.load_data(path)

    def load_data(self, path):
        """Read ``<path>/data.csv`` (reference MediaPlayer/model.py:51-128)."""
        self.path = path
        self.timestamps, self.camera_labels, self.images = [], {}, []
        with open(os.path.join(path, "data.csv"), newline="") as f:
            rows = list(csv.reader(f))
        if not rows:
            raise ValueError("%s/data.csv is empty" % path)
        self.header = rows[0]
        self.line_count = len(rows)
        capture, stamp, cam = None, None, 0
        for r in rows[1:]:
            if len(r) < 4:
                raise ValueError("malformed row in data.csv: %r" % (r,))
            cid, ts, label, name = int(r[0]), r[1], r[2], r[3]
            if cid != capture:
                if cid != len(self.images):
                    raise ValueError("capture ids must count 0, 1, 2, ...: got %d after %d captures"
                                     % (cid, len(self.images)))
                capture, stamp = cid, None
                self.images.append([])
                self.timestamps.append([])
            if ts != stamp:
                stamp, cam = ts, 0
                self.timestamps[capture].append(ts)
            if label not in self.camera_labels:
                self.camera_labels[label] = len(self.camera_labels)
            while len(self.images[capture]) <= cam:
                self.images[capture].append([])
            self.images[capture][cam].append(name)
            cam += 1
        if not self.images:
            self.timestamps, self.images = [[]], [[[]]]
        return self

    # -- the reference reader's accessors ---------------------------------------------------------
    def get_image(self, timestamp_idx, camera_idx, capture_idx):
        return self.images[capture_idx][camera_idx][timestamp_idx]

    # -- sizes --------------------------------------------------------------------------------------
    @property
    def n_captures(self):
        return len(self.images)

    def n_frames(self, capture=0):
        """Frame-sets of a capture in which every camera has an image."""
        cams = self.images[capture]
        return min(len(c) for c in cams) if cams and all(len(c) for c in cams) else 0

    def labels(self):
        return sorted(self.camera_labels, key=self.camera_labels.get)

    def image_path(self, capture, camera_label, idx):
        return os.path.join(self.path, "data", self.images[capture][self.camera_labels[camera_label]][idx])

    # -- decoding -----------------------------------------------------------------------------------
    def load_frame_set(self, capture, idx, flags=None):
        """``{camera_label: ndarray}`` of one timestamp (cv2.imread, like MediaPlayer/view.py:374)."""
        import cv2
        out = {}
        for label in self.labels():
            p = self.image_path(capture, label, idx)
            img = cv2.imread(p) if flags is None else cv2.imread(p, flags)
            if img is None:
                raise IOError("cannot read image %s" % p)
            out[label] = img
        return out

    def read_batch(self, capture, lo, hi, out=None, workers=None):
        """Decode frame-sets ``[lo, hi)`` into ``{label: uint8 tensor [hi-lo, H, W, C]}`` (pinned host
        memory when CUDA is available; pass ``out`` to reuse buffers)."""
        import cv2
        import torch
        labels = self.labels()
        n = hi - lo
        if n <= 0:
            raise ValueError("empty frame range [%d, %d)" % (lo, hi))
        if out is None:
            first = self.load_frame_set(capture, lo)
            pin = torch.cuda.is_available()
            out = {l: torch.empty((n,) + first[l].shape, dtype=torch.uint8, pin_memory=pin) for l in labels}
        views = {l: out[l].numpy() for l in labels}

        def job(args):
            label, f = args
            p = self.image_path(capture, label, lo + f)
            img = cv2.imread(p)
            if img is None:
                raise IOError("cannot read image %s" % p)
            if img.shape != views[label][f].shape:
                raise ValueError("%s: image shape %r differs from %r" % (p, img.shape, views[label][f].shape))
            views[label][f] = img

        jobs = [(l, f) for f in range(n) for l in labels]
        with ThreadPoolExecutor(max_workers=workers or min(16, os.cpu_count() or 1)) as pool:
            list(pool.map(job, jobs))
        return out


def write_capture(path, frame_sets, capture_id=0, timestamps=None, quality=80, img_format="jpg", prefix="abcd"):
    """Append one capture in the reference's format (data_capture_node.py:107-130, :171-179):
    ``frame_sets`` is a list of ``{camera_label: ndarray}``.  Returns the timestamps used."""
    import cv2
    os.makedirs(os.path.join(path, "data"), exist_ok=True)
    csv_file = os.path.join(path, "data.csv")
    if not os.path.isfile(csv_file):
        with open(csv_file, "a", newline="") as fd:
            csv.writer(fd).writerow(HEADER)
    stamps = []
    with open(csv_file, "a", newline="") as fd:
        w = csv.writer(fd)
        for i, fs in enumerate(frame_sets):
            ts = int(timestamps[i]) if timestamps is not None else 1565270000000 + 33 * i
            stamps.append(ts)
            for label in fs:
                name = "{}-{}_{}.{}".format(prefix, ts, label, img_format)
                params = [cv2.IMWRITE_JPEG_QUALITY, quality] if img_format in ("jpg", "jpeg") else []
                if not cv2.imwrite(os.path.join(path, "data", name), fs[label], params):
                    raise IOError("cannot write %s" % name)
                w.writerow([capture_id, ts, label, name])
    return stamps


class CaptureRunner(object):
    """Composites frame ranges of one capture: a ``SequencePipeline`` plus two sets of pinned decode buffers,
    allocated once (pinned allocations cost far more than compositing a short range) and reused by every
    :meth:`run`.  While the pipeline composites batch k from one buffer set, a helper thread decodes batch k + 1
    into the other (the reference's player replays the next frames while the current ones are being stitched,
    video_mapping_node.py:105-130)."""

    def __init__(self, stitcher, seq, capture=0, device=None, chunk=16, depth=3, batch=32, workers=None):
        import torch
        from .sequence import SequencePipeline
        self.stitcher, self.seq, self.capture = stitcher, seq, capture
        self.batch, self.workers = int(batch), workers
        self.labels = [str(l) for l in stitcher.img_labels]
        if sorted(self.labels) != sorted(seq.labels()):
            raise ValueError("stitcher cameras %r differ from the recording's %r" % (self.labels, seq.labels()))
        first = seq.load_frame_set(capture, 0)
        shapes = [first[l].shape for l in self.labels]
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.pipe = SequencePipeline(stitcher, shapes, self.device, chunk=chunk, depth=depth)
        self.sets = [None, None]

    def out_shape(self):
        return self.pipe.plan.out_shape()

    def _decode(self, i, s, e):
        """Decode frame-sets ``[s, e)`` into buffer set ``i & 1`` (allocated at the full batch size on first use)."""
        from .sequence import pinned_like
        cur = self.sets[i & 1]
        if cur is None:
            first = self.seq.load_frame_set(self.capture, s)
            cur = self.sets[i & 1] = {l: pinned_like((self.batch,) + tuple(first[l].shape)) for l in self.seq.labels()}
        view = {l: t[:e - s] for l, t in cur.items()}
        self.seq.read_batch(self.capture, s, e, out=view, workers=self.workers)
        return view

    def run(self, lo, hi, out=None):
        """Composite frame-sets ``[lo, hi)``; returns the pinned host tensor ``[hi - lo, H_out, W_out, C]``
        (``out`` when given)."""
        from .sequence import pinned_like
        if out is None:
            out = pinned_like((hi - lo,) + self.out_shape())
        ranges = [(s, min(hi, s + self.batch)) for s in range(lo, hi, self.batch)]
        with ThreadPoolExecutor(max_workers=1) as ahead:
            pending = ahead.submit(self._decode, 0, *ranges[0]) if ranges else None
            for i, (s, e) in enumerate(ranges):
                bufs = pending.result()
                if i + 1 < len(ranges):
                    # the other buffer set: batch i - 1 has left it (SequencePipeline.run blocks until its downloads landed)
                    pending = ahead.submit(self._decode, i + 1, *ranges[i + 1])
                self.pipe.run({l: bufs[l] for l in self.stitcher.img_labels}, out[s - lo:e - lo])
        return out


def stitch_capture(stitcher, seq, capture=0, lo=0, hi=None, device=None, chunk=16, depth=3, batch=32,
                   rank=0, world_size=1, workers=None):
    """Composite frame-sets ``[lo, hi)`` of a capture (this rank's share of them) and return
    ``(first_frame, panoramas)`` with ``panoramas`` a host uint8 tensor ``[n, H_out, W_out, C]``.
    One-shot form of :class:`CaptureRunner` (which keeps its pipeline and pinned buffers between calls)."""
    from .sequence import shard_range
    hi = seq.n_frames(capture) if hi is None else hi
    a, b = shard_range(hi - lo, world_size, rank)
    a, b = lo + a, lo + b
    if b <= a:
        return a, None
    runner = CaptureRunner(stitcher, seq, capture, device, chunk, depth, batch, workers)
    return a, runner.run(a, b)
