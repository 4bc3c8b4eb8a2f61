"""GPU: the host-facing sequence path (``sequence.SequencePipeline``, what ``bench.py`` times as
``e2e``) with window uploads: only the part of every camera that can reach the panorama is
copied to the device (``mcs_plan_source_windows`` + ``mcs_copy_window_u8``).  The device slots are
poisoned first, so a byte the kernel reads but the window does not cover would show up."""
import numpy as np
import pytest
import torch

from helpers import compare_u8, synthetic_chain
from multicamera_stitching_b200.sequence import SequencePipeline, pinned_like
from oracle import stitcher_ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,h,w,c", [(6, 270, 480, 3), (3, 180, 320, 1), (4, 120, 200, 4)])
def test_pipeline_with_window_uploads_equals_the_cv2_chain(cuda_device, n, h, w, c):
    st, states, labels, images = synthetic_chain(n, h, w, c, kind="noise")
    shapes = [images[l].shape for l in labels]
    F = 7
    sets = [synthetic_chain(n, h, w, c, kind="noise", frame_index=f)[3] for f in range(F)]
    host = {l: pinned_like((F,) + tuple(images[l].shape)) for l in labels}
    for l in labels:
        for f in range(F):
            host[l][f].copy_(torch.from_numpy(sets[f][l]))
    pipe = SequencePipeline(st, shapes, cuda_device, chunk=3, depth=2)
    h2d, d2h = pipe.bytes_per_frame()
    full = sum(int(np.prod(s)) for s in shapes)
    assert 0 < h2d < full                       # overlapping cameras: some part is always hidden
    wins = pipe.plan.handle.source_windows()
    assert len(wins) == n and wins[0] == (0, 0, w, h)   # camera 0 is pasted whole
    for k in range(n):                                   # the row-band spans refine the window
        spans = [s for s in pipe.plan.handle.source_spans(k, 16) if s[1] > s[0]]
        assert min(s[0] for s in spans) == wins[k][0] and max(s[1] for s in spans) == wins[k][2]
    for slot in pipe.slots:                     # poison what the windows do not cover
        for t in slot["src"]:
            t.fill_(0xAB)
    out = pinned_like((F,) + pipe.plan.out_shape())
    assert pipe.run(host, out) == F
    for f in range(F):
        assert compare_u8(out[f].numpy(), stitcher_ref.stitch_chain(states, labels, sets[f])) == (0, 1.0)
    # whole-frame uploads give the same panoramas
    whole = SequencePipeline(st, shapes, cuda_device, chunk=3, depth=2, windows=False)
    assert whole.bytes_per_frame()[0] == full
    out2 = pinned_like((F,) + pipe.plan.out_shape())
    whole.run(host, out2)
    assert torch.equal(out, out2)


def test_run_returns_with_the_panoramas_on_the_host(cuda_device):
    """``run()`` blocks until the last download has landed: the host buffer is read straight after the
    call, with no synchronize in between, on a batch long enough for the copies to still be in flight if
    the call only ordered streams; ``sync=False`` is the asynchronous form."""
    st, states, labels, images = synthetic_chain(6, 540, 960, 3, kind="noise")
    shapes = [images[l].shape for l in labels]
    F = 24
    host = {l: pinned_like((F,) + tuple(images[l].shape)) for l in labels}
    for l in labels:
        for f in range(F):
            host[l][f].copy_(torch.from_numpy(images[l]))
    pipe = SequencePipeline(st, shapes, cuda_device, chunk=4, depth=3)
    ref = stitcher_ref.stitch_chain(states, labels, images)
    out = pinned_like((F,) + pipe.plan.out_shape())
    for _ in range(3):
        out.fill_(0x5A)
        assert pipe.run(host, out) == F
        last = out[F - 1].numpy().copy()          # no torch.cuda.synchronize() before this read
        assert np.array_equal(last, ref)
    out.fill_(0x5A)
    pipe.run(host, out, sync=False)
    torch.cuda.current_stream().synchronize()     # the caller's stream was ordered behind the pipeline
    assert np.array_equal(out[F - 1].numpy(), ref) and np.array_equal(out[0].numpy(), ref)


def test_feather_mode_keeps_the_window_uploads_when_fused(cuda_device, monkeypatch):
    """Fused feather form (BAND tiles): the plan knows what the band samples read, so the pipeline still
    uploads windows only and the panoramas match the specification; the two-pass form needs whole frames."""
    import torch
    from oracle import feather_model
    st, states, labels, images = synthetic_chain(3, 120, 200, 3, kind="noise")
    st.feather_log2 = 2
    shapes = [images[l].shape for l in labels]
    whole = sum(int(np.prod(s)) for s in shapes)
    pipe = SequencePipeline(st, shapes, cuda_device, chunk=2, depth=2)
    assert 0 < pipe.bytes_per_frame()[0] < whole
    F = 3
    sets = [synthetic_chain(3, 120, 200, 3, kind="noise", frame_index=f)[3] for f in range(F)]
    host = {l: torch.from_numpy(np.stack([s[l] for s in sets])).pin_memory() for l in labels}
    out = torch.empty((F,) + pipe.plan.out_shape(), dtype=torch.uint8).pin_memory()
    for slot in pipe.slots:                      # stale bytes outside the windows must never show
        for t in slot["src"]:
            t.fill_(0xA5)
    pipe.run(host, out)
    for f in range(F):
        assert np.array_equal(out[f].numpy(), feather_model.feather_chain(states, labels, sets[f], 2))
    monkeypatch.setenv("MCS_TILED_BAND", "0")
    st2, _, _, _ = synthetic_chain(3, 120, 200, 3, kind="noise")
    st2.feather_log2 = 2
    assert SequencePipeline(st2, shapes, cuda_device, chunk=2, depth=2).bytes_per_frame()[0] == whole


def test_single_call_numpy_path_uploads_only_the_visible_windows(cuda_device):
    """``Stitcher.stitch(images_dic)`` with numpy frames sends the same windows; the reused device
    staging tensors keep stale bytes outside them, which must never reach a panorama."""
    st, states, labels, images = synthetic_chain(5, 200, 360, 3, kind="noise")
    assert compare_u8(st.stitch(images), stitcher_ref.stitch_chain(states, labels, images)) == (0, 1.0)
    eng = st._engine_()
    plan = st.plan([images[l].shape for l in labels], cuda_device)
    bands = plan.upload_bands()
    sent = sum(w["nbytes"] * w["rows"] for _, _, copies in bands.values() for w in copies)
    assert 0 < sent < sum(images[l].size for l in labels)
    for t in eng._staging.values():
        t.fill_(0x5A)
    for f in (1, 2):
        im = synthetic_chain(5, 200, 360, 3, kind="noise", frame_index=f)[3]
        assert compare_u8(st.stitch(im), stitcher_ref.stitch_chain(states, labels, im)) == (0, 1.0)


def test_ring_mode_cycles_the_host_buffers(cuda_device):
    """``run(..., ring=True)``: frame f of a long sequence lives in host slot f % R (config 5)."""
    st, states, labels, images = synthetic_chain(3, 120, 200, 3, kind="noise")
    shapes = [images[l].shape for l in labels]
    R = 5
    sets = [synthetic_chain(3, 120, 200, 3, kind="noise", frame_index=f)[3] for f in range(R)]
    host = {l: pinned_like((R,) + tuple(images[l].shape)) for l in labels}
    for l in labels:
        for f in range(R):
            host[l][f].copy_(torch.from_numpy(sets[f][l]))
    pipe = SequencePipeline(st, shapes, cuda_device, chunk=2, depth=2)
    out = pinned_like((R,) + pipe.plan.out_shape())
    out.zero_()
    assert pipe.run(host, out, 3, 3 + 2 * R + 1, ring=True) == 2 * R + 1     # frames 3 .. 13: every slot written
    for r in range(R):
        assert compare_u8(out[r].numpy(), stitcher_ref.stitch_chain(states, labels, sets[r])) == (0, 1.0)
