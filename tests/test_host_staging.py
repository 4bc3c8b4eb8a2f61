"""Host side of the reference's own call shape, `Stitcher.stitch(images_dic)` with numpy frames
(StitcherClass.py:114-136): the pinned result pool (a result is an array of its own, like the array cv2
allocates at :239, and its buffer is recycled only once the array and its views are gone) and the threaded
staging of pageable frames."""
import gc

import numpy as np
import pytest
import torch

from multicamera_stitching_b200.engine import PinnedResults


def _pageable_pool(budget):
    pool = PinnedResults(budget=budget)
    pool._alloc = lambda shape: torch.empty(shape, dtype=torch.uint8)   # no pinned memory without a device
    return pool


def test_result_pool_recycles_only_dead_arrays():
    pool = _pageable_pool(1 << 20)
    a = pool.as_array(pool.take((4, 5, 3)))
    b = pool.as_array(pool.take((4, 5, 3)))
    assert a.ctypes.data != b.ctypes.data
    a[:] = 7
    b[:] = 9
    assert int(a.max()) == 7 and int(a.min()) == 7
    pa = a.ctypes.data
    view = a[1:3]
    del a
    gc.collect()
    c = pool.take((4, 5, 3))            # the view still holds the first buffer
    assert c.data_ptr() != pa and c.data_ptr() != b.ctypes.data
    del view
    gc.collect()
    d = pool.take((4, 5, 3))
    assert d.data_ptr() == pa           # now it is reused
    assert int(b.min()) == 9


def test_result_pool_budget():
    pool = _pageable_pool(100)
    a = pool.take((60,))
    assert a is not None
    assert pool.take((60,)) is None     # over budget: the caller falls back to a pageable array
    pool.give_back(a)
    assert pool.take((60,)) is a
    pool.give_back(a)
    b = pool.take((8, 10))              # another shape: the free buffer of the old shape is dropped first
    assert b is not None and tuple(b.shape) == (8, 10)
    assert pool.take((0, 3)) is None


@pytest.mark.gpu
def test_numpy_results_do_not_alias(cuda_device):
    from helpers import synthetic_chain
    from oracle import stitcher_ref
    st, states, labels, images = synthetic_chain(3, 180, 320, 3, kind="noise")
    _, _, _, images2 = synthetic_chain(3, 180, 320, 3, kind="noise", frame_index=1)
    ref1 = stitcher_ref.stitch_chain(states, labels, images)
    ref2 = stitcher_ref.stitch_chain(states, labels, images2)
    got1 = st.stitch(images)
    got2 = st.stitch(images2)
    assert got1.ctypes.data != got2.ctypes.data
    assert np.array_equal(got1, ref1) and np.array_equal(got2, ref2)
    keep = [st.stitch(images) for _ in range(6)]            # results kept alive stay intact
    got3 = st.stitch(images2)
    assert all(np.array_equal(k, ref1) for k in keep) and np.array_equal(got3, ref2)
    p1 = got1.ctypes.data
    del got1
    gc.collect()
    again = st.stitch(images2)
    assert again.ctypes.data == p1 and np.array_equal(again, ref2)
    assert again.flags.writeable
    again[:] = 0                                             # the caller owns its result
    assert np.array_equal(got2, ref2)


@pytest.mark.gpu
def test_pinned_budget_spent_falls_back_to_pageable(cuda_device):
    from helpers import synthetic_chain
    from oracle import stitcher_ref
    st, states, labels, images = synthetic_chain(3, 180, 320, 3, kind="noise")
    ref = stitcher_ref.stitch_chain(states, labels, images)
    st._engine_().results.budget = ref.nbytes + 1
    keep = [st.stitch(images) for _ in range(3)]
    assert all(np.array_equal(k, ref) for k in keep)
    assert len({k.ctypes.data for k in keep}) == 3


@pytest.mark.gpu
@pytest.mark.parametrize("threads", ["1", "3"])
def test_staged_uploads(cuda_device, monkeypatch, threads):
    """Pageable frames staged through pinned memory: contiguous frames (visible windows only), strided
    views, and frames that change between calls while the staging buffers are reused."""
    from helpers import synthetic_chain
    from oracle import stitcher_ref
    monkeypatch.setenv("MCS_UPLOAD_THREADS", threads)
    st, states, labels, images = synthetic_chain(4, 240, 416, 3, kind="noise")
    for frame_index in (0, 1, 2):
        _, _, _, imgs = synthetic_chain(4, 240, 416, 3, kind="noise", frame_index=frame_index)
        ref = stitcher_ref.stitch_chain(states, labels, imgs)
        assert np.array_equal(st.stitch(imgs), ref)
        wide = {l: np.concatenate([imgs[l], imgs[l]], axis=1) for l in labels}
        views = {l: wide[l][:, :416] for l in labels}      # row-strided views of wider arrays
        assert not views[labels[0]].flags.c_contiguous
        assert np.array_equal(st.stitch(views), ref)
        mixed = dict(imgs)
        mixed[labels[1]] = torch.from_numpy(imgs[labels[1]])   # a CPU tensor among numpy frames
        assert np.array_equal(np.asarray(st.stitch(mixed)), ref)
        # views whose rows are not dense (channels reversed) or run backwards (flipped): staged through a copy
        swapped = {l: imgs[l][:, :, ::-1] for l in labels}
        flipped = {l: imgs[l][::-1] for l in labels}
        for views in (swapped, flipped):
            want = stitcher_ref.stitch_chain(states, labels, {l: np.ascontiguousarray(v) for l, v in views.items()})
            assert np.array_equal(st.stitch(views), want)


def test_staging_threads_copy_the_windows_without_a_device():
    """The host half of mcs_upload_pageable_u8 - the persistent thread pool that copies windows of pageable
    frames into the staging frames piece by piece - runs here, where the DMA it issues afterwards can only fail
    (no device): the call reports the CUDA error, the staging frames hold exactly the windows.  Many calls
    back to back with changing thread counts and piece sizes exercise the hand-over between jobs."""
    if torch.cuda.is_available():
        pytest.skip("the fake device pointers of this test are only safe where every DMA fails")
    from multicamera_stitching_b200 import _cabi
    rng = np.random.default_rng(7)
    h, row = 97, 403
    wide = rng.integers(0, 256, (h, row + 61), dtype=np.uint8)
    frames = [rng.integers(0, 256, (h, row), dtype=np.uint8), wide[:, :row]]   # dense rows, and a row-strided view
    for it in range(300):
        threads = 1 + it % 7
        piece = [64, 1000, 4096, 1 << 20][it % 4]
        stagings = [np.full((h, row), 255, dtype=np.uint8) for _ in frames]
        expect = [s.copy() for s in stagings]
        windows = []
        for f, s, e in zip(frames, stagings, expect):
            for _ in range(1 + it % 3):
                x0, y0 = int(rng.integers(0, row - 1)), int(rng.integers(0, h - 1))
                w, r = int(rng.integers(1, row - x0 + 1)), int(rng.integers(1, h - y0 + 1))
                windows.append((0x1000, f.ctypes.data, s.ctypes.data, row, f.strides[0], x0, y0, w, r))
                e[y0:y0 + r, x0:x0 + w] = f[y0:y0 + r, x0:x0 + w]
        with pytest.raises(_cabi.McsError, match="cudaMemcpy2DAsync"):
            _cabi.upload_pageable_u8(windows, piece, threads)
        for s, e in zip(stagings, expect):
            assert np.array_equal(s, e), it


def test_upload_pageable_rejects_bad_windows():
    from multicamera_stitching_b200 import _cabi
    with pytest.raises(_cabi.McsError, match="wider than a row"):
        _cabi.upload_pageable_u8([(0x1000, 0x1000, 0x1000, 64, 64, 60, 0, 16, 4)], 0, 2)
    with pytest.raises(_cabi.McsError, match="negative extent"):
        _cabi.upload_pageable_u8([(0x1000, 0x1000, 0x1000, 64, 64, 0, -1, 16, 4)], 0, 2)
    with pytest.raises(_cabi.McsError, match="NULL buffer"):
        _cabi.upload_pageable_u8([(0, 0x1000, 0x1000, 64, 64, 0, 0, 16, 4)], 0, 2)
    with pytest.raises(_cabi.McsError, match="staging thread"):
        _cabi.upload_pageable_u8([(0x1000, 0x1000, 0x1000, 64, 64, 0, 0, 16, 4)], 0, 0)
    _cabi.upload_pageable_u8([(0x1000, 0x1000, 0x1000, 64, 64, 0, 0, 0, 4)], 0, 2)    # empty window: nothing to do
