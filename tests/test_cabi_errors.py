"""Error behaviour of the C ABI (include/mcs.h): every entry point returns a negative MCS_ERR_*
code with a message in ``mcs_last_error()`` and never throws or aborts - the library-level
counterpart of the reference's log-and-carry-on convention (StitcherClass.py:124-128, :255-256).

The argument checks that run before any CUDA call are exercised on the CPU; the rest needs a device."""
import ctypes

import numpy as np
import pytest

from multicamera_stitching_b200 import _cabi

MCS_ERR_INVALID, MCS_ERR_CUDA, MCS_ERR_UNSUPPORTED = -1, -2, -3


def _last():
    return _cabi.load().mcs_last_error().decode()


def _plan_args(n=1, kind=(1,), hw=((60, 80),), origin=((0, 0),), rect=((0, 0, 80, 60),)):
    i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)  # noqa: E731
    k, s, o, r = i32(kind), i32(hw), i32(origin), i32(rect)
    h = np.ascontiguousarray(np.tile(np.eye(3).ravel(), (max(n, 1), 1)))
    p = lambda a: a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))  # noqa: E731
    return (k, s, o, r, h), (p(k), p(s), h.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), p(o), p(r))


def test_plan_create_rejects_bad_counts_before_touching_the_device():
    lib = _cabi.load()
    keep, (k, s, h, o, r) = _plan_args()
    handle = ctypes.c_void_p()
    for n_layers, channels, needle in ((0, 3, "n_layers"), (17, 3, "n_layers"), (1, 2, "channels"), (1, 5, "channels")):
        rc = lib.mcs_plan_create(ctypes.byref(handle), n_layers, channels, k, s, h, o, r, 80, 60)
        assert rc == MCS_ERR_INVALID and handle.value is None and needle in _last()
    assert lib.mcs_plan_create(None, 1, 3, k, s, h, o, r, 80, 60) == MCS_ERR_INVALID
    assert lib.mcs_plan_create(ctypes.byref(handle), 1, 3, k, s, h, o, r, 1 << 24, 60) == MCS_ERR_INVALID
    assert lib.mcs_plan_destroy(None) == 0          # destroying nothing is fine


def test_resize_and_copy_reject_bad_arguments_before_touching_the_device():
    lib = _cabi.load()
    assert lib.mcs_resize_linear_u8(None, 10, 10, 30, 0, None, 5, 5, 15, 0, 2, 1, None) == MCS_ERR_INVALID
    assert "channels" in _last()
    assert lib.mcs_resize_linear_u8(None, 0, 10, 30, 0, None, 5, 5, 15, 0, 3, 1, None) == MCS_ERR_INVALID
    assert lib.mcs_resize_linear_u8(None, 10, 10, 30, 0, None, 5, 5, 15, 0, 3, 0, None) == 0      # no frames: nothing to do
    assert lib.mcs_resize_linear_u8(None, 10, 10, 30, 0, None, 5, 5, 15, 0, 3, 1, None) == MCS_ERR_INVALID
    assert "NULL" in _last()
    assert lib.mcs_copy_window_u8(None, 64, 0, None, 64, 0, 0, 16, 0, 4, 1, None) == MCS_ERR_INVALID
    assert lib.mcs_upload_pageable_u8(0, None, None, None, None, None, None, 0, 4, None) == 0     # no windows: nothing to do
    assert lib.mcs_upload_pageable_u8(1, None, None, None, None, None, None, 0, 4, None) == MCS_ERR_INVALID
    assert "NULL" in _last()
    assert lib.mcs_upload_pageable_u8(-1, None, None, None, None, None, None, 0, 4, None) == MCS_ERR_INVALID
    assert lib.mcs_plan_source_windows(None, None) == MCS_ERR_INVALID
    assert lib.mcs_plan_set_feather(None, 1) == MCS_ERR_INVALID
    assert lib.mcs_plan_force_variant(None, 1) == MCS_ERR_INVALID
    assert lib.mcs_plan_owned_pixels(None, None, None) == MCS_ERR_INVALID
    assert lib.mcs_stitch_u8(None, None, None, None, 1, None, 0, 0, None) == MCS_ERR_INVALID


@pytest.mark.gpu
def test_plan_and_stitch_errors_on_the_device(cuda_device):
    import torch
    lib = _cabi.load()
    handle = ctypes.c_void_p()
    with torch.cuda.device(cuda_device):
        # a COPY layer must not read outside its source
        keep, (k, s, h, o, r) = _plan_args(kind=(0,), rect=((0, 0, 81, 60),))
        assert lib.mcs_plan_create(ctypes.byref(handle), 1, 3, k, s, h, o, r, 100, 60) == MCS_ERR_INVALID
        assert "layer 0 invalid" in _last()
        # a REMAP layer needs its map
        keep, (k, s, h, o, r) = _plan_args(kind=(2,))
        assert lib.mcs_plan_create(ctypes.byref(handle), 1, 3, k, s, h, o, r, 80, 60) == MCS_ERR_INVALID
        assert "REMAP" in _last()
        with pytest.raises(_cabi.McsError, match="exceeds"):     # and the map must cover the rectangle
            _cabi.Plan([2], [(60, 80)], np.eye(3).reshape(1, 9), [(0, 0)], [(0, 0, 80, 60)], 80, 60, 3,
                       maps=[(np.zeros((50, 80, 2), np.int16), None)])
        # a valid plan, then bad launches
        plan = _cabi.Plan([1], [(60, 80)], np.eye(3).reshape(1, 9), [(0, 0)], [(0, 0, 80, 60)], 80, 60, 3)
        src = torch.zeros((60, 80, 3), dtype=torch.uint8, device=cuda_device)
        dst = torch.zeros((60, 80, 3), dtype=torch.uint8, device=cuda_device)
        with pytest.raises(_cabi.McsError, match="dst pitch"):
            plan.stitch([src.data_ptr()], [240], [0], 1, dst.data_ptr(), 100, 0)
        with pytest.raises(_cabi.McsError, match="pitch"):
            plan.stitch([src.data_ptr()], [100], [0], 1, dst.data_ptr(), 240, 0)
        with pytest.raises(_cabi.McsError, match="n_frames"):
            plan.stitch([src.data_ptr()], [240], [0], -1, dst.data_ptr(), 240, 0)
        with pytest.raises(_cabi.McsError, match="NULL"):
            plan.stitch([0], [240], [0], 1, dst.data_ptr(), 240, 0)
        plan.stitch([src.data_ptr()], [240], [0], 0, 0, 240, 0)      # zero frames: nothing to do, no error
        # the tiled variant cannot take a pitch that is not a multiple of 16 bytes: forcing it says so
        src2 = torch.zeros((60, 81, 3), dtype=torch.uint8, device=cuda_device)[:, :80]
        plan.force_variant(2)
        with pytest.raises(_cabi.McsError, match="code -3"):
            plan.stitch([src2.data_ptr()], [243], [0], 1, dst.data_ptr(), 240, 0)
        plan.force_variant(0)
        plan.stitch([src2.data_ptr()], [243], [0], 1, dst.data_ptr(), 240, 0)   # automatic: the gather variant serves it
        assert plan.last_variant() == 1
        with pytest.raises(_cabi.McsError):
            plan.force_variant(7)
        with pytest.raises(_cabi.McsError, match="feather_log2"):
            plan.set_feather(13)
        torch.cuda.synchronize()


def test_set_blend_rejects_a_null_plan_before_touching_the_device():
    lib = _cabi.load()
    assert lib.mcs_plan_set_blend(None, 1, None, None, None) == MCS_ERR_INVALID and "plan is NULL" in _last()


@pytest.mark.gpu
def test_set_blend_errors_leave_the_plan_in_a_usable_mode(cuda_device):
    """mcs_plan_set_blend: bad arguments are refused with a message; a blend the plan cannot take in its fused
    form (weight maps with feather_log2 > 5) is MCS_ERR_UNSUPPORTED and the plan keeps the reference's overwrite."""
    import torch
    from helpers import synthetic_chain
    from oracle import stitcher_ref
    st, states, labels, images = synthetic_chain(3, 120, 200, 3, kind="noise")
    plan = st.plan([images[l].shape for l in labels], cuda_device)
    h = plan.handle
    n = h.n_layers
    lib = _cabi.load()
    assert lib.mcs_plan_set_blend(h._h, 13, None, None, None) == MCS_ERR_INVALID and "feather_log2" in _last()
    # weight maps without pitches
    m = np.full((120, 200), 4, np.uint8)
    ptrs = (ctypes.c_void_p * n)()
    ptrs[1] = m.ctypes.data
    assert lib.mcs_plan_set_blend(h._h, 3, None, ptrs, None) == MCS_ERR_INVALID and "pitches" in _last()
    # maps beyond the six weight bits of the overlay descriptors
    with pytest.raises(_cabi.McsError):
        h.set_blend(6, None, [None, m] + [None] * (n - 2))
    # a paste rectangle that does not contain the visible one
    paste = np.array([l.rect for l in plan.flat.layers], np.int32)
    paste[0, 2] -= 10
    with pytest.raises(_cabi.McsError):
        h.set_blend(2, paste, None)
    # after every refusal: the overwrite, bit for bit
    h.set_blend(0)
    dev = {l: torch.from_numpy(images[l][None]).to(cuda_device) for l in labels}
    assert np.array_equal(st.stitch_batch(dev)[0].cpu().numpy(), stitcher_ref.stitch_chain(states, labels, images))
    assert h.last_variant() == 2
