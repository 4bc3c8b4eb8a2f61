"""Strided window upload: copy engine (cudaMemcpy3DAsync) against SM reads of mapped pinned memory."""
import ctypes, sys, time, torch
sys.path.insert(0, "/root/repo")
from multicamera_stitching_b200 import _cabi
lib = _cabi.load()
lib.mcs_upload_window_u8.restype = ctypes.c_int
vp = ctypes.c_void_p
dev = torch.device("cuda", 0)
F, H, ROW = 16, 1080, 5760
host = torch.empty((F, H, ROW), dtype=torch.uint8, pin_memory=True); host.fill_(5)
dst = torch.zeros((F, H, ROW), dtype=torch.uint8, device=dev)
out_h = torch.empty((F, 1132, 22332), dtype=torch.uint8, pin_memory=True)
out_d = torch.empty((F, 1132, 22332), dtype=torch.uint8, device=dev)
b0, nb = 2304, 3456     # the 60 % window of cameras 1 / 3 / 5
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def bench(fn, label, with_d2h):
    for _ in range(2): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(10):
        with torch.cuda.stream(s1): fn()
        if with_d2h:
            with torch.cuda.stream(s2): out_h.copy_(out_d, non_blocking=True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print("%-28s %s  H2D %.1f GB/s%s" % (label, "with D2H" if with_d2h else "alone   ", 10 * F * H * nb / dt / 1e9,
                                        ("  D2H %.1f GB/s" % (10 * out_h.numel() / dt / 1e9)) if with_d2h else ""))
def ce():
    _cabi.copy_window_u8(dst.data_ptr(), ROW, H * ROW, host.data_ptr(), ROW, H * ROW, b0, nb, 0, H, F, torch.cuda.current_stream().cuda_stream)
def sm(ctas):
    def f():
        rc = lib.mcs_upload_window_u8(vp(dst.data_ptr()), ctypes.c_int64(ROW), ctypes.c_int64(H * ROW), vp(host.data_ptr()), ctypes.c_int64(ROW),
                                      ctypes.c_int64(H * ROW), ctypes.c_int64(b0), ctypes.c_int64(nb), 0, H, F, ctas, vp(torch.cuda.current_stream().cuda_stream))
        assert rc == 0, lib.mcs_last_error()
    return f
for d2h in (False, True):
    bench(ce, "copy engine (3D memcpy)", d2h)
    for ctas in (16, 64, 148, 592):
        bench(sm(ctas), "SM kernel, %d CTAs" % ctas, d2h)
assert torch.equal(dst[:, :, b0:b0 + nb].cpu(), host[:, :, b0:b0 + nb])
