"""Probe: does the KIND of pinned host memory matter for the strided window uploads of the sequence pipeline when
the other direction is busy?  Plain pinned memory (cudaHostAlloc default, what torch's pin_memory gives) against
write-combined pinned memory (cudaHostAllocWriteCombined: the device reads it without cache snoops; the CPU only
ever writes it, as a decoder would).  Strided 2-D windows (3 456 of 5 760 bytes per row) and whole frames up,
a contiguous panorama buffer down, both directions at once.  One line per case."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from multicamera_stitching_b200 import _cabi  # noqa: E402


def host_alloc(nbytes, flags):
    rt = ctypes.CDLL("libcudart.so.12")
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(flags))
    if rc != 0:
        raise RuntimeError("cudaHostAlloc failed: %d" % rc)
    buf = (ctypes.c_uint8 * nbytes).from_address(p.value)
    return torch.frombuffer(buf, dtype=torch.uint8), p


def main():
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    _cabi.load()
    F, H, ROW = 16, 1080, 5760
    n_cam = 8
    fs = H * ROW
    win_b0, win_n = 2304, 3456           # 60 % of every row
    d_src = torch.empty((n_cam, F, H, ROW), dtype=torch.uint8, device=dev)
    d_out = torch.empty(F * 33480000 // 1, dtype=torch.uint8, device=dev)
    h_out = torch.empty(d_out.numel(), dtype=torch.uint8, pin_memory=True)
    s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()
    n_up = int(os.environ.get("PROBE_UP_STREAMS", "1"))     # upload streams the cameras are dealt over
    ups = [s_up] + [torch.cuda.Stream() for _ in range(n_up - 1)]
    for kind, flags in (("pinned", 0), ("write-combined", 4)):
        h_src, keep = host_alloc(n_cam * F * fs, flags)
        h_src.fill_(7) if flags == 0 else h_src[::4096].fill_(7)
        for mode in ("windows", "whole"):
            for duplex in (False, True):
                def up():
                    for c in range(n_cam):
                        base = h_src.data_ptr() + c * F * fs
                        st = ups[c % n_up].cuda_stream
                        if mode == "whole":
                            _cabi.copy_window_u8(d_src[c].data_ptr(), ROW, fs, base, ROW, fs, 0, ROW, 0, H, F, st)
                        else:
                            for y0 in range(0, H, 64):
                                _cabi.copy_window_u8(d_src[c].data_ptr(), ROW, fs, base, ROW, fs, win_b0, win_n, y0,
                                                     min(64, H - y0), F, st)
                reps = 6
                e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
                torch.cuda.synchronize()
                for u in ups[1:]:
                    u.wait_stream(s_up)
                e[0].record(s_up)
                e[2].record(s_dn)
                for _ in range(reps):
                    up()
                    if duplex:
                        with torch.cuda.stream(s_dn):
                            h_out.copy_(d_out, non_blocking=True)
                for u in ups[1:]:
                    s_up.wait_stream(u)
                e[1].record(s_up)
                e[3].record(s_dn)
                torch.cuda.synchronize()
                up_bytes = reps * n_cam * F * H * (ROW if mode == "whole" else win_n)
                line = {"host_memory": kind, "upload": mode, "duplex": duplex, "upload_streams": n_up,
                        "h2d_gbs": up_bytes / e[0].elapsed_time(e[1]) / 1e6}
                if duplex:
                    line["d2h_gbs"] = reps * d_out.numel() / e[2].elapsed_time(e[3]) / 1e6
                print(json.dumps(line), flush=True)
        del h_src
        ctypes.CDLL("libcudart.so.12").cudaFreeHost(keep)


if __name__ == "__main__":
    main()
