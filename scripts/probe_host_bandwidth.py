import ctypes, time, torch, numpy as np
cudart = ctypes.CDLL("libcudart.so.12")
def host_alloc(nbytes, flags):
    p = ctypes.c_void_p()
    rc = cudart.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(nbytes), ctypes.c_uint(flags))
    assert rc == 0, rc
    buf = (ctypes.c_uint8 * nbytes).from_address(p.value)
    return torch.frombuffer(buf, dtype=torch.uint8), p
N = 1 << 30
dev = torch.device("cuda", 0)
d_in = torch.empty(N, dtype=torch.uint8, device=dev)
d_out = torch.empty(N, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def bw(h_in, h_out, label):
    for both in (False, True):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(5):
            with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
            if both:
                with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(label, "both" if both else "h2d only", "H2D %.1f GB/s" % (5 * N / dt / 1e9), ("D2H %.1f GB/s" % (5 * N / dt / 1e9)) if both else "")
pin_in = torch.empty(N, dtype=torch.uint8, pin_memory=True); pin_in.fill_(3)
pin_out = torch.empty(N, dtype=torch.uint8, pin_memory=True)
bw(pin_in, pin_out, "pinned     ")
wc_in, _p = host_alloc(N, 0x04)   # cudaHostAllocWriteCombined
wc_in.fill_(3)
bw(wc_in, pin_out, "write-comb.")
pm_in, _p2 = host_alloc(N, 0x01)  # portable
pm_in.fill_(3)
bw(pm_in, pin_out, "portable   ")
