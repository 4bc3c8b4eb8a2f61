"""Which part of the sequence pipeline keeps the host <-> device path below what plain copies reach?
Replays the copy pattern of `SequencePipeline` on config-2 geometry without any kernel:
  A  free-running: H2D chunks on one stream, D2H chunks on another, no dependencies
  B  + the slot events of the pipeline (depth 3)
  C  + a stand-in for the kernel between them (an empty stream hop)
"""
import sys, time, torch
F, CH = 32, int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda", 0)
cams = [torch.empty((F, 1080, 1920, 3), dtype=torch.uint8, pin_memory=True) for _ in range(6)]
for c in cams: c.fill_(7)
out = torch.empty((F, 1132, 7444, 3), dtype=torch.uint8, pin_memory=True)
depth = 3
slots = [dict(src=[torch.empty((CH, 1080, 1920, 3), dtype=torch.uint8, device=dev) for _ in range(6)],
              dst=torch.empty((CH, 1132, 7444, 3), dtype=torch.uint8, device=dev),
              ev_in=torch.cuda.Event(), ev_k=torch.cuda.Event(), ev_out=torch.cuda.Event()) for _ in range(depth)]
s_in, s_k, s_out = torch.cuda.Stream(), torch.cuda.Stream(), torch.cuda.Stream()
h2d = sum(c[0].numel() for c in cams); d2h = out[0].numel()

def run(mode, reps=4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    i = 0
    for _ in range(reps):
        for f0 in range(0, F, CH):
            slot = slots[i % depth]; used = i >= depth; i += 1
            with torch.cuda.stream(s_in):
                if mode != "A" and used: s_in.wait_event(slot["ev_k"])
                for c in range(6): slot["src"][c].copy_(cams[c][f0:f0 + CH], non_blocking=True)
                slot["ev_in"].record(s_in)
            if mode == "C":
                with torch.cuda.stream(s_k):
                    s_k.wait_event(slot["ev_in"])
                    if used: s_k.wait_event(slot["ev_out"])
                    slot["ev_k"].record(s_k)
            else:
                slot["ev_k"].record(s_in)
            with torch.cuda.stream(s_out):
                if mode != "A": s_out.wait_event(slot["ev_k"])
                out[f0:f0 + CH].copy_(slot["dst"], non_blocking=True)
                slot["ev_out"].record(s_out)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    n = reps * F
    print("mode %s chunk %d: %.0f frame-sets/s  H2D %.1f GB/s  D2H %.1f GB/s" % (mode, CH, n / dt, n * h2d / dt / 1e9, n * d2h / dt / 1e9))

for m in ("A", "B", "C", "A"):
    run(m)
