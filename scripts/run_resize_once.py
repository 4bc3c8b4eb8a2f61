"""One launch of each resize case of scripts/bench_prewarp.py (for ncu captures)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from multicamera_stitching_b200.engine import CompositeEngine
dev = torch.device("cuda", 0)
F = int(sys.argv[1]) if len(sys.argv) > 1 else 16
batch = torch.from_numpy(np.random.default_rng(0).integers(0, 256, size=(F, 1080, 1920, 3), dtype=np.uint8)).to(dev)
eng = CompositeEngine()
for hw in ((720, 1280), (2160, 3840), (540, 960)):
    out = eng.resize(batch, hw, batched=True)
    torch.cuda.synchronize()
    print(hw, tuple(out.shape))
