import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from helpers import synthetic_chain, compare_u8
from oracle import stitcher_ref
dev = torch.device("cuda", 0)
n, h, w, c = [int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (3, 144, 256, 3))]
st, states, labels, images = synthetic_chain(n, h, w, c, kind="noise")
plan = st.plan([images[l].shape for l in labels], dev)
torch.cuda.synchronize()
print("plan ok; tiled status: %r" % plan.handle.tiled_status(), flush=True)
srcs = {l: torch.from_numpy(images[l]).to(dev) for l in labels}
ref = stitcher_ref.stitch_chain(states, labels, images)
for variant in (1, 2):
    plan.handle.force_variant(variant)
    out = st.stitch(srcs)
    torch.cuda.synchronize()
    print("variant", variant, compare_u8(out.cpu().numpy(), ref), flush=True)
