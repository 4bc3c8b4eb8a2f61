cat > /tmp/san.py <<'PY'
import sys
sys.path.insert(0, 'tests')
import numpy as np, torch
from helpers import synthetic_chain
from oracle import feather_model, stitcher_ref
dev = torch.device('cuda', 0)
# overwrite, feather ramp (fused), weight maps, super mode - small geometry, a few frames
st, states, labels, images = synthetic_chain(3, 120, 200, 3, kind="noise", xoffset=2, yoffset=7)
assert np.array_equal(st.stitch(images), stitcher_ref.stitch_chain(states, labels, images))
st.feather_log2 = 3
assert np.array_equal(st.stitch(images), feather_model.feather_chain(states, labels, images, 3))
sets = [synthetic_chain(3, 120, 200, 3, kind="noise", frame_index=f)[3] for f in range(5)]
batch = {l: torch.from_numpy(np.stack([s[l] for s in sets])).to(dev) for l in labels}
out = st.stitch_batch(batch).cpu().numpy()
for f in range(5):
    assert np.array_equal(out[f], feather_model.feather_chain(states, labels, sets[f], 3))
print("sanitizer workload ok", st.plan([images[l].shape for l in labels], dev).handle.tiled_stats())
PY
timeout 500 compute-sanitizer --tool memcheck --error-exitcode 7 python /tmp/san.py 2>&1 | tail -6
echo "memcheck rc=$?"
timeout 500 compute-sanitizer --tool racecheck --error-exitcode 7 python /tmp/san.py 2>&1 | tail -6
echo "racecheck rc=$?"
