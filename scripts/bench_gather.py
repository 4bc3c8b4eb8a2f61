import sys; sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch
from helpers import synthetic_chain
dev = torch.device('cuda', 0)
st, states, labels, images = synthetic_chain(6, 1080, 1920, 3, kind='smooth')
F = 32
batch = {l: torch.from_numpy(np.stack([images[l]] * F)).to(dev) for l in labels}
plan = st.plan([images[l].shape for l in labels], dev)
out = plan.new_output(F, pitch_align=128)
import os
for v, legacy in ((2, ''), (1, ''), (1, 'MCS_GATHER_BYTES'), (1, 'MCS_GATHER_LEGACY')):
    plan.handle.force_variant(v)
    os.environ.pop('MCS_GATHER_LEGACY', None)
    os.environ.pop('MCS_GATHER_BYTES', None)
    if legacy:
        os.environ[legacy] = '1'   # read by the library at every call
    st.stitch_batch(batch, out=out); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): st.stitch_batch(batch, out=out)
    e0.record()
    for _ in range(10): st.stitch_batch(batch, out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    ab = plan.algorithmic_bytes()
    print('variant=%d mode=%s ms=%.3f frac=%.3f' % (plan.handle.last_variant(), legacy, ms, ab * F / ms / 1e6 / 6533.5))
