timeout 300 python -m pytest tests/test_feather.py -m gpu -x -q -k "large_source" 2>&1 | tail -12
python - <<'PY'
import sys
sys.path.insert(0,'tests')
import numpy as np, torch
from multicamera_stitching_b200 import Stitcher, synthetic
for scale in (0.62, 0.45):
    n,h,w,c=3,360,640,3
    images = synthetic.make_frames(n,h,w,c,0,"noise")
    st = Stitcher(images); labels=list(st.img_labels); shapeB=images[labels[0]].shape
    for k in range(n-1):
        H=np.array([[scale,0.02,shapeB[1]-0.4*w*scale],[-0.012,scale,9.0*(1 if k%2==0 else -1)],[1e-5,-0.5e-5,1.0]])
        st.stitchers[k].set_homography(H, shapeA=images[labels[k+1]].shape, shapeB=shapeB, xoffset=0, yoffset=0); shapeB=st.stitchers[k].result_shape()
    st.feather_log2=3
    st.stitch(images)
    plan=st.plan([images[l].shape for l in labels], torch.device('cuda',0))
    print(scale, plan.handle.tiled_stats(), plan.handle.last_variant(), plan.handle.tiled_ctas_per_sm(), repr(plan.handle.tiled_status()))
PY
