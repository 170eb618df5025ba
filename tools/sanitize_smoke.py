"""Small end-to-end run for compute-sanitizer (memcheck): every kernel once, odd sizes included."""
import sys
sys.path.insert(0, ".")
import numpy as np
import torch
from pl_vi_orbslam3_b200 import synth
from pl_vi_orbslam3_b200.frontend import FrontEnd

for (w, h, n) in ((752, 480, 3), (641, 479, 2), (333, 250, 2)):
    frames = np.stack([synth.frame_euroc(i, w, h) for i in range(n)])
    fe = FrontEnd(n, w=w, h=h)
    d = torch.from_numpy(frames).cuda()
    with torch.cuda.stream(fe.stream):
        fe.step(d)
    fe.stream.synchronize()
    o = fe.outputs()
    print(w, h, o["counts"].tolist(), o["line_counts"].tolist(), o["nmatches"].tolist(), o["line_nmatches"].tolist())
    fe.close()
print("done")
