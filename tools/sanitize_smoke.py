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

# entry points outside the batched front-end: stereo matching, in-radius search, line Fuse search
import oracle  # noqa: E402  (inputs only)
from pl_vi_orbslam3_b200 import FrameView, LineMatcher, ORBextractor, ORBmatcher, frame_grid  # noqa: E402
from pl_vi_orbslam3_b200.capi import QUERY_DTYPE  # noqa: E402

w, h = 641, 479
left = synth.frame_euroc(3, w, h)
right = np.roll(left, -9, axis=1)
el = ORBextractor(500, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=1)
er = ORBextractor(500, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=1)
dl, dr = torch.from_numpy(left[None]).cuda(), torch.from_numpy(right[None]).cuda()
lo, ro = el.extract_batch_device(dl), er.extract_batch_device(dr)
ur, dp, ns = el.stereo_matches(er, lo, ro, 0.11, 47.9)
torch.cuda.synchronize()
print("stereo", int(ns[0]))
el.close(); er.close()

r = oracle.orb_extract(left, nfeatures=500)
om = ORBmatcher(0.9, True, max_pairs=1, max_train=2048, max_query=2048)
q = np.zeros(len(r["keypoints"]), QUERY_DTYPE)
q["u"], q["v"], q["radius"] = r["keypoints"]["x"] + 1, r["keypoints"]["y"] - 1, 9.0
q["min_level"], q["max_level"] = r["keypoints"]["octave"] - 1, r["keypoints"]["octave"]
s2 = (1.0 / (1.2 ** np.arange(8)) ** 2).astype(np.float32)
print("radius", om.SearchInRadius(FrameView(r["keypoints"], r["descriptors"], frame_grid(0, w, 0, h)), q, r["descriptors"], s2)[0])
om.close()

ll = oracle.line_extract(left)
lm = LineMatcher(max_pairs=1, max_train=512, max_query=512)
kl = ll["keylines"]
lq = np.stack([kl["startPointX"], kl["startPointY"], kl["endPointX"], kl["endPointY"], np.full(len(kl), 12.0, np.float32),
               kl["octave"].astype(np.float32)], axis=1)
print("linefuse", lm.FuseSearch(kl, ll["descriptors"], lq, ll["descriptors"])[0])
lm.close()
print("done 2")
