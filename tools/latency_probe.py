"""Single-frame (blocking call) latency of the two extractors through the C ABI, host buffers."""
import sys, time
sys.path.insert(0, ".")
import numpy as np
from pl_vi_orbslam3_b200 import ORBextractor, Lineextractor, synth
img = synth.frame_euroc(0)
orb = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=1)
line = Lineextractor(200, 0, 0.8, 2, 2.0, 0, max_batch=1)
for name, fn in (("ORBextractor::operator()", lambda: orb(img)), ("Lineextractor::operator()", lambda: line(img))):
    for _ in range(3):
        fn()
    t = time.perf_counter()
    n = 10
    for _ in range(n):
        fn()
    print(f"{name}: {(time.perf_counter() - t) / n * 1e3:.2f} ms per blocking single-frame call")
