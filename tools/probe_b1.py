import sys; sys.path.insert(0, ".")
import numpy as np
from pl_vi_orbslam3_b200 import Lineextractor, synth
img = synth.frame_euroc(11)
line = Lineextractor(200, 0, 0.8, 2, 2.0, 0, max_batch=1)
for _ in range(4):
    line(img)
