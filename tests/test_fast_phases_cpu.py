"""CPU check of the argument behind k_fast's two threshold phases (csrc/orb_kernels.cu; the reference runs cv::FAST per
cell at iniThFAST and, if the cell stays empty, at minThFAST: src/ORBextractor.cc:809-817).

The FAST score does not depend on the detection threshold, and a corner's 3x3 non-maximum suppression only loses to
scores >= its own.  Hence cv::FAST(th) + NMS equals "pixels whose score is >= th and larger than the scores of their eight
neighbours", whether the neighbours' scores below th are counted or not -- which is what lets the kernel (a) detect at 20
first and keep those survivors as they are, (b) run the cells that stayed empty again at 7 with the scores of the first
phase still in the tile."""
import numpy as np

import oracle
from pl_vi_orbslam3_b200 import synth


def test_fast_nms_from_threshold_independent_scores():
    img = synth.frame_euroc(9)
    rng = np.random.RandomState(1)
    noise = rng.randint(0, 256, (480, 752)).astype(np.uint8)
    for trial in range(16):
        src = img if trial % 4 else noise
        x0, y0 = rng.randint(0, 752 - 40), rng.randint(0, 480 - 40)
        roi = np.ascontiguousarray(src[y0:y0 + 36, x0:x0 + 37])      # a FAST cell with its 3-pixel border
        sm = oracle.fast_score_map(roi)
        h, w = sm.shape
        for th in (20, 7):
            ref = {(int(r[0]), int(r[1]), int(r[2])) for r in oracle.fast_roi(roi, th)}
            for scores in (np.where(sm >= th, sm, 0), sm):            # neighbours below th left out / counted
                got = set()
                for y in range(3, h - 3):
                    for x in range(3, w - 3):
                        s = sm[y, x]
                        if s < th:
                            continue
                        nb = scores[y - 1:y + 2, x - 1:x + 2].copy()
                        nb[1, 1] = -1
                        if nb.max() < s:
                            got.add((x, y, int(s)))
                assert got == ref, (trial, th)
