import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, ctypes as C
from pl_vi_orbslam3_b200 import Lineextractor, synth, capi
lib = capi.lib()
for rounds in (1, 12):
    os.environ["PLVI_LSD_BR_ROUNDS"] = str(rounds)
    f = synth.frame_euroc(11)
    le = Lineextractor(200, 0, 0.8, 2, 2.0, 0, max_batch=2)
    try:
        kl, ld, eq, lc = le.extract_batch(f[None])
    except Exception as e:
        print("extract:", e)
    for o in (0, 1):
        stt = np.zeros(8 * 256, np.int32); cnt = C.c_int(0)
        rc = lib.plvi_line_read_lsd(le._h, 0, o, 8, stt.ctypes.data_as(C.c_void_p), 8 * 256, C.byref(cnt))
        s = stt[: 8 * cnt.value].reshape(-1, 8)
        print("rounds", rounds, "oct", o, "bands", cnt.value)
        print(" pixels ", s[:, 5].tolist())
        print(" records", [int(r[r[2]]) for r in s])
        print(" runs   ", s[:, 6].tolist())
        print(" kcycles", (s[:, 7] // 1000).tolist())
        print(" cyc/px ", (s[:, 7] // np.maximum(s[:, 5], 1)).tolist())
    le.close()
