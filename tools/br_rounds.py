"""Band-run diagnostics: rounds each (frame, octave) needs to reach the fixed point, fallbacks, and timing vs the number
of rounds launched.  usage: python tools/br_rounds.py [w h n]"""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import time
import numpy as np, ctypes as C
w, h, n = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (640, 480, 256)
os.environ["PLVI_LSD_BR_MAX"] = "1024"
from pl_vi_orbslam3_b200 import Lineextractor, synth, capi
lib = capi.lib()
frames = synth.seq_batch(n, w, h, base_seed=0)
for rounds in (12, 20, 36):
    os.environ["PLVI_LSD_BR_ROUNDS"] = str(rounds)
    le = Lineextractor(200, 0, 0.8, 2, 2.0, 0, max_width=w, max_height=h, max_batch=n)
    le.extract_batch(frames)
    t0 = time.perf_counter()
    for _ in range(3):
        le.extract_batch(frames)
    dt = (time.perf_counter() - t0) / 3
    need, fb = [], 0
    for f in range(n):
        for o in (0, 1):
            fl = np.zeros(40, np.int32); cnt = C.c_int(0)
            lib.plvi_line_read_lsd(le._h, f, o, 7, fl.ctypes.data_as(C.c_void_p), 40, C.byref(cnt))
            fb += int(fl[0])
            d = fl[3:3 + rounds]   # dirty flag of rounds 1..rounds
            need.append(int(np.max(np.nonzero(d)[0])) + 1 if d.any() else 0)
    need = np.array(need)
    print(f"{w}x{h} n={n} rounds={rounds}: host-buffer extract {dt * 1e3:.1f} ms, fallbacks {fb}/{2 * n}, rounds with work: "
          f"hist {np.bincount(need).tolist()}", flush=True)
    le.close()
