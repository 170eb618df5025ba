import sys, time
sys.path.insert(0, '.')
import numpy as np
import oracle
from pl_vi_orbslam3_b200 import Lineextractor, synth
e = Lineextractor(200, 0, 0.8, 2, 2.0, 0, max_batch=2)
img = synth.frame_euroc(0)
e.set_debug(True)
t = time.time(); kl, desc, eq = e(img); print('gpu call', time.time() - t, 'launches', e.last_launches)
ow, oh, sw, sh = e.octave_sizes(752, 480); print(ow, oh, sw, sh)
oct1 = oracle.resize_linear(img, int(ow[1]), int(oh[1]))
print('octave1 diff', (e.read_lsd(0, 1, 'octave', 752, 480) != oct1).sum())
for o, im in enumerate((img, oct1)):
    segs, dbg = oracle.lsd(im, 0.8, debug=True)
    sc = e.read_lsd(0, o, 'scaled', 752, 480); print(o, 'scaled diff', (sc != dbg['scaled']).sum(), np.abs(sc - dbg['scaled']).max())
    mg = e.read_lsd(0, o, 'modgrad', 752, 480); print(o, 'modgrad diff', (mg != dbg['modgrad']).sum())
    ang = e.read_lsd(0, o, 'angle_deg', 752, 480)
    ar = np.where(ang == -1024.0, -1024.0, ang.astype(np.float64) * (np.pi / 180)); print(o, 'angle diff', (ar != dbg['angles']).sum())
    got = e.read_lsd(0, o, 'segments', 752, 480); print(o, 'segs', len(got), len(segs))
    n = min(len(got), len(segs))
    if n: print(o, 'seg maxdiff', np.abs(got[:n] - segs[:n]).max(), 'exact', (got[:n] == segs[:n]).all(axis=1).mean())
ref = oracle.line_extract(img)
print('n', len(kl), len(ref['keylines']))
n = min(len(kl), len(ref['keylines']))
for f in oracle.KEYLINE_DTYPE.names:
    d = kl[f][:n] != ref['keylines'][f][:n]; print(f, d.sum(), end='; ')
print()
print('lbd bits differ', np.unpackbits(desc[:n] ^ ref['descriptors'][:n]).sum(), 'of', n * 256)
print('eq maxdiff', np.abs(eq[:n] - ref['line_eq'][:n]).max())
