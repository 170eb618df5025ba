import sys, time
sys.path.insert(0, '.')
import numpy as np
import oracle
from pl_vi_orbslam3_b200 import ORBextractor, synth
e = ORBextractor(1000, 1.2, 8, 20, 7, max_batch=4)
img = synth.frame_euroc(0)
mono, kps, desc = e(img)
ref = oracle.orb_extract(img, debug=True)
print('n', len(kps), len(ref['keypoints']), 'mono', mono, ref['mono_index'], 'launches', e.last_launches)
for l in range(8):
    got = e.read_level(0, l, 752, 480); print('pyr', l, (got != ref['pyramid'][l]).sum(), end=' | ')
    got = e.read_level(0, l, 752, 480, blurred=True); print('blur', (got != ref['blurred'][l]).sum(), end=' | ')
    c = e.read_candidates(0, l); rc = oracle.grid_fast(ref['pyramid'][l]).astype(np.int32)
    c = c[np.lexsort((c[:, 0], c[:, 1]))]; rc = rc[np.lexsort((rc[:, 0], rc[:, 1]))]
    print('cand', len(c), len(rc), np.array_equal(c, rc))
rk = ref['keypoints']
n = min(len(kps), len(rk))
for fld in ('x','y','size','response','octave','angle'):
    d = (kps[fld][:n] != rk[fld][:n]); print(fld, d.sum(), np.nonzero(d)[0][:5])
print('desc bits differ', np.unpackbits(desc[:n] ^ ref['descriptors'][:n]).sum(), 'of', n*256)
