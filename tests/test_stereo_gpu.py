"""GPU parity: Frame::ComputeStereoMatches (src/Frame.cc:1228-1406) on the device-resident pyramids of two ORB extractor
handles vs the CPU oracle.  Bar: mvuRight / mvDepth bit-exact (float arithmetic in the reference's evaluation order),
the set of stereo points identical."""
import numpy as np
import pytest

import oracle
from pl_vi_orbslam3_b200 import ORBextractor, synth

pytestmark = pytest.mark.gpu

MBF, FX = 47.90639384423901, 435.2046959714599      # Examples/Stereo-Line/EuRoC.yaml: Camera.bf, Camera.fx
MB = np.float32(np.float32(MBF) / np.float32(FX))     # mb = mbf / fx (src/Frame.cc:171)


def stereo_pair(seed, disparity, w=752, h=480):
    left = synth.frame_euroc(seed, w, h)
    right = np.empty_like(left)
    right[:, :w - disparity] = left[:, disparity:]
    right[:, w - disparity:] = left[:, -1:]
    rng = np.random.RandomState(seed + 77)
    right = np.clip(right.astype(np.int16) + rng.randint(-3, 4, right.shape), 0, 255).astype(np.uint8)
    return left, right


@pytest.mark.parametrize("w,h,nfeat", [(752, 480, 1000), (640, 480, 2000)])
def test_stereo_matches_bit_exact(gpu, w, h, nfeat):
    import torch
    pairs = [stereo_pair(30 + i, d, w, h) for i, d in enumerate((12, 3, 40, 25))]
    lf = np.stack([p[0] for p in pairs])
    rf = np.stack([p[1] for p in pairs])
    el = ORBextractor(nfeat, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=4)
    er = ORBextractor(nfeat, 1.2, 8, 20, 7, max_width=w, max_height=h, max_batch=4)
    try:
        dl, dr = torch.from_numpy(lf).cuda(), torch.from_numpy(rf).cuda()
        lo = el.extract_batch_device(dl)
        ro = er.extract_batch_device(dr)
        ur, dp, ns = el.stereo_matches(er, lo, ro, MB, MBF)
        torch.cuda.synchronize()
        ur, dp, ns = ur.cpu().numpy(), dp.cpu().numpy(), ns.cpu().numpy()
        total = 0
        for i in range(len(pairs)):
            a = oracle.orb_extract(lf[i], nfeatures=nfeat, debug=True)
            b = oracle.orb_extract(rf[i], nfeatures=nfeat, debug=True)
            rur, rdp, rn = oracle.stereo_matches(a["keypoints"], a["descriptors"], b["keypoints"], b["descriptors"],
                                                 a["pyramid"], b["pyramid"], a["plan"]["scale"], MB, MBF)
            n = len(a["keypoints"])
            assert np.array_equal(ur[i, :n].view(np.uint32), rur.view(np.uint32)), i
            assert np.array_equal(dp[i, :n].view(np.uint32), rdp.view(np.uint32)), i
            assert ns[i] == rn == int((rur >= 0).sum())
            if oracle.ref_available():   # the reference's own Frame::ComputeStereoMatches (Frame.cc compiled unmodified)
                fur, fdp, fn = oracle.ref_stereo_matches(a["keypoints"], a["descriptors"], b["keypoints"], b["descriptors"],
                                                         a["pyramid"], b["pyramid"], a["plan"]["scale"], MB, MBF)
                assert fn == ns[i] and np.array_equal(ur[i, :n].view(np.uint32), fur.view(np.uint32))
                assert np.array_equal(dp[i, :n].view(np.uint32), fdp.view(np.uint32))
            assert (ur[i, n:] == -1).all()
            total += rn
        assert total > 800          # the pairs really produce stereo points
    finally:
        el.close()
        er.close()


def test_stereo_matches_rejects_mismatched_handles(gpu):
    import torch
    from pl_vi_orbslam3_b200 import PlviError
    el = ORBextractor(1000, 1.2, 8, 20, 7, max_width=752, max_height=480, max_batch=1)
    er = ORBextractor(1000, 1.2, 8, 20, 7, max_width=640, max_height=480, max_batch=1)
    try:
        lo = el.extract_batch_device(torch.from_numpy(synth.frame_euroc(1)[None]).cuda())
        ro = er.extract_batch_device(torch.from_numpy(synth.frame_euroc(1, 640, 480)[None]).cuda())
        with pytest.raises(PlviError):
            el.stereo_matches(er, lo, ro, MB, MBF)
    finally:
        el.close()
        er.close()
