#!/usr/bin/env python3
"""Generates tests/golden/ref_outputs.npz: outputs of THE REFERENCE'S OWN CODE, run in the build container.

oracle/_ref/libplvi_ref.so is built by oracle/Makefile.ref from the unmodified sources under /root/reference
(src/ORBextractor.cc, src/LSD/lsd.cpp, src/LineExtractor.cc, Thirdparty/line_descriptor/src/LSDDetector_custom.cpp,
binary_descriptor_custom.cpp) against the OpenCV/Eigen stand-in of oracle/cvmini/ (OpenCV primitives = the
scalar models pinned against cv2; heap addresses monotone, see oracle/ref_glue.cpp).  These vectors pin the oracle
restatement -- and through it the CUDA path -- to what the reference's authors wrote, on boxes where
/root/reference does not exist.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import oracle  # noqa: E402
from pl_vi_orbslam3_b200 import synth  # noqa: E402

GOLD = ROOT / "tests" / "golden"


def frame(name):
    if name.startswith("synth_"):
        p = name.split("_")
        return synth.frame_euroc(int(p[1]), *(int(v) for v in p[2:4])) if len(p) > 2 else synth.frame_euroc(int(p[1]))
    return np.load(GOLD / f"frame_{name}.npz")["img"]


# (frame, ORB nfeatures, lapping, line nfeatures)
CASES = [("data2_1", 1000, (0, 0), 200), ("data2_3", 2000, (0, 1000), 200), ("data_1_gray", 1000, (0, 0), 200),
         ("synth_0", 1000, (0, 0), 0), ("synth_7_640_480", 2000, (100, 300), 200), ("synth_10", 1000, (0, 0), 0),
         ("synth_3_1280_720", 2000, (0, 0), 200)]


def main():
    assert oracle.ref_available(), "needs /root/reference"
    out = {}
    for name, nf, lap, lnf in CASES:
        img = frame(name)
        o = oracle.ref_orb_extract(img, nfeatures=nf, lapping=lap)
        out[f"{name}/orb_kp"] = o["keypoints"]
        out[f"{name}/orb_desc"] = o["descriptors"]
        out[f"{name}/orb_mono"] = np.int32(o["mono_index"])
        li = oracle.ref_line_extract(img, lsd_nfeatures=lnf)
        out[f"{name}/line_kl"] = li["keylines"]
        out[f"{name}/line_desc"] = li["descriptors"]
        out[f"{name}/line_eq"] = li["line_eq"]
        out[f"{name}/lsd_raw"] = oracle.ref_lsd(img)
        print(name, len(o["keypoints"]), len(li["keylines"]), len(out[f"{name}/lsd_raw"]))
    # Frame::ComputeBoW through the reference's own DBoW2 (loadFromTextFile + transform) on a seeded synthetic vocabulary
    import tempfile
    from pl_vi_orbslam3_b200.vocabulary import ORBVocabulary
    v = ORBVocabulary.random_tree(k=8, L=4, seed=21, stop_fraction=0.03, early_leaf_fraction=0.1)
    with tempfile.TemporaryDirectory() as td:
        path = Path(td) / "voc.txt"
        v.save_text(path)
        path.write_text(path.read_text().rstrip("\n"))      # see tests/test_oracle_vs_ref.py on the trailing newline
        b = oracle.ref_bow_transform(path, oracle.orb_extract(frame("synth_0"))["descriptors"], 4)
    out["bow/words"], out["bow/values"] = b["bow"]
    out["bow/fv_nodes"], out["bow/fv_start"], out["bow/fv_features"] = b["fv"]
    # LineMatcher.cpp of the reference (compiled with the stand-in SLAM classes) on the reference's own LBD descriptors
    sys.path.insert(0, str(ROOT / "tests"))
    d1 = oracle.ref_line_extract(frame("synth_0"))["descriptors"]
    d2 = oracle.ref_line_extract(frame("synth_1"))["descriptors"]
    n, m = oracle.ref_line_match(d1, d2, 0.75, "match")
    out["linematch/match_n"], out["linematch/match"] = np.int32(n), m
    n, m = oracle.ref_line_match_mad(d1, d2, 0.5)
    out["linematch/init_n"], out["linematch/init"] = np.int32(n), m
    rng = np.random.default_rng(5)
    h1, h2 = (rng.random(len(d1)) < 0.3).astype(np.uint8), (rng.random(len(d2)) < 0.3).astype(np.uint8)
    n, m = oracle.ref_line_match_mad(d1, d2, 0.1, h1, h2)
    out["linematch/has1"], out["linematch/has2"] = h1, h2
    out["linematch/tri_n"], out["linematch/tri"] = np.int32(n), m
    import test_oracle_vs_ref as TL
    kl, desc, sf, q, qd, bad = TL.line_fuse_case(1, 8.0)
    n, bi = oracle.ref_line_fuse(kl, desc, TL.LINE_BOUNDS, sf, q, qd, bad, 8.0)
    out["linematch/fuse_n"], out["linematch/fuse"] = np.int32(n), bi
    print("linefuse", n)
    print("linematch", len(d1), len(d2), out["linematch/match_n"], out["linematch/init_n"], out["linematch/tri_n"])
    # ORBmatcher.cc of the reference (compiled with the stand-in SLAM classes): the four searches on a warped frame pair
    sys.path.insert(0, str(ROOT / "tests"))
    import test_oracle_vs_ref_matchers as T
    f1, f2, A = synth.warp_pair(3)
    r1, r2 = oracle.ref_orb_extract(f1), oracle.ref_orb_extract(f2)
    c = T.mappoint_case(r1, r2, A, 1, 3.0)
    n, mt = oracle.ref_search_mappoints(r2["keypoints"], r2["descriptors"], T.GRID, T.SCALES, c["proj"], c["viewcos"], c["level"],
                                        c["flags"], r1["descriptors"], 3.0, 0.8, c["blocked"])
    out["orbmatch/mappoints_n"], out["orbmatch/mappoints"] = np.int32(n), mt
    k1 = r1["keypoints"]
    prev = np.stack([k1["x"], k1["y"]], 1).astype(np.float32)
    n, m12, pm = oracle.ref_search_init(k1, r1["descriptors"], r2["keypoints"], r2["descriptors"], T.GRID, prev, 100, 0.9, True)
    out["orbmatch/init_n"], out["orbmatch/init"], out["orbmatch/init_prev"] = np.int32(n), m12, pm
    fv1, fv2, mp1, mp2 = T.bow_case(r1, r2, 6, 3, 2, 9)
    n, mt = oracle.ref_search_bow_kf_f(k1, r1["descriptors"], mp1, fv1, r2["keypoints"], r2["descriptors"], fv2, 0.7, True)
    out["orbmatch/bow_n"], out["orbmatch/bow"] = np.int32(n), mt
    n, m = oracle.ref_search_bow_kfkf(k1, r1["descriptors"], mp1, fv1, r2["keypoints"], r2["descriptors"], mp2, fv2, 0.8, True)
    out["orbmatch/bowkf_n"], out["orbmatch/bowkf"] = np.int32(n), m
    c = T.frame_case(r1, r2, A, 1, 7.0)
    n, mt = oracle.ref_search_frame(r2["keypoints"], r2["descriptors"], T.GRID, T.BOUNDS, T.SCALES, k1, c["uv"], c["flags"], r1["descriptors"],
                                    7.0, True, c["blocked"])
    out["orbmatch/frame_n"], out["orbmatch/frame"] = np.int32(n), mt
    fv1, fv2, mp1, mp2, F12, ep, sg2 = T.triangulation_case(r1, r2, 6, 3, 2, 3)
    n, m = oracle.ref_search_triangulation(k1, r1["descriptors"], mp1, fv1, r2["keypoints"], r2["descriptors"], mp2, fv2, F12, ep, T.SCALES,
                                           sg2, sg2, False, True)
    out["orbmatch/tri_n"], out["orbmatch/tri"] = np.int32(n), m
    c = T.kf_case(r1, r2, A, 1, 4.0)
    n, bi = oracle.ref_fuse(r2["keypoints"], r2["descriptors"], T.GRID, T.BOUNDS, T.SCALES, T.INV_SIGMA2, c["uv"], c["level"], c["flags"],
                            r1["descriptors"], 4.0, False)
    out["orbmatch/fuse_n"], out["orbmatch/fuse"] = np.int32(n), bi
    n, bi = oracle.ref_fuse(r2["keypoints"], r2["descriptors"], T.GRID, T.BOUNDS, T.SCALES, T.INV_SIGMA2, c["uv"], c["level"], c["flags"],
                            r1["descriptors"], 4.0, True)
    out["orbmatch/fuse_sim3_n"], out["orbmatch/fuse_sim3"] = np.int32(n), bi
    c = T.kf_case(r1, r2, A, 1, 15.0)
    n, mt = oracle.ref_search_by_projection_kf(r2["keypoints"], r2["descriptors"], T.GRID, T.BOUNDS, T.SCALES, c["uv"], c["level"],
                                               c["flags"], r1["descriptors"], 15, 1.5, c["matched_in"])
    out["orbmatch/kf_n"], out["orbmatch/kf"] = np.int32(n), mt
    (uv1, l1, f1), (uv2, l2, f2) = T.sim3_case(r1, r2, A, 1)
    n, m = oracle.ref_search_by_sim3(k1, r1["descriptors"], uv1, l1, f1, r2["keypoints"], r2["descriptors"], uv2, l2, f2, T.GRID, T.BOUNDS,
                                     T.SCALES, 7.5)
    out["orbmatch/sim3_n"], out["orbmatch/sim3"] = np.int32(n), m
    print("sim3", n)
    # Frame.cc of the reference (its own Frame.h over stand-in collaborators)
    import test_oracle_vs_ref_frame as TF
    fa, fb = TF.stereo_pair(5, 12)
    ur, dp, n = oracle.ref_stereo_matches(fa["keypoints"], fa["descriptors"], fb["keypoints"], fb["descriptors"], fa["pyramid"],
                                          fb["pyramid"], fa["plan"]["scale"], TF.MB, TF.MBF)
    out["frame/stereo_ur"], out["frame/stereo_depth"], out["frame/stereo_n"] = ur, dp, np.int32(n)
    keys = oracle.ref_orb_extract(frame("synth_0"))["keypoints"]
    bounds = (0.0, 752.0, 0.0, 480.0)
    xyr, lv = TF.area_queries(keys, 0)
    lists = oracle.ref_features_in_area(keys, bounds, xyr, lv)
    out["frame/area_start"] = np.cumsum([0] + [len(x) for x in lists]).astype(np.int32)
    out["frame/area_items"] = np.concatenate(lists).astype(np.int32)
    out["frame/grid_start"], out["frame/grid_items"] = oracle.ref_assign_grid(keys, bounds)
    uk = oracle.ref_undistort_keypoints(keys)
    out["frame/undistorted_xy"] = np.stack([uk["x"], uk["y"]], 1)
    print("frame", n, len(out["frame/area_items"]))
    dd, dc = T.distinctive_case(7)
    out["mappoint/distinctive"] = np.stack([oracle.ref_distinctive_descriptor(dd[p, :dc[p]]) for p in range(len(dc))])
    print("orbmatch2", out["orbmatch/frame_n"], out["orbmatch/tri_n"], out["orbmatch/fuse_n"], out["orbmatch/fuse_sim3_n"], out["orbmatch/kf_n"])
    print("orbmatch", out["orbmatch/mappoints_n"], out["orbmatch/init_n"], out["orbmatch/bow_n"], out["orbmatch/bowkf_n"])
    out["cases"] = np.array([f"{n}|{nf}|{lap[0]}|{lap[1]}|{lnf}" for n, nf, lap, lnf in CASES])
    np.savez_compressed(GOLD / "ref_outputs.npz", **out)
    print((GOLD / "ref_outputs.npz").stat().st_size, "bytes")


if __name__ == "__main__":
    main()
