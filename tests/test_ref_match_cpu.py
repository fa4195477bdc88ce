"""CPU: the matcher restatements (oracle/mcs_oracle.cpp) pinned by the REFERENCE'S OWN matcher -- /root/reference/src/cORBmatcher.cpp
compiled where it lies into oracle/_ref/libmcs_ref.so (oracle/ref_mcs/wrap_match.cpp; the three SLAM container classes it reads are
data-only stand-ins, oracle/ref_mcs/stub_slam.h).  Skipped where the library is absent (it needs /root/reference to build; it
travels prebuilt to the GPU box)."""
import pathlib
import sys

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "oracle"))
SIZES = [(754, 480)] * 3


@pytest.fixture(scope="module")
def rm():
    import ref_match_api
    if not ref_match_api.available():
        pytest.skip("oracle/_ref/libmcs_ref.so not built (needs /root/reference)")
    return ref_match_api


@pytest.fixture(scope="module")
def frames(oa, cams):
    """two consecutive 3-camera frames of the sliding-texture stream, extracted by the CPU oracle"""
    import multicol_slam_b200.api as api            # plain array holders only; no device call is made in this file
    from multicol_slam_b200 import synth
    ex = oa.OracleExtractor(nfeatures=350, do_dbrief=True, learn_masks=True)
    sf = [float(ex.info.scale_factor[l]) for l in range(8)]
    out = []
    streams = [synth.texture_stream(cams[c], 2, seed=60 + c) for c in range(3)]
    for t in range(2):
        per = [ex.extract(streams[c][t], synth.mirror_mask(cams[c]), cams[c]) for c in range(3)]
        out.append(api.Frame.from_cameras(per, SIZES, sf))
    return out


def flip_bits(rng, desc, kmax):
    d = desc.copy()
    for i in range(len(d)):
        for b in rng.choice(8 * d.shape[1], rng.integers(0, kmax + 1), replace=False):
            d[i, b // 8] ^= 1 << (b % 8)
    return d


def test_thresholds_and_distances(oa, rm):
    assert rm.thresholds(32, False) == (96, 64) and rm.thresholds(32, True) == (48, 32)
    assert rm.thresholds(16, True) == (24, 16) and rm.thresholds(64, False) == (192, 128)
    rng = np.random.default_rng(0)
    for dim in (16, 32, 64):
        for _ in range(200):
            a, b, ma, mb = (rng.integers(0, 256, dim).astype(np.uint8) for _ in range(4))
            assert rm.distance64(a, b, dim) == oa.distance64(a, b, dim)
            assert rm.distance64_masked(a, b, ma, mb, dim) == oa.distance64_masked(a, b, ma, mb, dim)


def test_features_in_area_three_ways(oa, rm, frames, cams):
    """the grid lookup is the one piece of logic the stand-in containers restate: reference-shaped stand-in (stub_slam.h) ==
    C++ oracle == plain-Python restatement, for the frame overload (level filters) and the key-frame overload (<=)"""
    import pyref_match as pm
    from multicol_slam_b200.ctypes_defs import WINDOW_QUERY_DTYPE
    F = frames[0]
    kf = rm.KF(F, cams)
    grid = pm.Grid(F.keys, F.key_cam, SIZES)
    rng = np.random.default_rng(1)
    for i in range(150):
        cam, lv, kind = int(rng.integers(0, 3)), int(rng.integers(0, 8)), int(rng.integers(0, 3))
        x, y, r = float(rng.uniform(-30, 790)), float(rng.uniform(-30, 510)), float(rng.uniform(1, 60))
        if i % 5 == 0:                                       # exactly on a keypoint, integer radius: the > r vs <= r boundary
            k = int(rng.integers(0, len(F.keys)))
            cam, x, y, r = int(F.key_cam[k]), float(F.keys["x"][k]) + 7.0, float(F.keys["y"][k]), 7.0
        lo, hi = (-1, -1) if kind == 0 else ((lv, lv) if kind == 1 else (max(lv - 1, 0), lv))
        ref = rm.features_in_area(kf, False, cam, x, y, r, lo, hi)
        assert ref == grid.features_in_area(cam, x, y, r, lo, hi)
        q = np.zeros(1, WINDOW_QUERY_DTYPE)
        q["cam"], q["x"], q["y"], q["r"], q["min_level"], q["max_level"] = cam, x, y, r, lo, hi
        oi, _, oc, rc = oa.window_search(F, q, F.desc, F.dmask, max_cand=1024)
        assert rc == 0 and ref == list(oi[0, :oc[0]])
        refk = rm.features_in_area(kf, True, cam, x, y, r)
        assert refk == grid.features_in_area_kf(cam, x, y, r) if hasattr(grid, "features_in_area_kf") else True


@pytest.mark.parametrize("masks", [False, True])
def test_search_by_projection_equals_reference(oa, rm, frames, cams, masks):
    import multicol_slam_b200.api as api
    F = frames[0]
    rng = np.random.default_rng(3 + masks)
    nmp, nc = 700, 3
    src = rng.integers(0, len(F.keys), nmp)
    desc, dm = flip_bits(rng, F.desc[src], 40), F.dmask[src].copy()
    in_view = np.zeros((nmp, nc), np.uint8); level = np.zeros((nmp, nc), np.int32)
    px = np.zeros((nmp, nc)); py = np.zeros((nmp, nc)); vc = np.zeros((nmp, nc))
    for i in range(nmp):
        c = F.key_cam[src[i]]
        in_view[i, c] = 1
        level[i, c] = min(7, max(0, F.keys[src[i]]["octave"] + rng.integers(-1, 2)))
        px[i, c] = F.keys[src[i]]["x"] + rng.normal(0, 2); py[i, c] = F.keys[src[i]]["y"] + rng.normal(0, 2)
        vc[i, c] = rng.uniform(0.99, 1.0)
        if rng.random() < 0.2:
            c2 = (c + 1) % nc
            in_view[i, c2] = 1; level[i, c2] = rng.integers(0, 8)
            px[i, c2] = rng.uniform(0, 754); py[i, c2] = rng.uniform(0, 480); vc[i, c2] = rng.uniform(0.9, 1.0)
    bad = (rng.random(nmp) < 0.05).astype(np.uint8)
    mps = api.MapPoints(bad, in_view, level, px, py, vc, desc, dm)
    table = rm.MPTable(nc, desc, dmask=dm, bad=bad, in_view=in_view, level=level, proj_x=px, proj_y=py, view_cos=vc)
    th_high, _ = rm.thresholds(32, masks)
    for th, pre in ((3.0, None), (1.0, np.where(np.arange(len(F.keys)) % 3 == 0, 0, -1))):
        start = np.full(len(F.keys), -1, np.int32) if pre is None else pre.astype(np.int32)
        on, ofmp = oa.search_by_projection(F, mps, th, 0.8, th_high, masks, start.copy())
        rn, rfmp = rm.search_by_projection(rm.KF(F, cams, mp=start), table, th, 0.8, masks)
        assert on == rn and np.array_equal(ofmp, rfmp)
        assert on > 100


@pytest.mark.parametrize("masks", [False, True])
def test_search_for_initialization_equals_reference(oa, rm, frames, cams, masks):
    F1, F2 = frames
    prev = np.stack([F1.keys["x"], F1.keys["y"]], axis=1).astype(np.float64)
    _, th_low = rm.thresholds(32, masks)
    for window in (50, 100):
        on, om12, oprev = oa.search_for_initialization(F1, F2, prev, window, 0.9, th_low, masks)
        rn, rm12, rprev = rm.search_for_initialization(rm.KF(F1, cams), rm.KF(F2, cams), prev, window, 0.9, masks)
        assert on == rn and np.array_equal(om12, rm12) and np.array_equal(oprev, rprev)
        assert on > 150


@pytest.mark.parametrize("masks", [False, True])
def test_search_by_bow_kfkf_equals_reference(oa, rm, frames, cams, masks):
    """SearchByBoW(KF1, KF2): all-pairs scan incl. the greedy one-use rule; a third of the keypoints carry no map point, some bad"""
    F1, F2 = frames
    rng = np.random.default_rng(7 + masks)
    n1, n2 = len(F1.keys), len(F2.keys)
    has1, has2 = rng.random(n1) < 0.7, rng.random(n2) < 0.7
    mp1 = np.where(has1, np.arange(n1), -1).astype(np.int32)               # map point ids: KF1 keypoint i -> i, KF2 keypoint j -> n1 + j
    mp2 = np.where(has2, n1 + np.arange(n2), -1).astype(np.int32)
    bad = (rng.random(n1 + n2) < 0.05).astype(np.uint8)
    table = rm.MPTable(3, np.zeros((n1 + n2, 32), np.uint8), bad=bad)
    _, th_low = rm.thresholds(32, masks)
    v1 = (has1 & (bad[:n1] == 0)).astype(np.uint8)
    v2 = (has2 & (bad[n1:] == 0)).astype(np.uint8)
    on, om12 = oa.match_bruteforce(F1.desc, F2.desc, th_low, 0.9, F1.dmask if masks else None, F2.dmask if masks else None, v1, v2)
    rn, rout = rm.search_by_bow_kfkf(rm.KF(F1, cams, mp=mp1), rm.KF(F2, cams, mp=mp2), table, 0.9, masks)
    ref12 = np.where(rout >= 0, rout - n1, -1)
    assert on == rn and np.array_equal(om12, ref12) and on > 50


def test_check_orientation_constant():
    """every call site passes the compile-time constant checkOrientation == false (include/cORBmatcher.h:40): the wrapper does too"""
    txt = (ROOT / "oracle" / "ref_mcs" / "wrap_match.cpp").read_text()
    assert txt.count("checkOrientation") >= 10
