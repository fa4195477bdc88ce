"""Matcher half of the oracle (oracle/mcs_oracle.cpp) against a second, independent plain-Python restatement written from
the reference sources (oracle/pyref_match.py) -- the counterpart of oracle/pyref.py for the extractor.  CPU only."""
import pathlib
import sys

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "oracle"))


@pytest.fixture(scope="module")
def pm():
    import pyref_match
    return pyref_match


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
        out.append(api.Frame.from_cameras(per, [(754, 480)] * 3, sf))
    return out


def test_grid_and_window_search(oa, pm, frames):
    from multicol_slam_b200.ctypes_defs import WINDOW_QUERY_DTYPE
    F = frames[0]
    grid = pm.Grid(F.keys, F.key_cam, [(754, 480)] * 3)
    rng = np.random.default_rng(0)
    nq = 300
    qs = np.zeros(nq, WINDOW_QUERY_DTYPE)
    kind = rng.integers(0, 3, nq)
    lv = rng.integers(0, 8, nq)
    qs["cam"] = rng.integers(0, 3, nq)
    qs["min_level"] = np.where(kind == 0, -1, np.where(kind == 1, lv, np.maximum(lv - 1, 0)))
    qs["max_level"] = np.where(kind == 0, -1, lv)
    qs["desc_index"] = rng.integers(0, len(F.keys), nq)
    qs["x"] = rng.uniform(-30, 790, nq); qs["y"] = rng.uniform(-30, 510, nq); qs["r"] = rng.uniform(1, 60, nq)
    oi, od, oc, rc = oa.window_search(F, qs, F.desc, F.dmask, max_cand=512)
    assert rc == 0 and oc.max() > 5
    for i in range(nq):
        cand = grid.features_in_area(int(qs["cam"][i]), qs["x"][i], qs["y"][i], qs["r"][i], int(qs["min_level"][i]), int(qs["max_level"][i]))
        assert cand == list(oi[i, :oc[i]]), i
        qd, qm = F.desc[qs["desc_index"][i]], F.dmask[qs["desc_index"][i]]
        assert [pm.distance(qd, F.desc[k], qm, F.dmask[k]) for k in cand] == list(od[i, :oc[i]])


def test_frame_prepare_rays_and_grid(oa, pm, frames, cams):
    F = frames[0]
    rays, start, items = oa.frame_prepare(F.keys, F.key_cam, cams)
    grid = pm.Grid(F.keys, F.key_cam, [(754, 480)] * 3)
    flat = []
    for c in range(3):
        for ix in range(pm.GRID_COLS):
            for iy in range(pm.GRID_ROWS):
                cell = (c * pm.GRID_COLS + ix) * pm.GRID_ROWS + iy
                assert list(items[start[cell]:start[cell + 1]]) == grid.cells[c][ix][iy]
                flat += grid.cells[c][ix][iy]
    assert len(flat) == len(items)
    for i in range(0, len(F.keys), 7):
        r = pm.img_to_world(cams[int(F.key_cam[i])], float(F.keys["x"][i]), float(F.keys["y"][i]))
        assert tuple(rays[i]) == r


@pytest.mark.parametrize("masks", [False, True])
def test_search_by_projection(oa, pm, frames, masks):
    import multicol_slam_b200.api as api
    F = frames[0]
    grid = pm.Grid(F.keys, F.key_cam, [(754, 480)] * 3)
    rng = np.random.default_rng(3 + masks)
    nmp, nc = 700, 3
    src = rng.integers(0, len(F.keys), nmp)
    desc = F.desc[src].copy()
    for i in range(nmp):
        for b in rng.choice(256, rng.integers(0, 41), replace=False):
            desc[i, b // 8] ^= 1 << (b % 8)
    dm = F.dmask[src].copy()
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
    th_high = 48 if masks else 96
    for th, pre in ((3.0, None), (1.0, np.where(np.arange(len(F.keys)) % 3 == 0, 0, -1))):
        start = np.full(len(F.keys), -1, np.int32) if pre is None else pre.astype(np.int32)
        on, ofmp = oa.search_by_projection(F, mps, th, 0.8, th_high, masks, start.copy())
        pn, pfmp = pm.search_by_projection(F, grid, mps, th, 0.8, th_high, masks, start.copy())
        assert on == pn and np.array_equal(ofmp, pfmp)
        assert on > 100


@pytest.mark.parametrize("masks", [False, True])
def test_search_for_initialization(oa, pm, frames, masks):
    F1, F2 = frames
    grid2 = pm.Grid(F2.keys, F2.key_cam, [(754, 480)] * 3)
    prev = np.stack([F1.keys["x"], F1.keys["y"]], axis=1).astype(np.float64)
    th_low = 32 if masks else 64
    on, om12, oprev = oa.search_for_initialization(F1, F2, prev, 50, 0.9, th_low, masks)
    pn, pm12, pprev = pm.search_for_initialization(F1, F2, grid2, prev, 50, 0.9, th_low, masks)
    assert on == pn and np.array_equal(om12, pm12) and np.array_equal(oprev, pprev)
    assert on > 150


@pytest.mark.parametrize("masks", [False, True])
def test_bruteforce_and_triangulation(oa, pm, frames, cams, masks):
    F1, F2 = frames
    rng = np.random.default_rng(5 + masks)
    s1 = rng.permutation(len(F1.keys))[:160]; s2 = rng.permutation(len(F2.keys))[:200]
    d1, d2 = F1.desc[s1], F2.desc[s2]
    m1, m2 = (F1.dmask[s1], F2.dmask[s2]) if masks else (None, None)
    v1 = (rng.random(len(d1)) < 0.8).astype(np.uint8); v2 = (rng.random(len(d2)) < 0.8).astype(np.uint8)
    th_low = 32 if masks else 64
    on, om = oa.match_bruteforce(d1, d2, th_low, 0.9, m1, m2, v1, v2)
    pn, pmm = pm.search_by_bow_kf(d1, d2, th_low, 0.9, m1, m2, v1, v2)
    assert on == pn and np.array_equal(om, pmm)
    # triangulation search: same-camera pairs, epipolar test with a small relative pose
    rays1 = np.array([pm.img_to_world(cams[int(F1.key_cam[i])], float(F1.keys["x"][i]), float(F1.keys["y"][i])) for i in s1])
    rays2 = np.array([pm.img_to_world(cams[int(F2.key_cam[i])], float(F2.keys["x"][i]), float(F2.keys["y"][i])) for i in s2])
    c1, c2 = F1.key_cam[s1], F2.key_cam[s2]
    t = np.array([0.02, -0.01, 0.005])
    tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
    E = np.zeros((3, 3, 3, 3))
    for a in range(3):
        for b in range(3):
            E[a, b] = tx * (1.0 + 0.1 * a - 0.05 * b)
    f1 = (rng.random(len(d1)) < 0.9).astype(np.uint8); f2 = (rng.random(len(d2)) < 0.9).astype(np.uint8)
    for thr in (1e-2, 1e-5):
        on, om = oa.search_for_triangulation(d1, m1, c1, f1, rays1, d2, m2, c2, f2, rays2, E, th_low, thr)
        pn, pmm = pm.search_for_triangulation(d1, m1, c1, f1, rays1, d2, m2, c2, f2, rays2, E, th_low, thr)
        assert on == pn and np.array_equal(om, pmm)


def test_project_mappoints(oa, pm, cams):
    from multicol_slam_b200 import synth
    rng = np.random.default_rng(8)
    nc, n = 3, 400
    masks = np.stack([synth.mirror_mask(c) for c in cams])
    mtmc, mtmc_inv = [], []
    for c in range(nc):
        a = rng.normal(0, 0.4, 3)
        th = np.linalg.norm(a); k = a / th
        K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
        R = np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K
        M = np.eye(4); M[:3, :3] = R; M[:3, 3] = rng.normal(0, 0.1, 3)
        mtmc.append(M); mtmc_inv.append(np.linalg.inv(M))
    pos = rng.normal(0, 1, (n, 3)); pos *= (rng.uniform(1.0, 6.0, n) / np.linalg.norm(pos, axis=1))[:, None]
    nrm = rng.normal(0, 1, (n, 3)); nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    dmin = rng.uniform(0.5, 2.0, n); dmax = dmin * rng.uniform(1.5, 6.0, n)
    sf = [1.2000000476837158 ** l for l in range(8)]
    in_view, level, px, py, vc = oa.project_mappoints(np.array(mtmc_inv), np.array(mtmc), cams, masks, pos, nrm, dmin, dmax, sf)
    seen = 0
    for i in range(n):
        for c in range(nc):
            r = pm.is_in_frustum(mtmc_inv[c], mtmc[c], cams[c], masks[c], pos[i], nrm[i], dmin[i], dmax[i], sf)
            assert bool(in_view[i, c]) == (r is not None), (i, c)
            if r is not None:
                seen += 1
                assert (px[i, c], py[i, c], level[i, c], vc[i, c]) == r, (i, c)
    assert seen > 100


@pytest.mark.parametrize("masks", [False, True])
def test_m4_window_searches(oa, pm, frames, masks):
    """WindowSearch (:326), SearchByProjection(F1,F2,win) (:476), SearchByProjection(Current,Last,th) (:1990): restatements of the
    three reference functions vs the oracle's generic rule loop (mcso_search_windows) fed with the queries the mirror builds."""
    from multicol_slam_b200.api import RULE_BEST, RULE_RATIO, _queries
    F1, F2 = frames
    g2 = pm.Grid(F2.keys, F2.key_cam, [(754, 480)] * 3)
    rng = np.random.default_rng(8 + masks)
    valid1 = (rng.random(len(F1.keys)) < 0.8).astype(np.uint8)
    th_high = 48 if masks else 96
    qm = F1.dmask if masks else None
    lv = F1.keys["octave"]
    free2 = np.full(len(F2.keys), -1, np.int32)

    sel = np.flatnonzero((valid1 != 0) & (lv >= 1) & (lv <= 5))
    q = _queries(F1.key_cam[sel], F1.keys["x"][sel].astype(np.float64), F1.keys["y"][sel].astype(np.float64), 60.0, -1, -1, sel)
    on, om21 = oa.search_windows(F2, q, F1.desc, qm, sel, RULE_RATIO, 0.8, th_high, free2.copy())
    pn, pm21 = pm.window_search(F1, F2, g2, 60, valid1, 0.8, th_high, masks, 1, 5)
    assert on == pn and np.array_equal(om21, pm21) and on > 100

    uv = np.zeros((len(F1.keys), 3, 2)); in_mask = np.zeros((len(F1.keys), 3), np.uint8)
    for c in range(3):
        uv[:, c, 0] = F1.keys["x"] - 3.0 + rng.normal(0, 1, len(F1.keys))
        uv[:, c, 1] = F1.keys["y"] - 2.0 + rng.normal(0, 1, len(F1.keys))
        in_mask[:, c] = (F1.key_cam == c) | (rng.random(len(F1.keys)) < 0.1)
    pre = free2.copy(); pre[::7] = 0
    i1, c = np.nonzero((valid1 != 0)[:, None] & (in_mask != 0))
    q = _queries(c, uv[i1, c, 0], uv[i1, c, 1], 40.0, lv[i1], lv[i1], i1)
    on2, oa2 = oa.search_windows(F2, q, F1.desc, qm, i1, RULE_RATIO, 0.8, th_high, pre.copy())
    pn2, pa2 = pm.search_by_projection_frames(F1, F2, g2, 40, valid1, uv, in_mask, pre.copy(), 0.8, th_high, masks)
    assert on2 == pn2 and np.array_equal(oa2, pa2) and on2 > 100

    uvl = np.stack([F1.keys["x"] - 3.0, F1.keys["y"] - 2.0], axis=1).astype(np.float64)
    inm = (rng.random(len(F1.keys)) < 0.95).astype(np.uint8)
    sel = np.flatnonzero((valid1 != 0) & (inm != 0))
    l3 = lv[sel]
    q = _queries(F1.key_cam[sel], uvl[sel, 0], uvl[sel, 1], 50.0 * F2.scale_factors[l3], l3 - 1, l3 + 1, sel)
    on3, oa3 = oa.search_windows(F2, q, F1.desc, qm, sel, RULE_BEST, 0.8, th_high, free2.copy())
    pn3, pa3 = pm.search_by_projection_last(F2, g2, F1, 50.0, valid1, uvl, inm, free2.copy(), th_high, masks)
    assert on3 == pn3 and np.array_equal(oa3, pa3) and on3 > 100


@pytest.mark.parametrize("masks", [False, True])
def test_fuse_candidates(oa, pm, frames, masks):
    """matching core of Fuse / SearchBySim3 vs the oracle's stateless best rule (the mirror's FuseCandidates query layout)"""
    from multicol_slam_b200.api import RULE_BEST_FREE, _queries
    KF = frames[1]
    grid = pm.Grid(KF.keys, KF.key_cam, [(754, 480)] * 3)
    rng = np.random.default_rng(21 + masks)
    n, nc = 500, 3
    src = rng.integers(0, len(KF.keys), n)
    desc = KF.desc[src].copy()
    flip = rng.integers(0, 256, desc.shape, dtype=np.uint8) & rng.integers(0, 256, desc.shape, dtype=np.uint8) & rng.integers(0, 256, desc.shape, dtype=np.uint8)
    desc ^= flip
    dm = KF.dmask[src].copy()
    uv = np.zeros((n, nc, 2)); in_mask = np.zeros((n, nc), np.uint8); level = np.zeros((n, nc), np.int32)
    for c in range(nc):
        own = KF.key_cam[src] == c
        uv[:, c, 0] = np.where(own, KF.keys["x"][src] + rng.normal(0, 1.5, n), rng.uniform(0, 754, n))
        uv[:, c, 1] = np.where(own, KF.keys["y"][src] + rng.normal(0, 1.5, n), rng.uniform(0, 480, n))
        in_mask[:, c] = own | (rng.random(n) < 0.15)
        level[:, c] = np.clip(KF.keys["octave"][src] + rng.integers(0, 2, n), 0, 7)
    th_low = 32 if masks else 64
    i, c = np.nonzero(in_mask != 0)
    lv = level[i, c]
    q = _queries(c, uv[i, c, 0], uv[i, c, 1], 3.0 * KF.scale_factors[lv], lv - 1, lv, i)
    on, res = oa.search_windows(KF, q, desc, dm if masks else None, np.zeros(len(q), np.int32), RULE_BEST_FREE, 0.8, th_low,
                                np.full(max(len(q), len(KF.keys)), -1, np.int32))
    got = np.full(in_mask.shape, -1, np.int64); got[i, c] = res[:len(q)]
    ref = pm.fuse_candidates(KF, grid, uv, in_mask, level, 3.0, desc, dm, th_low, masks)
    assert np.array_equal(got, ref) and (ref >= 0).sum() == on and on > 150
