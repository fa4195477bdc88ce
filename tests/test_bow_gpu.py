"""GPU parity of the bag-of-words row: mcs_bow_transform / mcs_bow_vectors / mcs_bow_score /
mcs_search_by_bow through the C ABI vs the CPU oracle and vs the golden outputs of the reference's own DBoW2.
Word / node ids and feature lists bit-exact, weights and scores == (same double operation order)."""
import pathlib
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = pathlib.Path(__file__).resolve().parents[1]
GOLD = ROOT / "tests" / "golden"
sys.path.insert(0, str(GOLD))


@pytest.fixture(scope="module")
def voc():
    return np.load(GOLD / "voc_small_9_6.npz")


def same_transform(A, B):
    return (np.array_equal(A[0], B[0]) and np.array_equal(A[1], B[1]) and np.array_equal(A[2][0], B[2]) and np.array_equal(A[2][1], B[3]) and
            np.array_equal(A[2][2], B[4]))


def test_reference_golden(api, voc):
    g = np.load(GOLD / "bow_small_voc.npz")
    for ci in range(int(g["n_cases"][0])):
        sc, wg, levelsup, n, _ = (int(x) for x in g[f"c{ci}_cfg"])
        v = api.ORBVocabulary(voc, sc, wg)
        d = g[f"c{ci}_desc"]
        bw, bv, (fn, fo, ff) = v.transform(d, levelsup)
        assert np.array_equal(bw, g[f"c{ci}_bow_words"]) and np.array_equal(bv, g[f"c{ci}_bow_values"])
        assert np.array_equal(fn, g[f"c{ci}_fv_nodes"]) and np.array_equal(fo, g[f"c{ci}_fv_off"]) and np.array_equal(ff, g[f"c{ci}_fv_feat"])
        w, wt, _ = v.transform_features(d, levelsup)
        assert np.array_equal(w, g[f"c{ci}_word"]) and np.array_equal(wt, g[f"c{ci}_weight"])
        a = v.transform(d[:n // 2], levelsup); b = v.transform(d[n // 2:], levelsup)
        assert v.score(a[:2], b[:2]) == g[f"c{ci}_score"][0]


@pytest.mark.parametrize("scoring,weighting", [(0, 0), (1, 1), (2, 2), (3, 3), (4, 0), (5, 1), (5, 3)])
def test_transform_vs_oracle(api, oa, voc, scoring, weighting):
    from make_bow_golden import descriptors
    v = api.ORBVocabulary(voc, scoring, weighting); o = oa.OracleVocabulary(voc, scoring, weighting)
    d = descriptors(voc, 6000, 40 + scoring)
    for levelsup in (4, 1, 0, 6, 8):                # 1 and 0: leaves shallower than the node level -> documented rule
        assert same_transform(v.transform(d, levelsup), o.transform(d, levelsup)), levelsup
        w, wt, nd = v.transform_features(d, levelsup); w2, wt2, nd2 = o.transform_features(d, levelsup)
        assert np.array_equal(w, w2) and np.array_equal(wt, wt2) and np.array_equal(nd, nd2)
    a = v.transform(d[:3000]); b = v.transform(d[3000:])
    assert v.score(a[:2], b[:2]) == o.score(a[0], a[1], b[0], b[1])
    assert v.score(a[:2], a[:2]) == o.score(a[0], a[1], a[0], a[1])


def test_transform_extracted_descriptors(api, oa, voc, cams):
    """the real call: descriptors of all cameras of a frame, concatenated (ref src/cMultiFrame.cpp:356-363)"""
    from multicol_slam_b200 import synth
    ex = api.mdBRIEFextractorOct(nfeatures=2000, do_dBrief=True, learnMasks=True)
    descs = []
    for c in range(3):
        kps, d, m = ex(synth.frame(cams[c], 3 * 16 + c), synth.mirror_mask(cams[c]), cams[c])
        descs.append(d)
    d = np.concatenate(descs)
    assert len(d) > 5000
    v = api.ORBVocabulary(voc); o = oa.OracleVocabulary(voc)
    A = v.transform(d, 4)
    assert same_transform(A, o.transform(d, 4))
    assert abs(A[1].sum() - 1.0) < 1e-12 and (np.diff(A[0]) > 0).all() and (np.diff(A[2][0]) > 0).all()
    assert sorted(A[2][2].tolist()) == list(range(len(d))) or (A[1] > 0).all()


def test_loaders_and_edge_cases(api, oa, voc, tmp_path):
    from make_bow_golden import descriptors
    sys.path.insert(0, str(ROOT / "tools"))
    from extract_vocabulary import write_text
    write_text({k: voc[k] for k in voc.files}, tmp_path / "voc.txt")
    vt = api.ORBVocabulary.loadFromTextFile(tmp_path / "voc.txt")
    v = api.ORBVocabulary(voc)
    d = descriptors(voc, 500, 77)
    A, B = vt.transform(d), v.transform(d)
    assert np.array_equal(A[0], B[0]) and np.array_equal(A[1], B[1]) and all(np.array_equal(x, y) for x, y in zip(A[2], B[2]))
    assert vt.size() == v.size() == 6999
    # empty input: empty vectors
    bw, bv, (fn, fo, ff) = v.transform(np.zeros((0, 32), np.uint8))
    assert len(bw) == 0 and len(fn) == 0 and list(fo) == [0]
    # one descriptor; a node descriptor maps to itself along the way
    one = voc["desc"][voc["word_node"][123]][None, :]
    w, wt, nd = v.transform_features(one)
    ow, owt, ond = oa.OracleVocabulary(voc).transform_features(one)
    assert w[0] == ow[0] and wt[0] == owt[0] and nd[0] == ond[0]
    with pytest.raises(ValueError):
        v.transform(np.zeros((4, 16), np.uint8))
    # malformed trees are refused
    bad = {k: voc[k].copy() for k in voc.files}
    bad["parent"][5] = 5
    with pytest.raises(api.McsError):
        api.ORBVocabulary(bad)
    # wide and deep toy trees: k = 40 children (more than one pass of the 16-lane group), ties between identical children
    rng = np.random.default_rng(5)
    k = 40
    parent = np.concatenate([[0], np.zeros(k, np.int32), np.repeat(np.arange(1, k + 1), 3)]).astype(np.int32)
    n = len(parent)
    desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    desc[7] = desc[3]; desc[k + 5] = desc[k + 4]                      # duplicates: the first one must win
    leaf = np.ones(n, bool); leaf[parent[1:]] = False; leaf[0] = False
    toy = dict(k=k, L=2, scoring=0, weighting=0, parent=parent, weight=np.where(leaf, rng.random(n) + 0.1, 0.0), desc=desc,
               word_node=np.nonzero(leaf)[0].astype(np.int32))
    tv = api.ORBVocabulary(toy); to = oa.OracleVocabulary(toy)
    q = np.concatenate([desc[rng.integers(1, n, 300)], rng.integers(0, 256, (300, 32), dtype=np.uint8)]).astype(np.uint8)
    for levelsup in (1, 0, 2):
        assert same_transform(tv.transform(q, levelsup), to.transform(q, levelsup))


@pytest.mark.parametrize("masked", [False, True])
def test_search_by_bow_frame(api, oa, voc, cams, masked):
    """SearchByBoW(KF, F): key frame = frame t, frame = frame t+1 of the synthetic stream (same texture, shifted)"""
    from multicol_slam_b200 import synth
    ex = api.mdBRIEFextractorOct(nfeatures=1500, do_dBrief=True, learnMasks=True)
    imgs = [synth.texture_stream(cams[c], 2, seed=4 + c) for c in range(3)]
    D, M = [], []
    for f in range(2):
        ds, ms = [], []
        for c in range(3):
            kps, d, m = ex(imgs[c][f], synth.mirror_mask(cams[c]), cams[c])
            ds.append(d); ms.append(m)
        D.append(np.concatenate(ds)); M.append(np.concatenate(ms))
    v = api.ORBVocabulary(voc)
    fv1 = v.transform(D[0])[2]; fv2 = v.transform(D[1])[2]
    rng = np.random.default_rng(2)
    valid1 = (rng.random(len(D[0])) < 0.7).astype(np.uint8)
    mt = api.cORBmatcher(0.9, False, 32, masked)
    n_g, m_g = mt.SearchByBoWFrame(D[0], fv1, D[1], fv2, M[0] if masked else None, M[1] if masked else None, valid1)
    n_o, m_o = oa.search_by_bow(D[0], M[0] if masked else None, valid1, fv1, D[1], M[1] if masked else None, fv2, mt.TH_LOW_, 0.9)
    assert n_g == n_o and np.array_equal(m_g, m_o)
    assert n_g > 100 and (m_g >= 0).sum() == n_g
    assert all(valid1[i] for i in m_g[m_g >= 0])


def test_search_by_bow_one_big_node_chunks(api, oa):
    """all keypoints in one node: 3400 x 3400 distances > the 8 M staging buffer -> several GPU chunks, same greedy result"""
    rng = np.random.default_rng(9)
    n = 3400
    d1 = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    d2 = d1[rng.permutation(n)].copy()
    d2 ^= (rng.integers(0, 256, d2.shape, dtype=np.uint8) & rng.integers(0, 256, d2.shape, dtype=np.uint8) & rng.integers(0, 256, d2.shape, dtype=np.uint8))
    fv = (np.array([17], np.int32), np.array([0, n], np.int32), np.arange(n, dtype=np.int32))
    fv2 = (np.array([3, 17], np.int32), np.array([0, 0, n], np.int32), rng.permutation(n).astype(np.int32))
    mt = api.cORBmatcher(0.8, False, 32, False)
    n_g, m_g = mt.SearchByBoWFrame(d1, fv, d2, fv2)
    n_o, m_o = oa.search_by_bow(d1, None, None, fv, d2, None, fv2, mt.TH_LOW_, 0.8)
    assert n_g == n_o and np.array_equal(m_g, m_o) and n_g > 3000
