"""GPU parity: CUDA extractor (through the C ABI) vs the CPU oracle and the golden fixtures.
Bit-exact for every integer/byte/index output; keypoint floats (IC angle, scaled coordinates) are compared
bit-for-bit as well (tolerance of the north star: 1e-5 on angles -- we require 0)."""
import glob
import pathlib
import zlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = pathlib.Path(__file__).resolve().parent / "golden"
MODES = {"orb": dict(do_dbrief=False, learn_masks=False), "dbrief": dict(do_dbrief=True, learn_masks=False),
         "mdbrief": dict(do_dbrief=True, learn_masks=True)}


def gpu_extractor(api, nf, mode, **kw):
    m = MODES[mode]
    return api.mdBRIEFextractorOct(nfeatures=nf, do_dBrief=m["do_dbrief"], learnMasks=m["learn_masks"], **kw)


def compare_all(api, oa, cam, img, mask, nf, mode, levels=True, **kw):
    ex = gpu_extractor(api, nf, mode, **kw)
    okw = {}
    if "fastThreshold" in kw:
        okw["fast_threshold"] = kw["fastThreshold"]
    if "descSize" in kw:
        okw["desc_size"] = kw["descSize"]
    if "nlevels" in kw:
        okw["nlevels"] = kw["nlevels"]
    if "scaleFactor" in kw:
        okw["scale_factor"] = kw["scaleFactor"]
    oe = oa.OracleExtractor(nfeatures=nf, **MODES[mode], **okw)
    k, d, m = ex(img, mask, cam)
    ok, od, om = oe.extract(img, mask, cam)
    if levels:
        for l in range(ex.GetLevels()):
            assert np.array_equal(ex.debug_read(l, 0), oe.debug_read(l, 0)), f"pyramid level {l}"
            assert np.array_equal(ex.debug_read(l, 1), oe.debug_read(l, 1)) or len(ok[ok["octave"] == l]) == 0, f"blur level {l}"
            assert np.array_equal(ex.debug_read(l, 2), oe.debug_read(l, 2)), f"mask level {l}"
            assert np.array_equal(ex.debug_read(l, 3), oe.debug_read(l, 3)), f"raw corners level {l}"
    assert len(k) == len(ok)
    for f in ("x", "y", "size", "response", "octave", "class_id"):
        assert np.array_equal(k[f], ok[f]), f"keypoint field {f}"
    assert np.array_equal(k["angle"].view(np.uint32), ok["angle"].view(np.uint32)), "IC angle bits"
    assert np.array_equal(d, od), "descriptor bits"
    assert np.array_equal(m, om), "descriptor mask bits"
    return k, d, m


@pytest.mark.parametrize("mode", ["orb", "dbrief", "mdbrief"])
def test_lafida_cam0_all_stages(api, oa, cams, mode):
    from multicol_slam_b200 import synth
    cam = cams[0]
    compare_all(api, oa, cam, synth.frame(cam, 0), synth.mirror_mask(cam), 1000, mode)


@pytest.mark.parametrize("path", sorted(glob.glob(str(GOLD / "extract_lafida_*.npz"))))
def test_golden_fixtures(api, cams, path):
    from multicol_slam_b200 import synth
    g = np.load(path)
    cam = cams[int(g["cam_index"])]
    img = synth.frame(cam, int(g["seed"]))
    ex = api.mdBRIEFextractorOct(nfeatures=int(g["nfeatures"]), do_dBrief=bool(g["do_dbrief"]), learnMasks=bool(g["learn_masks"]))
    k, d, m = ex(img, synth.mirror_mask(cam), cam)
    assert k.tobytes() == g["kps"].tobytes()
    assert np.array_equal(d, g["desc"]) and np.array_equal(m, g["dmask"])
    for l in range(8):
        assert zlib.crc32(ex.debug_read(l, 0).tobytes()) == int(g["level_crc"][l])
        assert zlib.crc32(ex.debug_read(l, 3).tobytes()) == int(g["raw_crc"][l])


@pytest.mark.parametrize("nf,ci,seed", [(2000, 1, 21), (400, 2, 5), (4000, 0, 9)])
def test_feature_budgets(api, oa, cams, nf, ci, seed):
    from multicol_slam_b200 import synth
    cam = cams[ci]
    compare_all(api, oa, cam, synth.frame(cam, seed), synth.mirror_mask(cam), nf, "mdbrief", levels=False)


def test_init_extractor_threshold5(api, oa, cams):
    # the tracker's init extractor: 2*nFeatures, FAST threshold 5 (ref src/cTracking.cpp:152-158)
    from multicol_slam_b200 import synth
    cam = cams[0]
    compare_all(api, oa, cam, synth.frame(cam, 33), synth.mirror_mask(cam), 800, "orb", fastThreshold=5)


@pytest.mark.parametrize("w,h", [(1280, 720), (1920, 1080), (333, 211)])
def test_other_image_sizes(api, oa, cams, w, h):
    from multicol_slam_b200 import synth
    cam = synth.scaled_cam(cams[0], w, h)
    compare_all(api, oa, cam, synth.frame(cam, w), synth.mirror_mask(cam), 1000 if w < 1900 else 4000, "mdbrief",
                levels=(w < 1900), nlevels=8 if w > 400 else 4)


@pytest.mark.parametrize("ds", [16, 64])
def test_descriptor_sizes(api, oa, cams, ds):
    from multicol_slam_b200 import synth
    cam = cams[1]
    compare_all(api, oa, cam, synth.frame(cam, 77), synth.mirror_mask(cam), 500, "mdbrief", levels=False, descSize=ds)


def test_edge_cases(api, oa, cams):
    from multicol_slam_b200 import synth
    cam = dict(cams[0])
    mask = synth.mirror_mask(cam)
    ex = api.mdBRIEFextractorOct(nfeatures=500)
    # empty image: silent return (ref :1252-1253)
    assert ex(np.zeros((0, 0), np.uint8), mask, cam) is None
    # flat image: zero keypoints (ref :1270-1274 releases the outputs)
    k, d, m = ex(np.full((480, 754), 128, np.uint8), mask, cam)
    assert len(k) == 0 and d.shape == (0, 32)
    # all-zero mask: corners detected but all filtered
    k, d, m = ex(synth.frame(cam, 1), np.zeros_like(mask), cam)
    assert len(k) == 0
    # all-ones mask (Camera.mirrorMask: 0) and a strided (non-contiguous rows) image
    cam["mirror_mask"] = 0
    ones = api.mirror_mask(cam)
    assert ones.min() == 1 and ones.max() == 1
    big = np.zeros((480, 800), np.uint8)
    big[:, :754] = synth.frame(cam, 2)
    k, d, m = ex(big[:, :754], ones, cam)
    ok, od, om = oa.OracleExtractor(nfeatures=500).extract(np.ascontiguousarray(big[:, :754]), ones, cam)
    assert k.tobytes() == ok.tobytes() and np.array_equal(d, od)


def test_batch_equals_single(api, oa, cams):
    from multicol_slam_b200 import synth
    imgs, coi = [], []
    for f in range(3):
        for c in range(3):
            imgs.append(synth.frame(cams[c], 16 * f + c))
            coi.append(c)
    masks = np.stack([synth.mirror_mask(c) for c in cams])
    ex = api.mdBRIEFextractorOct(nfeatures=1000, do_dBrief=True, learnMasks=True)
    kps, desc, dmask, counts = ex.extract_batch(np.stack(imgs), masks, cams, coi)
    oe = oa.OracleExtractor(nfeatures=1000, do_dbrief=True, learn_masks=True)
    for i in range(9):
        ok, od, om = oe.extract(imgs[i], masks[coi[i]], cams[coi[i]])
        n = counts[i]
        assert n == len(ok)
        assert kps[i, :n].tobytes() == ok.tobytes()
        assert np.array_equal(desc[i, :n], od) and np.array_equal(dmask[i, :n], om)


def test_full_size_properties(api, cams):
    """BASELINE config 2 sizes (2000 feat/cam, mdBRIEF) -- size-independent properties on a larger batch."""
    from multicol_slam_b200 import synth
    B = 24
    imgs = np.stack([synth.frame(cams[i % 3], 100 + i) for i in range(B)])
    masks = np.stack([synth.mirror_mask(c) for c in cams])
    coi = [i % 3 for i in range(B)]
    ex = api.mdBRIEFextractorOct(nfeatures=2000, do_dBrief=True, learnMasks=True)
    kps, desc, dmask, counts = ex.extract_batch(imgs, masks, cams, coi)
    quotas = list(ex.info.features_per_level[:8])
    for i in range(B):
        n = counts[i]
        k = kps[i, :n]
        assert 1900 <= n <= 2016
        assert np.all(np.diff(k["octave"]) >= 0)                       # level-major output order
        for l in range(8):
            assert (k["octave"] == l).sum() <= quotas[l] + 2           # the octree over-delivers by at most 2
        xi, yi = np.rint(k["x"]).astype(int), np.rint(k["y"]).astype(int)
        # the mask is tested on the LEVEL's nearest-neighbour copy (ref :1219-1247); at level 0 resolution a border keypoint may round outside
        assert np.mean(masks[coi[i]][np.clip(yi, 0, 479), np.clip(xi, 0, 753)] > 0) > 0.99
        assert np.all((k["angle"] >= 0) & (k["angle"] < 360.0001))
        assert np.all(k["response"] >= 20)
    # idempotence: same input, same bytes
    kps2, desc2, dmask2, counts2 = ex.extract_batch(imgs, masks, cams, coi)
    assert np.array_equal(counts, counts2)
    for i in range(B):
        assert kps[i, :counts[i]].tobytes() == kps2[i, :counts[i]].tobytes()
        assert np.array_equal(desc[i, :counts[i]], desc2[i, :counts[i]])


def test_device_api_pitched_and_unaligned(api, oa, cams):
    """mcs_extract_batch_device: 16-byte aligned pitch (128-bit staging loads) and an unaligned caller image
    (byte staging path) give the same bytes as the oracle."""
    import torch
    from multicol_slam_b200 import synth
    imgs = np.stack([synth.frame(cams[c], 50 + c) for c in range(3)])
    masks = np.stack([synth.mirror_mask(c) for c in cams])
    ex = api.mdBRIEFextractorOct(nfeatures=700, do_dBrief=True, learnMasks=True)
    oe = oa.OracleExtractor(nfeatures=700, do_dbrief=True, learn_masks=True)
    ref = [oe.extract(imgs[c], masks[c], cams[c]) for c in range(3)]
    dev = torch.device("cuda", 0)
    tight = torch.from_numpy(imgs).to(dev)                                   # stride 754: unaligned rows
    pitched = torch.zeros((3, 480, 768), dtype=torch.uint8, device=dev)
    pitched[:, :, :754] = tight
    st = torch.cuda.Stream(dev)
    for t, w in ((tight, None), (pitched, 754)):
        with torch.cuda.stream(st):
            out = ex.extract_batch_device(t, masks, cams, [0, 1, 2], stream=st, width=w)
        torch.cuda.synchronize(dev)
        counts = out["counts"].cpu().numpy()
        kps = out["kps"].cpu().numpy().view(api.KEYPOINT_DTYPE).reshape(3, -1)
        for c in range(3):
            n = counts[c]
            assert n == len(ref[c][0]) and kps[c, :n].tobytes() == ref[c][0].tobytes()
            assert np.array_equal(out["desc"][c, :n].cpu().numpy(), ref[c][1])
            assert np.array_equal(out["dmask"][c, :n].cpu().numpy(), ref[c][2])


def test_div25_magic():
    # K1's blur uses ((S+12)*5243)>>17 for (S+12)/25
    s = np.arange(0, 25 * 255 + 1, dtype=np.int64)
    assert np.array_equal(((s + 12) * 5243) >> 17, (s + 12) // 25)


@pytest.mark.parametrize("sf,nl", [(1.5, 5), (2.0, 3), (1.1, 8), (1.33, 6)])
def test_other_scale_factors(api, oa, cams, sf, nl):
    """Other pyramid geometries: wide source regions (the 16-lane staging path for scale factors > 1.5), many / few levels."""
    from multicol_slam_b200 import synth
    cam = cams[2]
    compare_all(api, oa, cam, synth.frame(cam, 123), synth.mirror_mask(cam), 800, "mdbrief", scaleFactor=sf, nlevels=nl)


def test_odd_sizes_batch(api, oa, cams):
    """Widths that are not multiples of the tile / vector sizes, in one batch of several images."""
    from multicol_slam_b200 import synth
    for (w, h) in [(641, 479), (700, 350)]:
        cam = synth.scaled_cam(cams[1], w, h)
        mask = synth.mirror_mask(cam)
        imgs = np.stack([synth.frame(cam, 500 + i) for i in range(4)])
        ex = api.mdBRIEFextractorOct(nfeatures=600, do_dBrief=True, learnMasks=True, nlevels=6)
        kps, desc, dmask, counts = ex.extract_batch(imgs, mask[None], [cam], [0, 0, 0, 0])
        oe = oa.OracleExtractor(nfeatures=600, do_dbrief=True, learn_masks=True, nlevels=6)
        for i in range(4):
            ok, od, om = oe.extract(imgs[i], mask, cam)
            n = counts[i]
            assert n == len(ok) and kps[i, :n].tobytes() == ok.tobytes()
            assert np.array_equal(desc[i, :n], od) and np.array_equal(dmask[i, :n], om)


def test_k3_tier_statistics(api, oa, cams):
    """K3 evaluates the distorted patterns in three tiers (fp32 relative to the keypoint / FP64 polynomial / exact): the result is
    bit-exact whatever tier decides (checked against the oracle here), and the cheap tier must carry almost all of the load"""
    from multicol_slam_b200 import synth
    cam = cams[1]
    img, mask = synth.frame(cam, 123), synth.mirror_mask(cam)
    ex = api.mdBRIEFextractorOct(nfeatures=2000, do_dBrief=True, learnMasks=True)
    ex.tier_stats(True)
    k, d, m = ex(img, mask, cam)
    t = ex.tier_stats(False)
    ok, od, om = oa.OracleExtractor(nfeatures=2000, do_dbrief=True, learn_masks=True).extract(img, mask, cam)
    assert k.tobytes() == ok.tobytes() and np.array_equal(d, od) and np.array_equal(m, om)
    assert t.sum() == 3 * len(k)
    print("K3 tiers (fp32, fp32 + FP64 repair, fp64 polynomial, exact):", t.tolist(), t / t.sum())
    assert t[0] > 0.85 * t.sum() and t[0] + t[1] > 0.93 * t.sum() and t[3] < 0.01 * t.sum()


@pytest.mark.gpu
def test_per_frame_call_is_graph_replayed(api, oa, cams):
    """mcs_extract_batch with a few images (the per-frame call of cMultiFrame's constructor) is served by one cached CUDA graph once
    the static inputs repeat; the result is the oracle's whether a call was run eagerly, captured or replayed, and a change of the
    masks, the camera table or the batch size drops the graph instead of replaying stale state"""
    from multicol_slam_b200 import synth
    masks = np.stack([synth.mirror_mask(c) for c in cams])
    ex = api.mdBRIEFextractorOct(nfeatures=600, do_dBrief=True, learnMasks=True)
    oe = oa.OracleExtractor(nfeatures=600, do_dbrief=True, learn_masks=True)

    def check(frame, coi, msk):
        imgs = np.stack([synth.frame(cams[c], 40 * frame + c) for c in coi])
        kps, desc, dmask, counts = ex.extract_batch(imgs, msk, cams, coi)
        for i, c in enumerate(coi):
            ok, od, om = oe.extract(imgs[i], msk[c], cams[c])
            n = counts[i]
            assert n == len(ok) and kps[i, :n].tobytes() == ok.tobytes()
            assert np.array_equal(desc[i, :n], od) and np.array_equal(dmask[i, :n], om)

    for f in range(4):
        check(f, [0, 1, 2], masks)
    assert ex.graph_replays() >= 2                      # first call eager (+ capture), later ones replayed
    r = ex.graph_replays()
    masks2 = masks.copy(); masks2[1, :200, :] = 0       # another mask: must not replay with the old tile flags / mask pyramid
    check(5, [0, 1, 2], masks2)
    assert ex.graph_replays() == r
    check(6, [0, 1, 2], masks2)
    assert ex.graph_replays() == r + 1
    check(7, [2, 1, 0], masks2)                         # another camera table
    assert ex.graph_replays() == r + 1
    check(8, [0, 1], masks2)                            # another batch size
    check(9, [0, 1], masks2)
    assert ex.graph_replays() == r + 2
    # a device-path call in between touches the shared buffers: the next per-frame call runs eagerly again
    import torch
    t = torch.zeros((2, 480, 768), dtype=torch.uint8, device="cuda")
    ex.extract_batch_device(t, masks2, cams, [0, 1], width=754)
    torch.cuda.synchronize()
    check(10, [0, 1], masks2)
    assert ex.graph_replays() == r + 2
    check(11, [0, 1], masks2)
    assert ex.graph_replays() == r + 3
