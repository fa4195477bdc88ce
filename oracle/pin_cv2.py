"""pin_cv2.py -- pins oracle/mcs_oracle.cpp against the real OpenCV (cv2 4.13.0 here) and writes the
golden fixtures under tests/golden/.  Run in the build container:  python oracle/pin_cv2.py

1. primitive level: resize INTER_LINEAR / INTER_NEAREST, boxFilter 5x5, fastAtan2, FAST-9 per cell
   (cv2.FastFeatureDetector incl. mask filter) -- cv2 vs the hand-written C++ restatement.
2. pipeline level: oracle/pyref.py (cv2 primitives + independent Python restatement of the
   reference-specific logic) vs the C++ oracle, for ORB / dBRIEF / mdBRIEF on the 3 Lafida cameras.
3. goldens: small .npz files holding seeds + expected outputs, checked on CPU by tests/test_oracle_golden.py
   and on the GPU by tests/test_extract_gpu.py.
"""
import pathlib
import sys
import zlib

import cv2
import numpy as np

ROOT = pathlib.Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "oracle"))
import oracle_api as oa  # noqa: E402
import pyref  # noqa: E402
from multicol_slam_b200 import synth  # noqa: E402

GOLD = ROOT / "tests" / "golden"
GOLD.mkdir(parents=True, exist_ok=True)
cv2.setNumThreads(1)


def check(name, ok):
    print(f"[{'ok' if ok else 'FAIL'}] {name}")
    if not ok:
        raise SystemExit(1)


def primitives():
    cams = synth.lafida_cams()
    img = synth.frame(cams[0], 7)
    rng = np.random.default_rng(1)
    # resize chains (Lafida level sizes + odd sizes)
    src = img
    for (w, h) in [(628, 400), (524, 333), (436, 278), (364, 231), (303, 193), (253, 161), (210, 134)]:
        ref = cv2.resize(src, (w, h), interpolation=cv2.INTER_LINEAR)
        check(f"resize_linear {src.shape[::-1]}->{(w, h)}", np.array_equal(ref, oa.resize_linear(src, w, h)))
        m = (rng.integers(0, 2, size=src.shape) * 255).astype(np.uint8)
        check(f"resize_nearest ->{(w, h)}", np.array_equal(cv2.resize(m, (w, h), interpolation=cv2.INTER_NEAREST),
                                                           oa.resize_nearest(m, w, h)))
        src = ref
    for (sw, sh, dw, dh) in [(101, 77, 84, 64), (1920, 1080, 1600, 900), (1280, 720, 1067, 600), (64, 64, 53, 53)]:
        s = rng.integers(0, 256, size=(sh, sw)).astype(np.uint8)
        check(f"resize_linear rand {sw}x{sh}->{dw}x{dh}",
              np.array_equal(cv2.resize(s, (dw, dh), interpolation=cv2.INTER_LINEAR), oa.resize_linear(s, dw, dh)))
    # box filter
    for s in [img, rng.integers(0, 256, size=(61, 87)).astype(np.uint8)]:
        ref = cv2.boxFilter(s, -1, (5, 5), None, (-1, -1), True, cv2.BORDER_REFLECT_101)
        check(f"boxFilter {s.shape}", np.array_equal(ref, oa.box5(s)))
    # in-place on the ROI of a bordered buffer == isolated result (reference usage :1301)
    b = cv2.copyMakeBorder(img, 25, 25, 25, 25, cv2.BORDER_REFLECT_101)
    roi = b[25:-25, 25:-25]
    cv2.boxFilter(roi, -1, (5, 5), roi, (-1, -1), True, cv2.BORDER_REFLECT_101)
    check("boxFilter in-place ROI", np.array_equal(roi, oa.box5(img)))
    # fastAtan2
    ys = rng.integers(-200000, 200000, size=20000).astype(np.float32)
    xs = rng.integers(-200000, 200000, size=20000).astype(np.float32)
    ok = all(np.float32(cv2.fastAtan2(float(y), float(x))) == np.float32(oa.fast_atan2(float(y), float(x)))
             for y, x in zip(ys, xs))
    check("fastAtan2 20000 random + known answers", ok and
          np.float32(oa.fast_atan2(1, 1)) == np.float32(44.990455627441406) and
          np.float32(oa.fast_atan2(3, -4)) == np.float32(143.13629150390625) and
          np.float32(oa.fast_atan2(-7, 2)) == np.float32(285.94793701171875) and
          oa.fast_atan2(0, 0) == 0 and oa.fast_atan2(0, -5) == 180)
    # FAST per cell with mask
    fd = cv2.FastFeatureDetector_create(20, True, 2)
    tot = 0
    for t in range(40):
        h, w = int(rng.integers(8, 45)), int(rng.integers(8, 45))
        y0, x0 = int(rng.integers(0, 400)), int(rng.integers(0, 700))
        cell = img[y0:y0 + h, x0:x0 + w]
        mask = (rng.integers(0, 4, size=img.shape) > 0).astype(np.uint8) * 255
        mc = mask[y0:y0 + h, x0:x0 + w]
        kps = fd.detect(cell, mc)
        ref = np.array([[int(k.pt[0]), int(k.pt[1]), int(k.response)] for k in kps], np.int32).reshape(-1, 3)
        got = oa.fast9(cell, mc, 20)
        tot += len(ref)
        if not np.array_equal(ref, got):
            check(f"FAST cell {t}", False)
    for th in (5, 20, 40):
        fd2 = cv2.FastFeatureDetector_create(th, True, 2)
        kps = fd2.detect(img, None)
        ref = np.array([[int(k.pt[0]), int(k.pt[1]), int(k.response)] for k in kps], np.int32).reshape(-1, 3)
        check(f"FAST whole image th={th} ({len(ref)} kps)", np.array_equal(ref, oa.fast9(img, None, th)))
    check(f"FAST 40 random masked cells ({tot} kps)", True)


def kps_equal(pk, ck):
    if len(pk) != len(ck):
        return False
    for a, b in zip(pk, ck):
        if not (np.float32(a[0]) == b["x"] and np.float32(a[1]) == b["y"] and a[2] == b["size"] and
                np.float32(a[3]) == b["angle"] and a[4] == b["response"] and a[5] == b["octave"]):
            return False
    return True


def pipeline():
    cams = synth.lafida_cams()
    configs = [("orb", dict(do_dbrief=False, learn_masks=False)),
               ("dbrief", dict(do_dbrief=True, learn_masks=False)),
               ("mdbrief", dict(do_dbrief=True, learn_masks=True))]
    for ci, cam in enumerate(cams):
        img = synth.frame(cam, 16 * 0 + ci)
        mask = synth.mirror_mask(cam)
        pc = pyref.Cam(cam["c"], cam["d"], cam["e"], cam["u0"], cam["v0"], cam["pol"], cam["inv_pol"], cam["width"],
                       cam["height"], cam["mirror_mask"])
        check(f"mirror mask cam{ci}", np.array_equal(pc.mirror_mask_img(), mask))
        for name, kw in configs:
            if ci > 0 and name != "mdbrief":
                continue
            nf = 1000
            pe = pyref.Extractor(nfeatures=nf, **kw)
            r = pe(img, mask, pc)
            oe = oa.OracleExtractor(nfeatures=nf, **kw)
            k, d, m = oe.extract(img, mask, cam)
            check(f"cam{ci} {name}: quotas", list(oe.info.features_per_level[:8]) == pe.quota)
            for l in range(8):
                check(f"cam{ci} {name}: pyramid L{l}", np.array_equal(r["pyr"][l], oe.debug_read(l, 0)))
                check(f"cam{ci} {name}: mask pyr L{l}", np.array_equal(r["mpyr"][l], oe.debug_read(l, 2)))
                raw = np.array([[int(a[0]) + 22, int(a[1]) + 22, int(a[2])] for a in r["raws"][l]], np.int32).reshape(-1, 3)
                check(f"cam{ci} {name}: raw corners L{l} ({len(raw)})", np.array_equal(raw, oe.debug_read(l, 3)))
                check(f"cam{ci} {name}: blurred L{l}", np.array_equal(r["blur"][l], oe.debug_read(l, 1)))
            check(f"cam{ci} {name}: {len(k)} keypoints", kps_equal(r["kps"], k))
            check(f"cam{ci} {name}: descriptors", np.array_equal(r["desc"], d))
            if kw["learn_masks"]:
                check(f"cam{ci} {name}: masks", np.array_equal(r["dmask"], m))
            np.savez_compressed(GOLD / f"extract_lafida_cam{ci}_{name}_nf{nf}.npz", seed=np.int64(ci), cam_index=np.int64(ci),
                                nfeatures=np.int64(nf), do_dbrief=np.int64(kw["do_dbrief"]),
                                learn_masks=np.int64(kw["learn_masks"]), kps=k, desc=d, dmask=m,
                                image_crc=np.int64(zlib.crc32(img.tobytes())),
                                level_crc=np.array([zlib.crc32(oe.debug_read(l, 0).tobytes()) for l in range(8)], np.int64),
                                blur_crc=np.array([zlib.crc32(oe.debug_read(l, 1).tobytes()) for l in range(8)], np.int64),
                                raw_counts=np.array([len(oe.debug_read(l, 3)) for l in range(8)], np.int64),
                                raw_crc=np.array([zlib.crc32(oe.debug_read(l, 3).tobytes()) for l in range(8)], np.int64))


def pipeline_extra():
    """More parameter corners of the same pipeline, pyref (real cv2 primitives) vs the C++ oracle: other scale factors
    (other resize LUTs), level counts, FAST thresholds, feature budgets, descriptor sizes and sensor sizes.  Written as
    tests/golden/pin_*.npz and checked on CPU only (tests/test_oracle_golden.py::test_pinned_parameter_corners)."""
    cams = synth.lafida_cams()
    big = synth.scaled_cam(cams[1], 1280, 720)
    small = synth.scaled_cam(cams[2], 333, 211)
    cases = [
        ("init_th5_nf2000", cams[0], 31, dict(nfeatures=2000, fast_threshold=5, do_dbrief=True, learn_masks=True)),
        ("sf15_l5_dbrief", cams[1], 32, dict(nfeatures=800, scale_factor=1.5, nlevels=5, do_dbrief=True, learn_masks=False)),
        ("sf20_l3_orb", cams[2], 33, dict(nfeatures=600, scale_factor=2.0, nlevels=3)),
        ("sf11_l8_mdbrief", cams[0], 34, dict(nfeatures=500, scale_factor=1.1, nlevels=8, do_dbrief=True, learn_masks=True)),
        ("desc16_mdbrief", cams[1], 35, dict(nfeatures=400, do_dbrief=True, learn_masks=True, desc_size=16)),
        ("desc64_dbrief", cams[2], 36, dict(nfeatures=300, do_dbrief=True, learn_masks=False, desc_size=64)),
        ("1280x720_mdbrief", big, 37, dict(nfeatures=1200, do_dbrief=True, learn_masks=True)),
        ("333x211_l4_th40_orb", small, 38, dict(nfeatures=300, nlevels=4, fast_threshold=40)),
    ]
    for name, cam, seed, kw in cases:
        img = synth.frame(cam, seed)
        mask = synth.mirror_mask(cam)
        pc = pyref.Cam(cam["c"], cam["d"], cam["e"], cam["u0"], cam["v0"], cam["pol"], cam["inv_pol"], cam["width"],
                       cam["height"], cam["mirror_mask"])
        pe = pyref.Extractor(**kw)
        r = pe(img, mask, pc)
        oe = oa.OracleExtractor(**kw)
        k, d, m = oe.extract(img, mask, cam)
        L = pe.nlevels
        ok = kps_equal(r["kps"], k) and np.array_equal(r["desc"], d) and (not kw.get("learn_masks") or np.array_equal(r["dmask"], m))
        for l in range(L):
            ok = ok and np.array_equal(r["pyr"][l], oe.debug_read(l, 0)) and np.array_equal(r["blur"][l], oe.debug_read(l, 1))
            raw = np.array([[int(a[0]) + 22, int(a[1]) + 22, int(a[2])] for a in r["raws"][l]], np.int32).reshape(-1, 3)
            ok = ok and np.array_equal(raw, oe.debug_read(l, 3))
        check(f"extra {name}: {len(k)} keypoints, all stages", ok)
        np.savez_compressed(GOLD / f"pin_{name}.npz", seed=np.int64(seed), nlevels=np.int64(L),
                            cam_json=np.frombuffer(__import__("json").dumps(cam).encode(), np.uint8),
                            params_json=np.frombuffer(__import__("json").dumps(kw).encode(), np.uint8),
                            kps=k, desc=d, dmask=m, image_crc=np.int64(zlib.crc32(img.tobytes())),
                            level_crc=np.array([zlib.crc32(oe.debug_read(l, 0).tobytes()) for l in range(L)], np.int64),
                            blur_crc=np.array([zlib.crc32(oe.debug_read(l, 1).tobytes()) for l in range(L)], np.int64),
                            raw_crc=np.array([zlib.crc32(oe.debug_read(l, 3).tobytes()) for l in range(L)], np.int64))


if __name__ == "__main__":
    if "--extra-only" not in sys.argv:
        primitives()
        pipeline()
    pipeline_extra()
    print("all pinned")
