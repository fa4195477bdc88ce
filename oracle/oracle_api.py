"""ctypes wrapper of oracle/libmcs_oracle.so -- TEST INFRASTRUCTURE (see mcs_oracle.cpp header).
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/--impl reference legs import this."""
import ctypes as C
import pathlib
import subprocess
import sys

import numpy as np

_HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(_HERE.parent))
from multicol_slam_b200.ctypes_defs import (ExtractorInfo, ExtractorParams, FrameView, KEYPOINT_DTYPE,  # noqa: E402
                                            MapPointView, Ocam, make_ocam, make_params)

_lib = None


def build():
    subprocess.check_call(["make", "-s", "-C", str(_HERE), "CXX=g++"])


def lib():
    global _lib
    if _lib is None:
        so = _HERE / "libmcs_oracle.so"
        if not so.exists():
            build()
        _lib = C.CDLL(str(so))
        _lib.mcso_extractor_create.restype = C.c_void_p
        _lib.mcso_fast_atan2.restype = C.c_float
        _lib.mcso_fast_atan2.argtypes = [C.c_float, C.c_float]
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


class OracleExtractor:
    def __init__(self, **kw):
        self.params = make_params(**kw)
        self.h = C.c_void_p(lib().mcso_extractor_create(C.byref(self.params)))
        if not self.h:
            raise ValueError("unsupported extractor params")
        self.info = ExtractorInfo()
        lib().mcso_extractor_get_info(self.h, C.byref(self.info))

    def __del__(self):
        if getattr(self, "h", None):
            lib().mcso_extractor_destroy(self.h)
            self.h = None

    def extract(self, image, mask, cam):
        """-> (kps structured array, desc [n,ds] u8, dmask [n,ds] u8)"""
        image = np.ascontiguousarray(image, np.uint8)
        mask = np.ascontiguousarray(mask, np.uint8)
        cap, ds = self.info.capacity, self.info.desc_size
        kps = np.zeros(cap, KEYPOINT_DTYPE)
        desc = np.zeros((cap, ds), np.uint8)
        dmask = np.zeros((cap, ds), np.uint8)
        n = C.c_int(0)
        oc = cam if isinstance(cam, Ocam) else make_ocam(cam)
        h, w = image.shape
        st = lib().mcso_extract(self.h, _p(image), w, h, image.strides[0], _p(mask), mask.strides[0], C.byref(oc),
                                _p(kps), _p(desc), _p(dmask), cap, C.byref(n))
        if st != 0:
            raise RuntimeError(f"oracle extract failed {st}")
        return kps[:n.value].copy(), desc[:n.value].copy(), dmask[:n.value].copy()

    def debug_read(self, level, what):
        w, h = C.c_int(0), C.c_int(0)
        buf = np.zeros(1 << 24, np.uint8)
        st = lib().mcso_debug_read(self.h, level, what, _p(buf), buf.nbytes, C.byref(w), C.byref(h))
        if st != 0:
            raise RuntimeError(f"debug_read {st}")
        if what == 3:
            return buf[:w.value * 12].view(np.int32).reshape(-1, 3).copy()
        return buf[:w.value * h.value].reshape(h.value, w.value).copy()


def resize_linear(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros((dh, dw), np.uint8)
    lib().mcso_resize_linear(_p(src), src.shape[1], src.shape[0], _p(dst), dw, dh)
    return dst


def resize_nearest(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros((dh, dw), np.uint8)
    lib().mcso_resize_nearest(_p(src), src.shape[1], src.shape[0], _p(dst), dw, dh)
    return dst


def box5(src):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros_like(src)
    lib().mcso_box5_reflect101(_p(src), src.shape[1], src.shape[0], _p(dst))
    return dst


def fast_atan2(y, x):
    return lib().mcso_fast_atan2(C.c_float(y), C.c_float(x))


def fast9(img, mask, threshold):
    img = np.ascontiguousarray(img, np.uint8)
    out = np.zeros((img.size, 3), np.int32)
    n = lib().mcso_fast9(_p(img), img.shape[1], img.shape[0], img.strides[0], _p(mask),
                         mask.strides[0] if mask is not None else 0, threshold, _p(out), out.shape[0])
    return out[:n].copy()


def octree(xyr, minX, maxX, minY, maxY, N):
    xyr = np.ascontiguousarray(xyr, np.float32)
    out = np.zeros((max(len(xyr), 1), 3), np.float32)
    n = lib().mcso_octree(_p(xyr), len(xyr), minX, maxX, minY, maxY, N, _p(out), out.shape[0])
    return out[:n].copy()


# ---- matcher oracles (take multicol_slam_b200.api.Frame / MapPoints, which are plain array holders) ----
def match_bruteforce(d1, d2, th_low, nnratio, m1=None, m2=None, valid1=None, valid2=None):
    d1 = np.ascontiguousarray(d1, np.uint8)
    d2 = np.ascontiguousarray(d2, np.uint8)
    m1 = None if m1 is None else np.ascontiguousarray(m1, np.uint8)
    m2 = None if m2 is None else np.ascontiguousarray(m2, np.uint8)
    valid1 = None if valid1 is None else np.ascontiguousarray(valid1, np.uint8)
    valid2 = None if valid2 is None else np.ascontiguousarray(valid2, np.uint8)
    m12 = np.zeros(len(d1), np.int32)
    n = C.c_int(0)
    lib().mcso_match_bruteforce(_p(d1), _p(m1), _p(valid1), len(d1), _p(d2), _p(m2), _p(valid2), len(d2), d1.shape[1],
                                th_low, C.c_double(nnratio), _p(m12), C.byref(n))
    return n.value, m12


def window_search(frame, queries, qdesc, qmask=None, max_cand=64):
    from multicol_slam_b200.ctypes_defs import WINDOW_QUERY_DTYPE
    queries = np.ascontiguousarray(queries, WINDOW_QUERY_DTYPE)
    qdesc = np.ascontiguousarray(qdesc, np.uint8)
    qmask = None if qmask is None else np.ascontiguousarray(qmask, np.uint8)
    nq = len(queries)
    idx = np.zeros((nq, max_cand), np.int32)
    dist = np.zeros((nq, max_cand), np.int32)
    cnt = np.zeros(nq, np.int32)
    fv = frame.view()
    rc = lib().mcso_window_search(C.byref(fv), _p(queries), nq, _p(qdesc), _p(qmask), max_cand, _p(idx), _p(dist), _p(cnt))
    return idx, dist, cnt, rc


def search_by_projection(frame, mps, th, nnratio, th_high, having_masks, frame_mp=None):
    if frame_mp is None:
        frame_mp = np.full(len(frame.keys), -1, np.int32)
    frame_mp = np.ascontiguousarray(frame_mp, np.int32).copy()
    n = C.c_int(0)
    fv, mv = frame.view(), mps.view()
    lib().mcso_search_by_projection(C.byref(fv), C.byref(mv), C.c_double(th), C.c_double(nnratio), th_high, int(having_masks),
                                    _p(frame_mp), C.byref(n))
    return n.value, frame_mp


def search_for_initialization(f1, f2, prev_matched, window, nnratio, th_low, having_masks):
    prev = np.ascontiguousarray(prev_matched, np.float64).copy()
    m12 = np.zeros(len(f1.keys), np.int32)
    n = C.c_int(0)
    v1, v2 = f1.view(), f2.view()
    lib().mcso_search_for_initialization(C.byref(v1), C.byref(v2), _p(prev), window, C.c_double(nnratio), th_low,
                                         int(having_masks), _p(m12), C.byref(n))
    return n.value, m12, prev


def distance64(a, b, dim=32):
    a = np.ascontiguousarray(a, np.uint8)
    b = np.ascontiguousarray(b, np.uint8)
    return lib().mcso_descriptor_distance64(_p(a), _p(b), dim)


def distance64_masked(a, b, ma, mb, dim=32):
    a, b, ma, mb = (np.ascontiguousarray(v, np.uint8) for v in (a, b, ma, mb))
    return lib().mcso_descriptor_distance64_masked(_p(a), _p(b), _p(ma), _p(mb), dim)


def search_windows(frame, queries, qdesc, qmask, query_tag, rule, nnratio, threshold, assigned):
    from multicol_slam_b200.ctypes_defs import WINDOW_QUERY_DTYPE
    queries = np.ascontiguousarray(queries, WINDOW_QUERY_DTYPE)
    qdesc = np.ascontiguousarray(qdesc, np.uint8)
    qmask = None if qmask is None else np.ascontiguousarray(qmask, np.uint8)
    tags = np.ascontiguousarray(query_tag, np.int32)
    assigned = np.ascontiguousarray(assigned, np.int32).copy()
    n = C.c_int(0)
    fv = frame.view()
    if qmask is None:
        fv.dmask = None
    lib().mcso_search_windows(C.byref(fv), _p(queries), len(queries), _p(qdesc), _p(qmask), _p(tags), rule, C.c_double(nnratio),
                              threshold, _p(assigned), C.byref(n))
    return n.value, assigned


def stream_mt(images, masks, cams, n_threads, nfeatures=2000, nlevels=8, do_dbrief=True, learn_masks=True, th_low=32, nnratio=0.9):
    """images [F,C,H,W]: multi-threaded C++ driver (std::thread) of extraction + previous-frame brute-force matching.
    Returns (n_features, n_matches).  Used by bench.py's CPU baseline / reference arm."""
    from multicol_slam_b200.ctypes_defs import Ocam
    images = np.ascontiguousarray(images, np.uint8)
    masks = np.ascontiguousarray(masks, np.uint8)
    F, Cn, H, W = images.shape
    p = make_params(nfeatures=nfeatures, nlevels=nlevels, do_dbrief=do_dbrief, learn_masks=learn_masks)
    ocs = (Ocam * Cn)(*[c if isinstance(c, Ocam) else make_ocam(c) for c in cams])
    nm = C.c_long(0)
    lib().mcso_stream_mt.restype = C.c_long
    n = lib().mcso_stream_mt(C.byref(p), n_threads, F, Cn, _p(images), W, H, _p(masks), ocs, th_low, C.c_double(nnratio), C.byref(nm))
    return n, nm.value


def project_mappoints(mtmc_inv, mtmc, cams, masks, world_pos, normal, min_dist, max_dist, scale_factors):
    from multicol_slam_b200.ctypes_defs import Ocam
    mi = np.ascontiguousarray(mtmc_inv, np.float64); mm = np.ascontiguousarray(mtmc, np.float64)
    masks = np.ascontiguousarray(masks, np.uint8)
    pos = np.ascontiguousarray(world_pos, np.float64); nrm = np.ascontiguousarray(normal, np.float64)
    dmin = np.ascontiguousarray(min_dist, np.float64); dmax = np.ascontiguousarray(max_dist, np.float64)
    sf = np.ascontiguousarray(scale_factors, np.float64)
    nc, n = len(cams), len(pos)
    ocs = (Ocam * nc)(*[c if isinstance(c, Ocam) else make_ocam(c) for c in cams])
    in_view = np.zeros((n, nc), np.uint8); level = np.zeros((n, nc), np.int32)
    px, py, vc = np.zeros((n, nc)), np.zeros((n, nc)), np.zeros((n, nc))
    lib().mcso_project_mappoints(nc, _p(mi), _p(mm), ocs, _p(masks), n, _p(pos), _p(nrm), _p(dmin), _p(dmax), _p(sf), len(sf),
                                 _p(in_view), _p(level), _p(px), _p(py), _p(vc))
    return in_view, level, px, py, vc


def search_for_triangulation(d1, m1, c1, f1, r1, d2, m2, c2, f2, r2, E, th_low, epi_thresh=1e-2):
    d1, d2 = np.ascontiguousarray(d1, np.uint8), np.ascontiguousarray(d2, np.uint8)
    m1 = None if m1 is None else np.ascontiguousarray(m1, np.uint8)
    m2 = None if m2 is None else np.ascontiguousarray(m2, np.uint8)
    c1, c2 = np.ascontiguousarray(c1, np.int32), np.ascontiguousarray(c2, np.int32)
    f1, f2 = np.ascontiguousarray(f1, np.uint8), np.ascontiguousarray(f2, np.uint8)
    r1, r2 = np.ascontiguousarray(r1, np.float64), np.ascontiguousarray(r2, np.float64)
    Em = np.ascontiguousarray(E, np.float64)
    m12 = np.zeros(len(d1), np.int32)
    n = C.c_int(0)
    lib().mcso_search_for_triangulation(_p(d1), _p(m1), _p(c1), _p(f1), _p(r1), len(d1), _p(d2), _p(m2), _p(c2), _p(f2), _p(r2), len(d2),
                                        d1.shape[1], th_low, _p(Em), Em.shape[0], C.c_double(epi_thresh), _p(m12), C.byref(n))
    return n.value, m12


def frame_prepare(keys, key_cam, cams):
    from multicol_slam_b200.ctypes_defs import Ocam
    keys = np.ascontiguousarray(keys, KEYPOINT_DTYPE)
    key_cam = np.ascontiguousarray(key_cam, np.int32)
    nc, n = len(cams), len(keys)
    ocs = (Ocam * nc)(*[c if isinstance(c, Ocam) else make_ocam(c) for c in cams])
    rays = np.zeros((n, 3)); start = np.zeros(nc * 64 * 48 + 1, np.int32); items = np.zeros(max(n, 1), np.int32)
    ning = C.c_int(0)
    lib().mcso_frame_prepare(_p(keys), _p(key_cam), n, ocs, nc, _p(rays), _p(start), _p(items), C.byref(ning))
    return rays, start, items[:ning.value].copy()


# ---- bag of words (restatement of the DBoW2 vocabulary as the reference uses it) ----------------------------
class OracleVocabulary:
    """voc: dict / npz with k, L, scoring, weighting, parent, weight, desc, node_order, word_node (tools/extract_vocabulary.py)."""

    def __init__(self, voc, scoring=None, weighting=None):
        self.scoring = int(voc["scoring"] if scoring is None else scoring)
        self.weighting = int(voc["weighting"] if weighting is None else weighting)
        par = np.ascontiguousarray(voc["parent"], np.int32); wt = np.ascontiguousarray(voc["weight"], np.float64)
        ds = np.ascontiguousarray(voc["desc"], np.uint8); wn = np.ascontiguousarray(voc["word_node"], np.int32)
        order = voc["node_order"] if "node_order" in voc else None
        order = None if order is None else np.ascontiguousarray(order, np.int32)
        lib().mcso_voc_create.restype = C.c_void_p
        self.h = C.c_void_p(lib().mcso_voc_create(int(voc["k"]), int(voc["L"]), self.scoring, self.weighting, len(par), _p(par), _p(wt),
                                                  _p(ds), _p(order), len(wn), _p(wn)))

    def __del__(self):
        if getattr(self, "h", None):
            lib().mcso_voc_destroy(self.h); self.h = None

    def transform_features(self, desc, levelsup=4):
        desc = np.ascontiguousarray(desc, np.uint8); n = len(desc)
        w = np.zeros(n, np.int32); wt = np.zeros(n, np.float64); nd = np.zeros(n, np.int32)
        lib().mcso_bow_transform(self.h, _p(desc), n, levelsup, _p(w), _p(wt), _p(nd))
        return w, wt, nd

    def transform(self, desc, levelsup=4):
        """-> (bow_words, bow_values, fv_nodes, fv_offsets, fv_features)"""
        desc = np.ascontiguousarray(desc, np.uint8); n = len(desc)
        bw = np.zeros(max(n, 1), np.int32); bv = np.zeros(max(n, 1), np.float64); nb = C.c_int(0)
        fn = np.zeros(max(n, 1), np.int32); fo = np.zeros(n + 2, np.int32); nf = C.c_int(0); ff = np.zeros(max(n, 1), np.int32)
        lib().mcso_bow_vectors(self.h, _p(desc), n, levelsup, _p(bw), _p(bv), C.byref(nb), _p(fn), _p(fo), C.byref(nf), _p(ff))
        return bw[:nb.value].copy(), bv[:nb.value].copy(), fn[:nf.value].copy(), fo[:nf.value + 1].copy(), ff[:fo[nf.value]].copy()

    def score(self, w1, v1, w2, v2):
        w1 = np.ascontiguousarray(w1, np.int32); w2 = np.ascontiguousarray(w2, np.int32)
        v1 = np.ascontiguousarray(v1, np.float64); v2 = np.ascontiguousarray(v2, np.float64)
        lib().mcso_bow_score.restype = C.c_double
        return lib().mcso_bow_score(self.scoring, _p(w1), _p(v1), len(w1), _p(w2), _p(v2), len(w2))


def search_by_bow(d1, m1, valid1, fv1, d2, m2, fv2, th_low, nnratio):
    """fv = (nodes, offsets, features) -> (nmatches, match_of_2)"""
    d1, d2 = np.ascontiguousarray(d1, np.uint8), np.ascontiguousarray(d2, np.uint8)
    m1 = None if m1 is None else np.ascontiguousarray(m1, np.uint8)
    m2 = None if m2 is None else np.ascontiguousarray(m2, np.uint8)
    valid1 = None if valid1 is None else np.ascontiguousarray(valid1, np.uint8)
    a = [np.ascontiguousarray(x, np.int32) for x in fv1]; b = [np.ascontiguousarray(x, np.int32) for x in fv2]
    out = np.zeros(len(d2), np.int32); n = C.c_int(0)
    lib().mcso_search_by_bow(_p(d1), _p(m1), _p(valid1), len(d1), _p(a[0]), _p(a[1]), len(a[0]), _p(a[2]), _p(d2), _p(m2), len(d2),
                             _p(b[0]), _p(b[1]), len(b[0]), _p(b[2]), d1.shape[1], th_low, C.c_double(nnratio), _p(out), C.byref(n))
    return n.value, out
