"""ctypes wrapper of the matcher half of oracle/_ref/libmcs_ref.so: the REFERENCE's own src/cORBmatcher.cpp (with its
cam_system_omni.cpp / cam_model_omni.cpp / cConverter.cpp / misc.cpp / DBoW2 FeatureVector.cpp) compiled where it lies by
`make -C oracle ref`; the three SLAM container classes it reads are data-only stand-ins (oracle/ref_mcs/stub_slam.h).
TEST INFRASTRUCTURE: pins the matcher restatements and generates tests/golden/ref_match_*.npz."""
import ctypes as C
import pathlib
import sys

import numpy as np

_HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(_HERE.parent))
from multicol_slam_b200.ctypes_defs import FrameView, Ocam, make_ocam  # noqa: E402
import ref_mcs_api as _ra  # noqa: E402

available = _ra.available
lib = _ra.lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p).value if a is not None else None


class _KF(C.Structure):
    _fields_ = [("view", FrameView), ("rays", C.c_void_p), ("mp", C.c_void_p), ("outlier", C.c_void_p), ("cams", C.c_void_p),
                ("M_c", C.c_void_p), ("M_t", C.c_void_p), ("fv_n", C.c_int32), ("fv_nodes", C.c_void_p), ("fv_offsets", C.c_void_p),
                ("fv_features", C.c_void_p)]


class _MPS(C.Structure):
    _fields_ = [("n", C.c_int32), ("n_cams", C.c_int32), ("dim", C.c_int32), ("bad", C.c_void_p), ("world_pos", C.c_void_p),
                ("normal", C.c_void_p), ("min_dist", C.c_void_p), ("max_dist", C.c_void_p), ("desc", C.c_void_p), ("dmask", C.c_void_p),
                ("in_view", C.c_void_p), ("level", C.c_void_p), ("proj_x", C.c_void_p), ("proj_y", C.c_void_p), ("view_cos", C.c_void_p),
                ("obs_kf", C.c_void_p), ("obs_idx", C.c_void_p)]


class KF:
    """A cMultiFrame / cMultiKeyFrame for the reference matcher.  frame: multicol_slam_b200.api.Frame (keys, key_cam, desc, dmask,
    camera sizes, scale factors); cams: list of camera dicts; M_c [n_cams,4,4], M_t [4,4] (identity by default);
    mp [n_keys] map point index per keypoint; rays [n_keys,3]; outlier [n_keys]; featvec = (nodes, offsets, features)."""

    def __init__(self, frame, cams, M_c=None, M_t=None, mp=None, rays=None, outlier=None, featvec=None):
        self.frame = frame
        nc = len(frame.cam_w)
        self.cams = (Ocam * nc)(*[make_ocam(c) for c in cams])
        self.M_c = np.ascontiguousarray(np.tile(np.eye(4), (nc, 1, 1)) if M_c is None else M_c, np.float64)
        self.M_t = np.ascontiguousarray(np.eye(4) if M_t is None else M_t, np.float64)
        self.mp = None if mp is None else np.ascontiguousarray(mp, np.int32)
        self.rays = None if rays is None else np.ascontiguousarray(rays, np.float64)
        self.outlier = None if outlier is None else np.ascontiguousarray(outlier, np.uint8)
        self.fv = None if featvec is None else [np.ascontiguousarray(a, np.int32) for a in featvec]

    def struct(self, with_masks=True):
        k = _KF()
        k.view = self.frame.view()
        if not with_masks:
            k.view.dmask = None
        k.rays, k.mp, k.outlier = _p(self.rays), _p(self.mp), _p(self.outlier)
        k.cams = C.cast(self.cams, C.c_void_p).value
        k.M_c, k.M_t = _p(self.M_c), _p(self.M_t)
        if self.fv is not None:
            k.fv_n, k.fv_nodes, k.fv_offsets, k.fv_features = len(self.fv[0]), _p(self.fv[0]), _p(self.fv[1]), _p(self.fv[2])
        return k


class MPTable:
    """The map-point table of one call (all arrays optional except desc)."""
    FIELDS = dict(bad=np.uint8, world_pos=np.float64, normal=np.float64, min_dist=np.float64, max_dist=np.float64, desc=np.uint8,
                  dmask=np.uint8, in_view=np.uint8, level=np.int32, proj_x=np.float64, proj_y=np.float64, view_cos=np.float64,
                  obs_kf=np.int32, obs_idx=np.int32)

    def __init__(self, n_cams, desc, **kw):
        self.n_cams = n_cams
        self.a = {k: None for k in self.FIELDS}
        kw["desc"] = desc
        for k, v in kw.items():
            self.a[k] = None if v is None else np.ascontiguousarray(v, self.FIELDS[k])
        if self.a["bad"] is None:
            self.a["bad"] = np.zeros(len(self.a["desc"]), np.uint8)

    def struct(self):
        m = _MPS()
        m.n, m.n_cams, m.dim = len(self.a["desc"]), self.n_cams, self.a["desc"].shape[1]
        for k in self.FIELDS:
            setattr(m, k, _p(self.a[k]))
        return m


def thresholds(feat_dim, having_masks):
    hi, lo = C.c_int(), C.c_int()
    lib().mcsref_thresholds(feat_dim, int(having_masks), C.byref(hi), C.byref(lo))
    return hi.value, lo.value


def distance64(a, b, dim=32):
    a, b = np.ascontiguousarray(a, np.uint8), np.ascontiguousarray(b, np.uint8)
    return lib().mcsref_descriptor_distance64(C.c_void_p(_p(a)), C.c_void_p(_p(b)), dim)


def distance64_masked(a, b, ma, mb, dim=32):
    a, b, ma, mb = (np.ascontiguousarray(x, np.uint8) for x in (a, b, ma, mb))
    return lib().mcsref_descriptor_distance64_masked(C.c_void_p(_p(a)), C.c_void_p(_p(b)), C.c_void_p(_p(ma)), C.c_void_p(_p(mb)), dim)


def _rc(n):
    if n <= -1000:
        raise RuntimeError("reference matcher raised")
    return n


def features_in_area(kf, keyframe, cam, x, y, r, min_level=-1, max_level=-1):
    out = np.zeros(len(kf.frame.keys) + 1, np.int32)
    k = kf.struct()
    n = _rc(lib().mcsref_features_in_area(C.byref(k), int(keyframe), cam, C.c_double(x), C.c_double(y), C.c_double(r), min_level, max_level,
                                          C.c_void_p(_p(out)), len(out)))
    return out[:n].tolist()


def search_by_projection(F, mps, th, nnratio, masks):
    out = np.full(len(F.frame.keys), -1, np.int32)
    k, m = F.struct(masks), mps.struct()
    n = _rc(lib().mcsref_search_by_projection(C.byref(k), C.byref(m), C.c_double(th), C.c_double(nnratio), int(masks), C.c_void_p(_p(out))))
    return n, out


def search_for_initialization(F1, F2, prev, window, nnratio, masks):
    prev = np.ascontiguousarray(prev, np.float64).copy()
    m12 = np.full(len(F1.frame.keys), -1, np.int32)
    a, b = F1.struct(masks), F2.struct(masks)
    n = _rc(lib().mcsref_search_for_initialization(C.byref(a), C.byref(b), C.c_void_p(_p(prev)), window, C.c_double(nnratio), int(masks),
                                                   C.c_void_p(_p(m12))))
    return n, m12, prev


def search_by_bow_kfkf(K1, K2, mps, nnratio, masks):
    out = np.full(len(K1.frame.keys), -1, np.int32)
    a, b, m = K1.struct(masks), K2.struct(masks), mps.struct()
    n = _rc(lib().mcsref_search_by_bow_kfkf(C.byref(a), C.byref(b), C.byref(m), C.c_double(nnratio), int(masks), C.c_void_p(_p(out))))
    return n, out


def search_by_bow_kff(K, F, mps, nnratio, masks):
    out = np.full(len(F.frame.keys), -1, np.int32)
    a, b, m = K.struct(masks), F.struct(masks), mps.struct()
    n = _rc(lib().mcsref_search_by_bow_kff(C.byref(a), C.byref(b), C.byref(m), C.c_double(nnratio), int(masks), C.c_void_p(_p(out))))
    return n, out


def window_search(F1, F2, mps, window, min_level, max_level, nnratio, masks):
    out = np.full(len(F2.frame.keys), -1, np.int32)
    a, b, m = F1.struct(masks), F2.struct(masks), mps.struct()
    n = _rc(lib().mcsref_window_search(C.byref(a), C.byref(b), C.byref(m), window, min_level, max_level, C.c_double(nnratio), int(masks),
                                       C.c_void_p(_p(out))))
    return n, out


def search_by_projection_frames(F1, F2, mps, window, nnratio, masks):
    out = np.full(len(F2.frame.keys), -1, np.int32)
    a, b, m = F1.struct(masks), F2.struct(masks), mps.struct()
    n = _rc(lib().mcsref_search_by_projection_frames(C.byref(a), C.byref(b), C.byref(m), window, C.c_double(nnratio), int(masks),
                                                     C.c_void_p(_p(out))))
    return n, out


def search_by_projection_last(Cur, Last, mps, th, nnratio, masks):
    out = np.full(len(Cur.frame.keys), -1, np.int32)
    a, b, m = Cur.struct(masks), Last.struct(masks), mps.struct()
    n = _rc(lib().mcsref_search_by_projection_last(C.byref(a), C.byref(b), C.byref(m), C.c_double(th), C.c_double(nnratio), int(masks),
                                                   C.c_void_p(_p(out))))
    return n, out


def search_by_projection_reloc(Cur, K, mps, already_found, th, orb_dist, nnratio, masks):
    out = np.full(len(Cur.frame.keys), -1, np.int32)
    af = None if already_found is None else np.ascontiguousarray(already_found, np.uint8)
    a, b, m = Cur.struct(masks), K.struct(masks), mps.struct()
    n = _rc(lib().mcsref_search_by_projection_reloc(C.byref(a), C.byref(b), C.byref(m), C.c_void_p(_p(af)), C.c_double(th), orb_dist,
                                                    C.c_double(nnratio), int(masks), C.c_void_p(_p(out))))
    return n, out


def search_by_projection_scw(K, mps, Scw, points, matched, th, nnratio, masks):
    Scw = np.ascontiguousarray(Scw, np.float64)
    points = np.ascontiguousarray(points, np.int32)
    matched = np.ascontiguousarray(matched, np.int32).copy()
    a, m = K.struct(masks), mps.struct()
    n = _rc(lib().mcsref_search_by_projection_scw(C.byref(a), C.byref(m), C.c_void_p(_p(Scw)), C.c_void_p(_p(points)), len(points), int(th),
                                                  C.c_double(nnratio), int(masks), C.c_void_p(_p(matched))))
    return n, matched


def search_for_triangulation_raw(K1, K2, mps, nnratio, masks):
    pairs = np.zeros((len(K1.frame.keys) + 1, 2), np.int32)
    a, b, m = K1.struct(masks), K2.struct(masks), mps.struct()
    n = _rc(lib().mcsref_search_for_triangulation_raw(C.byref(a), C.byref(b), C.byref(m), C.c_double(nnratio), int(masks),
                                                      C.c_void_p(_p(pairs)), len(pairs)))
    return n, pairs[:n].copy()


def search_for_triangulation_between(K1, mps, cam1, cam2, nnratio, masks):
    pairs = np.zeros((len(K1.frame.keys) + 1, 2), np.int32)
    a, m = K1.struct(masks), mps.struct()
    n = _rc(lib().mcsref_search_for_triangulation_between(C.byref(a), C.byref(m), cam1, cam2, C.c_double(nnratio), int(masks),
                                                          C.c_void_p(_p(pairs)), len(pairs)))
    return n, pairs[:n].copy()


def search_by_sim3(K1, K2, mps, s12, R12, t12, th, matches12, nnratio, masks):
    R12, t12 = np.ascontiguousarray(R12, np.float64), np.ascontiguousarray(t12, np.float64)
    m12 = np.ascontiguousarray(matches12, np.int32).copy()
    a, b, m = K1.struct(masks), K2.struct(masks), mps.struct()
    n = _rc(lib().mcsref_search_by_sim3(C.byref(a), C.byref(b), C.byref(m), C.c_double(s12), C.c_void_p(_p(R12)), C.c_void_p(_p(t12)),
                                        C.c_double(th), C.c_double(nnratio), int(masks), C.c_void_p(_p(m12))))
    return n, m12


def fuse(variant, K, mps, points, th, nnratio, masks, CurK=None, Scw=None):
    """-> (nFused, ops [k,3]: (0, mp, keypoint idx) = AddObservation + AddMapPoint, (1, mp, other mp) = Replace)"""
    points = np.ascontiguousarray(points, np.int32)
    Scw = None if Scw is None else np.ascontiguousarray(Scw, np.float64)
    cap = 4 * len(points) * K.frame.view().n_cams + 16
    ops = np.zeros((cap, 3), np.int32)
    nops = C.c_int32(0)
    a, m = K.struct(masks), mps.struct()
    c = CurK.struct(masks) if CurK is not None else None
    n = _rc(lib().mcsref_fuse(variant, C.byref(a), C.byref(c) if c is not None else None, C.byref(m), C.c_void_p(_p(Scw)),
                              C.c_void_p(_p(points)), len(points), C.c_double(th), C.c_double(nnratio), int(masks), C.c_void_p(_p(ops)), cap,
                              C.byref(nops)))
    return n, ops[:nops.value].copy()
