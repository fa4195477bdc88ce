"""The reference's configuration files as inputs of this library (data formats on the caller's side of the path).

  read_opencv_yaml(path)            flat "key: value" maps as cv::FileStorage reads them (YAML 1.0 header, '#' comments)
  load_ocam(path)                   Camera.* interior orientation -> camera dict / mcs_ocam          (ref src/cSystem.cpp:146-170)
  load_rig(directory)               MultiCamSys_Calibration.yaml + InteriorOrientationFisheye<c>.yaml -> (M_c list, cameras)
                                    with cayley2hom (ref src/cSystem.cpp:125-180, include/misc.h:133-224)
  rig_matrices(M_t, M_c)            MtMc[c] = M_t * M_c[c] and its rigid inverse, the arrays mcs_project_mappoints takes
                                    (ref src/cam_system_omni.cpp, cConverter::invMat src/cConverter.cpp:31-44)
  compute_E(T1, T2)                 essential matrix for mcs_search_for_triangulation              (ref src/misc.cpp:71-85)
  extractor_settings(path)          extractor.* keys -> keyword arguments of the tracking extractor and of the initialisation
                                    extractor (2 x nFeatures, FAST threshold 5)                  (ref src/cTracking.cpp:108-159)
Host arithmetic only (double, same operation order as the reference's cv::Matx expressions)."""
import math
import pathlib
import re

import numpy as np


def read_opencv_yaml(path):
    out = {}
    for line in pathlib.Path(path).read_text(encoding="latin-1").splitlines():
        line = line.split("#", 1)[0].strip()
        if not line or line.startswith("%") or line == "---":
            continue
        m = re.match(r"^([A-Za-z_][\w.]*)\s*:\s*(.*)$", line)
        if not m:
            continue
        key, val = m.group(1), m.group(2).strip().strip('"')
        try:
            out[key] = int(val)
        except ValueError:
            try:
                out[key] = float(val)
            except ValueError:
                out[key] = val
    return out


def load_ocam(path):
    kv = read_opencv_yaml(path)
    nrpol, nrinv = int(kv["Camera.nrpol"]), int(kv["Camera.nrinvpol"])
    if nrpol > 5 or nrinv > 12:
        raise ValueError("the camera model holds at most 5 forward and 12 inverse polynomial coefficients (ref src/cSystem.cpp:150-155)")
    return dict(c=float(kv["Camera.c"]), d=float(kv["Camera.d"]), e=float(kv["Camera.e"]), u0=float(kv["Camera.u0"]), v0=float(kv["Camera.v0"]),
                pol=[float(kv[f"Camera.a{i}"]) for i in range(nrpol)], inv_pol=[float(kv[f"Camera.pol{i}"]) for i in range(nrinv)],
                width=int(kv["Camera.Iw"]), height=int(kv["Camera.Ih"]), mirror_mask=int(kv.get("Camera.mirrorMask", 0)))


def cayley2rot(c1, c2, c3):
    """include/misc.h:133-160"""
    c1s, c2s, c3s = c1 * c1, c2 * c2, c3 * c3
    scale = 1.0 + c1s + c2s + c3s
    R = np.array([[1 + c1s - c2s - c3s, 2 * (c1 * c2 - c3), 2 * (c1 * c3 + c2)],
                  [2 * (c1 * c2 + c3), 1 - c1s + c2s - c3s, 2 * (c2 * c3 - c1)],
                  [2 * (c1 * c3 - c2), 2 * (c2 * c3 + c1), 1 - c1s - c2s + c3s]], np.float64)
    return (1 / scale) * R


def cayley2hom(p6):
    """include/misc.h:211-224"""
    M = np.eye(4)
    M[:3, :3] = cayley2rot(float(p6[0]), float(p6[1]), float(p6[2]))
    M[:3, 3] = [float(p6[3]), float(p6[4]), float(p6[5])]
    return M


def load_rig(directory):
    """-> (M_c [n_cams, 4, 4], list of camera dicts)"""
    d = pathlib.Path(directory)
    kv = read_opencv_yaml(d / "MultiCamSys_Calibration.yaml")
    n = int(kv["CameraSystem.nrCams"])
    M_c = np.stack([cayley2hom([kv[f"CameraSystem.cam{c + 1}_{p}"] for p in range(1, 7)]) for c in range(n)])
    cams = [load_ocam(d / f"InteriorOrientationFisheye{c}.yaml") for c in range(n)]
    return M_c, cams


def _matmul(a, b):
    """cv::Matx product: s = 0; s += a(i,k) * b(k,j) in index order"""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    out = np.zeros((a.shape[0], b.shape[1]))
    for i in range(a.shape[0]):
        for j in range(b.shape[1]):
            s = 0.0
            for k in range(a.shape[1]):
                s += a[i, k] * b[k, j]
            out[i, j] = s
    return out


def inv_rigid(M):
    """cConverter::invMat (src/cConverter.cpp:31-44): [R t]^-1 = [R^T  -R^T t]"""
    M = np.asarray(M, np.float64)
    Rt = M[:3, :3].T.copy()
    t = -_matmul(Rt, M[:3, 3:4])[:, 0]
    out = np.eye(4)
    out[:3, :3], out[:3, 3] = Rt, t
    return out


def rig_matrices(M_t, M_c):
    """-> (mtmc [n,4,4], mtmc_inv [n,4,4]) for the pose M_t of the rig"""
    mtmc = np.stack([_matmul(M_t, M_c[c]) for c in range(len(M_c))])
    return mtmc, np.stack([inv_rigid(m) for m in mtmc])


def skew(v):
    return np.array([[0.0, -v[2], v[1]], [v[2], 0.0, -v[0]], [-v[1], v[0], 0.0]])


def compute_E(T1, T2):
    """ComputeE(T1, T2) (src/misc.cpp:71-85); SearchForTriangulationRaw passes Get_MtMc_inv(i) of KF1 and Get_MtMc(j) of KF2"""
    T1, T2 = np.asarray(T1, np.float64), np.asarray(T2, np.float64)
    R1, R2, t1, t2 = T1[:3, :3], T2[:3, :3], T1[:3, 3], T2[:3, 3]
    R12 = _matmul(R1, R2.T)
    t12 = _matmul(_matmul(-R1, R2.T), t2.reshape(3, 1))[:, 0] + t1
    t12 = t12 / math.sqrt(t12[0] * t12[0] + t12[1] * t12[1] + t12[2] * t12[2])
    return _matmul(skew(t12), R12)


def essential_matrices(mtmc_inv_1, mtmc_2):
    """Es[i][j] of SearchForTriangulationRaw (src/cORBmatcher.cpp:988-1001)"""
    n = len(mtmc_inv_1)
    return np.stack([np.stack([compute_E(mtmc_inv_1[i], mtmc_2[j]) for j in range(n)]) for i in range(n)])


def extractor_settings(path):
    """-> (kwargs of the tracking extractor, kwargs of the initialisation extractor) for api.mdBRIEFextractorOct
    (src/cTracking.cpp:108-159: edgeThreshold 25, firstLevel 0, patchSize 32 are hard-wired there)."""
    kv = read_opencv_yaml(path)
    score = int(kv["extractor.nScoreType"])
    if score not in (0, 1):
        raise ValueError("extractor.nScoreType must be 0 or 1")
    desc = int(kv["extractor.descSize"])
    if desc not in (16, 32, 64):
        raise ValueError("extractor.descSize must be 16, 32 or 64")
    base = dict(scaleFactor=float(np.float32(kv["extractor.scaleFactor"])), nlevels=int(kv["extractor.nLevels"]), edgeThreshold=25, firstLevel=0,
                scoreType=score, patchSize=32, useAgast=bool(int(kv["extractor.useAgast"])), fastAgastType=int(kv["extractor.fastAgastType"]),
                do_dBrief=bool(int(kv["extractor.usemdBRIEF"])), learnMasks=bool(int(kv["extractor.masks"])), descSize=desc)
    nf = int(kv["extractor.nFeatures"])
    track = dict(base, nfeatures=nf, fastThreshold=int(kv["extractor.fastTh"]))
    init = dict(base, nfeatures=2 * nf, fastThreshold=5)
    return track, init
