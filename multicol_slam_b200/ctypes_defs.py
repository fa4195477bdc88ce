"""ctypes mirrors of the POD types in include/mcs_b200.h (shared by the product wrapper and by the
test-only oracle wrapper)."""
import ctypes as C

MAX_LEVELS = 16


class Ocam(C.Structure):
    _fields_ = [("c", C.c_double), ("d", C.c_double), ("e", C.c_double), ("u0", C.c_double), ("v0", C.c_double),
                ("pol", C.c_double * 5), ("inv_pol", C.c_double * 12),
                ("width", C.c_int32), ("height", C.c_int32), ("mirror_mask", C.c_int32), ("_pad", C.c_int32)]


class ExtractorParams(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("edge_threshold", C.c_int32), ("first_level", C.c_int32), ("score_type", C.c_int32),
                ("patch_size", C.c_int32), ("fast_threshold", C.c_int32), ("use_agast", C.c_int32),
                ("fast_agast_type", C.c_int32), ("do_dbrief", C.c_int32), ("learn_masks", C.c_int32),
                ("desc_size", C.c_int32)]


class ExtractorInfo(C.Structure):
    _fields_ = [("nlevels", C.c_int32), ("capacity", C.c_int32), ("desc_size", C.c_int32),
                ("features_per_level", C.c_int32 * MAX_LEVELS),
                ("scale_factor", C.c_double * MAX_LEVELS), ("inv_scale_factor", C.c_double * MAX_LEVELS)]


class FrameView(C.Structure):
    _fields_ = [("n_cams", C.c_int32), ("n_keys", C.c_int32), ("keys", C.c_void_p), ("key_cam", C.c_void_p),
                ("desc", C.c_void_p), ("dmask", C.c_void_p), ("cam_width", C.c_void_p), ("cam_height", C.c_void_p),
                ("dim", C.c_int32), ("n_levels", C.c_int32), ("scale_factors", C.c_void_p)]


class WindowQuery(C.Structure):
    _fields_ = [("cam", C.c_int32), ("min_level", C.c_int32), ("max_level", C.c_int32), ("desc_index", C.c_int32),
                ("x", C.c_double), ("y", C.c_double), ("r", C.c_double)]


class MapPointView(C.Structure):
    _fields_ = [("n_points", C.c_int32), ("bad", C.c_void_p), ("in_view", C.c_void_p), ("level", C.c_void_p),
                ("proj_x", C.c_void_p), ("proj_y", C.c_void_p), ("view_cos", C.c_void_p), ("desc", C.c_void_p),
                ("dmask", C.c_void_p)]


import numpy as np

# numpy dtype binary-compatible with mcs_keypoint / cv::KeyPoint (28 bytes)
KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"), ("response", "<f4"),
                           ("octave", "<i4"), ("class_id", "<i4")])
WINDOW_QUERY_DTYPE = np.dtype([("cam", "<i4"), ("min_level", "<i4"), ("max_level", "<i4"), ("desc_index", "<i4"),
                               ("x", "<f8"), ("y", "<f8"), ("r", "<f8")])
assert KEYPOINT_DTYPE.itemsize == 28 and WINDOW_QUERY_DTYPE.itemsize == 40


def make_params(nfeatures=1000, scale_factor=1.2, nlevels=8, fast_threshold=20, do_dbrief=False, learn_masks=False,
                desc_size=32, use_agast=False, fast_agast_type=2):
    """Same defaults as the reference constructor (include/mdBRIEFextractorOct.h:340-352)."""
    return ExtractorParams(nfeatures, scale_factor, nlevels, 25, 0, 0, 32, fast_threshold, int(use_agast),
                           fast_agast_type, int(do_dbrief), int(learn_masks), desc_size)


def make_ocam(d):
    o = Ocam()
    o.c, o.d, o.e, o.u0, o.v0 = d["c"], d["d"], d["e"], d["u0"], d["v0"]
    for i in range(5):
        o.pol[i] = d["pol"][i] if i < len(d["pol"]) else 0.0
    for i in range(12):
        o.inv_pol[i] = d["inv_pol"][i] if i < len(d["inv_pol"]) else 0.0
    o.width, o.height, o.mirror_mask = d["width"], d["height"], d.get("mirror_mask", 1)
    return o
