"""CPU: the per-camera distortion tables behind tiers 1 and 2 of the descriptor kernel (mcs_cam_distort_table, built on the host in
multicol_slam_b200/csrc/describe_kernel.cu: build_distort_lut) checked against the camera model itself, independently of the kernel,
in numpy long double:  R(r) = rho(atan(-z / r)),  rho = the inverse polynomial of cCamModelGeneral_::WorldToImg
(ref src/cam_model_omni.cpp:49-67), z = -a0.  The kernel's rounding decisions rest on these error bounds:
  tier 2   degree-9 polynomial of R on [i - 22.5, i + 22.5]                      |error| < 2e-8 px
  tier 1   s-form  R(i + s) - R(i) = s' q(s'), s' = s / 32                       |error| < 1e-6 px
           m-form  G(c_i + hw_i t) - G(c_i) = t P(t), G(m) = R(sqrt m) / sqrt m  |error| (i + 1) < 1e-6 px
with the float coefficients exactly as the kernel reads them."""
import numpy as np
import pytest

LD = np.longdouble
REACH = 22.5


def horner(c, x):
    r = np.zeros_like(x)
    for k in range(len(c) - 1, -1, -1):
        r = r * x + LD(c[k])
    return r


def R_exact(cam, r):
    z = -LD(cam["pol"][0])
    return horner([LD(v) for v in cam["inv_pol"]], np.arctan(-z / r))


def floats_of(row, first, count):
    return np.frombuffer(np.ascontiguousarray(row[first:first + (count + 1) // 2]).tobytes(), np.float32)[:count].astype(LD)


@pytest.mark.parametrize("which", [0, 1, 2, "hd"])
def test_tables_against_the_camera_model(api, cams, which):
    from multicol_slam_b200 import synth
    cam = synth.scaled_cam(cams[1], 1920, 1080) if which == "hd" else cams[which]
    T = api.distort_table(cam)
    n, stride = T.shape
    assert stride == 24 and 200 < n <= 8192
    t = np.linspace(-1, 1, 41).astype(LD) * LD(0.999)
    n2 = n1s = n1m = 0
    for i in list(range(0, 130)) + list(range(130, n, 7)):
        row = T[i]
        lo, hi = max(LD(0), LD(i) - LD(REACH)), LD(i) + LD(REACH)
        if not np.isnan(row[0]):                                   # ---- tier 2
            r = (lo + hi) / 2 + (hi - lo) / 2 * t
            r = r[r > 0]
            tau = r * LD(row[1]) + LD(row[0])
            err = np.abs(horner(row[2:12], tau) - R_exact(cam, r)).max()
            assert err < 2.5e-8, (i, float(err))
            n2 += 1
        if row[17] == 1.0:                                         # ---- tier 1, s-form
            assert i >= 39 and not np.isnan(row[0])
            s = LD(REACH) * t
            s = s[np.abs(s) > 1e-6]
            sp = s / LD(32)
            q = np.concatenate([[LD(row[13])], floats_of(row, 14, 5)])
            err = np.abs(sp * horner(q, sp) - (R_exact(cam, LD(i) + s) - R_exact(cam, np.array([LD(i)]))[0])).max()
            assert abs(float(R_exact(cam, np.array([LD(i)]))[0]) - row[12]) < 1e-12 * max(1.0, abs(row[12]))
            assert err < 1.2e-6, (i, float(err))
            n1s += 1
        if row[23] == 1.0:                                         # ---- tier 1, m-form
            assert row[17] == 1.0 and i > REACH + 1
            c, hw = LD(i) * LD(i) + LD(REACH) ** 2, LD(2 * REACH) * LD(i)
            G = lambda m: R_exact(cam, np.sqrt(m)) / np.sqrt(m)    # noqa: E731
            tt = t[np.abs(t) > 1e-6]
            a = floats_of(row, 19, 8)
            err = np.abs(tt * horner(a, tt) - (G(c + hw * tt) - G(np.array([c]))[0])).max()
            assert abs(float(G(np.array([c]))[0]) - row[18]) < 1e-14
            assert err * (i + 1) < 1.2e-6, (i, float(err * (i + 1)))
            n1m += 1
    # coverage: tier 2 everywhere, the s-form from radius 40..54 on and the m-form from 62..67 on (Lafida cameras), never below 40
    assert n2 > 100 and n1s > 80 and n1m > 60
    assert not np.isnan(T[:, 0]).any()
    assert all(T[i, 17] == 1.0 for i in range(56, n - 1)) and all(T[i, 23] == 1.0 for i in range(72, n - 1))
    assert not any(T[i, 17] == 1.0 or T[i, 23] == 1.0 for i in range(0, 39))


def test_table_argument_validation(api, cams):
    import ctypes as C
    lib = api.lib()
    n, row = C.c_int32(0), C.c_int32(0)
    assert lib.mcs_cam_distort_table(None, None, 0, C.byref(n), C.byref(row)) == api.MCS_ERR_INVALID
    oc = api.as_ocam(cams[0])
    assert lib.mcs_cam_distort_table(C.byref(oc), None, 0, None, C.byref(row)) == api.MCS_ERR_INVALID
    assert lib.mcs_cam_distort_table(C.byref(oc), None, 0, C.byref(n), C.byref(row)) == api.MCS_OK and n.value > 200 and row.value == 24
