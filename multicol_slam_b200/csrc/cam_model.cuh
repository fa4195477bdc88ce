// cam_model.cuh -- Scaramuzza/OCam projection, the exact double arithmetic of the reference
// (ref src/cam_model_omni.cpp:49-67, :146-161; include/cam_model_omni.h:127-145; include/misc.h:115-122).
// Host and device share this code; build with -fmad=false / -ffp-contract=off.
#pragma once
#include <math.h>
#include "../../include/mcs_b200.h"

#ifdef __CUDACC__
#define MCS_HD __host__ __device__ __forceinline__
#else
#define MCS_HD inline
#endif

namespace mcs {

MCS_HD double cam_horner(const double* c, int n, double x) {
    double r = 0.0;
    for (int i = n - 1; i >= 0; --i) r = r * x + c[i];
    return r;
}

MCS_HD void cam_world_to_img(const mcs_ocam& cam, double x, double y, double z, double& u, double& v) {
    double norm = sqrt(x * x + y * y);
    if (norm == 0.0) norm = 1e-14;
    const double theta = atan(-z / norm);
    const double rho = cam_horner(cam.inv_pol, 12, theta);
    const double uu = x / norm * rho;
    const double vv = y / norm * rho;
    u = uu * cam.c + vv * cam.d + cam.u0;
    v = uu * cam.e + vv + cam.v0;
}

MCS_HD void cam_img_to_world(const mcs_ocam& cam, double u, double v, double& x, double& y, double& z) {
    const double inv_affine = cam.c - cam.d * cam.e;
    const double u_t = u - cam.u0;
    const double v_t = v - cam.v0;
    x = (u_t - cam.d * v_t) / inv_affine;
    y = (-cam.e * u_t + cam.c * v_t) / inv_affine;
    const double X2 = x * x, Y2 = y * y;
    z = -cam_horner(cam.pol, 5, sqrt(X2 + Y2));
    const double norm = sqrt(X2 + Y2 + z * z);
    x /= norm; y /= norm; z /= norm;
}

MCS_HD void cam_undistort(const mcs_ocam& cam, double px, double py, double s, double& ox, double& oy) {
    double x, y, z;
    cam_img_to_world(cam, px, py, x, y, z);
    ox = -x / z * s;
    oy = -y / z * s;
}

}  // namespace mcs
