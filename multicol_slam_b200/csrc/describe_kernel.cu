// describe_kernel.cu -- K3: orientation + descriptor, one warp per selected keypoint.
//
//   IC_Angle            ref src/mdBRIEFextractorOct.cpp:221-248   integer moments over the 845-px disc, fastAtan2
//   undistortPointsOcam ref :1306-1317, include/cam_model_omni.h:127-138
//   rotate[AndDistort]Pattern  ref :250-301
//   compute_ORB / compute_dBRIEF / compute_mdBRIEF   ref :303-554
//   output assembly     ref :1327-1335  (pt *= scale for levels > 0)
//
// Layout of the work: lane b of the warp owns descriptor byte b, i.e. the 16 pattern points 16b..16b+15 (8 test
// pairs).  For the distorted variants the lane keeps the 16 projected points of one pattern in registers, the
// warp reduces their sum (the reference subtracts the mean of all 512 projected points, :262-281), then every
// lane rounds and samples its own points: no shuffles, no second projection pass.
//
// Fisheye projection cost.  cCamModelGeneral_::WorldToImg evaluates, per pattern point, sqrt + 3 divisions +
// atan + a 12-term Horner in double (~250 FP64 instructions; 1536 points per mdBRIEF keypoint).  Because the
// third coordinate is the per-camera constant z = -a0, u = x*g(r)*c + y*g(r)*d + u0 with
//     g(r) = R(r) / r,   R(r) = rho(atan(-z / r)),      r = sqrt(x^2 + y^2),
// and R is a smooth 1-D function of r ("the distortion baked into a LUT" of the north star; g itself is not
// tabulated because the fitted inverse polynomial leaves a tiny rho(-pi/2) != 0, i.e. a 1/r pole).  The host
// tabulates R per camera as degree-5 polynomials on 1-px intervals fitted in long double (error < 1e-13 px, the
// rounding noise of the reference's own double evaluation); the device evaluates rsqrt + 5 FMA + the affine map.
// Only cvRound(u - mean) has to agree with the reference; a projected coordinate closer than 1e-7 px to a
// rounding tie (p ~ 2e-4 per pattern) makes the warp recompute that pattern with the reference's exact
// operation sequence (cam_model.cuh).  Radii outside the table take the exact path as well.
#include <cstdlib>
#include <vector>

#include "cam_model.cuh"
#include "kernels.h"
#include "mcs_common.cuh"

namespace mcs {

__constant__ signed char c_pairs[2048];          // learned_pattern_64_ORB (ref include/mdBRIEFextractorOct.h:44-47)
__constant__ signed char c_disc_u[848], c_disc_v[848];   // c_disc_u[0..16] = umax[] of the IC_Angle disc (ref :187-202); rest unused

// cv::fastAtan2, evaluated without FMA
__device__ __forceinline__ float fast_atan2_deg(float y, float x) {
    const float s = (float)(180.0 / 3.14159265358979323846);
    const float p1 = 0.9997878412794807f * s, p3 = -0.3258083974640975f * s;
    const float p5 = 0.1555786518463281f * s, p7 = -0.04432655554792128f * s;
    const float ax = fabsf(x), ay = fabsf(y);
    float a, c, c2;
    if (ax >= ay) {
        c = __fdiv_rn(ay, __fadd_rn(ax, (float)2.2204460492503131e-16));
        c2 = __fmul_rn(c, c);
        a = __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c);
    } else {
        c = __fdiv_rn(ax, __fadd_rn(ay, (float)2.2204460492503131e-16));
        c2 = __fmul_rn(c, c);
        a = __fsub_rn(90.f, __fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(__fadd_rn(__fmul_rn(p7, c2), p5), c2), p3), c2), p1), c));
    }
    if (x < 0) a = __fsub_rn(180.f, a);
    if (y < 0) a = __fsub_rn(360.f, a);
    return a;
}

__device__ __forceinline__ int sample_px(const uint8_t* bimg, const uint8_t* uimg, const LevelGeom& g, int row, int col) {
    // blurred ROI; outside the ROI the reference reads the un-blurred REFLECT_101 ring of its buffer
    if ((unsigned)row < (unsigned)g.h && (unsigned)col < (unsigned)g.w) return bimg[(size_t)row * g.pitch + col];
    row = min(max(row, -kEdge), g.h - 1 + kEdge); col = min(max(col, -kEdge), g.w - 1 + kEdge);   // memory-safety clamp
    return uimg[(size_t)reflect101(row, g.h) * g.pitch + reflect101(col, g.w)];
}

constexpr int kDescWarps = 4;
constexpr int kPatchR = 25;                       // staged patch: rows/cols ky/kx -25 .. +25 (keypoints are >= 25 px inside the ROI)
constexpr int kPatchS = 64;                       // bytes per staged patch row (4-byte aligned start + 51 columns)
constexpr int kLutDeg = 9;                        // degree of the per-centre polynomial of R(r)  (kernels.h: DistortLut)
#ifndef MCS_K3_HALVES
#define MCS_K3_HALVES 0
#endif
#ifndef MCS_K3_MINB
#define MCS_K3_MINB 4                // resident CTAs per SM the register budget is cut for (128 registers)
#endif
constexpr int kLutStride = 12;                    // doubles per centre: tau offset, tau scale, kLutDeg + 1 coefficients
constexpr double kLutReach = 22.5;                // half-width of a centre's interval: pattern radius 15*sqrt(2) + 0.5 + margin

// Rare path: one pattern of one keypoint with the reference's exact operation sequence (two projection passes;
// per-lane partial sums + butterfly: within ~1e-13 of the reference's sequential sum).
// Returns the descriptor byte(s) of this lane for that pattern, byte bb in bits [8bb, 8bb+8).
template <int PPL>
__device__ __noinline__ unsigned exact_pattern(const mcs_ocam* cam, const char2* s_pat, double ca, double sa, double ukx,
                                               double uky, int lane, int ds, const uint8_t* bimg, const uint8_t* uimg,
                                               const LevelGeom* g, int kx, int ky) {
    const double z = -cam->pol[0];
    double su = 0.0, sv = 0.0;
    for (int j = 0; j < PPL; ++j) {
        const char2 pp = s_pat[j * 32 + lane];
        const double px = (double)pp.x, py = (double)pp.y;
        const double xr = px * ca - py * sa + ukx;
        const double yr = px * sa + py * ca + uky;
        double u, v;
        cam_world_to_img(*cam, xr, yr, z, u, v);
        if (lane + 32 * (j >> 4) < ds) { su += u; sv += v; }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        su += __shfl_xor_sync(0xffffffffu, su, o);
        sv += __shfl_xor_sync(0xffffffffu, sv, o);
    }
    const double mu = su / (double)(16 * ds), mv = sv / (double)(16 * ds);
    unsigned out = 0;
    for (int j = 0; j < PPL; j += 2) {
        int smp[2];
        for (int e = 0; e < 2; ++e) {
            const char2 pp = s_pat[(j + e) * 32 + lane];
            const double px = (double)pp.x, py = (double)pp.y;
            const double xr = px * ca - py * sa + ukx;
            const double yr = px * sa + py * ca + uky;
            double u, v;
            cam_world_to_img(*cam, xr, yr, z, u, v);
            smp[e] = sample_px(bimg, uimg, *g, ky + __double2int_rn(v - mv), kx + __double2int_rn(u - mu));
        }
        out |= (unsigned)(smp[0] < smp[1]) << (j >> 1);
    }
    return out;
}

// ORB rotation (ref :285-301) and generic sampling straight from global memory; used for ORB when an offset leaves
// the staged patch (never for sane inputs) -- keeps the reference's read semantics (blurred ROI / reflected ring).
template <int PPL>
__device__ __noinline__ unsigned orb_pattern_global(const char2* s_pat, double ca, double sa, int lane, const uint8_t* bimg,
                                                    const uint8_t* uimg, const LevelGeom* g, int kx, int ky) {
    unsigned out = 0;
    for (int j = 0; j < PPL; j += 2) {
        int smp[2];
        for (int e = 0; e < 2; ++e) {
            const char2 pp = s_pat[(j + e) * 32 + lane];
            const double px = (double)pp.x, py = (double)pp.y;
            smp[e] = sample_px(bimg, uimg, *g, ky + __double2int_rn(px * sa + py * ca), kx + __double2int_rn(px * ca - py * sa));
        }
        out |= (unsigned)(smp[0] < smp[1]) << (j >> 1);
    }
    return out;
}

template <int PPL /* pattern points per lane: 16 for descSize <= 32, 32 for descSize 64 */, int MINB = MCS_K3_MINB>
__global__ void __launch_bounds__(kDescWarps * 32, MINB)
describe_kernel(const PyramidGeom* __restrict__ geom, const DescribeArgs args, const mcs_ocam* __restrict__ cams,
                const DistortLut* __restrict__ luts, const int* __restrict__ cam_of_image,
                const uint32_t* __restrict__ sel_xys, const int* __restrict__ sel_count,
                mcs_keypoint* __restrict__ kps_out, uint8_t* __restrict__ desc_out, uint8_t* __restrict__ dmask_out,
                int* __restrict__ counts_out, const int capacity, const int n_images) {
    __shared__ char2 s_pat[PPL * 32];            // [j][lane] : point 16*byte + k with byte = lane + 32*(j/16), k = j%16
    __shared__ __align__(16) double2 s_patd[PPL * 32];   // same, as doubles (int->double conversions run on the slow XU pipe)
    __shared__ mcs_ocam s_cam[kDescWarps];
#if MCS_K3_HALVES
    __shared__ int2 s_park[kDescWarps][PPL == 16 ? PPL * 32 : 1];    // projected coordinates of the current pattern, [point][lane]
#endif
    __shared__ __align__(16) uint8_t s_patch[kDescWarps][(2 * kPatchR + 1) * kPatchS];
    const int ds = geom->desc_size;
    for (int i = threadIdx.x; i < PPL * 32; i += blockDim.x) {
        const int j = i >> 5, ln = i & 31;
        const int byte = ln + 32 * (j >> 4), pt = 16 * byte + (j & 15);
        s_pat[i] = byte < ds ? make_char2(c_pairs[2 * pt], c_pairs[2 * pt + 1]) : make_char2(0, 0);
        s_patd[i] = make_double2((double)s_pat[i].x, (double)s_pat[i].y);
    }
    __syncthreads();

    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
    // one warp per slot (a persistent grid-stride variant measured 4 % slower: static imbalance + 60 B more spills)
    const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int sel_total = geom->sel_total;
    const int b = warp_global / sel_total;
    if (b >= n_images) return;
    const int slot = warp_global - b * sel_total;
    const int L = geom->nlevels;
    int level = 0, off = 0;
    for (int l = 0; l < L; ++l)
        if (slot >= geom->lv[l].sel_off) level = l;
    for (int l = 0; l < level; ++l) off += sel_count[b * L + l];
    const LevelGeom& g = geom->lv[level];
    const int p = slot - g.sel_off;
    const int cnt = sel_count[b * L + level];
    if (slot == 0 && lane == 0) {
        int tot = 0;
        for (int l = 0; l < L; ++l) tot += sel_count[b * L + l];
        counts_out[b] = min(tot, capacity);
    }
    if (p >= cnt) return;
    const int oidx = off + p;
    if (oidx >= capacity) return;
    const uint32_t c = sel_xys[(size_t)b * sel_total + slot];
    const int kx = corner_x(c), ky = corner_y(c);
    const uint8_t* uimg = args.lvl[level] + (size_t)b * g.img_bytes;
    const uint8_t* bimg = args.blur[level] + (size_t)b * g.img_bytes;
    const int pitch = g.pitch;

    // ---- stage the blurred 51x51 patch: 2 rows per warp instruction, 15 aligned words per row ----
    uint8_t* patch = s_patch[wib];
    const int x0 = (kx - kPatchR) & ~3;                    // >= 0: keypoints lie >= 25 px inside the level
    {
        const int half = lane >> 4, w = lane & 15;
        const uint32_t* gp = (const uint32_t*)(bimg + (size_t)(ky - kPatchR + half) * pitch + x0) + w;
        uint32_t* sp = (uint32_t*)(patch + half * kPatchS) + w;
        // all loads of a batch are issued before the first store (memory-level parallelism)
#pragma unroll
        for (int k0 = 0; k0 < 26; k0 += 13) {
            uint32_t t[13];
#pragma unroll
            for (int k = 0; k < 13; ++k) {
                const int r = half + 2 * (k0 + k);
                t[k] = (w < 15 && r < 2 * kPatchR + 1) ? __ldg(gp + (size_t)(k0 + k) * (pitch / 2)) : 0u;
            }
#pragma unroll
            for (int k = 0; k < 13; ++k) {
                const int r = half + 2 * (k0 + k);
                if (w < 15 && r < 2 * kPatchR + 1) sp[(k0 + k) * (2 * kPatchS / 4)] = t[k];
            }
        }
    }
    const int pofs = kPatchR * kPatchS + (kx - x0);        // patch byte offset of the keypoint itself

    // ---- IC_Angle (ref :221-248): integer moments over the 845-pixel disc, lane = column u = lane-16 (+ u = 16) ----
    int m10 = 0, m01 = 0;
    {
        const uint8_t* ctr = uimg + (size_t)ky * pitch + kx;
        const int u = lane - kHalfPatch;                    // -16 .. 15
        const int au = u < 0 ? -u : u;
        // |v| <= vmax(u): the disc is symmetric (umax table, ref :187-202); vmax(|u|) = umax[|u|]
        const int vm = c_disc_u[au];                        // c_disc_u[0..16] doubles as umax[] (see upload_constants)
#pragma unroll 11
        for (int v = -kHalfPatch; v <= kHalfPatch; ++v) {
            const int val = (v >= -vm && v <= vm) ? (int)ctr[v * pitch + u] : 0;
            m10 += u * val;
            m01 += v * val;
        }
        if (lane < 2 * c_disc_u[kHalfPatch] + 1) {          // column u = +16: rows |v| <= umax[16]
            const int v = lane - c_disc_u[kHalfPatch];
            const int val = ctr[v * pitch + kHalfPatch];
            m10 += kHalfPatch * val;
            m01 += v * val;
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            m10 += __shfl_xor_sync(0xffffffffu, m10, o);
            m01 += __shfl_xor_sync(0xffffffffu, m01, o);
        }
    }
    const float angle = fast_atan2_deg((float)m01, (float)m10);

    const int ci = cam_of_image[b];
    const bool masks = geom->learn_masks != 0, dbrief = geom->do_dbrief != 0 || masks;
    const float scale = g.scale;
    const int npat = masks ? 3 : 1;
    double ca[3], sa[3], a_base, a_rot;
    {
        double a0;
        if (masks) a0 = (double)__fdiv_rn(angle, 57.2957763671875f);              // angle / RHOf      (ref :425)
        else a0 = (double)__fmul_rn(angle, 0.01745329238474369f);                   // angle * DEG2RADf  (ref :313,367)
        const double rot = 20.0 / (180.0 / 3.1415926535897932384626433832795);      // 20 / RHOd         (ref :424)
        sincos(a0, &sa[0], &ca[0]);
        a_base = a0; a_rot = rot;
        // angle +- 20 deg by the addition theorems (1e-16 away from cos/sin(a0 +- rot); the exact path below uses the
        // reference's own cos(angle +- rot) / sin(angle +- rot))
        const double c20 = 0.93969262078590838405, s20 = 0.34202014332566873304;
        ca[1] = fma(ca[0], c20, -sa[0] * s20); sa[1] = fma(sa[0], c20, ca[0] * s20);
        ca[2] = fma(ca[0], c20, sa[0] * s20);  sa[2] = fma(sa[0], c20, -ca[0] * s20);
    }
    constexpr int BPL = PPL / 16;                 // descriptor bytes per lane
    unsigned val[3][BPL];
    __syncwarp();                                 // patch staged
    if (!dbrief) {
        // ---- ORB: rotatePattern (ref :285-301) ----
        int ix[PPL], iy[PPL];
        bool far = false;
#pragma unroll
        for (int j = 0; j < PPL; ++j) {
            const char2 pp = s_pat[j * 32 + lane];
            const double px = (double)pp.x, py = (double)pp.y;
            ix[j] = __double2int_rn(px * ca[0] - py * sa[0]);
            iy[j] = __double2int_rn(px * sa[0] + py * ca[0]);
            far |= (unsigned)(ix[j] + kPatchR) > 2u * kPatchR || (unsigned)(iy[j] + kPatchR) > 2u * kPatchR;
        }
        if (__any_sync(0xffffffffu, far)) {
            const unsigned e = orb_pattern_global<PPL>(s_pat, ca[0], sa[0], lane, bimg, uimg, &g, kx, ky);
#pragma unroll
            for (int bb = 0; bb < BPL; ++bb) val[0][bb] = (e >> (8 * bb)) & 0xFFu;
        } else {
#pragma unroll
            for (int bb = 0; bb < BPL; ++bb) {
                unsigned v = 0;
#pragma unroll
                for (int bit = 0; bit < 8; ++bit) {
                    const int j0 = 16 * bb + 2 * bit;
                    const int s0 = patch[pofs + iy[j0] * kPatchS + ix[j0]], s1 = patch[pofs + iy[j0 + 1] * kPatchS + ix[j0 + 1]];
                    v |= (unsigned)(s0 < s1) << bit;
                }
                val[0][bb] = v;
            }
        }
    } else {
        // ---- dBRIEF / mdBRIEF: rotateAndDistortPattern (ref :250-283) ----
        if (lane < (int)(sizeof(mcs_ocam) / 8)) ((double*)&s_cam[wib])[lane] = ((const double*)&cams[ci])[lane];
        __syncwarp();
        const mcs_ocam& cam = s_cam[wib];
        const DistortLut lut = luts[ci];
        double ukx, uky;   // undistortPointsOcam(pt.x*scale, pt.y*scale, a0)  (ref :1306-1317)
        cam_undistort(cam, (double)__fmul_rn((float)kx, scale), (double)__fmul_rn((float)ky, scale), cam.pol[0], ukx, uky);
        // R(r) around this keypoint: one degree-9 polynomial in tau = (r - m)/hw, valid for every pattern point
        // (|r - rk| <= 21.3), picked by the keypoint's own undistorted radius; the coefficients live in registers
        const double rk = sqrt(ukx * ukx + uky * uky);
        const bool have_lut = rk < (double)(lut.n - 1);           // false for NaN as well -> exact path
        double P[kLutDeg + 1], t_off, t_scale;
        {
            const double2* cp = (const double2*)(lut.coef + (size_t)(have_lut ? __double2int_rn(rk) : 0) * kLutStride);
            const double2 h = __ldg(cp);
            t_off = h.x; t_scale = h.y;
#pragma unroll
            for (int k = 0; k < (kLutDeg + 1) / 2; ++k) {
                const double2 cc = __ldg(cp + 1 + k);
                P[2 * k] = cc.x; P[2 * k + 1] = cc.y;
            }
        }
        const double inv_n = 1.0 / (double)(16 * ds);
        const bool lane_valid = (PPL == 32) || (lane < ds);      // descSize 16: lanes 16..31 own no byte
        for (int q = 0; q < npat; ++q) {
#if MCS_K3_HALVES
            if constexpr (PPL == 16) {
            // Variant (compile-time, not the default): the 16 (32) points of a lane are processed in groups of 8.
            // Pass 1 projects a group and parks its coordinates -- relative to the keypoint's own pixel, as 24-bit fixed point: the
            // low word of (x + 1.5*2^28) is rn(x * 2^24) for |x| < 128 -- in shared memory; pass 2 reads them back, subtracts the
            // mean, rounds, and does the bit tests of the group.  Live set and code size are half of the unrolled 16-point form.
            // Error budget: two roundings of 2^-25 px + polynomial (< 2e-8) << the 4.8e-7 px tie guard (16 units below).
            constexpr double kMagicF = 402653184.0;
            constexpr int GRP = 8;
            int2* park = s_park[wib];
            double su = 0.0, sv = 0.0;
            bool need_exact = false;
            int worst_tau = 0, worst_rng = 0;          // high words of max |tau| and of max |relative coordinate|
            const double du0 = cam.u0 - (double)__fmul_rn((float)kx, scale), dv0 = cam.v0 - (double)__fmul_rn((float)ky, scale);
#pragma unroll 1
            for (int g0 = 0; g0 < PPL; g0 += GRP) {
#pragma unroll
                for (int jj = 0; jj < GRP; ++jj) {
                    const int j = g0 + jj;
                    const double2 pp = s_patd[j * 32 + lane];
                    const double xr = fma(pp.x, ca[q], fma(-pp.y, sa[q], ukx));
                    const double yr = fma(pp.x, sa[q], fma(pp.y, ca[q], uky));
                    const double s2 = fma(xr, xr, yr * yr);
                    double y0;
                    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(s2));
                    const double e = fma(-(s2 * y0), y0, 1.0);
                    const double rinv = fma(y0 * e, fma(0.375, e, 0.5), y0), r = s2 * rinv;
                    const double tau = fma(r, t_scale, t_off);
                    double gg = P[kLutDeg];
#pragma unroll
                    for (int k = kLutDeg - 1; k >= 0; --k) gg = fma(gg, tau, P[k]);
                    worst_tau = max(worst_tau, __double2hiint(tau) & 0x7fffffff);
                    gg *= rinv;
                    const double uu = xr * gg, vv = yr * gg;
                    const double ur = fma(uu, cam.c, fma(vv, cam.d, du0));
                    const double vr = fma(uu, cam.e, vv + dv0);
                    if (lane_valid) { su += ur; sv += vr; }
                    worst_rng = max(worst_rng, max(__double2hiint(ur) & 0x7fffffff, __double2hiint(vr) & 0x7fffffff));
                    park[j * 32 + lane] = make_int2(__double2loint(ur + kMagicF), __double2loint(vr + kMagicF));
                }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                su += __shfl_xor_sync(0xffffffffu, su, o);
                sv += __shfl_xor_sync(0xffffffffu, sv, o);
            }
            // every |ur|, |vr| < 32 (checked below) -> |mean| < 32: the differences below stay far inside 32 bits.
            // t = (x - mean) * 2^24 + 2^23 + 8: t >> 24 is rn(x - mean) unless the low 24 bits are < 16, i.e. x - mean lies
            // within 8 * 2^-24 = 4.8e-7 px of a rounding tie (then the exact path decides).
            const int cu = __double2loint(su * inv_n + kMagicF) - ((1 << 23) + 8), cv = __double2loint(sv * inv_n + kMagicF) - ((1 << 23) + 8);
            unsigned worst_tie = 0xFFFFFFFFu, worst_ofs = 0;
            unsigned bits[BPL];
#pragma unroll
            for (int bb = 0; bb < BPL; ++bb) bits[bb] = 0;
#pragma unroll 1
            for (int g0 = 0; g0 < PPL; g0 += GRP) {
                int ix[GRP], iy[GRP];
#pragma unroll
                for (int jj = 0; jj < GRP; ++jj) {
                    const int2 f = park[(g0 + jj) * 32 + lane];     // this lane's own slot: no synchronisation needed
                    const int tu = f.x - cu, tv = f.y - cv;
                    worst_tie = min(worst_tie, min((unsigned)tu & 0xFFFFFFu, (unsigned)tv & 0xFFFFFFu));
                    ix[jj] = tu >> 24; iy[jj] = tv >> 24;
                    worst_ofs = max(worst_ofs, max((unsigned)(ix[jj] + kPatchR), (unsigned)(iy[jj] + kPatchR)));
                    // keep the gathers inside the staged patch even when this pattern is going to be redone exactly
                    ix[jj] = min(max(ix[jj], -kPatchR), kPatchR); iy[jj] = min(max(iy[jj], -kPatchR), kPatchR);
                }
                unsigned v = 0;
#pragma unroll
                for (int bit = 0; bit < GRP / 2; ++bit) {
                    const int s0 = patch[pofs + iy[2 * bit] * kPatchS + ix[2 * bit]], s1 = patch[pofs + iy[2 * bit + 1] * kPatchS + ix[2 * bit + 1]];
                    v |= (unsigned)(s0 < s1) << bit;
                }
                // group g0 holds bits (g0 % 16) / 2 .. +3 of byte g0 / 16
                if (BPL == 1) bits[0] |= v << ((g0 & 15) >> 1);
                else { if (g0 < 16) bits[0] |= v << ((g0 & 15) >> 1); else bits[BPL - 1] |= v << ((g0 & 15) >> 1); }
            }
            need_exact |= !have_lut || (lane_valid && (worst_tie < 16u || worst_ofs > 2u * kPatchR || worst_rng >= __double2hiint(32.0) ||
                                                       worst_tau >= __double2hiint(1.0)));
            if (__any_sync(0xffffffffu, need_exact)) {
                const double aq = q == 0 ? a_base : (q == 1 ? a_base + a_rot : a_base - a_rot);
                const unsigned e = exact_pattern<PPL>(&cam, s_pat, cos(aq), sin(aq), ukx, uky, lane, ds, bimg, uimg, &g, kx, ky);
#pragma unroll
                for (int bb = 0; bb < BPL; ++bb) val[q][bb] = (e >> (8 * bb)) & 0xFFu;
                continue;
            }
#pragma unroll
            for (int bb = 0; bb < BPL; ++bb) val[q][bb] = bits[bb];
            } else
#endif
            {
            double us[PPL], vs[PPL];
            double su = 0.0, sv = 0.0;
            bool need_exact = false;
            int worst_tau = 0;                         // high word of max |tau|
            // round-to-nearest-even through the 1.5*2^52 trick: no F2I/I2F (XU pipe), same result as lrint
            constexpr double kMagic = 6755399441055744.0;
#pragma unroll
            for (int j = 0; j < PPL; ++j) {
                const double2 pp = s_patd[j * 32 + lane];
                const double xr = fma(pp.x, ca[q], fma(-pp.y, sa[q], ukx));
                const double yr = fma(pp.x, sa[q], fma(pp.y, ca[q], uky));
                const double s2 = fma(xr, xr, yr * yr);
                // 1/sqrt(s2): hardware approximation (~1e-7) + one third-order step -> < 1e-16 relative
                double y0;
                asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(s2));
                const double e = fma(-(s2 * y0), y0, 1.0);
                const double rinv = fma(y0 * e, fma(0.375, e, 0.5), y0), r = s2 * rinv;
                // R(r) by Horner; |tau| > 1 (a point outside the fitted interval, or NaN from s2 == 0) is caught below
                const double tau = fma(r, t_scale, t_off);
                double gg = P[kLutDeg];
#pragma unroll
                for (int k = kLutDeg - 1; k >= 0; --k) gg = fma(gg, tau, P[k]);
                worst_tau = max(worst_tau, __double2hiint(tau) & 0x7fffffff);
                gg *= rinv;
                const double uu = xr * gg, vv = yr * gg;
                us[j] = fma(uu, cam.c, fma(vv, cam.d, cam.u0));
                vs[j] = fma(uu, cam.e, vv + cam.v0);
                if (lane_valid) { su += us[j]; sv += vs[j]; }
            }
#pragma unroll
            for (int o = 16; o; o >>= 1) {
                su += __shfl_xor_sync(0xffffffffu, su, o);
                sv += __shfl_xor_sync(0xffffffffu, sv, o);
            }
            const double mu = su * inv_n, mv = sv * inv_n;
            int ix[PPL], iy[PPL];
            // Closeness to a rounding tie and the patch range are tracked as integer maxima: the high word of |frac|
            // orders like the double itself (non-negative), a NaN / huge value has a larger high word than any fraction.
            int worst_frac = 0;
            unsigned worst_ofs = 0;
#pragma unroll
            for (int j = 0; j < PPL; ++j) {
                const double du = us[j] - mu, dv = vs[j] - mv;
                const double tu = du + kMagic, tv = dv + kMagic;
                ix[j] = __double2loint(tu); iy[j] = __double2loint(tv);
                const int hu = __double2hiint(du - (tu - kMagic)) & 0x7fffffff, hv = __double2hiint(dv - (tv - kMagic)) & 0x7fffffff;
                worst_frac = max(worst_frac, max(hu, hv));
                worst_ofs = max(worst_ofs, max((unsigned)(ix[j] + kPatchR), (unsigned)(iy[j] + kPatchR)));
            }
            // closer than ~7e-7 px to a rounding tie (high word of 0.5 - 5e-7), or outside the staged patch -> exact path
            // (inside the table window |u| is bounded by the fitted polynomial, so the magic-number rounding cannot alias;
            //  a NaN shows up as a huge high word of the fraction)
            need_exact |= !have_lut || (lane_valid && (worst_frac >= __double2hiint(0.5 - 5e-7) || worst_ofs > 2u * kPatchR ||
                                                       worst_tau >= __double2hiint(1.0)));
            if (__any_sync(0xffffffffu, need_exact)) {
                const double aq = q == 0 ? a_base : (q == 1 ? a_base + a_rot : a_base - a_rot);
                const unsigned e = exact_pattern<PPL>(&cam, s_pat, cos(aq), sin(aq), ukx, uky, lane, ds, bimg, uimg, &g, kx, ky);
#pragma unroll
                for (int bb = 0; bb < BPL; ++bb) val[q][bb] = (e >> (8 * bb)) & 0xFFu;
                continue;
            }
#pragma unroll
            for (int bb = 0; bb < BPL; ++bb) {
                unsigned v = 0;
#pragma unroll
                for (int bit = 0; bit < 8; ++bit) {
                    const int j0 = 16 * bb + 2 * bit;
                    const int s0 = patch[pofs + iy[j0] * kPatchS + ix[j0]], s1 = patch[pofs + iy[j0 + 1] * kPatchS + ix[j0 + 1]];
                    v |= (unsigned)(s0 < s1) << bit;
                }
                val[q][bb] = v;
            }
            }
        }
    }
#pragma unroll
    for (int bb = 0; bb < BPL; ++bb) {
        const int byte = lane + 32 * bb;
        if (byte < ds) {
            desc_out[((size_t)b * capacity + oidx) * ds + byte] = (uint8_t)val[0][bb];
            if (dmask_out) {
                // stable bit <=> both +-20 degree re-tests agree with the bit (ref :449-451 ...)
                const unsigned m = masks ? (~((val[1][bb] ^ val[0][bb]) | (val[2][bb] ^ val[0][bb])) & 0xFFu) : 0u;
                dmask_out[((size_t)b * capacity + oidx) * ds + byte] = (uint8_t)m;
            }
        }
    }
    if (lane == 0) {
        mcs_keypoint k;
        k.x = level ? __fmul_rn((float)kx, scale) : (float)kx;      // pt *= scale for l > 0 (ref :1327-1332)
        k.y = level ? __fmul_rn((float)ky, scale) : (float)ky;
        k.size = g.patch_size; k.angle = angle; k.response = (float)corner_s(c);
        k.octave = level; k.class_id = -1;
        kps_out[(size_t)b * capacity + oidx] = k;
    }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
cudaError_t upload_constants(const signed char* pairs, const signed char* du, const signed char* dv) {
    cudaError_t e = cudaMemcpyToSymbol(c_pairs, pairs, 2048);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(c_disc_u, du, 848);
    if (e != cudaSuccess) return e;
    return cudaMemcpyToSymbol(c_disc_v, dv, 848);
}

// Tabulate R(r) = rho(atan(-z/r)) on [0, n) px in unit intervals: 6 monomial coefficients in
// tau = 2 (r - idx) - 1, interpolating R at the 6 Chebyshev nodes of the interval, all in long double.
void build_distort_lut(const mcs_ocam& cam, std::vector<double>& coef, int& n_out) {
    // table extent: largest undistorted radius of any pixel the mirror mask keeps, plus the pattern reach
    const double scaleF = cam.pol[0];
    double rmax = 64.0;
    const double mcx = cam.u0, mcy = cam.v0, mrad = cam.v0 + 22.0;
    for (int y = 0; y < cam.height; y += 4)
        for (int x = 0; x < cam.width; x += 4) {
            if (cam.mirror_mask == 1 && std::hypot(x - mcx, y - mcy) > mrad + 4.0) continue;
            double wx, wy, wz;
            cam_img_to_world(cam, (double)x, (double)y, wx, wy, wz);
            if (!(wz * scaleF < 0.0)) continue;            // behind / on the undistortion plane: no finite image
            const double ux = -wx / wz * scaleF, uy = -wy / wz * scaleF;
            const double r = std::hypot(ux, uy);
            if (std::isfinite(r) && r > rmax) rmax = r;
        }
    int n = (int)std::min(8192.0, std::ceil(rmax + 64.0));
    n_out = n;
    coef.assign((size_t)n * kLutStride, 0.0);
    const long double z = -(long double)cam.pol[0];
    auto R_exact = [&](long double r) {
        const long double theta = atanl(-z / r);
        long double rho = 0.0L;
        for (int t = 11; t >= 0; --t) rho = rho * theta + (long double)cam.inv_pol[t];
        return rho;
    };
    constexpr int NC = kLutDeg + 1;
    long double nodes[NC];
    for (int k = 0; k < NC; ++k) nodes[k] = cosl((2 * k + 1) * 3.14159265358979323846264338327950288L / (2.0L * NC));
    for (int i = 0; i < n; ++i) {
        // centre i serves keypoints with rn(rk) == i: their pattern points lie in [i - 21.8, i + 21.8] and r >= 0
        const long double lo = std::max(0.0L, (long double)i - (long double)kLutReach), hi = (long double)i + (long double)kLutReach;
        const long double m = 0.5L * (lo + hi), hw = 0.5L * (hi - lo);
        long double A[NC][NC + 1];
        for (int k = 0; k < NC; ++k) {
            long double pw = 1.0L;
            for (int t = 0; t < NC; ++t) { A[k][t] = pw; pw *= nodes[k]; }
            A[k][NC] = R_exact(m + hw * nodes[k]);
        }
        for (int col = 0; col < NC; ++col) {          // Gauss-Jordan with partial pivoting
            int piv = col;
            for (int r2 = col + 1; r2 < NC; ++r2) if (fabsl(A[r2][col]) > fabsl(A[piv][col])) piv = r2;
            for (int t = 0; t <= NC; ++t) std::swap(A[col][t], A[piv][t]);
            for (int r2 = 0; r2 < NC; ++r2) {
                if (r2 == col) continue;
                const long double f = A[r2][col] / A[col][col];
                for (int t = col; t <= NC; ++t) A[r2][t] -= f * A[col][t];
            }
        }
        double* e = coef.data() + (size_t)i * kLutStride;
        const double scale = (double)(1.0L / hw);
        e[0] = (double)(-m / hw); e[1] = scale;
        for (int t = 0; t < NC; ++t) e[2 + t] = (double)(A[t][NC] / A[t][t]);
        // check the fit as the kernel evaluates it (double Horner); a centre whose error could move a rounding
        // decision past the kernel's 5e-7 px tie guard is disabled (NaN coefficients -> exact path)
        double worst = 0.0;
        for (int sidx = 0; sidx <= 96; ++sidx) {
            const double r = (double)(lo + (hi - lo) * ((long double)sidx + 0.37L) / 97.0L);
            const double tau = std::fma(r, e[1], e[0]);
            double gv = e[2 + kLutDeg];
            for (int t = kLutDeg - 1; t >= 0; --t) gv = std::fma(gv, tau, e[2 + t]);
            worst = std::max(worst, (double)fabsl((long double)gv - R_exact((long double)r)));
        }
        if (!(worst < 2e-8))
            for (int t = 0; t < kLutStride; ++t) e[t] = std::nan("");
    }
}

cudaError_t launch_describe(const PyramidGeom& G, const PyramidGeom* G_dev, int n_images, const DescribeArgs& args,
                            const mcs_ocam* cams, const DistortLut* luts, const int* cam_of_image, const uint32_t* sel_xys,
                            const int* sel_count, mcs_keypoint* kps, uint8_t* desc, uint8_t* dmask, int* counts, int capacity,
                            cudaStream_t st) {
    const long long warps = (long long)n_images * G.sel_total;
    const int blocks = (int)std::max<long long>(1, (warps + kDescWarps - 1) / kDescWarps);
    // 128 registers (4 CTAs of 4 warps per SM): measured faster than 96 / 80 registers with more warps (spills), see DESIGN.md
    if (G.desc_size <= 32)
        describe_kernel<16><<<blocks, kDescWarps * 32, 0, st>>>(G_dev, args, cams, luts, cam_of_image, sel_xys, sel_count, kps, desc,
                                                               dmask, counts, capacity, n_images);
    else
        describe_kernel<32><<<blocks, kDescWarps * 32, 0, st>>>(G_dev, args, cams, luts, cam_of_image, sel_xys, sel_count, kps, desc,
                                                               dmask, counts, capacity, n_images);
    return cudaGetLastError();
}

}  // namespace mcs
