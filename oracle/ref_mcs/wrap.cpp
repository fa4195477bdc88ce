// wrap.cpp -- C entry points around the REFERENCE'S OWN extractor / camera model, compiled from
//     /root/reference/src/mdBRIEFextractorOct.cpp, src/cam_model_omni.cpp, src/misc.cpp
// where they lie (oracle/Makefile target `ref`) against the stand-in OpenCV header in stub/.   TEST INFRASTRUCTURE.
// Nothing here restates reference logic: each function builds the reference's objects the way src/cSystem.cpp:144-170 and
// src/cTracking.cpp:150-159 do and calls the reference's methods.  The structs are those of include/mcs_b200.h so that the
// oracle restatement, the CUDA path and this library are driven with identical arguments.
#define protected public       // test access to mvImagePyramid / mvMaskPyramid / DistributeOctTree of the reference class
#include "mdBRIEFextractorOct.h"
#undef protected
#include "cam_model_omni.h"
#include "misc.h"

#include "../../include/mcs_b200.h"

using namespace MultiColSLAM;

// ---- deterministic heap for the duration of one reference call -----------------------------------------------------------
// DistributeOctTree sorts pair<int, ExtractorNode*> (src/mdBRIEFextractorOct.cpp:782): equally sized nodes are ordered by the
// HEAP ADDRESS of their std::list node, so the reference's keypoint selection depends on the allocator's state (two runs of the
// stock build on the same image can return different keypoints).  While a reference call runs, every allocation made by this
// library comes from a bump arena that is rewound at the start of the call and never reuses memory inside it: addresses grow
// with creation order, so "ties by pointer" is exactly "the node created later is expanded first" -- the rule the oracle
// restatement and the CUDA octree implement -- and the reference becomes reproducible.  (-Wl,-Bsymbolic binds this library's
// own operator new/delete references to these definitions; memory that did not come from the arena is freed normally.)
#include <atomic>
#include <new>
#include <sys/mman.h>
namespace arena {
static const size_t kCap = (size_t)3 << 30;                 // virtual reservation; pages are touched on demand
static char* base = nullptr;
static std::atomic<size_t> off(0);
static std::atomic<int> on(0);
static std::atomic<int> deterministic(1);     // 0: plain malloc (re-entrant; the timing arm of bench.py runs many extractors at once)
static void ensure() {
    if (!base) {
        void* p = mmap(nullptr, kCap, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (p == MAP_FAILED) throw std::bad_alloc();
        base = (char*)p;
    }
}
static inline bool owns(const void* p) { return base && (const char*)p >= base && (const char*)p < base + kCap; }
struct Scope {
    bool active;
    Scope() : active(deterministic.load() != 0) { if (active) { ensure(); off.store(0); on.store(1); } }
    ~Scope() { if (active) on.store(0); }
};
}  // namespace arena
void* operator new(size_t n) {
    if (arena::on.load(std::memory_order_relaxed)) {
        const size_t a = (n + 15) & ~(size_t)15;
        const size_t o = arena::off.fetch_add(a);
        if (o + a <= arena::kCap) return arena::base + o;
    }
    void* p = std::malloc(n ? n : 1);
    if (!p) throw std::bad_alloc();
    return p;
}
void* operator new[](size_t n) { return operator new(n); }
void operator delete(void* p) noexcept { if (p && !arena::owns(p)) std::free(p); }
void operator delete[](void* p) noexcept { operator delete(p); }
void operator delete(void* p, size_t) noexcept { operator delete(p); }
void operator delete[](void* p, size_t) noexcept { operator delete(p); }

namespace {
// cCamModelGeneral_ built like cSystem::LoadMCS (src/cSystem.cpp:144-170): 5x1 / 12x1 polynomials, mirror masks from the
// reference's CreateMirrorMask (4 levels) or all ones
cCamModelGeneral_ make_cam(const mcs_ocam* c) {
    cv::Mat_<double> poly = cv::Mat::zeros(5, 1, CV_64F);
    for (int i = 0; i < 5; ++i) poly.at<double>(i, 0) = c->pol[i];
    cv::Mat_<double> invpoly = cv::Mat::zeros(12, 1, CV_64F);
    for (int i = 0; i < 12; ++i) invpoly.at<double>(i, 0) = c->inv_pol[i];
    double cdeu0v0[5] = {c->c, c->d, c->e, c->u0, c->v0};
    cCamModelGeneral_ cam(cdeu0v0, poly, invpoly, c->width, c->height);
    std::vector<cv::Mat> masks;
    if (c->mirror_mask == 1) CreateMirrorMask(cam, 4, masks);
    else masks.push_back(cv::Mat::ones(cv::Size(c->width, c->height), CV_8UC1));
    cam.SetMirrorMasks(masks);
    return cam;
}
}  // namespace

struct mcsref_extractor {
    mdBRIEFextractorOct* ex;
    mcs_extractor_params p;
};

extern "C" {

mcsref_extractor* mcsref_extractor_create(const mcs_extractor_params* p) {
    mcsref_extractor* h = new mcsref_extractor;
    h->p = *p;
    h->ex = new mdBRIEFextractorOct(p->nfeatures, p->scale_factor, p->nlevels, p->edge_threshold, p->first_level, p->score_type,
                                    p->patch_size, p->fast_threshold, p->use_agast != 0, p->fast_agast_type, p->do_dbrief != 0,
                                    p->learn_masks != 0, p->desc_size);
    return h;
}
void mcsref_extractor_destroy(mcsref_extractor* h) { if (h) { delete h->ex; delete h; } }

// 1 (default): monotonic heap during a call -> reproducible pointer tie-break, one call at a time.  0: the stock allocator --
// the extractor is then re-entrant (one instance per thread) and its tie-break is whatever the heap gives, like the stock build.
void mcsref_set_deterministic(int on) { arena::deterministic.store(on ? 1 : 0); }

int mcsref_extractor_tables(mcsref_extractor* h, int* quotas, double* sf, double* isf, int* umax17) {
    for (int l = 0; l < h->p.nlevels; ++l) {
        quotas[l] = h->ex->mnFeaturesPerLevel[l]; sf[l] = h->ex->mvScaleFactor[l]; isf[l] = h->ex->mvInvScaleFactor[l];
    }
    for (int i = 0; i < 17; ++i) umax17[i] = h->ex->umax[i];
    return 0;
}

// mdBRIEFextractorOct::operator() of the reference on one image (src/mdBRIEFextractorOct.cpp:1244-1337)
int mcsref_extract(mcsref_extractor* h, const uint8_t* image, int w, int hgt, int stride, const uint8_t* mask, int mstride,
                   const mcs_ocam* cam_in, mcs_keypoint* kps, uint8_t* desc, uint8_t* dmask, int capacity, int* n_out) {
    try {
        // buffers of the previous call live in the arena: let go of them before it is rewound
        for (auto& m : h->ex->mvImagePyramid) m = cv::Mat();
        for (auto& m : h->ex->mvMaskPyramid) m = cv::Mat();
        arena::Scope heap;
        cCamModelGeneral_ cam = make_cam(cam_in);
        cv::Mat img(hgt, w, CV_8UC1, (void*)image, (size_t)stride);
        cv::Mat msk(hgt, w, CV_8UC1, (void*)mask, (size_t)mstride);
        std::vector<cv::KeyPoint> keys;
        cv::Mat d, m;
        (*h->ex)(img, msk, keys, cam, d, m);
        const int n = (int)keys.size(), ds = h->p.desc_size;
        *n_out = n;
        if (n > capacity) return MCS_ERR_CAPACITY;
        static_assert(sizeof(cv::KeyPoint) == sizeof(mcs_keypoint), "cv::KeyPoint layout");
        for (int i = 0; i < n; ++i) {
            std::memcpy(&kps[i], &keys[i], sizeof(mcs_keypoint));
            std::memcpy(desc + (size_t)i * ds, d.ptr<uint8_t>(i), (size_t)ds);
            if (dmask) std::memcpy(dmask + (size_t)i * ds, m.ptr<uint8_t>(i), (size_t)ds);
        }
        return MCS_OK;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "mcsref_extract: %s\n", e.what());
        return MCS_ERR_INVALID;
    }
}

// intermediates of the last call: what = 0 image level (ROI; blurred if the level had keypoints, like the reference leaves
// it), 1 mask level, 2 image level including the 25-px ring
int mcsref_debug_read(mcsref_extractor* h, int level, int what, uint8_t* out, size_t out_bytes, int* w_out, int* h_out) {
    if (level < 0 || level >= h->p.nlevels) return MCS_ERR_INVALID;
    const cv::Mat& m = what == 1 ? h->ex->mvMaskPyramid[level] : h->ex->mvImagePyramid[level];
    if (m.empty()) return MCS_ERR_INVALID;
    const int ring = what == 2 ? 25 : 0;
    const int w = m.cols + 2 * ring, hh = m.rows + 2 * ring;
    *w_out = w; *h_out = hh;
    if ((size_t)w * hh > out_bytes) return MCS_ERR_CAPACITY;
    for (int y = 0; y < hh; ++y) std::memcpy(out + (size_t)y * w, m.parentPtr(y - ring, -ring), (size_t)w);
    return MCS_OK;
}

// DistributeOctTree of the reference (src/mdBRIEFextractorOct.cpp:631-861) on a caller-given corner list (x, y, response)
int mcsref_octree(mcsref_extractor* h, const float* xyr, int n, int minX, int maxX, int minY, int maxY, int N, float* out, int cap) {
    for (auto& m : h->ex->mvImagePyramid) m = cv::Mat();
    for (auto& m : h->ex->mvMaskPyramid) m = cv::Mat();
    arena::Scope heap;
    std::vector<cv::KeyPoint> in((size_t)n);
    for (int i = 0; i < n; ++i) in[i] = cv::KeyPoint(xyr[3 * i], xyr[3 * i + 1], 7.f, -1, xyr[3 * i + 2]);
    std::vector<cv::KeyPoint> r = h->ex->DistributeOctTree(in, minX, maxX, minY, maxY, N, 0);
    if ((int)r.size() > cap) return -1;
    for (size_t i = 0; i < r.size(); ++i) { out[3 * i] = r[i].pt.x; out[3 * i + 1] = r[i].pt.y; out[3 * i + 2] = r[i].response; }
    return (int)r.size();
}

// camera model of the reference (src/cam_model_omni.cpp)
void mcsref_cam_world_to_img(const mcs_ocam* c, double x, double y, double z, double* u, double* v) {
    make_cam(c).WorldToImg(x, y, z, *u, *v);
}
void mcsref_cam_img_to_world(const mcs_ocam* c, double u, double v, double* x, double* y, double* z) {
    make_cam(c).ImgToWorld(*x, *y, *z, u, v);
}
void mcsref_cam_undistort(const mcs_ocam* c, double px, double py, double* ox, double* oy) {
    cCamModelGeneral_ cam = make_cam(c);
    cam.undistortPointsOcam(px, py, cam.Get_P().at<double>(0), *ox, *oy);
}
int mcsref_cam_mirror_mask(const mcs_ocam* c, uint8_t* out) {
    cv::Mat m = make_cam(c).GetMirrorMask(0);
    for (int y = 0; y < m.rows; ++y) std::memcpy(out + (size_t)y * m.cols, m.ptr<uint8_t>(y), (size_t)m.cols);
    return MCS_OK;
}
// batched isPointInMirrorMask(u, v, 0) (src/cam_model_omni.cpp:163-178)
void mcsref_cam_points_in_mask(const mcs_ocam* c, const double* uv, int n, uint8_t* out) {
    cCamModelGeneral_ cam = make_cam(c);
    for (int i = 0; i < n; ++i) out[i] = cam.isPointInMirrorMask(uv[2 * i], uv[2 * i + 1], 0) ? 1 : 0;
}

// src/misc.cpp
int mcsref_check_epipolar(const double* ray1, const double* ray2, const double* E, double thresh) {
    cv::Matx33d Em;
    for (int i = 0; i < 9; ++i) Em.val[i] = E[i];
    return CheckDistEpipolarLine(cv::Vec3d(ray1[0], ray1[1], ray1[2]), cv::Vec3d(ray2[0], ray2[1], ray2[2]), Em, thresh) ? 1 : 0;
}
void mcsref_compute_E(const double* T1, const double* T2, double* E) {
    cv::Matx44d a, b;
    for (int i = 0; i < 16; ++i) { a.val[i] = T1[i]; b.val[i] = T2[i]; }
    cv::Matx33d e = ComputeE(a, b);
    for (int i = 0; i < 9; ++i) E[i] = e.val[i];
}
double mcsref_const(int which) { return which == 0 ? (double)RHOf : RHOd; }

// ---- the stand-in's own OpenCV primitives, exported so that they can be pinned against the real cv2 ----
void mcsref_cv_resize(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh, int interpolation) {
    cv::Mat s(sh, sw, CV_8UC1, (void*)src), d(dh, dw, CV_8UC1, (void*)dst);
    cv::resize(s, d, cv::Size(dw, dh), 0, 0, interpolation);
}
// copyMakeBorder of the whole image with `b` pixels; reflect != 0 -> BORDER_REFLECT_101 else BORDER_CONSTANT(0)
void mcsref_cv_make_border(const uint8_t* src, int w, int h, int b, int reflect, uint8_t* dst) {
    cv::Mat s(h, w, CV_8UC1, (void*)src), d(h + 2 * b, w + 2 * b, CV_8UC1, (void*)dst);
    cv::copyMakeBorder(s, d, b, b, b, b, reflect ? cv::BORDER_REFLECT_101 : cv::BORDER_CONSTANT);
}
// boxFilter 5x5 in place on the ROI (x0, y0, w, h) of a `bw` x `bh` buffer (the reference's use, :1301), or isolated
void mcsref_cv_box5(uint8_t* buf, int bw, int bh, int x0, int y0, int w, int h) {
    cv::Mat whole(bh, bw, CV_8UC1, (void*)buf);
    cv::Mat roi = whole(cv::Rect(x0, y0, w, h));
    cv::boxFilter(roi, roi, roi.depth(), cv::Size(5, 5), cv::Point(-1, -1), true, cv::BORDER_REFLECT_101);
}
float mcsref_cv_fast_atan2(float y, float x) { return cv::fastAtan2(y, x); }
// FastFeatureDetector::create(threshold, true, TYPE_9_16)->detect(image ROI, keypoints, mask ROI); returns the count
int mcsref_cv_fast(const uint8_t* img, int w, int h, int stride, const uint8_t* mask, int mstride, int threshold, float* xyr, int cap) {
    cv::Mat im(h, w, CV_8UC1, (void*)img, (size_t)stride), mk;
    if (mask) mk = cv::Mat(h, w, CV_8UC1, (void*)mask, (size_t)mstride);
    std::vector<cv::KeyPoint> k;
    cv::FastFeatureDetector::create(threshold, true, 2)->detect(im, k, mk);
    if ((int)k.size() > cap) return -1;
    for (size_t i = 0; i < k.size(); ++i) { xyr[3 * i] = k[i].pt.x; xyr[3 * i + 1] = k[i].pt.y; xyr[3 * i + 2] = k[i].response; }
    return (int)k.size();
}
int mcsref_cv_round(double v) { return cvRound(v); }

}  // extern "C"
