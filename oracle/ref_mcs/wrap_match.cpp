// wrap_match.cpp -- C entry points around the REFERENCE'S OWN matcher, /root/reference/src/cORBmatcher.cpp compiled where it
// lies (oracle/Makefile target `ref`) together with the reference's cam_system_omni.cpp, cam_model_omni.cpp, cConverter.cpp,
// misc.cpp and DBoW2/FeatureVector.cpp.   TEST INFRASTRUCTURE.
// The matcher reads three SLAM container classes that cannot be compiled here; stub_slam.h supplies data-only stand-ins for
// them (see its header).  This file only marshals flat arrays into those containers, calls the reference's cORBmatcher
// methods and flattens the results: no matching logic lives here.
#include "cORBmatcher.h"
#include "cConverter.h"
#include "misc.h"

#include "../../include/mcs_b200.h"

#include <memory>

using namespace MultiColSLAM;

extern "C" {
// a cMultiFrame / cMultiKeyFrame as flat arrays
typedef struct mcsref_kf {
    mcs_frame_view view;          // keys, key_cam, desc, dmask, camera sizes, dim, levels, scale factors
    const double* rays;           // [n_keys*3] bearing rays (mvKeysRays) or NULL
    const int32_t* mp;            // [n_keys] index of the map point of each keypoint, -1 = none (mvpMapPoints) or NULL
    const uint8_t* outlier;       // [n_keys] mvbOutlier or NULL
    const mcs_ocam* cams;         // [n_cams] interior orientations
    const double* M_c;            // [n_cams*16] row-major camera-to-MCS transformations
    const double* M_t;            // [16] MCS pose
    int32_t fv_n;                 // DBoW2 feature vector as CSR: node ids, offsets [fv_n+1], feature indices
    const int32_t* fv_nodes;
    const int32_t* fv_offsets;
    const int32_t* fv_features;
} mcsref_kf;
// the map-point table shared by the frames of one call
typedef struct mcsref_mps {
    int32_t n, n_cams, dim;
    const uint8_t* bad;
    const double* world_pos;      // [n*3]
    const double* normal;         // [n*3] or NULL
    const double* min_dist;       // [n]
    const double* max_dist;
    const uint8_t* desc;          // [n*dim]
    const uint8_t* dmask;         // [n*dim] or NULL
    const uint8_t* in_view;       // tracking fields [n*n_cams], all NULL when unused
    const int32_t* level;
    const double* proj_x;
    const double* proj_y;
    const double* view_cos;
    const int32_t* obs_kf;        // [n] -1 none / 0 / 1: the point is observed in the first / second key-frame argument (or NULL)
    const int32_t* obs_idx;       // [n] keypoint index there
} mcsref_mps;
}

namespace {
cCamModelGeneral_ make_cam(const mcs_ocam* c) {                        // as src/cSystem.cpp:144-170
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
cv::Matx44d m44(const double* p) { cv::Matx44d m; for (int i = 0; i < 16; ++i) m.val[i] = p[i]; return m; }

struct World {
    std::vector<std::unique_ptr<cMapPoint>> mps;
    RefMutationLog log;
    cMapPoint* mp(int i) { return i >= 0 && i < (int)mps.size() ? mps[i].get() : nullptr; }

    void build_points(const mcsref_mps* t, cMultiKeyFrame* kf0, cMultiKeyFrame* kf1) {
        if (!t) return;
        const int w = t->dim / 8;
        for (int i = 0; i < t->n; ++i) {
            std::unique_ptr<cMapPoint> p(new cMapPoint);
            p->id = i; p->log = &log;
            p->bad = t->bad && t->bad[i];
            if (t->world_pos) p->worldPos = cv::Vec3d(t->world_pos[3 * i], t->world_pos[3 * i + 1], t->world_pos[3 * i + 2]);
            if (t->normal) p->normal = cv::Vec3d(t->normal[3 * i], t->normal[3 * i + 1], t->normal[3 * i + 2]);
            if (t->min_dist) { p->minDist = t->min_dist[i]; p->maxDist = t->max_dist[i]; }
            p->desc.assign(w, 0); p->dmask.assign(w, 0);
            if (t->desc) std::memcpy(p->desc.data(), t->desc + (size_t)i * t->dim, t->dim);
            if (t->dmask) std::memcpy(p->dmask.data(), t->dmask + (size_t)i * t->dim, t->dim);
            if (t->in_view) {
                const int nc = t->n_cams;
                p->mbTrackInView.resize(nc); p->mnTrackScaleLevel.resize(nc); p->mTrackViewCos.resize(nc);
                p->mTrackProjX.resize(nc); p->mTrackProjY.resize(nc);
                for (int c = 0; c < nc; ++c) {
                    p->mbTrackInView[c] = t->in_view[(size_t)i * nc + c] != 0; p->mnTrackScaleLevel[c] = t->level[(size_t)i * nc + c];
                    p->mTrackViewCos[c] = t->view_cos[(size_t)i * nc + c];
                    p->mTrackProjX[c] = t->proj_x[(size_t)i * nc + c]; p->mTrackProjY[c] = t->proj_y[(size_t)i * nc + c];
                }
            }
            if (t->obs_kf && t->obs_kf[i] >= 0) {
                cMultiKeyFrame* k = t->obs_kf[i] == 0 ? kf0 : kf1;
                if (k) p->obs[k].push_back((size_t)t->obs_idx[i]);
            }
            mps.push_back(std::move(p));
        }
    }

    template <class F> void build_frame(F& f, const mcsref_kf* k) {
        const mcs_frame_view& v = k->view;
        const int nc = v.n_cams, n = v.n_keys, dim = v.dim;
        std::vector<cv::Matx44d> Mc(nc);
        std::vector<cCamModelGeneral_> cams;
        for (int c = 0; c < nc; ++c) { Mc[c] = m44(k->M_c + 16 * c); cams.push_back(make_cam(&k->cams[c])); }
        f.camSystem = cMultiCamSys_(m44(k->M_t), Mc, cams);
        f.mvKeys.resize(n); f.mvKeysRays.resize(n);
        std::vector<int> cnt(nc, 0), w(nc), h(nc);
        for (int c = 0; c < nc; ++c) { w[c] = v.cam_width[c]; h[c] = v.cam_height[c]; }
        for (int i = 0; i < n; ++i) {
            static_assert(sizeof(cv::KeyPoint) == sizeof(mcs_keypoint), "cv::KeyPoint layout");
            std::memcpy(&f.mvKeys[i], &v.keys[i], sizeof(mcs_keypoint));
            if (k->rays) f.mvKeysRays[i] = cv::Vec3d(k->rays[3 * i], k->rays[3 * i + 1], k->rays[3 * i + 2]);
            const int c = v.key_cam[i];
            f.keypoint_to_cam[i] = c;                                   // src/cMultiFrame.cpp:173-174
            f.cont_idx_to_local_cam_idx[i] = cnt[c]++;
        }
        f.mDescriptors.resize(nc); f.mDescriptorMasks.resize(nc);
        // rows [0, cnt[c]) are camera c's descriptors; the matrices are allocated n + 1 rows tall and zero beyond that, so that
        // the reference's SearchByProjection(pKF, Scw, ...) -- which indexes a camera's matrix with the CONTIGUOUS keypoint id
        // (src/cORBmatcher.cpp:2367,2372) and therefore reads past the real matrix for every camera but the first -- stays
        // inside allocated memory (its result there is undefined in the reference; tests only use the defined range)
        for (int c = 0; c < nc; ++c) {
            f.mDescriptors[c] = cv::Mat::zeros(n + 1, dim, CV_8UC1);
            f.mDescriptorMasks[c] = cv::Mat::zeros(n + 1, dim, CV_8UC1);
        }
        std::vector<int> row(nc, 0);
        for (int i = 0; i < n; ++i) {
            const int c = v.key_cam[i], r = row[c]++;
            std::memcpy(f.mDescriptors[c].template ptr<uint8_t>(r), v.desc + (size_t)i * dim, dim);
            if (v.dmask) std::memcpy(f.mDescriptorMasks[c].template ptr<uint8_t>(r), v.dmask + (size_t)i * dim, dim);
        }
        f.mvpMapPoints.assign(n, nullptr);
        if (k->mp) for (int i = 0; i < n; ++i) f.mvpMapPoints[i] = mp(k->mp[i]);
        f.mnScaleLevels = v.n_levels;
        f.mvScaleFactors.assign(v.scale_factors, v.scale_factors + v.n_levels);
        for (int j = 0; j < k->fv_n; ++j)
            for (int q = k->fv_offsets[j]; q < k->fv_offsets[j + 1]; ++q) f.mFeatVec.addFeature(k->fv_nodes[j], k->fv_features[q]);
        f.grid.build(f.mvKeys, f.keypoint_to_cam, w, h);
    }
    void build(cMultiFrame& f, const mcsref_kf* k) {
        build_frame(f, k);
        f.mvbOutlier.assign(k->view.n_keys, false);
        if (k->outlier) for (int i = 0; i < k->view.n_keys; ++i) f.mvbOutlier[i] = k->outlier[i] != 0;
    }
    void build(cMultiKeyFrame& f, const mcsref_kf* k) { build_frame(f, k); }

    template <class V> void ids_out(const V& v, int32_t* out) { for (size_t i = 0; i < v.size(); ++i) out[i] = v[i] ? v[i]->id : -1; }
    int log_out(int32_t* ops, int cap) {
        const int n = (int)log.ops.size();
        for (int i = 0; i < n && i < cap; ++i) { ops[3 * i] = log.ops[i].kind; ops[3 * i + 1] = log.ops[i].mp; ops[3 * i + 2] = log.ops[i].other_or_idx; }
        return n;
    }
};
}  // namespace

#define GUARD(...) try { __VA_ARGS__ } catch (const std::exception& e) { std::fprintf(stderr, "mcsref matcher: %s\n", e.what()); return -1000; }

extern "C" {

int mcsref_thresholds(int feat_dim, int having_masks, int* th_high, int* th_low) {
    cORBmatcher m(0.6, false, feat_dim, having_masks != 0);
    *th_high = m.TH_HIGH_; *th_low = m.TH_LOW_;
    return 0;
}
int mcsref_descriptor_distance64(const uint64_t* a, const uint64_t* b, int dim) { return DescriptorDistance64(a, b, dim); }
int mcsref_descriptor_distance64_masked(const uint64_t* a, const uint64_t* b, const uint64_t* ma, const uint64_t* mb, int dim) {
    return DescriptorDistance64Masked(a, b, ma, mb, dim);
}

// GetFeaturesInArea of the stand-in frames (restated in stub_slam.h; exported so that tests can pin it against the independent
// restatements in oracle/mcs_oracle.cpp and oracle/pyref_match.py).  keyframe != 0 -> the cMultiKeyFrame overload.
int mcsref_features_in_area(const mcsref_kf* k, int keyframe, int cam, double x, double y, double r, int min_level, int max_level,
                            int32_t* out, int cap) {
    GUARD(
        World w;
        std::vector<size_t> v;
        if (keyframe) { cMultiKeyFrame f; w.build(f, k); v = f.GetFeaturesInArea(cam, x, y, r); }
        else { cMultiFrame f; w.build(f, k); v = f.GetFeaturesInArea(cam, x, y, r, min_level, max_level); }
        for (size_t i = 0; i < v.size() && (int)i < cap; ++i) out[i] = (int32_t)v[i];
        return (int)v.size();
    )
}

// cORBmatcher::SearchByProjection(cMultiFrame&, vector<cMapPoint*>&, th)   src/cORBmatcher.cpp:67-166
int mcsref_search_by_projection(const mcsref_kf* F, const mcsref_mps* mps, double th, double nnratio, int having_masks, int32_t* frame_mp) {
    GUARD(
        World w; w.build_points(mps, nullptr, nullptr);
        cMultiFrame f; w.build(f, F);
        std::vector<cMapPoint*> v;
        for (int i = 0; i < mps->n; ++i) v.push_back(w.mp(i));
        cORBmatcher m(nnratio, checkOrientation, F->view.dim, having_masks != 0);
        const int n = m.SearchByProjection(f, v, th);
        w.ids_out(f.mvpMapPoints, frame_mp);
        return n;
    )
}

// SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize)   :579-726
int mcsref_search_for_initialization(const mcsref_kf* F1, const mcsref_kf* F2, double* prev_matched, int window, double nnratio,
                                     int having_masks, int32_t* matches12) {
    GUARD(
        World w;
        cMultiFrame f1, f2; w.build(f1, F1); w.build(f2, F2);
        std::vector<cv::Vec2d> prev(F1->view.n_keys);
        for (int i = 0; i < F1->view.n_keys; ++i) prev[i] = cv::Vec2d(prev_matched[2 * i], prev_matched[2 * i + 1]);
        std::vector<int> m12;
        cORBmatcher m(nnratio, checkOrientation, F1->view.dim, having_masks != 0);
        const int n = m.SearchForInitialization(f1, f2, prev, m12, window);
        for (size_t i = 0; i < m12.size(); ++i) matches12[i] = m12[i];
        for (int i = 0; i < F1->view.n_keys; ++i) { prev_matched[2 * i] = prev[i](0); prev_matched[2 * i + 1] = prev[i](1); }
        return n;
    )
}

// SearchByBoW(KF1, KF2, vpMatches12)   :885-966.  out[i1] = map point id matched to keypoint i1 of KF1 (-1 none)
int mcsref_search_by_bow_kfkf(const mcsref_kf* K1, const mcsref_kf* K2, const mcsref_mps* mps, double nnratio, int having_masks, int32_t* out) {
    GUARD(
        World w; w.build_points(mps, nullptr, nullptr);
        cMultiKeyFrame k1, k2; w.build(k1, K1); w.build(k2, K2);
        std::vector<cMapPoint*> v;
        cORBmatcher m(nnratio, checkOrientation, K1->view.dim, having_masks != 0);
        const int n = m.SearchByBoW(&k1, &k2, v);
        w.ids_out(v, out);
        return n;
    )
}

// SearchByBoW(KF, F, vpMapPointMatches)   :179-324.  out[iF] = map point id assigned to frame keypoint iF
int mcsref_search_by_bow_kff(const mcsref_kf* K, const mcsref_kf* F, const mcsref_mps* mps, double nnratio, int having_masks, int32_t* out) {
    GUARD(
        World w; w.build_points(mps, nullptr, nullptr);
        cMultiKeyFrame k; w.build(k, K);
        cMultiFrame f; w.build(f, F);
        std::vector<cMapPoint*> v;
        cORBmatcher m(nnratio, checkOrientation, K->view.dim, having_masks != 0);
        const int n = m.SearchByBoW(&k, f, v);
        w.ids_out(v, out);
        return n;
    )
}

// WindowSearch(F1, F2, windowSize, vpMapPointMatches2, minOctave, maxOctave)   :326-474
int mcsref_window_search(const mcsref_kf* F1, const mcsref_kf* F2, const mcsref_mps* mps, int window, int min_level, int max_level,
                         double nnratio, int having_masks, int32_t* out2) {
    GUARD(
        World w; w.build_points(mps, nullptr, nullptr);
        cMultiFrame f1, f2; w.build(f1, F1); w.build(f2, F2);
        std::vector<cMapPoint*> v;
        cORBmatcher m(nnratio, checkOrientation, F1->view.dim, having_masks != 0);
        const int n = m.WindowSearch(f1, f2, window, v, min_level, max_level);
        w.ids_out(v, out2);
        return n;
    )
}

// SearchByProjection(F1, F2, windowSize, vpMapPointMatches2)   :476-577
int mcsref_search_by_projection_frames(const mcsref_kf* F1, const mcsref_kf* F2, const mcsref_mps* mps, int window, double nnratio,
                                       int having_masks, int32_t* out2) {
    GUARD(
        World w; w.build_points(mps, nullptr, nullptr);
        cMultiFrame f1, f2; w.build(f1, F1); w.build(f2, F2);
        std::vector<cMapPoint*> v((size_t)F2->view.n_keys, nullptr);
        if (F2->mp) for (int i = 0; i < F2->view.n_keys; ++i) v[i] = w.mp(F2->mp[i]);
        cORBmatcher m(nnratio, checkOrientation, F1->view.dim, having_masks != 0);
        const int n = m.SearchByProjection(f1, f2, window, v);
        w.ids_out(v, out2);
        return n;
    )
}

// SearchByProjection(CurrentFrame, LastFrame, th)   :1990-2118
int mcsref_search_by_projection_last(const mcsref_kf* Cur, const mcsref_kf* Last, const mcsref_mps* mps, double th, double nnratio,
                                     int having_masks, int32_t* cur_mp) {
    GUARD(
        World w; w.build_points(mps, nullptr, nullptr);
        cMultiFrame cur, last; w.build(cur, Cur); w.build(last, Last);
        cORBmatcher m(nnratio, checkOrientation, Cur->view.dim, having_masks != 0);
        const int n = m.SearchByProjection(cur, last, th);
        w.ids_out(cur.mvpMapPoints, cur_mp);
        return n;
    )
}

// SearchByProjection(CurrentFrame, pKF, sAlreadyFound, th, ORBdist)   :2120-2263
int mcsref_search_by_projection_reloc(const mcsref_kf* Cur, const mcsref_kf* K, const mcsref_mps* mps, const uint8_t* already_found,
                                      double th, int orb_dist, double nnratio, int having_masks, int32_t* cur_mp) {
    GUARD(
        World w; w.build_points(mps, nullptr, nullptr);
        cMultiFrame cur; w.build(cur, Cur);
        cMultiKeyFrame k; w.build(k, K);
        std::set<cMapPoint*> found;
        if (already_found) for (int i = 0; i < mps->n; ++i) if (already_found[i]) found.insert(w.mp(i));
        cORBmatcher m(nnratio, checkOrientation, Cur->view.dim, having_masks != 0);
        const int n = m.SearchByProjection(cur, &k, found, th, orb_dist);
        w.ids_out(cur.mvpMapPoints, cur_mp);
        return n;
    )
}

// SearchByProjection(pKF, Scw, vpPoints, vpMatched, th)   :2265-2392.  points[np]: map point ids (-1 = NULL entry);
// matched[n_keys]: in/out map point id per keypoint
int mcsref_search_by_projection_scw(const mcsref_kf* K, const mcsref_mps* mps, const double* Scw, const int32_t* points, int np, int th,
                                    double nnratio, int having_masks, int32_t* matched) {
    GUARD(
        World w; w.build_points(mps, nullptr, nullptr);
        cMultiKeyFrame k; w.build(k, K);
        std::vector<cMapPoint*> pts, mt((size_t)K->view.n_keys, nullptr);
        for (int i = 0; i < np; ++i) pts.push_back(w.mp(points[i]));
        for (int i = 0; i < K->view.n_keys; ++i) mt[i] = w.mp(matched[i]);
        cORBmatcher m(nnratio, checkOrientation, K->view.dim, having_masks != 0);
        const int n = m.SearchByProjection(&k, m44(Scw), pts, mt, th);
        w.ids_out(mt, matched);
        return n;
    )
}

// SearchForTriangulationRaw(KF1, KF2, ...)   :968-1156.  pairs [cap*2] (idx1, idx2) in output order
int mcsref_search_for_triangulation_raw(const mcsref_kf* K1, const mcsref_kf* K2, const mcsref_mps* mps, double nnratio, int having_masks,
                                        int32_t* pairs, int cap) {
    GUARD(
        World w; w.build_points(mps, nullptr, nullptr);
        cMultiKeyFrame k1, k2; w.build(k1, K1); w.build(k2, K2);
        std::vector<cv::KeyPoint> a, b; std::vector<cv::Vec3d> ra, rb; std::vector<std::pair<size_t, size_t>> p;
        cORBmatcher m(nnratio, checkOrientation, K1->view.dim, having_masks != 0);
        const int n = m.SearchForTriangulationRaw(&k1, &k2, a, ra, b, rb, p);
        for (size_t i = 0; i < p.size() && (int)i < cap; ++i) { pairs[2 * i] = (int32_t)p[i].first; pairs[2 * i + 1] = (int32_t)p[i].second; }
        return n;
    )
}

// SearchForTriangulationBetweenCameras(KF1, cam1, cam2, ...)   :1158-1263
int mcsref_search_for_triangulation_between(const mcsref_kf* K1, const mcsref_mps* mps, int cam1, int cam2, double nnratio, int having_masks,
                                            int32_t* pairs, int cap) {
    GUARD(
        World w; w.build_points(mps, nullptr, nullptr);
        cMultiKeyFrame k1; w.build(k1, K1);
        std::vector<cv::KeyPoint> a, b; std::vector<cv::Vec3d> ra, rb; std::vector<std::pair<size_t, size_t>> p;
        cORBmatcher m(nnratio, checkOrientation, K1->view.dim, having_masks != 0);
        const int n = m.SearchForTriangulationBetweenCameras(&k1, cam1, cam2, a, ra, b, rb, p);
        for (size_t i = 0; i < p.size() && (int)i < cap; ++i) { pairs[2 * i] = (int32_t)p[i].first; pairs[2 * i + 1] = (int32_t)p[i].second; }
        return n;
    )
}

// SearchBySim3(KF1, KF2, vpMatches12, s12, R12, t12, th)   :1721-1988.  matches12 [n1]: in = already matched map point ids
// (-1 none), out = the reference's vpMatches12
int mcsref_search_by_sim3(const mcsref_kf* K1, const mcsref_kf* K2, const mcsref_mps* mps, double s12, const double* R12, const double* t12,
                          double th, double nnratio, int having_masks, int32_t* matches12) {
    GUARD(
        World w;
        cMultiKeyFrame k1, k2;
        w.build_points(mps, &k1, &k2);
        w.build(k1, K1); w.build(k2, K2);
        std::vector<cMapPoint*> v((size_t)K1->view.n_keys, nullptr);
        for (int i = 0; i < K1->view.n_keys; ++i) v[i] = w.mp(matches12[i]);
        cv::Matx33d R; for (int i = 0; i < 9; ++i) R.val[i] = R12[i];
        cORBmatcher m(nnratio, checkOrientation, K1->view.dim, having_masks != 0);
        const int n = m.SearchBySim3(&k1, &k2, v, s12, R, cv::Vec3d(t12[0], t12[1], t12[2]), th);
        w.ids_out(v, matches12);
        return n;
    )
}

// the three Fuse overloads.  variant 0: Fuse(pKF, curKF, vpMapPoints, th) :1265 (points[i] belongs to keypoint i of curKF);
// 1: Fuse(pKF, vpMapPoints, th) :1420;  2: Fuse(pKF, Scw, vpPoints, th) :1570.  ops [cap*3]: the map mutations in call order
// (kind 0 = AddObservation+AddMapPoint(mp, idx), 1 = Replace(mp -> other)); *n_ops their number.  Returns nFused.
int mcsref_fuse(int variant, const mcsref_kf* K, const mcsref_kf* CurK, const mcsref_mps* mps, const double* Scw, const int32_t* points, int np,
                double th, double nnratio, int having_masks, int32_t* ops, int cap, int32_t* n_ops) {
    GUARD(
        World w;
        cMultiKeyFrame k, cur;
        w.build_points(mps, &k, &cur);
        w.build(k, K);
        if (CurK) w.build(cur, CurK);
        std::vector<cMapPoint*> pts;
        for (int i = 0; i < np; ++i) pts.push_back(w.mp(points[i]));
        cORBmatcher m(nnratio, checkOrientation, K->view.dim, having_masks != 0);
        int n = 0;
        if (variant == 0) n = m.Fuse(&k, &cur, pts, th);
        else if (variant == 1) n = m.Fuse(&k, pts, th);
        else n = m.Fuse(&k, m44(Scw), pts, th);
        *n_ops = w.log_out(ops, cap);
        return n;
    )
}

}  // extern "C"
