// mcs_adapters.hpp -- the cORBmatcher methods of the reference, re-expressed over libmcs_b200's C ABI for the reference's OWN
// container classes.  Header-only templates: they compile against anything that has the members the reference's matcher reads
// (cMultiFrame / cMultiKeyFrame / cMapPoint of /root/reference/include, or the data-only stand-ins the test suite uses), so a
// maintainer of the reference can keep cTracking / cLocalMapping / cLoopClosing untouched and forward each matcher call:
//
//     int cORBmatcher::SearchByProjection(cMultiFrame& F, const std::vector<cMapPoint*>& vpMapPoints, const double th)
//     { return mcs_adapt::SearchByProjection(F, vpMapPoints, th, mfNNratio, TH_HIGH_, havingMasks); }
//
// Each adapter (a) flattens the containers into the plain arrays of include/mcs_b200.h, (b) calls the C entry point (GPU), and
// (c) writes the result back into the containers exactly where the reference's method does (mvpMapPoints, vnMatches12,
// vbPrevMatched, vpMatches12).  tests/cpp/adapter_check.cpp runs them next to the reference's own cORBmatcher (compiled from
// /root/reference/src/cORBmatcher.cpp) on the same containers.  Needs only <vector>, the container headers and mcs_b200.h.
#ifndef MCS_ADAPTERS_HPP
#define MCS_ADAPTERS_HPP

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "mcs_b200.h"

namespace mcs_adapt {

inline void check(int rc) {
    if (rc != MCS_OK) throw std::runtime_error(std::string("libmcs_b200: ") + mcs_last_error());
}

// cMultiFrame / cMultiKeyFrame as the flat frame view of the C ABI.  The reference keeps one descriptor matrix per camera
// (mDescriptors[c], row = cont_idx_to_local_cam_idx[i]); the C ABI wants rows in contiguous keypoint order.
struct FlatFrame {
    std::vector<mcs_keypoint> keys;
    std::vector<int32_t> key_cam, cam_w, cam_h;
    std::vector<uint8_t> desc, dmask;
    std::vector<double> scale_factors;
    int dim = 32;
    mcs_frame_view view(bool masks) const {
        mcs_frame_view v;
        v.n_cams = (int32_t)cam_w.size(); v.n_keys = (int32_t)keys.size();
        v.keys = keys.data(); v.key_cam = key_cam.data(); v.desc = desc.data(); v.dmask = masks ? dmask.data() : nullptr;
        v.cam_width = cam_w.data(); v.cam_height = cam_h.data();
        v.dim = dim; v.n_levels = (int32_t)scale_factors.size(); v.scale_factors = scale_factors.data();
        return v;
    }
};

// Works for cMultiFrame (public mvKeys / mDescriptors / mvScaleFactors) -- the members src/cORBmatcher.cpp itself reads.
template <class Frame>
FlatFrame flatten_frame(Frame& F, int dim) {
    FlatFrame f;
    f.dim = dim;
    const int n = (int)F.mvKeys.size(), nc = F.camSystem.GetNrCams();
    static_assert(sizeof(F.mvKeys[0]) == sizeof(mcs_keypoint), "cv::KeyPoint layout");
    f.keys.resize(n); f.key_cam.resize(n);
    f.desc.assign((size_t)n * dim, 0); f.dmask.assign((size_t)n * dim, 0);
    for (int i = 0; i < n; ++i) {
        std::memcpy(&f.keys[i], &F.mvKeys[i], sizeof(mcs_keypoint));
        const int c = F.keypoint_to_cam.find(i)->second, r = F.cont_idx_to_local_cam_idx.find(i)->second;
        f.key_cam[i] = c;
        std::memcpy(&f.desc[(size_t)i * dim], F.mDescriptors[c].template ptr<uint8_t>(r), dim);
        if (!F.mDescriptorMasks.empty() && !F.mDescriptorMasks[c].empty())
            std::memcpy(&f.dmask[(size_t)i * dim], F.mDescriptorMasks[c].template ptr<uint8_t>(r), dim);
    }
    for (int c = 0; c < nc; ++c) {
        auto cam = F.camSystem.GetCamModelObj(c);
        f.cam_w.push_back((int32_t)cam.GetWidth()); f.cam_h.push_back((int32_t)cam.GetHeight());
    }
    f.scale_factors.assign(F.mvScaleFactors.begin(), F.mvScaleFactors.end());
    return f;
}

// cORBmatcher::SearchByProjection(cMultiFrame& F, const vector<cMapPoint*>& vpMapPoints, th)   ref src/cORBmatcher.cpp:67-166
template <class Frame, class MapPoint>
int SearchByProjection(Frame& F, const std::vector<MapPoint*>& vpMapPoints, const double th, double nnratio, int th_high, bool havingMasks,
                       int featDim = 32) {
    const FlatFrame f = flatten_frame(F, featDim);
    const int nc = (int)f.cam_w.size(), nmp = (int)vpMapPoints.size();
    std::vector<uint8_t> bad(nmp), in_view((size_t)nmp * nc, 0), desc((size_t)nmp * featDim), dmask((size_t)nmp * featDim, 0);
    std::vector<int32_t> level((size_t)nmp * nc, 0);
    std::vector<double> px((size_t)nmp * nc, 0), py((size_t)nmp * nc, 0), vc((size_t)nmp * nc, 0);
    for (int i = 0; i < nmp; ++i) {
        MapPoint* p = vpMapPoints[i];
        bad[i] = p->isBad() ? 1 : 0;
        std::memcpy(&desc[(size_t)i * featDim], p->GetDescriptorPtr(), featDim);
        if (havingMasks) std::memcpy(&dmask[(size_t)i * featDim], p->GetDescriptorMaskPtr(), featDim);
        for (int c = 0; c < nc && c < (int)p->mbTrackInView.size(); ++c) {
            const size_t k = (size_t)i * nc + c;
            in_view[k] = p->mbTrackInView[c] ? 1 : 0; level[k] = p->mnTrackScaleLevel[c];
            px[k] = p->mTrackProjX[c]; py[k] = p->mTrackProjY[c]; vc[k] = p->mTrackViewCos[c];
        }
    }
    // F.mvpMapPoints as indices into vpMapPoints (-1 = NULL, -2 = some other map point: occupied)
    std::vector<int32_t> frame_mp(f.keys.size(), -1);
    for (size_t i = 0; i < f.keys.size(); ++i)
        if (F.mvpMapPoints[i]) frame_mp[i] = nmp;            // occupied by a point outside this call's list: any value >= 0
    const mcs_frame_view fv = f.view(havingMasks);
    mcs_mappoint_view mv;
    mv.n_points = nmp; mv.bad = bad.data(); mv.in_view = in_view.data(); mv.level = level.data(); mv.proj_x = px.data(); mv.proj_y = py.data();
    mv.view_cos = vc.data(); mv.desc = desc.data(); mv.dmask = havingMasks ? dmask.data() : nullptr;
    int32_t n = 0;
    check(mcs_search_by_projection(&fv, &mv, th, nnratio, th_high, havingMasks ? 1 : 0, frame_mp.data(), &n));
    for (size_t i = 0; i < f.keys.size(); ++i)
        if (frame_mp[i] >= 0 && frame_mp[i] < nmp) F.mvpMapPoints[i] = vpMapPoints[frame_mp[i]];
    return n;
}

// cORBmatcher::SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize)   ref :579-726
template <class Frame, class Vec2>
int SearchForInitialization(Frame& F1, Frame& F2, std::vector<Vec2>& vbPrevMatched, std::vector<int>& vnMatches12, int windowSize, double nnratio,
                            int th_low, bool havingMasks, int featDim = 32) {
    const FlatFrame a = flatten_frame(F1, featDim), b = flatten_frame(F2, featDim);
    std::vector<double> prev(2 * a.keys.size());
    for (size_t i = 0; i < a.keys.size(); ++i) { prev[2 * i] = vbPrevMatched[i](0); prev[2 * i + 1] = vbPrevMatched[i](1); }
    std::vector<int32_t> m12(a.keys.size(), -1);
    const mcs_frame_view va = a.view(havingMasks), vb = b.view(havingMasks);
    int32_t n = 0;
    check(mcs_search_for_initialization(&va, &vb, prev.data(), windowSize, nnratio, th_low, havingMasks ? 1 : 0, m12.data(), &n));
    vnMatches12.assign(m12.begin(), m12.end());
    for (size_t i = 0; i < a.keys.size(); ++i) { vbPrevMatched[i](0) = prev[2 * i]; vbPrevMatched[i](1) = prev[2 * i + 1]; }
    return n;
}

// cORBmatcher::SearchByBoW(cMultiKeyFrame* pKF1, cMultiKeyFrame* pKF2, vector<cMapPoint*>& vpMatches12)   ref :885-966
template <class KeyFrame, class MapPoint>
int SearchByBoW(KeyFrame* pKF1, KeyFrame* pKF2, std::vector<MapPoint*>& vpMatches12, double nnratio, int th_low, bool havingMasks,
                int featDim = 32) {
    const std::vector<MapPoint*> mp1 = pKF1->GetMapPointMatches(), mp2 = pKF2->GetMapPointMatches();
    const int n1 = (int)mp1.size(), n2 = (int)mp2.size();
    std::vector<uint8_t> d1((size_t)n1 * featDim), m1((size_t)n1 * featDim, 0), d2((size_t)n2 * featDim), m2((size_t)n2 * featDim, 0), v1(n1), v2(n2);
    auto fill = [&](KeyFrame* kf, const std::vector<MapPoint*>& mp, std::vector<uint8_t>& d, std::vector<uint8_t>& m, std::vector<uint8_t>& v) {
        for (size_t i = 0; i < mp.size(); ++i) {
            const int c = kf->keypoint_to_cam.find(i)->second, r = kf->cont_idx_to_local_cam_idx.find(i)->second;
            std::memcpy(&d[i * featDim], kf->GetDescriptorRowPtr(c, r), featDim);
            if (havingMasks) std::memcpy(&m[i * featDim], kf->GetDescriptorMaskRowPtr(c, r), featDim);
            v[i] = (mp[i] && !mp[i]->isBad()) ? 1 : 0;
        }
    };
    fill(pKF1, mp1, d1, m1, v1); fill(pKF2, mp2, d2, m2, v2);
    std::vector<int32_t> m12(n1, -1);
    int32_t n = 0;
    check(mcs_match_bruteforce(d1.data(), havingMasks ? m1.data() : nullptr, v1.data(), n1, d2.data(), havingMasks ? m2.data() : nullptr, v2.data(), n2,
                               featDim, th_low, nnratio, m12.data(), &n));
    vpMatches12.assign(n1, static_cast<MapPoint*>(nullptr));
    for (int i = 0; i < n1; ++i)
        if (m12[i] >= 0) vpMatches12[i] = mp2[m12[i]];
    return n;
}

}  // namespace mcs_adapt
#endif  // MCS_ADAPTERS_HPP
