// mcs_shim.hpp -- host-side C++ mirror of the reference's operator / matcher interface for the hot path,
// implemented on top of the C ABI (mcs_b200.h).  Header-only; link with -lmcs_b200.
//
//   MultiColSLAM::mdBRIEFextractorOct   ref include/mdBRIEFextractorOct.h:333-421
//   MultiColSLAM::cORBmatcher           ref include/cORBmatcher.h:55-178  (the three searches named in BASELINE.json)
//   MultiColSLAM::cCamModelGeneral_     ref include/cam_model_omni.h:44-254 (the members the path reads)
//   DescriptorDistance64[Masked]        ref include/cORBmatcher.h:43-52
//
// Same class names, constructor arguments (order + defaults), method names, argument meaning and error behaviour:
//   * operator(): empty image -> silent return, outputs untouched (ref src/mdBRIEFextractorOct.cpp:1252-1253);
//     zero keypoints -> both outputs released (:1270-1274); a missing mask throws (the reference's rowRange on an
//     empty Mat throws cv::Exception, :913-917) -- here std::invalid_argument / mcs::Error;
//   * matchers return the int match count.
//
// The reference passes OpenCV types.  OpenCV C++ is not installed in the build image, so the default build uses the
// minimal binary-compatible stand-ins below (mcs::KeyPoint has cv::KeyPoint's 28-byte layout, mcs::Mat8 is a
// continuous CV_8UC1 matrix).  With -DMCS_WITH_OPENCV the very same classes take cv::InputArray / cv::OutputArray /
// std::vector<cv::KeyPoint> (INTEGRATION.md shows that build).
#ifndef MCS_SHIM_HPP
#define MCS_SHIM_HPP

#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "mcs_b200.h"

#ifdef MCS_WITH_OPENCV
#include <opencv2/core/core.hpp>
#endif

namespace mcs {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error("libmcs_b200 error " + std::to_string(c) + ": " + m), code(c) {}
};
inline void check(int rc) {
    if (rc != MCS_OK) throw Error(rc, mcs_last_error());
}

#ifdef MCS_WITH_OPENCV
typedef cv::KeyPoint KeyPoint;
#else
struct Point2f { float x, y; };
struct KeyPoint {              // layout of cv::KeyPoint
    Point2f pt;
    float size, angle, response;
    int octave, class_id;
};
#endif
static_assert(sizeof(KeyPoint) == sizeof(mcs_keypoint), "KeyPoint must be binary compatible with mcs_keypoint");

// continuous 8-bit single-channel matrix (stand-in for cv::Mat of type CV_8UC1)
struct Mat8 {
    int rows = 0, cols = 0;
    std::vector<uint8_t> data;
    Mat8() {}
    Mat8(int r, int c, uint8_t v = 0) : rows(r), cols(c), data((size_t)r * c, v) {}
    bool empty() const { return rows == 0 || cols == 0; }
    void create(int r, int c) { rows = r; cols = c; data.assign((size_t)r * c, 0); }
    void release() { rows = cols = 0; data.clear(); }
    uint8_t* ptr(int r = 0) { return data.data() + (size_t)r * cols; }
    const uint8_t* ptr(int r = 0) const { return data.data() + (size_t)r * cols; }
    // the matchers read rows as 64-bit words (ref src/cORBmatcher.cpp:125)
    const uint64_t* ptr64(int r) const { return reinterpret_cast<const uint64_t*>(ptr(r)); }
};

}  // namespace mcs

namespace MultiColSLAM {

inline int DescriptorDistance64(const uint64_t* descr_i, const uint64_t* descr_j, const int& dim) {
    return mcs_descriptor_distance64(descr_i, descr_j, dim);
}
inline int DescriptorDistance64Masked(const uint64_t* descr_i, const uint64_t* descr_j, const uint64_t* mask_i,
                                      const uint64_t* mask_j, const int& dim) {
    return mcs_descriptor_distance64_masked(descr_i, descr_j, mask_i, mask_j, dim);
}

// ---- camera model: the members the extractor / matchers read --------------------------------------------
class cCamModelGeneral_ {
public:
    cCamModelGeneral_() { std::memset(&oc_, 0, sizeof(oc_)); oc_.c = 1; }
    // cdeu0v0[5], forward polynomial p (nrpol <= 5), inverse polynomial invP (nrinvpol <= 12), image size
    cCamModelGeneral_(const double cdeu0v0[5], const std::vector<double>& p, const std::vector<double>& invP, double Iw, double Ih,
                      bool mirrorMask = true) {
        std::memset(&oc_, 0, sizeof(oc_));
        oc_.c = cdeu0v0[0]; oc_.d = cdeu0v0[1]; oc_.e = cdeu0v0[2]; oc_.u0 = cdeu0v0[3]; oc_.v0 = cdeu0v0[4];
        for (size_t i = 0; i < p.size() && i < 5; ++i) oc_.pol[i] = p[i];
        for (size_t i = 0; i < invP.size() && i < 12; ++i) oc_.inv_pol[i] = invP[i];
        oc_.width = (int)Iw; oc_.height = (int)Ih; oc_.mirror_mask = mirrorMask ? 1 : 0;
        mask_.create(oc_.height, oc_.width);
        mcs::check(mcs_cam_mirror_mask(&oc_, mask_.ptr()));      // CreateMirrorMask level 0, src/cam_model_omni.cpp:181-220
    }
    void WorldToImg(const double& x, const double& y, const double& z, double& u, double& v) const {
        mcs_cam_world_to_img(&oc_, x, y, z, &u, &v);
    }
    void ImgToWorld(double& x, double& y, double& z, const double& u, const double& v) const {
        mcs_cam_img_to_world(&oc_, u, v, &x, &y, &z);
    }
    const mcs::Mat8& GetMirrorMask(int /*pyrLevel*/ = 0) const { return mask_; }   // only level 0 is ever consumed
    double Get_u0() const { return oc_.u0; }
    double Get_v0() const { return oc_.v0; }
    double GetWidth() const { return oc_.width; }
    double GetHeight() const { return oc_.height; }
    const mcs_ocam& raw() const { return oc_; }

private:
    mcs_ocam oc_;
    mcs::Mat8 mask_;
};

// ---- extractor ---------------------------------------------------------------------------------------
class mdBRIEFextractorOct {
public:
    enum { HARRIS_SCORE = 0, FAST_SCORE = 1 };

    mdBRIEFextractorOct(int _nfeatures = 1000, float _scaleFactor = 1.2f, int _nlevels = 8, int _edgeThreshold = 25,
                        int _firstLevel = 0, int _scoreType = HARRIS_SCORE, int _patchSize = 32, int _fastThreshold = 20,
                        bool _useAgast = false, int _fastAgastType = 2, bool _do_dBrief = false, bool _learnMasks = false,
                        int _descSize = 32)
        : h_(nullptr) {
        p_ = mcs_extractor_params{_nfeatures, _scaleFactor, _nlevels, _edgeThreshold, _firstLevel, _scoreType, _patchSize,
                                  _fastThreshold, _useAgast ? 1 : 0, _fastAgastType, _do_dBrief ? 1 : 0, _learnMasks ? 1 : 0, _descSize};
        mcs::check(mcs_extractor_create(&p_, &h_));
        mcs::check(mcs_extractor_get_info(h_, &info_));
    }
    ~mdBRIEFextractorOct() { mcs_extractor_destroy(h_); }
    mdBRIEFextractorOct(const mdBRIEFextractorOct&) = delete;
    mdBRIEFextractorOct& operator=(const mdBRIEFextractorOct&) = delete;

    // operator()(image, mask, keypoints&, camModel&, descriptors, descriptorMasks)   ref :1244-1337
    void operator()(const mcs::Mat8& image, const mcs::Mat8& mask, std::vector<mcs::KeyPoint>& keypoints,
                    const cCamModelGeneral_& camModel, mcs::Mat8& descriptors, mcs::Mat8& descriptorMasks) {
        if (image.empty()) return;
        if (mask.empty()) throw std::invalid_argument("mdBRIEFextractorOct: a mask is mandatory (ref :913-917 throws)");
        extract(image.ptr(), image.cols, image.rows, image.cols, mask.ptr(), mask.cols, camModel.raw(), keypoints, descriptors,
                descriptorMasks);
    }
#ifdef MCS_WITH_OPENCV
    void operator()(cv::InputArray _image, cv::InputArray _mask, std::vector<cv::KeyPoint>& _keypoints,
                    const cCamModelGeneral_& camModel, cv::OutputArray _descriptors, cv::OutputArray _descriptorMasks) {
        if (_image.empty()) return;
        cv::Mat image = _image.getMat(), mask = _mask.getMat();
        CV_Assert(image.type() == CV_8UC1 && !mask.empty());
        mcs::Mat8 d, m;
        extract(image.data, image.cols, image.rows, (int)image.step, mask.data, (int)mask.step, camModel.raw(), _keypoints, d, m);
        if (_keypoints.empty()) { _descriptors.release(); _descriptorMasks.release(); return; }
        _descriptors.create((int)_keypoints.size(), info_.desc_size, CV_8U);
        _descriptorMasks.create((int)_keypoints.size(), info_.desc_size, CV_8U);
        std::memcpy(_descriptors.getMat().data, d.ptr(), d.data.size());
        std::memcpy(_descriptorMasks.getMat().data, m.ptr(), m.data.size());
    }
#endif
    int GetLevels() { return info_.nlevels; }
    double GetScaleFactor() { return (double)p_.scale_factor; }
    bool GetMasksLearned() { return p_.learn_masks != 0; }
    int GetDescriptorSize() { return info_.desc_size; }
    mcs_extractor* handle() { return h_; }

private:
    void extract(const uint8_t* img, int w, int h, int stride, const uint8_t* mask, int mstride, const mcs_ocam& cam,
                 std::vector<mcs::KeyPoint>& keypoints, mcs::Mat8& descriptors, mcs::Mat8& descriptorMasks) {
        const int cap = info_.capacity, ds = info_.desc_size;
        kp_.resize(cap); d_.resize((size_t)cap * ds); m_.resize((size_t)cap * ds);
        int n = 0;
        mcs::check(mcs_extract(h_, img, w, h, stride, mask, mstride, &cam, reinterpret_cast<mcs_keypoint*>(kp_.data()), d_.data(),
                               m_.data(), cap, &n));
        keypoints.clear();
        if (n == 0) { descriptors.release(); descriptorMasks.release(); return; }     // ref :1270-1274
        keypoints.assign(kp_.begin(), kp_.begin() + n);
        descriptors.create(n, ds); descriptorMasks.create(n, ds);
        std::memcpy(descriptors.ptr(), d_.data(), (size_t)n * ds);
        std::memcpy(descriptorMasks.ptr(), m_.data(), (size_t)n * ds);
    }
    mcs_extractor_params p_;
    mcs_extractor_info info_;
    mcs_extractor* h_;
    std::vector<mcs::KeyPoint> kp_;
    std::vector<uint8_t> d_, m_;
};

// ---- the frame / map-point fields the matchers read, flattened (cMultiFrame, cMapPoint stay host objects) ----
struct FrameFields {               // ref include/cMultiFrame.h:90-175
    std::vector<mcs::KeyPoint> mvKeys;            // contiguous, camera-major (src/cMultiFrame.cpp:168-184)
    std::vector<int> keypoint_to_cam;
    mcs::Mat8 descriptors, descriptorMasks;       // rows in contiguous index order
    std::vector<int> camWidth, camHeight;         // mnMaxX - mnMinX, mnMaxY - mnMinY per camera
    std::vector<double> mvScaleFactors;
    bool masksLearned = false;
    mcs_frame_view view() const {
        mcs_frame_view v;
        v.n_cams = (int)camWidth.size(); v.n_keys = (int)mvKeys.size();
        v.keys = reinterpret_cast<const mcs_keypoint*>(mvKeys.data()); v.key_cam = keypoint_to_cam.data();
        v.desc = descriptors.ptr(); v.dmask = masksLearned ? descriptorMasks.ptr() : nullptr;
        v.cam_width = camWidth.data(); v.cam_height = camHeight.data();
        v.dim = descriptors.cols; v.n_levels = (int)mvScaleFactors.size(); v.scale_factors = mvScaleFactors.data();
        return v;
    }
};
struct MapPointFields {            // ref include/cMapPoint.h (tracking fields) as parallel arrays, [point * nCams + cam]
    std::vector<uint8_t> bad, mbTrackInView;
    std::vector<int> mnTrackScaleLevel;
    std::vector<double> mTrackProjX, mTrackProjY, mTrackViewCos;
    mcs::Mat8 descriptors, descriptorMasks;
    mcs_mappoint_view view(bool masks) const {
        mcs_mappoint_view v;
        v.n_points = (int)bad.size(); v.bad = bad.data(); v.in_view = mbTrackInView.data(); v.level = mnTrackScaleLevel.data();
        v.proj_x = mTrackProjX.data(); v.proj_y = mTrackProjY.data(); v.view_cos = mTrackViewCos.data();
        v.desc = descriptors.ptr(); v.dmask = masks ? descriptorMasks.ptr() : nullptr;
        return v;
    }
};

// ---- ORBVocabulary (ref include/cORBVocabulary.h:34 = DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>) ----
namespace DBoW2 {
// BowVector / FeatureVector keep DBoW2's container types (ref ThirdParty/DBoW2/DBoW2/BowVector.h:62, FeatureVector.h:24)
typedef unsigned int WordId;
typedef double WordValue;
typedef unsigned int NodeId;
class BowVector : public std::map<WordId, WordValue> {};
class FeatureVector : public std::map<NodeId, std::vector<unsigned int> > {};
}  // namespace DBoW2

class ORBVocabulary {
public:
    ORBVocabulary() : h_(nullptr), n_words_(0) {}
    ~ORBVocabulary() { if (h_) mcs_vocabulary_destroy(h_); }
    ORBVocabulary(const ORBVocabulary&) = delete;
    ORBVocabulary& operator=(const ORBVocabulary&) = delete;
    // DBoW2 text layout (ref TemplatedVocabulary.h:1338-1425): "k L scoring weighting", then one line per node
    // "parent isLeaf d0 .. d31 weight"; node ids and word ids follow the line order.
    bool loadFromTextFile(const std::string& filename) {
        std::ifstream f(filename.c_str());
        if (!f) return false;
        int k, L, sc, wg;
        std::string line;
        if (!std::getline(f, line)) return false;
        { std::stringstream ss(line); ss >> k >> L >> sc >> wg; }
        if (k < 0 || k > 20 || L < 1 || L > 10 || sc < 0 || sc > 5 || wg < 0 || wg > 3) return false;     // same sanity check as ref :1359-1363
        std::vector<int> parent(1, -1), words;
        std::vector<double> weight(1, 0.0);
        std::vector<uint8_t> desc(32, 0);
        while (std::getline(f, line)) {
            if (line.find_first_not_of(" \t\r\n") == std::string::npos) continue;
            std::stringstream ss(line);
            int pid = 0, leaf = 0; double w = 0;
            ss >> pid >> leaf;
            const int nid = (int)parent.size();
            for (int i = 0; i < 32; ++i) { int b = 0; ss >> b; desc.push_back((uint8_t)b); }
            ss >> w;
            parent.push_back(pid); weight.push_back(w);
            if (leaf > 0) words.push_back(nid);
        }
        return create(k, L, sc, wg, parent, weight, desc, words);
    }
    bool create(int k, int L, int scoring, int weighting, const std::vector<int>& parent, const std::vector<double>& weight,
                const std::vector<uint8_t>& desc, const std::vector<int>& word_node, const int* node_order = nullptr) {
        if (h_) { mcs_vocabulary_destroy(h_); h_ = nullptr; }
        mcs::check(mcs_vocabulary_create(k, L, scoring, weighting, (int)parent.size(), parent.data(), weight.data(), desc.data(), node_order,
                                         (int)word_node.size(), word_node.data(), &h_));
        n_words_ = (unsigned)word_node.size();
        return true;
    }
    unsigned int size() const { return n_words_; }
    bool empty() const { return n_words_ == 0; }
    // transform(features, BowVector&, FeatureVector&, levelsup)  ref :1126-1194; features = descriptor rows of all cameras
    void transform(const mcs::Mat8& features, DBoW2::BowVector& v, DBoW2::FeatureVector& fv, int levelsup) const {
        v.clear(); fv.clear();
        if (!h_ || features.rows == 0) return;
        const int n = features.rows;
        std::vector<int> bw(n), fn(n), fo(n + 1), ff(n);
        std::vector<double> bv(n);
        int nb = 0, nf = 0;
        mcs::check(mcs_bow_vectors(h_, features.ptr(), n, levelsup, bw.data(), bv.data(), &nb, fn.data(), fo.data(), &nf, ff.data()));
        for (int i = 0; i < nb; ++i) v.insert(v.end(), std::make_pair((DBoW2::WordId)bw[i], bv[i]));
        for (int i = 0; i < nf; ++i)
            fv.insert(fv.end(), std::make_pair((DBoW2::NodeId)fn[i], std::vector<unsigned int>(ff.begin() + fo[i], ff.begin() + fo[i + 1])));
    }
    double score(const DBoW2::BowVector& a, const DBoW2::BowVector& b) const {
        std::vector<int> wa, wb; std::vector<double> va, vb;
        for (const auto& e : a) { wa.push_back((int)e.first); va.push_back(e.second); }
        for (const auto& e : b) { wb.push_back((int)e.first); vb.push_back(e.second); }
        double s = 0;
        mcs::check(mcs_bow_score(h_, wa.data(), va.data(), (int)wa.size(), wb.data(), vb.data(), (int)wb.size(), &s));
        return s;
    }
    const mcs_vocabulary* handle() const { return h_; }

private:
    mcs_vocabulary* h_;
    unsigned n_words_;
};

class cORBmatcher {
public:
    // ref src/cORBmatcher.cpp:46-64
    cORBmatcher(double nnratio = 0.6, bool checkOri = true, const int featDim = 32, bool havingMasks_ = false)
        : mfNNratio(nnratio), mbCheckOrientation(checkOri), mbFeatDim(featDim), havingMasks(havingMasks_) {
        if (havingMasks_) { TH_HIGH_ = (int)std::floor(1.5 * featDim); TH_LOW_ = (int)std::floor((double)featDim); }
        else { TH_HIGH_ = 3 * featDim; TH_LOW_ = 2 * featDim; }
    }
    // SearchByProjection(cMultiFrame&, vector<cMapPoint*>&, th)  ref :67-166.  mvpMapPoints: index of the map point
    // assigned to each keypoint (-1 = NULL), updated in place.
    int SearchByProjection(const FrameFields& F, const MapPointFields& vpMapPoints, std::vector<int>& mvpMapPoints, const double th) {
        int n = 0;
        mcs_frame_view fv = F.view(); mcs_mappoint_view mv = vpMapPoints.view(havingMasks);
        mcs::check(mcs_search_by_projection(&fv, &mv, th, mfNNratio, TH_HIGH_, havingMasks ? 1 : 0, mvpMapPoints.data(), &n));
        return n;
    }
    // SearchForInitialization(F1, F2, vbPrevMatched, vnMatches12, windowSize)  ref :579-726
    int SearchForInitialization(const FrameFields& F1, const FrameFields& F2, std::vector<double>& vbPrevMatched /* x0,y0,x1,y1,.. */,
                                std::vector<int>& vnMatches12, int windowSize = 10) {
        int n = 0;
        vnMatches12.assign(F1.mvKeys.size(), -1);
        mcs_frame_view a = F1.view(), b = F2.view();
        mcs::check(mcs_search_for_initialization(&a, &b, vbPrevMatched.data(), windowSize, mfNNratio, TH_LOW_, havingMasks ? 1 : 0,
                                                 vnMatches12.data(), &n));
        return n;
    }
    // SearchByBoW(cMultiKeyFrame*, cMultiKeyFrame*, vpMatches12)  ref :885-966 (all-pairs scan).  valid1/valid2 flag the
    // keypoints that carry a good map point; vpMatches12[i] = matched index in KF2 or -1.
    int SearchByBoW(const mcs::Mat8& desc1, const mcs::Mat8& mask1, const std::vector<uint8_t>& valid1, const mcs::Mat8& desc2,
                    const mcs::Mat8& mask2, const std::vector<uint8_t>& valid2, std::vector<int>& vpMatches12) {
        int n = 0;
        vpMatches12.assign(desc1.rows, -1);
        mcs::check(mcs_match_bruteforce(desc1.ptr(), havingMasks ? mask1.ptr() : nullptr, valid1.empty() ? nullptr : valid1.data(), desc1.rows,
                                        desc2.ptr(), havingMasks ? mask2.ptr() : nullptr, valid2.empty() ? nullptr : valid2.data(), desc2.rows,
                                        mbFeatDim, TH_LOW_, mfNNratio, vpMatches12.data(), &n));
        return n;
    }
    // SearchByBoW(cMultiKeyFrame* pKF, cMultiFrame& F, vpMapPointMatches)  ref :179-324: matching inside common vocabulary
    // nodes.  validKF flags key-frame keypoints with a good map point; vpMatchesF[i] = key-frame keypoint matched to frame
    // keypoint i, or -1 (the caller maps it to pKF->GetMapPointMatches()[...]).
    int SearchByBoW(const mcs::Mat8& descKF, const mcs::Mat8& maskKF, const std::vector<uint8_t>& validKF,
                    const DBoW2::FeatureVector& featVecKF, const mcs::Mat8& descF, const mcs::Mat8& maskF,
                    const DBoW2::FeatureVector& featVecF, std::vector<int>& vpMatchesF) {
        std::vector<int> n1, o1, f1, n2, o2, f2;
        flatten(featVecKF, n1, o1, f1); flatten(featVecF, n2, o2, f2);
        int n = 0;
        vpMatchesF.assign(descF.rows, -1);
        mcs::check(mcs_search_by_bow(descKF.ptr(), havingMasks ? maskKF.ptr() : nullptr, validKF.empty() ? nullptr : validKF.data(), descKF.rows,
                                     n1.data(), o1.data(), (int)n1.size(), f1.data(), descF.ptr(), havingMasks ? maskF.ptr() : nullptr, descF.rows,
                                     n2.data(), o2.data(), (int)n2.size(), f2.data(), mbFeatDim, TH_LOW_, mfNNratio, vpMatchesF.data(), &n));
        return n;
    }
    int TH_LOW_, TH_HIGH_;

protected:
    static void flatten(const DBoW2::FeatureVector& fv, std::vector<int>& nodes, std::vector<int>& off, std::vector<int>& feat) {
        off.push_back(0);
        for (const auto& e : fv) {
            nodes.push_back((int)e.first);
            for (unsigned int i : e.second) feat.push_back((int)i);
            off.push_back((int)feat.size());
        }
    }
    double mfNNratio;
    bool mbCheckOrientation;      // the reference compiles the orientation check out (include/cORBmatcher.h:40)
    int mbFeatDim;
    bool havingMasks;
};

}  // namespace MultiColSLAM
#endif  // MCS_SHIM_HPP
