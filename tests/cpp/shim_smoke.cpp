// Drives the reference-shaped C++ classes of include/mcs_shim.hpp: extraction through operator() on a synthetic
// image, SearchByBoW between the frame and a bit-flipped copy.  Prints a line the pytest wrapper parses.
// usage: shim_smoke <image.raw> <mask.raw> <w> <h> [vocabulary.txt]   (cam = Lafida camera 0, hard-coded from the fixture)
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include "../../include/mcs_shim.hpp"
using namespace MultiColSLAM;

static mcs::Mat8 load(const char* path, int w, int h) {
    mcs::Mat8 m(h, w);
    std::ifstream f(path, std::ios::binary);
    f.read((char*)m.ptr(), (std::streamsize)w * h);
    return m;
}

int main(int argc, char** argv) {
    if (argc < 5) return 2;
    const int w = atoi(argv[3]), h = atoi(argv[4]);
    try {
        const double cde[5] = {0.999626131079017, -0.0034775192597376, 0.00385134991673147, 392.219508388648, 243.494438476351};
        std::vector<double> p = {-209.200757992065, 0.0, 0.00213741670953883, -4.2203617319086e-06, 1.77146086919594e-08};
        std::vector<double> ip = {293.667187375663, 149.982043337335, -10.448650568161, 28.2295300683376, 7.13365723186292, 0.056303218962532,
                                  10.4144677485333, 0.166354960773665, -5.86858687381081, 1.18165998645705, 3.1108311354746, 0.810799620714366};
        cCamModelGeneral_ cam(cde, p, ip, w, h);
        mcs::Mat8 image = load(argv[1], w, h), mask = load(argv[2], w, h);
        bool same_mask = image.rows == mask.rows && std::memcmp(mask.ptr(), cam.GetMirrorMask(0).ptr(), (size_t)w * h) == 0;
        mdBRIEFextractorOct extractor(1000, 1.2f, 8, 25, 0, 0, 32, 20, false, 2, true, true, 32);
        std::vector<mcs::KeyPoint> kps;
        mcs::Mat8 desc, dmask;
        extractor(image, cam.GetMirrorMask(0), kps, cam, desc, dmask);
        unsigned long long sum = 0;
        for (size_t i = 0; i < desc.data.size(); ++i) sum = sum * 1315423911ull + desc.data[i] + 7ull * dmask.data[i];
        // empty image: silent return, outputs untouched
        std::vector<mcs::KeyPoint> kps2 = kps; mcs::Mat8 d2 = desc, m2 = dmask;
        extractor(mcs::Mat8(), mask, kps2, cam, d2, m2);
        bool untouched = kps2.size() == kps.size() && d2.data == desc.data;
        // loop-closure style brute force against a copy with one flipped bit per row
        mcs::Mat8 other = desc;
        for (int r = 0; r < other.rows; ++r) other.ptr(r)[r % 32] ^= 1;
        cORBmatcher matcher(0.9, false, 32, true);
        std::vector<int> m12;
        int n = matcher.SearchByBoW(desc, dmask, {}, other, dmask, {}, m12);
        int self = 0;
        for (int r = 0; r < desc.rows; ++r) self += m12[r] == r;
        // bag of words (optional 5th argument: vocabulary in DBoW2 text layout): transform both frames, then the
        // feature-vector guided SearchByBoW(KF, F) and the L1 score
        int bow = -1, fvn = -1, bm = -1; unsigned long long bowhash = 0; double score = -1;
        if (argc > 5) {
            ORBVocabulary voc;
            if (!voc.loadFromTextFile(argv[5])) { printf("SHIM error: vocabulary\n"); return 1; }
            DBoW2::BowVector bv1, bv2; DBoW2::FeatureVector fv1, fv2;
            voc.transform(desc, bv1, fv1, 4); voc.transform(other, bv2, fv2, 4);
            bow = (int)bv1.size(); fvn = (int)fv1.size();
            for (const auto& e : fv1) { bowhash = bowhash * 1315423911ull + e.first; for (unsigned i : e.second) bowhash = bowhash * 1315423911ull + i; }
            for (const auto& e : bv1) bowhash = bowhash * 1315423911ull + e.first;
            score = voc.score(bv1, bv2);
            std::vector<int> mf;
            bm = matcher.SearchByBoW(desc, dmask, {}, fv1, other, dmask, fv2, mf);
        }
        printf("SHIM nkp=%zu levels=%d ds=%d hash=%llu same_mask=%d untouched=%d matches=%d self=%d d01=%d bow=%d fvn=%d bowhash=%llu bm=%d score=%.17g\n",
               kps.size(), extractor.GetLevels(), extractor.GetDescriptorSize(), sum, (int)same_mask, (int)untouched, n, self,
               DescriptorDistance64(desc.ptr64(0), other.ptr64(0), 32), bow, fvn, bowhash, bm, score);
    } catch (const std::exception& e) {
        printf("SHIM error: %s\n", e.what());
        return 1;
    }
    return 0;
}
