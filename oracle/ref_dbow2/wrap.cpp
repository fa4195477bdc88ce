// C entry points around the REFERENCE's own DBoW2 (compiled in place from /root/reference/ThirdParty/DBoW2 by
// oracle/Makefile target `ref`) -- used by tests/ to pin oracle/mcs_oracle.cpp's BoW restatement and to generate
// tests/golden/bow_*.npz.  TEST INFRASTRUCTURE, never linked into the product.
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "DBoW2/FORB.h"
#include "DBoW2/TemplatedVocabulary.h"

typedef DBoW2::TemplatedVocabulary<DBoW2::FORB::TDescriptor, DBoW2::FORB> ORBVocabulary;   // ref include/cORBVocabulary.h:34

extern "C" {

void* refbow_load_text(const char* path) {
    ORBVocabulary* v = new ORBVocabulary();
    if (!v->loadFromTextFile(path)) { delete v; return nullptr; }
    return v;
}
void refbow_free(void* h) { delete (ORBVocabulary*)h; }
int refbow_size(void* h) { return (int)((ORBVocabulary*)h)->size(); }
void refbow_set_types(void* h, int scoring, int weighting) {
    ORBVocabulary* v = (ORBVocabulary*)h;
    v->setScoringType((DBoW2::ScoringType)scoring);
    // setWeightingType would recompute TF weights; the tests only flip between the flavours whose node weights are given
    v->setWeightingType((DBoW2::WeightingType)weighting);
}

static std::vector<cv::Mat> rows_of(const uint8_t* desc, int n) {       // as cConverter::toDescriptorVector (ref src/cConverter.cpp:58-76)
    cv::Mat all(n > 0 ? n : 1, 32, CV_8U);
    if (n > 0) std::memcpy(all.ptr<uint8_t>(), desc, (size_t)n * 32);
    std::vector<cv::Mat> v;
    for (int i = 0; i < n; ++i) v.push_back(all.row(i));
    return v;
}

// mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, levelsup)  (ref src/cMultiFrame.cpp:356-363)
// outputs: bow (word, value) in map order; feature vector as CSR (node, offsets, features)
int refbow_transform(void* h, const uint8_t* desc, int n, int levelsup, int32_t* bow_words, double* bow_values, int* n_bow,
                     int32_t* fv_nodes, int32_t* fv_off, int* n_fv, int32_t* fv_feat) {
    ORBVocabulary* v = (ORBVocabulary*)h;
    DBoW2::BowVector bv; DBoW2::FeatureVector fv;
    v->transform(rows_of(desc, n), bv, fv, levelsup);
    int k = 0;
    for (auto& e : bv) { bow_words[k] = (int32_t)e.first; bow_values[k] = e.second; ++k; }
    *n_bow = k;
    int f = 0, o = 0;
    for (auto& e : fv) {
        fv_nodes[f] = (int32_t)e.first; fv_off[f] = o;
        for (unsigned int i : e.second) fv_feat[o++] = (int32_t)i;
        ++f;
    }
    fv_off[f] = o; *n_fv = f;
    return 0;
}

// per-feature word id / weight (ref TemplatedVocabulary.h:1050-1062, :1040-1046)
void refbow_words(void* h, const uint8_t* desc, int n, int32_t* word, double* weight) {
    ORBVocabulary* v = (ORBVocabulary*)h;
    std::vector<cv::Mat> r = rows_of(desc, n);
    for (int i = 0; i < n; ++i) { const DBoW2::WordId w = v->transform(r[i]); word[i] = (int32_t)w; weight[i] = v->getWordWeight(w); }
}

double refbow_score(void* h, const int32_t* w1, const double* v1, int n1, const int32_t* w2, const double* v2, int n2) {
    DBoW2::BowVector a, b;
    for (int i = 0; i < n1; ++i) a.insert(a.end(), std::make_pair((DBoW2::WordId)w1[i], v1[i]));
    for (int i = 0; i < n2; ++i) b.insert(b.end(), std::make_pair((DBoW2::WordId)w2[i], v2[i]));
    return ((ORBVocabulary*)h)->score(a, b);
}

}  // extern "C"
