// mcs_match_api.cu -- C-ABI host layer of the matcher entry points (include/mcs_b200.h).
//
// The GPU evaluates every Hamming distance and every grid-window candidate search; the host code here
// only (a) flattens the caller's frame into the CSR grid of cMultiFrame (ref src/cMultiFrame.cpp:168-184,
// :342-353) and (b) replays the reference's order-dependent greedy bookkeeping over the GPU results
// (ref src/cORBmatcher.cpp:121-164, :627-682, :922-961).
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "kernels.h"
#include "dev_scratch.h"
#include "mcs_common.cuh"

using namespace mcs;

namespace {

int mfail(int code, const std::string& msg);   // records the message for mcs_last_error() (mcs_api.cu)

#define MCK(expr)                                                                                      \
    do {                                                                                               \
        cudaError_t e__ = (expr);                                                                      \
        if (e__ != cudaSuccess) {                                                                      \
            cudaGetLastError();                                                                        \
            return mfail(e__ == cudaErrorNoDevice || e__ == cudaErrorInsufficientDriver ? MCS_ERR_NO_DEVICE : MCS_ERR_CUDA, \
                         std::string(#expr) + ": " + cudaGetErrorString(e__));                         \
        }                                                                                              \
    } while (0)


inline int cv_round(double v) { return (int)lrint(v); }

// frame uploaded for window searches
struct FrameDev {
    Dev kx, ky, koct, desc, dmask, cell_start, cell_items, winv, hinv, cam_first;
    WindowFrameDev view;
};

int upload_frame(const mcs_frame_view* f, FrameDev& d, cudaStream_t st) {
    // keypoints go up as they are; SoA copies, grid cell of every keypoint and the CSR grid are built on the device
    const int n = f->n_keys, nc = f->n_cams;
    for (int i = 0; i < n; ++i)
        if (f->key_cam[i] < 0 || f->key_cam[i] >= nc) return mfail(MCS_ERR_INVALID, "key_cam out of range");
    std::vector<mcs_ocam> cams(nc);
    std::memset(cams.data(), 0, sizeof(mcs_ocam) * nc);
    for (int c = 0; c < nc; ++c) { cams[c].width = f->cam_width[c]; cams[c].height = f->cam_height[c]; }
    const int ncell = nc * MCS_FRAME_GRID_COLS * MCS_FRAME_GRID_ROWS;
    const size_t db = (size_t)n * f->dim;
    Dev dkeys, dkc, dcams, dcell, dcur;
    MCK(dkeys.alloc((size_t)n * sizeof(mcs_keypoint))); MCK(dkc.alloc((size_t)n * 4)); MCK(dcams.alloc(nc * sizeof(mcs_ocam)));
    MCK(dcell.alloc((size_t)n * 4)); MCK(dcur.alloc((size_t)ncell * 4));
    MCK(d.kx.alloc(n * 4)); MCK(d.ky.alloc(n * 4)); MCK(d.koct.alloc(n * 4)); MCK(d.desc.alloc(db));
    MCK(d.cell_start.alloc((ncell + 1) * 4)); MCK(d.cell_items.alloc((size_t)std::max(n, 1) * 4));
    MCK(d.winv.alloc(nc * 8)); MCK(d.hinv.alloc(nc * 8));
    MCK(cudaMemcpyAsync(dkeys.p, f->keys, (size_t)n * sizeof(mcs_keypoint), cudaMemcpyHostToDevice, st));
    MCK(cudaMemcpyAsync(dkc.p, f->key_cam, (size_t)n * 4, cudaMemcpyHostToDevice, st));
    MCK(cudaMemcpyAsync(dcams.p, cams.data(), nc * sizeof(mcs_ocam), cudaMemcpyHostToDevice, st));
    MCK(cudaMemcpyAsync(d.desc.p, f->desc, db, cudaMemcpyHostToDevice, st));
    if (f->dmask) {
        MCK(d.dmask.alloc(db));
        MCK(cudaMemcpyAsync(d.dmask.p, f->dmask, db, cudaMemcpyHostToDevice, st));
    }
    MCK(launch_frame_prepare(dkeys.as<mcs_keypoint>(), dkc.as<int>(), n, dcams.as<mcs_ocam>(), nc, d.kx.as<float>(), d.ky.as<float>(),
                             d.koct.as<int>(), nullptr, dcell.as<int>(), dcur.as<int>(), d.cell_start.as<int>(), d.cell_items.as<int>(),
                             d.winv.as<double>(), d.hinv.as<double>(), st));
    MCK(cudaStreamSynchronize(st));    // host staging vectors and the scratch buffers die with this scope
    d.view = WindowFrameDev{nc, n, f->dim, d.kx.as<float>(), d.ky.as<float>(), d.koct.as<int>(), d.desc.as<uint8_t>(),
                            f->dmask ? d.dmask.as<uint8_t>() : nullptr, d.cell_start.as<int>(), d.cell_items.as<int>(),
                            d.winv.as<double>(), d.hinv.as<double>()};
    return MCS_OK;
}

// run the window-search kernel for host-side queries; grows max_cand until nothing overflows
int window_search_host(const FrameDev& fd, const std::vector<mcs_window_query>& qs, const uint8_t* qdesc, const uint8_t* qmask,
                       size_t qdesc_rows, int dim, std::vector<int>& cidx, std::vector<int>& cdist, std::vector<int>& ccount,
                       int& max_cand, bool allow_grow, cudaStream_t st) {
    const int nq = (int)qs.size();
    ccount.assign(nq, 0);
    if (nq == 0) return MCS_OK;
    Dev dq, dqd, dqm;
    MCK(dq.alloc(sizeof(mcs_window_query) * nq));
    MCK(dqd.alloc(qdesc_rows * dim));
    MCK(cudaMemcpyAsync(dq.p, qs.data(), sizeof(mcs_window_query) * nq, cudaMemcpyHostToDevice, st));
    MCK(cudaMemcpyAsync(dqd.p, qdesc, qdesc_rows * dim, cudaMemcpyHostToDevice, st));
    const bool masked = qmask && fd.view.dmask;
    if (masked) {
        MCK(dqm.alloc(qdesc_rows * dim));
        MCK(cudaMemcpyAsync(dqm.p, qmask, qdesc_rows * dim, cudaMemcpyHostToDevice, st));
    }
    for (;;) {
        Dev di, dd, dc;
        MCK(di.alloc((size_t)nq * max_cand * 4)); MCK(dd.alloc((size_t)nq * max_cand * 4)); MCK(dc.alloc((size_t)nq * 4));
        MCK(cudaMemsetAsync(di.p, 0xFF, (size_t)nq * max_cand * 4, st));      // unused list entries read back as -1
        MCK(cudaMemsetAsync(dd.p, 0, (size_t)nq * max_cand * 4, st));
        MCK(launch_window_search(fd.view, dq.as<mcs_window_query>(), nq, dqd.as<uint8_t>(), masked ? dqm.as<uint8_t>() : nullptr,
                                 max_cand, di.as<int>(), dd.as<int>(), dc.as<int>(), st));
        cidx.resize((size_t)nq * max_cand); cdist.resize((size_t)nq * max_cand);
        MCK(cudaMemcpyAsync(ccount.data(), dc.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
        MCK(cudaMemcpyAsync(cidx.data(), di.p, (size_t)nq * max_cand * 4, cudaMemcpyDeviceToHost, st));
        MCK(cudaMemcpyAsync(cdist.data(), dd.p, (size_t)nq * max_cand * 4, cudaMemcpyDeviceToHost, st));
        MCK(cudaStreamSynchronize(st));
        const int mx = *std::max_element(ccount.begin(), ccount.end());
        if (mx <= max_cand) return MCS_OK;
        if (!allow_grow) return MCS_ERR_CAPACITY;
        max_cand = (mx + 31) & ~31;
    }
}

// The same search with the candidates returned as ONE dense array in query order: candidates of query q are
// cidx / cdist [coff[q] .. coff[q + 1]).  Only the counts and the candidates that exist cross PCIe.
int window_search_compact(const FrameDev& fd, const mcs_window_query* qs, const int nq, const uint8_t* qdesc, const uint8_t* qmask,
                          size_t qdesc_rows, int dim, std::vector<int>& cidx, std::vector<int>& cdist, std::vector<int>& coff, cudaStream_t st) {
    coff.assign(nq + 1, 0);
    cidx.clear(); cdist.clear();
    if (nq == 0) return MCS_OK;
    Dev dq, dqd, dqm, dc, doff;
    MCK(dq.alloc(sizeof(mcs_window_query) * nq));
    MCK(dqd.alloc(qdesc_rows * dim));
    MCK(dc.alloc((size_t)nq * 4)); MCK(doff.alloc((size_t)nq * 4));
    MCK(cudaMemcpyAsync(dq.p, qs, sizeof(mcs_window_query) * nq, cudaMemcpyHostToDevice, st));
    MCK(cudaMemcpyAsync(dqd.p, qdesc, qdesc_rows * dim, cudaMemcpyHostToDevice, st));
    const bool masked = qmask && fd.view.dmask;
    if (masked) {
        MCK(dqm.alloc(qdesc_rows * dim));
        MCK(cudaMemcpyAsync(dqm.p, qmask, qdesc_rows * dim, cudaMemcpyHostToDevice, st));
    }
    std::vector<int> count(nq);
    int max_cand = 32;
    for (;;) {
        Dev di, dd;
        MCK(di.alloc((size_t)nq * max_cand * 4)); MCK(dd.alloc((size_t)nq * max_cand * 4));
        MCK(launch_window_search(fd.view, dq.as<mcs_window_query>(), nq, dqd.as<uint8_t>(), masked ? dqm.as<uint8_t>() : nullptr,
                                 max_cand, di.as<int>(), dd.as<int>(), dc.as<int>(), st));
        MCK(cudaMemcpyAsync(count.data(), dc.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
        MCK(cudaStreamSynchronize(st));
        const int mx = *std::max_element(count.begin(), count.end());
        if (mx > max_cand) { max_cand = (mx + 31) & ~31; continue; }        // a list overflowed: once more with room for the longest
        for (int q = 0; q < nq; ++q) coff[q + 1] = coff[q] + count[q];
        const int total = coff[nq];
        if (total == 0) return MCS_OK;
        Dev doi, dod;
        MCK(doi.alloc((size_t)total * 4)); MCK(dod.alloc((size_t)total * 4));
        MCK(cudaMemcpyAsync(doff.p, coff.data(), (size_t)nq * 4, cudaMemcpyHostToDevice, st));
        MCK(launch_compact_lists(di.as<int>(), dd.as<int>(), dc.as<int>(), doff.as<int>(), nq, max_cand, doi.as<int>(), dod.as<int>(), st));
        cidx.resize(total); cdist.resize(total);
        MCK(cudaMemcpyAsync(cidx.data(), doi.p, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
        MCK(cudaMemcpyAsync(cdist.data(), dod.p, (size_t)total * 4, cudaMemcpyDeviceToHost, st));
        MCK(cudaStreamSynchronize(st));
        return MCS_OK;
    }
}

}  // namespace

void mcs_set_error_(const std::string& msg);   // mcs_api.cu
namespace { int mfail(int code, const std::string& msg) { mcs_set_error_(msg); return code; } }

extern "C" {

int mcs_descriptor_distance64(const uint64_t* a, const uint64_t* b, int32_t dim) {   // ref :2438-2450
    uint64_t d = 0;
    for (int i = 0; i < dim / 8; ++i) d += (uint64_t)__builtin_popcountll(a[i] ^ b[i]);
    return (int)d;
}

int mcs_descriptor_distance64_masked(const uint64_t* a, const uint64_t* b, const uint64_t* ma, const uint64_t* mb,
                                     int32_t dim) {   // ref :2452-2474
    uint64_t d = 0;
    for (int i = 0; i < dim / 8; ++i) {
        const uint64_t x = a[i] ^ b[i];
        d += (uint64_t)__builtin_popcountll(x & ma[i]);
        d += (uint64_t)__builtin_popcountll(x & mb[i]);
    }
    return (int)(d / 2);
}

int mcs_hamming_topk_device(const uint8_t* q_dev, const uint8_t* qmask_dev, int32_t nq, const uint8_t* d_dev,
                            const uint8_t* dmask_dev, int32_t nd, const uint8_t* db_skip_dev, int32_t dim, int32_t K,
                            int32_t* topk_idx_dev, int32_t* topk_dist_dev, void* stream) {
    if (!q_dev || !d_dev || !topk_idx_dev || !topk_dist_dev) return mfail(MCS_ERR_INVALID, "null argument");
    if (nq < 0 || nd < 0 || K < 1 || K > 8) return mfail(MCS_ERR_INVALID, "bad sizes (K must be 1..8)");
    if (dim != 16 && dim != 32 && dim != 64) return mfail(MCS_ERR_INVALID, "dim must be 16, 32 or 64");
    if (nd >= (1 << 21)) return mfail(MCS_ERR_UNSUPPORTED, "more than 2^21 - 1 database descriptors in one call (split the database)");
    MCK(launch_hamming_topk(q_dev, qmask_dev, nq, d_dev, dmask_dev, nd, db_skip_dev, dim, K, 0xFFFFFFFFu, topk_idx_dev, topk_dist_dev,
                            (cudaStream_t)stream));
    return MCS_OK;
}

int mcs_hamming_topk(const uint8_t* q, const uint8_t* qmask, int32_t nq, const uint8_t* d, const uint8_t* dmask, int32_t nd,
                     const uint8_t* db_skip, int32_t dim, int32_t K, int32_t* topk_idx, int32_t* topk_dist) {
    if (!q || !d || !topk_idx || !topk_dist) return mfail(MCS_ERR_INVALID, "null argument");
    if (nq <= 0) return MCS_OK;
    const bool masked = qmask && dmask;
    Dev dq, dqm, dd, ddm, ds, di, dt;
    MCK(dq.alloc((size_t)nq * dim)); MCK(dd.alloc((size_t)nd * dim));
    MCK(di.alloc((size_t)nq * K * 4)); MCK(dt.alloc((size_t)nq * K * 4));
    MCK(cudaMemcpy(dq.p, q, (size_t)nq * dim, cudaMemcpyHostToDevice));
    MCK(cudaMemcpy(dd.p, d, (size_t)nd * dim, cudaMemcpyHostToDevice));
    if (masked) {
        MCK(dqm.alloc((size_t)nq * dim)); MCK(ddm.alloc((size_t)nd * dim));
        MCK(cudaMemcpy(dqm.p, qmask, (size_t)nq * dim, cudaMemcpyHostToDevice));
        MCK(cudaMemcpy(ddm.p, dmask, (size_t)nd * dim, cudaMemcpyHostToDevice));
    }
    if (db_skip) { MCK(ds.alloc(nd)); MCK(cudaMemcpy(ds.p, db_skip, nd, cudaMemcpyHostToDevice)); }
    int rc = mcs_hamming_topk_device(dq.as<uint8_t>(), masked ? dqm.as<uint8_t>() : nullptr, nq, dd.as<uint8_t>(),
                                     masked ? ddm.as<uint8_t>() : nullptr, nd, db_skip ? ds.as<uint8_t>() : nullptr, dim, K,
                                     di.as<int>(), dt.as<int>(), nullptr);
    if (rc) return rc;
    MCK(cudaMemcpy(topk_idx, di.p, (size_t)nq * K * 4, cudaMemcpyDeviceToHost));
    MCK(cudaMemcpy(topk_dist, dt.p, (size_t)nq * K * 4, cudaMemcpyDeviceToHost));
    return MCS_OK;
}

// SearchByBoW(KF1, KF2) (ref :885-966) with the descriptors already on the device: K best unmatched database entries per query
// on the GPU, sequential greedy replay on the host, further rounds for queries whose list was used up by earlier matches
static thread_local int g_bf_rounds = 0;
int mcs_last_bruteforce_rounds(void) { return g_bf_rounds; }

// seg[0..n_seg]: query segments (e.g. the key frames of a batch) whose "database entry already matched" state is independent, as
// it is between separate SearchByBoW(KF1, KF2) calls of the reference; seg == nullptr: one segment [0, nq).  One K-best launch
// serves all queries of all segments (they start from the same database state); the ordered acceptance (ref :899-961) then runs
// on the device as well, one CTA per segment (bruteforce_replay_kernel): a query whose list was used up by matches accepted
// earlier in its segment is rescanned exactly inside the kernel.  The host only moves the small validity / offset arrays in and
// the matches out.
static int bruteforce_core(const uint8_t* q_dev, const uint8_t* qm_dev, const uint8_t* valid1, int nq, const uint8_t* d_dev,
                           const uint8_t* dm_dev, const uint8_t* valid2, int nd, int dim, int th_low, double nnratio, int* matches12,
                           int* nmatches, cudaStream_t st, const int* seg = nullptr, int n_seg = 1) {
    const bool masked = qm_dev && dm_dev;
    // K = 8: with the relevance bound only entries that matter are ever inserted, so long lists cost nothing in the distance loop
    // and leave few queries to rescan (BASELINE config 4: 207 of 4000 per key frame at K = 4, 60 at K = 8)
    constexpr int K = 8;
    const int one_seg[2] = {0, nq};
    if (!seg) { seg = one_seg; n_seg = 1; }
    if (nd >= (1 << 21)) return mfail(MCS_ERR_UNSUPPORTED, "more than 2^21 database descriptors in one call");
    Dev ds, di, dt, dv1, dseg, dm12, dnm;
    MCK(ds.alloc(nd)); MCK(di.alloc((size_t)nq * K * 4)); MCK(dt.alloc((size_t)nq * K * 4));
    MCK(dv1.alloc(nq)); MCK(dseg.alloc((size_t)(n_seg + 1) * 4)); MCK(dm12.alloc((size_t)nq * 4)); MCK(dnm.alloc((size_t)n_seg * 4));
    std::vector<uint8_t> skip(nd, 0);
    if (valid2) for (int i = 0; i < nd; ++i) skip[i] = valid2[i] ? 0 : 1;
    g_bf_rounds = 1;
    MCK(cudaMemcpyAsync(ds.p, skip.data(), nd, cudaMemcpyHostToDevice, st));
    if (valid1) MCK(cudaMemcpyAsync(dv1.p, valid1, nq, cudaMemcpyHostToDevice, st));
    MCK(cudaMemcpyAsync(dseg.p, seg, (size_t)(n_seg + 1) * 4, cudaMemcpyHostToDevice, st));
    // Entries at or beyond the relevance bound of (th_low, nnratio) stay out of the lists: a list shorter than K then means "every
    // entry that can influence the decision is here".
    MCK(launch_hamming_topk(q_dev, masked ? qm_dev : nullptr, nq, d_dev, masked ? dm_dev : nullptr, nd, ds.as<uint8_t>(), dim, K,
                            greedy_dist_bound(th_low, nnratio), di.as<int>(), dt.as<int>(), st));
    // the replay kernel wants "valid" bytes for the database (1 = usable): reuse the skip buffer inverted on the fly is not worth a
    // kernel -- upload the caller's array when there is one
    Dev dv2;
    if (valid2) { MCK(dv2.alloc(nd)); MCK(cudaMemcpyAsync(dv2.p, valid2, nd, cudaMemcpyHostToDevice, st)); }
    MCK(launch_bruteforce_replay(di.as<int>(), dt.as<int>(), K, q_dev, masked ? qm_dev : nullptr, valid1 ? dv1.as<uint8_t>() : nullptr,
                                 dseg.as<int>(), n_seg, nq, d_dev, masked ? dm_dev : nullptr, valid2 ? dv2.as<uint8_t>() : nullptr, nd, dim, th_low,
                                 nnratio, dm12.as<int>(), dnm.as<int>(), st));
    MCK(cudaMemcpyAsync(matches12, dm12.p, (size_t)nq * 4, cudaMemcpyDeviceToHost, st));
    MCK(cudaMemcpyAsync(nmatches, dnm.p, (size_t)n_seg * 4, cudaMemcpyDeviceToHost, st));
    MCK(cudaStreamSynchronize(st));
    return MCS_OK;
}

int mcs_match_bruteforce_batch_device(const uint8_t* q_dev, const uint8_t* qmask_dev, const uint8_t* valid1, const int32_t* seg_start,
                                      int32_t n_seg, const uint8_t* d_dev, const uint8_t* dmask_dev, const uint8_t* valid2, int32_t nd,
                                      int32_t dim, int32_t th_low, double nnratio, int32_t* matches12, int32_t* nmatches, void* stream) {
    if (!q_dev || !d_dev || !matches12 || !nmatches || !seg_start) return mfail(MCS_ERR_INVALID, "null argument");
    if (dim != 16 && dim != 32 && dim != 64) return mfail(MCS_ERR_INVALID, "dim must be 16, 32 or 64");
    if (n_seg < 1 || seg_start[0] != 0) return mfail(MCS_ERR_INVALID, "seg_start must begin at 0 and hold n_seg + 1 offsets");
    for (int s = 0; s < n_seg; ++s)
        if (seg_start[s + 1] < seg_start[s]) return mfail(MCS_ERR_INVALID, "seg_start must be non-decreasing");
    const int nq = seg_start[n_seg];
    for (int s = 0; s < n_seg; ++s) nmatches[s] = 0;
    for (int i = 0; i < nq; ++i) matches12[i] = -1;
    if (nq <= 0 || nd <= 0) return MCS_OK;
    return bruteforce_core(q_dev, qmask_dev, valid1, nq, d_dev, dmask_dev, valid2, nd, dim, th_low, nnratio, matches12, nmatches,
                           (cudaStream_t)stream, seg_start, n_seg);
}

int mcs_match_bruteforce_device(const uint8_t* q_dev, const uint8_t* qmask_dev, const uint8_t* valid1, int32_t nq, const uint8_t* d_dev,
                                const uint8_t* dmask_dev, const uint8_t* valid2, int32_t nd, int32_t dim, int32_t th_low, double nnratio,
                                int32_t* matches12, int32_t* nmatches, void* stream) {
    if (!q_dev || !d_dev || !matches12 || !nmatches) return mfail(MCS_ERR_INVALID, "null argument");
    if (dim != 16 && dim != 32 && dim != 64) return mfail(MCS_ERR_INVALID, "dim must be 16, 32 or 64");
    *nmatches = 0;
    for (int i = 0; i < nq; ++i) matches12[i] = -1;
    if (nq <= 0 || nd <= 0) return MCS_OK;
    return bruteforce_core(q_dev, qmask_dev, valid1, nq, d_dev, dmask_dev, valid2, nd, dim, th_low, nnratio, matches12, nmatches,
                           (cudaStream_t)stream);
}

int mcs_match_bruteforce(const uint8_t* q, const uint8_t* qmask, const uint8_t* valid1, int32_t nq, const uint8_t* d,
                         const uint8_t* dmask, const uint8_t* valid2, int32_t nd, int32_t dim, int32_t th_low, double nnratio,
                         int32_t* matches12, int32_t* nmatches) {
    if (!q || !d || !matches12 || !nmatches) return mfail(MCS_ERR_INVALID, "null argument");
    if (dim != 16 && dim != 32 && dim != 64) return mfail(MCS_ERR_INVALID, "dim must be 16, 32 or 64");
    *nmatches = 0;
    for (int i = 0; i < nq; ++i) matches12[i] = -1;
    if (nq <= 0 || nd <= 0) return MCS_OK;
    const bool masked = qmask && dmask;
    Dev dq, dqm, dd, ddm;
    MCK(dq.alloc((size_t)nq * dim)); MCK(dd.alloc((size_t)nd * dim));
    MCK(cudaMemcpy(dq.p, q, (size_t)nq * dim, cudaMemcpyHostToDevice));
    MCK(cudaMemcpy(dd.p, d, (size_t)nd * dim, cudaMemcpyHostToDevice));
    if (masked) {
        MCK(dqm.alloc((size_t)nq * dim)); MCK(ddm.alloc((size_t)nd * dim));
        MCK(cudaMemcpy(dqm.p, qmask, (size_t)nq * dim, cudaMemcpyHostToDevice));
        MCK(cudaMemcpy(ddm.p, dmask, (size_t)nd * dim, cudaMemcpyHostToDevice));
    }
    return bruteforce_core(dq.as<uint8_t>(), masked ? dqm.as<uint8_t>() : nullptr, valid1, nq, dd.as<uint8_t>(),
                           masked ? ddm.as<uint8_t>() : nullptr, valid2, nd, dim, th_low, nnratio, matches12, nmatches, nullptr);
}

// CheckDistEpipolarLine (ref src/misc.cpp:53-69): cv::Matx products accumulate s = 0; s += a*b in index order
static bool epipolar_ok(const double* r1, const double* r2, const double* E, double thresh) {
    double t[3];                                       // ray2^T * E  (1x3)
    for (int j = 0; j < 3; ++j) { double s = 0; for (int i = 0; i < 3; ++i) s += r2[i] * E[3 * i + j]; t[j] = s; }
    double nom = 0; for (int j = 0; j < 3; ++j) nom += t[j] * r1[j];
    double ex1[3], etx2[3];
    for (int i = 0; i < 3; ++i) { double s = 0; for (int k = 0; k < 3; ++k) s += E[3 * i + k] * r1[k]; ex1[i] = s; }
    for (int i = 0; i < 3; ++i) { double s = 0; for (int k = 0; k < 3; ++k) s += E[3 * k + i] * r2[k]; etx2[i] = s; }
    const double den = ex1[0] * ex1[0] + ex1[1] * ex1[1] + ex1[2] * ex1[2] + etx2[0] * etx2[0] + etx2[1] * etx2[1] + etx2[2] * etx2[2];
    if (den == 0.0) return false;
    return (nom * nom) / den < thresh;
}

int mcs_search_for_triangulation(const uint8_t* desc1, const uint8_t* mask1, const int32_t* cam1, const uint8_t* free1,
                                 const double* rays1, int32_t n1, const uint8_t* desc2, const uint8_t* mask2, const int32_t* cam2,
                                 const uint8_t* free2, const double* rays2, int32_t n2, int32_t dim, int32_t th_low, const double* E,
                                 int32_t n_cams, double epi_thresh, int32_t* matches12, int32_t* nmatches) {
    if (!desc1 || !cam1 || !free1 || !rays1 || !desc2 || !cam2 || !free2 || !rays2 || !E || !matches12 || !nmatches)
        return mfail(MCS_ERR_INVALID, "null argument");
    if (dim != 16 && dim != 32 && dim != 64) return mfail(MCS_ERR_INVALID, "dim must be 16, 32 or 64");
    *nmatches = 0;
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    if (n1 <= 0 || n2 <= 0) return MCS_OK;
    const bool masked = mask1 && mask2;
    constexpr int K = 8;
    Dev dd, ddm, ds;
    MCK(dd.alloc((size_t)n2 * dim)); MCK(ds.alloc(n2));
    MCK(cudaMemcpy(dd.p, desc2, (size_t)n2 * dim, cudaMemcpyHostToDevice));
    if (masked) { MCK(ddm.alloc((size_t)n2 * dim)); MCK(cudaMemcpy(ddm.p, mask2, (size_t)n2 * dim, cudaMemcpyHostToDevice)); }
    std::vector<uint8_t> matched2(n2, 0), skip(n2);
    int nm = 0;
    for (int c = 0; c < n_cams; ++c) {                 // queries of camera c only ever see database entries of camera c (ref :1043-1045)
        std::vector<int> q_idx;
        for (int i = 0; i < n1; ++i) if (cam1[i] == c && free1[i]) q_idx.push_back(i);
        if (q_idx.empty()) continue;
        const int nq = (int)q_idx.size();
        std::vector<uint8_t> qd((size_t)nq * dim), qmk(masked ? (size_t)nq * dim : 0);
        for (int k = 0; k < nq; ++k) {
            std::memcpy(&qd[(size_t)k * dim], desc1 + (size_t)q_idx[k] * dim, dim);
            if (masked) std::memcpy(&qmk[(size_t)k * dim], mask1 + (size_t)q_idx[k] * dim, dim);
        }
        Dev dq, dqm, di, dt;
        MCK(dq.alloc((size_t)nq * dim)); MCK(di.alloc((size_t)nq * K * 4)); MCK(dt.alloc((size_t)nq * K * 4));
        MCK(cudaMemcpy(dq.p, qd.data(), (size_t)nq * dim, cudaMemcpyHostToDevice));
        if (masked) { MCK(dqm.alloc((size_t)nq * dim)); MCK(cudaMemcpy(dqm.p, qmk.data(), (size_t)nq * dim, cudaMemcpyHostToDevice)); }
        std::vector<int> tidx((size_t)nq * K), tdist((size_t)nq * K);
        const double* Ecc = E + ((size_t)c * n_cams + c) * 9;
        // one GPU pass: the K best database entries (distance, index) of every query under the skip state at the start
        for (int i = 0; i < n2; ++i) skip[i] = (matched2[i] || !free2[i] || cam2[i] != c) ? 1 : 0;
        MCK(cudaMemcpy(ds.p, skip.data(), n2, cudaMemcpyHostToDevice));
        int rc = mcs_hamming_topk_device(dq.as<uint8_t>(), masked ? dqm.as<uint8_t>() : nullptr, nq, dd.as<uint8_t>(),
                                         masked ? ddm.as<uint8_t>() : nullptr, n2, ds.as<uint8_t>(), dim, K, di.as<int>(), dt.as<int>(), nullptr);
        if (rc) return rc;
        MCK(cudaMemcpy(tidx.data(), di.p, (size_t)nq * K * 4, cudaMemcpyDeviceToHost));
        MCK(cudaMemcpy(tdist.data(), dt.p, (size_t)nq * K * 4, cudaMemcpyDeviceToHost));
        std::vector<int> pi(K), pd(K), marked;
        for (int k = 0; k < nq; ++k) {                   // sequential replay (ref :1015-1107)
            const double* r1 = rays1 + 3 * (size_t)q_idx[k];
            const int* li = &tidx[(size_t)k * K];
            const int* ld = &tdist[(size_t)k * K];
            int best = -1, dist_th = 0, found = -1;
            bool done = false;
            auto walk = [&](const int* idx, const int* dst) {      // returns true when the list was consumed without a decision
                for (int e = 0; e < K; ++e) {
                    if (idx[e] < 0 || dst[e] > th_low) { done = true; return false; }     // no candidate with dist <= TH_LOW_ left
                    if (matched2[idx[e]]) continue;                                       // taken by an earlier query of this pass
                    if (best < 0) { best = dst[e]; dist_th = (int)lrint(2.0 * best); }
                    if (dst[e] > dist_th) { done = true; return false; }
                    if (epipolar_ok(r1, rays2 + 3 * (size_t)idx[e], Ecc, epi_thresh)) { found = idx[e]; done = true; return false; }
                    skip[idx[e]] = 2; marked.push_back(idx[e]);                           // examined by this query (paging only)
                }
                return true;
            };
            if (walk(li, ld)) {
                // more than K candidates inside both thresholds: page through the rest for this one query on the GPU
                std::vector<uint8_t> sk(n2);
                while (!done) {
                    for (int i = 0; i < n2; ++i) sk[i] = (matched2[i] || !free2[i] || cam2[i] != c || skip[i] == 2) ? 1 : 0;
                    MCK(cudaMemcpy(ds.p, sk.data(), n2, cudaMemcpyHostToDevice));
                    rc = mcs_hamming_topk_device(dq.as<uint8_t>() + (size_t)k * dim, masked ? dqm.as<uint8_t>() + (size_t)k * dim : nullptr, 1,
                                                 dd.as<uint8_t>(), masked ? ddm.as<uint8_t>() : nullptr, n2, ds.as<uint8_t>(), dim, K,
                                                 di.as<int>(), dt.as<int>(), nullptr);
                    if (rc) return rc;
                    MCK(cudaMemcpy(pi.data(), di.p, K * 4, cudaMemcpyDeviceToHost));
                    MCK(cudaMemcpy(pd.data(), dt.p, K * 4, cudaMemcpyDeviceToHost));
                    if (!walk(pi.data(), pd.data())) break;
                }
            }
            for (int e : marked) skip[e] = 0;                                             // clear this query's paging marks
            marked.clear();
            if (found >= 0) { matches12[q_idx[k]] = found; matched2[found] = 1; ++nm; }
        }
    }
    *nmatches = nm;
    return MCS_OK;
}

int mcs_window_search(const mcs_frame_view* frame, const mcs_window_query* queries, int32_t nq, const uint8_t* qdesc,
                      const uint8_t* qmask, int32_t max_cand, int32_t* cand_idx, int32_t* cand_dist, int32_t* cand_count) {
    if (!frame || !queries || !qdesc || !cand_idx || !cand_dist || !cand_count || max_cand < 1)
        return mfail(MCS_ERR_INVALID, "null argument");
    if (nq <= 0) return MCS_OK;
    FrameDev fd;
    int rc = upload_frame(frame, fd, nullptr);
    if (rc) return rc;
    int rows = 0;
    for (int i = 0; i < nq; ++i) {
        if (queries[i].cam < 0 || queries[i].cam >= frame->n_cams) return mfail(MCS_ERR_INVALID, "query camera out of range");
        if (queries[i].desc_index < 0) return mfail(MCS_ERR_INVALID, "negative query descriptor index");
        rows = std::max(rows, queries[i].desc_index + 1);
    }
    std::vector<mcs_window_query> qs(queries, queries + nq);
    std::vector<int> ci, cd, cc;
    int mc = max_cand;
    rc = window_search_host(fd, qs, qdesc, qmask, rows, frame->dim, ci, cd, cc, mc, false, nullptr);
    if (rc != MCS_OK && rc != MCS_ERR_CAPACITY) return rc;
    std::memcpy(cand_count, cc.data(), sizeof(int) * nq);
    std::memcpy(cand_idx, ci.data(), sizeof(int) * (size_t)nq * max_cand);
    std::memcpy(cand_dist, cd.data(), sizeof(int) * (size_t)nq * max_cand);
    if (rc == MCS_ERR_CAPACITY) return mfail(MCS_ERR_CAPACITY, "max_cand too small for at least one query");
    return MCS_OK;
}

int mcs_frame_prepare(const mcs_keypoint* keys, const int32_t* key_cam, int32_t n_keys, const mcs_ocam* cams, int32_t n_cams,
                      double* rays_out, int32_t* cell_start_out, int32_t* cell_items_out, int32_t* n_in_grid) {
    if (!keys || !key_cam || !cams || !rays_out || !cell_start_out || !cell_items_out || !n_in_grid)
        return mfail(MCS_ERR_INVALID, "null argument");
    if (n_keys < 0 || n_cams < 1) return mfail(MCS_ERR_INVALID, "bad sizes");
    for (int i = 0; i < n_keys; ++i)
        if (key_cam[i] < 0 || key_cam[i] >= n_cams) return mfail(MCS_ERR_INVALID, "key_cam out of range");
    const int ncell = n_cams * MCS_FRAME_GRID_COLS * MCS_FRAME_GRID_ROWS;
    Dev dkeys, dkc, dcams, dcell, dcur, dkx, dky, dko, drays, dstart, ditems, dw, dh;
    const size_t n = (size_t)std::max(n_keys, 1);
    MCK(dkeys.alloc(n * sizeof(mcs_keypoint))); MCK(dkc.alloc(n * 4)); MCK(dcams.alloc(n_cams * sizeof(mcs_ocam)));
    MCK(dcell.alloc(n * 4)); MCK(dcur.alloc((size_t)ncell * 4)); MCK(dkx.alloc(n * 4)); MCK(dky.alloc(n * 4)); MCK(dko.alloc(n * 4));
    MCK(drays.alloc(n * 24)); MCK(dstart.alloc((size_t)(ncell + 1) * 4)); MCK(ditems.alloc(n * 4)); MCK(dw.alloc(n_cams * 8)); MCK(dh.alloc(n_cams * 8));
    MCK(cudaMemcpy(dkeys.p, keys, (size_t)n_keys * sizeof(mcs_keypoint), cudaMemcpyHostToDevice));
    MCK(cudaMemcpy(dkc.p, key_cam, (size_t)n_keys * 4, cudaMemcpyHostToDevice));
    MCK(cudaMemcpy(dcams.p, cams, n_cams * sizeof(mcs_ocam), cudaMemcpyHostToDevice));
    MCK(launch_frame_prepare(dkeys.as<mcs_keypoint>(), dkc.as<int>(), n_keys, dcams.as<mcs_ocam>(), n_cams, dkx.as<float>(), dky.as<float>(),
                             dko.as<int>(), drays.as<double>(), dcell.as<int>(), dcur.as<int>(), dstart.as<int>(), ditems.as<int>(),
                             dw.as<double>(), dh.as<double>(), nullptr));
    MCK(cudaMemcpy(rays_out, drays.p, (size_t)n_keys * 24, cudaMemcpyDeviceToHost));
    MCK(cudaMemcpy(cell_start_out, dstart.p, (size_t)(ncell + 1) * 4, cudaMemcpyDeviceToHost));
    MCK(cudaMemcpy(cell_items_out, ditems.p, (size_t)n_keys * 4, cudaMemcpyDeviceToHost));
    *n_in_grid = cell_start_out[ncell];
    return MCS_OK;
}

int mcs_project_mappoints(int32_t n_cams, const double* mtmc_inv, const double* mtmc, const mcs_ocam* cams, const uint8_t* masks,
                          int32_t n_points, const double* world_pos, const double* normal, const double* min_dist,
                          const double* max_dist, const double* scale_factors, int32_t n_levels, uint8_t* in_view, int32_t* level,
                          double* proj_x, double* proj_y, double* view_cos) {
    if (!mtmc_inv || !mtmc || !cams || !masks || !world_pos || !normal || !min_dist || !max_dist || !scale_factors || !in_view ||
        !level || !proj_x || !proj_y || !view_cos)
        return mfail(MCS_ERR_INVALID, "null argument");
    if (n_cams < 1 || n_points < 0 || n_levels < 1) return mfail(MCS_ERR_INVALID, "bad sizes");
    if (n_points == 0) return MCS_OK;
    size_t mask_bytes = 0;
    for (int c = 0; c < n_cams; ++c) {
        if (cams[c].width != cams[0].width || cams[c].height != cams[0].height)
            return mfail(MCS_ERR_UNSUPPORTED, "cameras of different image sizes");
        mask_bytes += (size_t)cams[c].width * cams[c].height;
    }
    const size_t n = (size_t)n_points * n_cams;
    Dev dmi, dm, dc, dk, dp, dn, dmn, dmx, dsf, div, dlv, dpx, dpy, dvc;
    MCK(dmi.alloc(n_cams * 128)); MCK(dm.alloc(n_cams * 128)); MCK(dc.alloc(n_cams * sizeof(mcs_ocam))); MCK(dk.alloc(mask_bytes));
    MCK(dp.alloc((size_t)n_points * 24)); MCK(dn.alloc((size_t)n_points * 24)); MCK(dmn.alloc((size_t)n_points * 8));
    MCK(dmx.alloc((size_t)n_points * 8)); MCK(dsf.alloc(n_levels * 8));
    MCK(div.alloc(n)); MCK(dlv.alloc(n * 4)); MCK(dpx.alloc(n * 8)); MCK(dpy.alloc(n * 8)); MCK(dvc.alloc(n * 8));
    MCK(cudaMemcpy(dmi.p, mtmc_inv, n_cams * 128, cudaMemcpyHostToDevice));
    MCK(cudaMemcpy(dm.p, mtmc, n_cams * 128, cudaMemcpyHostToDevice));
    MCK(cudaMemcpy(dc.p, cams, n_cams * sizeof(mcs_ocam), cudaMemcpyHostToDevice));
    MCK(cudaMemcpy(dk.p, masks, mask_bytes, cudaMemcpyHostToDevice));
    MCK(cudaMemcpy(dp.p, world_pos, (size_t)n_points * 24, cudaMemcpyHostToDevice));
    MCK(cudaMemcpy(dn.p, normal, (size_t)n_points * 24, cudaMemcpyHostToDevice));
    MCK(cudaMemcpy(dmn.p, min_dist, (size_t)n_points * 8, cudaMemcpyHostToDevice));
    MCK(cudaMemcpy(dmx.p, max_dist, (size_t)n_points * 8, cudaMemcpyHostToDevice));
    MCK(cudaMemcpy(dsf.p, scale_factors, n_levels * 8, cudaMemcpyHostToDevice));
    MCK(launch_frustum(n_cams, dmi.as<double>(), dm.as<double>(), dc.as<mcs_ocam>(), dk.as<uint8_t>(), n_points, dp.as<double>(),
                       dn.as<double>(), dmn.as<double>(), dmx.as<double>(), dsf.as<double>(), n_levels, div.as<uint8_t>(),
                       dlv.as<int>(), dpx.as<double>(), dpy.as<double>(), dvc.as<double>(), nullptr));
    MCK(cudaMemcpy(in_view, div.p, n, cudaMemcpyDeviceToHost));
    MCK(cudaMemcpy(level, dlv.p, n * 4, cudaMemcpyDeviceToHost));
    MCK(cudaMemcpy(proj_x, dpx.p, n * 8, cudaMemcpyDeviceToHost));
    MCK(cudaMemcpy(proj_y, dpy.p, n * 8, cudaMemcpyDeviceToHost));
    MCK(cudaMemcpy(view_cos, dvc.p, n * 8, cudaMemcpyDeviceToHost));
    return MCS_OK;
}

int mcs_search_windows(const mcs_frame_view* f, const mcs_window_query* queries, int32_t nq, const uint8_t* qdesc,
                       const uint8_t* qmask, const int32_t* query_tag, int32_t rule, double nnratio, int32_t threshold,
                       int32_t* assigned, int32_t* nmatches) {
    if (!f || !queries || !qdesc || !query_tag || !assigned || !nmatches) return mfail(MCS_ERR_INVALID, "null argument");
    if (rule < 0 || rule > MCS_RULE_SCW) return mfail(MCS_ERR_INVALID, "unknown rule");
    *nmatches = 0;
    if (nq <= 0) return MCS_OK;
    int rows = 0;
    for (int i = 0; i < nq; ++i) {
        if (queries[i].cam < 0 || queries[i].cam >= f->n_cams) return mfail(MCS_ERR_INVALID, "query camera out of range");
        if (query_tag[i] < 0) return mfail(MCS_ERR_INVALID, "query tags must be >= 0");
        if (queries[i].desc_index < 0) return mfail(MCS_ERR_INVALID, "negative query descriptor index");
        rows = std::max(rows, queries[i].desc_index + 1);
    }
    FrameDev fd;
    int rc = upload_frame(f, fd, nullptr);
    if (rc) return rc;
    if (rule == MCS_RULE_SCW) {
        // descriptor row of candidate idx for a query of camera c = row idx of camera c's matrix (ref :2367,2372): needs the
        // camera-major keypoint order of src/cMultiFrame.cpp:168-184
        std::vector<int> first(f->n_cams + 1, 0);
        for (int i = 0; i < f->n_keys; ++i) {
            if (i && f->key_cam[i] < f->key_cam[i - 1]) return mfail(MCS_ERR_INVALID, "MCS_RULE_SCW needs camera-major keypoint order");
            ++first[f->key_cam[i] + 1];
        }
        for (int c = 0; c < f->n_cams; ++c) first[c + 1] += first[c];
        MCK(fd.cam_first.alloc(first.size() * 4));
        MCK(cudaMemcpy(fd.cam_first.p, first.data(), first.size() * 4, cudaMemcpyHostToDevice));
        fd.view.cam_first = fd.cam_first.as<int>();
    }
    std::vector<int> ci, cd, co;
    rc = window_search_compact(fd, queries, nq, qdesc, qmask, rows, f->dim, ci, cd, co, nullptr);
    if (rc) return rc;
    int nm = 0;
    for (int qi = 0; qi < nq; ++qi) {       // sequential greedy replay over the GPU-computed candidate lists
        const int n = co[qi + 1] - co[qi];
        const bool stateless = rule == MCS_RULE_BEST_FREE || rule == MCS_RULE_FIRST_FREE;
        if (n == 0) { if (stateless) assigned[qi] = -1; continue; }
        if (rule == MCS_RULE_FIRST_FREE) {
            // Fuse(pKF, vpMapPoints, th) computes the distance and throws it away (ref :1506-1514: `dist` stays 0), so the first
            // candidate that passes the level filter wins with distance 0 <= TH_LOW_
            assigned[qi] = 0 <= threshold ? ci[co[qi]] : -1;
            nm += assigned[qi] >= 0;
            continue;
        }
        int bestDist = INT_MAX, bestLevel = -1, bestDist2 = INT_MAX, bestLevel2 = -1, bestIdx = -1;
        for (int k = 0; k < n; ++k) {
            const int idx = ci[co[qi] + k];
            if (!stateless && assigned[idx] >= 0) continue;
            const int dist = cd[co[qi] + k];
            if (dist < bestDist) {
                bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel;
                bestLevel = f->keys[idx].octave; bestIdx = idx;
            } else if (dist < bestDist2) {
                bestLevel2 = f->keys[idx].octave; bestDist2 = dist;
            }
        }
        if (rule == MCS_RULE_BEST_FREE) {          // stateless: the answer of query qi
            const bool hit = bestIdx >= 0 && bestDist <= threshold;
            assigned[qi] = hit ? bestIdx : -1;
            nm += hit;
            continue;
        }
        bool ok;
        if (rule == MCS_RULE_RATIO) ok = (double)bestDist <= (double)bestDist2 * nnratio && bestDist <= threshold;
        else if (rule == MCS_RULE_BEST) ok = bestDist <= threshold;
        else if (rule == MCS_RULE_SCW) ok = bestDist <= threshold && bestIdx > 0;          // `bestIdx > 0` as written at ref :2385
        else ok = bestDist <= threshold && !(bestLevel == bestLevel2 && bestDist > nnratio * bestDist2);
        if (ok && bestIdx >= 0) {
            assigned[bestIdx] = query_tag[qi];
            ++nm;
        }
    }
    *nmatches = nm;
    return MCS_OK;
}

int mcs_search_by_projection(const mcs_frame_view* f, const mcs_mappoint_view* mps, double th, double nnratio, int32_t th_high,
                             int32_t having_masks, int32_t* frame_mp, int32_t* nmatches) {
    if (!f || !mps || !frame_mp || !nmatches) return mfail(MCS_ERR_INVALID, "null argument");
    if (having_masks && (!f->dmask || !mps->dmask)) return mfail(MCS_ERR_INVALID, "masks requested but not supplied");
    *nmatches = 0;
    // queries in the reference's visiting order: map point outer, camera inner (ref :74-98)
    std::vector<mcs_window_query> qs;
    std::vector<int> tags;
    {
        size_t n_view = 0;
        const size_t n_all = (size_t)mps->n_points * f->n_cams;
        for (size_t k = 0; k < n_all; ++k) n_view += mps->in_view[k] != 0;
        qs.reserve(n_view); tags.reserve(n_view);
    }
    const bool bFactor = th != 1.0;
    for (int i = 0; i < mps->n_points; ++i) {
        if (mps->bad && mps->bad[i]) continue;
        for (int cam = 0; cam < f->n_cams; ++cam) {
            const size_t k = (size_t)i * f->n_cams + cam;
            if (!mps->in_view[k]) continue;
            const int lvl = mps->level[k];
            if (lvl < 0 || lvl >= f->n_levels) return mfail(MCS_ERR_INVALID, "map point scale level out of range");
            double r = mps->view_cos[k] > 0.998 ? 2.5 : 4.0;   // RadiusByViewingCos (ref :169-175)
            if (bFactor) r *= th;
            mcs_window_query q;
            q.cam = cam; q.min_level = lvl - 1; q.max_level = lvl; q.desc_index = i;
            q.x = mps->proj_x[k]; q.y = mps->proj_y[k]; q.r = r * f->scale_factors[lvl];
            qs.push_back(q);
            tags.push_back(i);
        }
    }
    if (qs.empty()) return MCS_OK;
    mcs_frame_view fv = *f;
    if (!having_masks) fv.dmask = nullptr;
    return mcs_search_windows(&fv, qs.data(), (int)qs.size(), mps->desc, having_masks ? mps->dmask : nullptr, tags.data(),
                              MCS_RULE_LEVEL_RATIO, nnratio, th_high, frame_mp, nmatches);
}

int mcs_search_for_initialization(const mcs_frame_view* f1, const mcs_frame_view* f2, double* prev_matched, int32_t window_size,
                                  double nnratio, int32_t th_low, int32_t having_masks, int32_t* matches12, int32_t* nmatches) {
    if (!f1 || !f2 || !prev_matched || !matches12 || !nmatches) return mfail(MCS_ERR_INVALID, "null argument");
    if (f1->dim != f2->dim) return mfail(MCS_ERR_INVALID, "descriptor sizes differ");
    if (having_masks && (!f1->dmask || !f2->dmask)) return mfail(MCS_ERR_INVALID, "masks requested but not supplied");
    *nmatches = 0;
    const int n1 = f1->n_keys;
    for (int i = 0; i < n1; ++i) matches12[i] = -1;
    if (n1 == 0 || f2->n_keys == 0) return MCS_OK;
    std::vector<mcs_window_query> qs(n1);
    for (int i1 = 0; i1 < n1; ++i1) {
        mcs_window_query& q = qs[i1];
        q.cam = f1->key_cam[i1];
        if (q.cam < 0 || q.cam >= f2->n_cams) return mfail(MCS_ERR_INVALID, "key_cam out of range");
        q.min_level = q.max_level = f1->keys[i1].octave;
        q.desc_index = i1;
        q.x = prev_matched[2 * i1]; q.y = prev_matched[2 * i1 + 1]; q.r = (double)window_size;
    }
    FrameDev fd;
    int rc = upload_frame(f2, fd, nullptr);
    if (rc) return rc;
    std::vector<int> ci, cd, co;
    rc = window_search_compact(fd, qs.data(), n1, f1->desc, having_masks ? f1->dmask : nullptr, n1, f1->dim, ci, cd, co, nullptr);
    if (rc) return rc;
    int nm = 0;
    std::vector<int> matchedDist(f2->n_keys, INT_MAX), matches21(f2->n_keys, -1);
    for (int i1 = 0; i1 < n1; ++i1) {            // ref :599-682
        const int n = co[i1 + 1] - co[i1];
        if (n == 0) continue;
        int bestDist = INT_MAX, bestDist2 = INT_MAX, bestIdx2 = -1;
        for (int k = 0; k < n; ++k) {
            const int i2 = ci[co[i1] + k], dist = cd[co[i1] + k];
            if (matchedDist[i2] <= dist) continue;
            if (dist < bestDist) { bestDist2 = bestDist; bestDist = dist; bestIdx2 = i2; }
            else if (dist < bestDist2) bestDist2 = dist;
        }
        if (bestDist <= th_low && bestDist < (double)bestDist2 * nnratio) {
            if (matches21[bestIdx2] >= 0) { matches12[matches21[bestIdx2]] = -1; --nm; }
            matches12[i1] = bestIdx2;
            matches21[bestIdx2] = i1;
            matchedDist[bestIdx2] = bestDist;
            ++nm;
        }
    }
    for (int i1 = 0; i1 < n1; ++i1)              // ref :717-720
        if (matches12[i1] >= 0) {
            prev_matched[2 * i1] = f2->keys[matches12[i1]].x;
            prev_matched[2 * i1 + 1] = f2->keys[matches12[i1]].y;
        }
    *nmatches = nm;
    return MCS_OK;
}

}  // extern "C"
