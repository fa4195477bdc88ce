// bow_kernels.cu -- bag-of-words kernels
//   bow_descend_kernel : TemplatedVocabulary::transform(feature, word, weight, nid, levelsup)
//                        (ref ThirdParty/DBoW2/DBoW2/TemplatedVocabulary.h:1218-1261, FORB::distance FORB.cpp:84-104):
//                        16 lanes per descriptor, lane j scores child j of the current node, the group takes the first minimum.
//   group_distance_kernel : the distance half of cORBmatcher::SearchByBoW(KF, F) (ref src/cORBmatcher.cpp:205-262): one warp
//                        per key-frame keypoint, distances to every frame keypoint of the same vocabulary node, in list order.
#include "mcs_common.cuh"
#include "kernels.h"

namespace mcs {

constexpr int kBowGroup = 16;

__global__ void __launch_bounds__(256)
bow_descend_kernel(const VocabularyDev v, const uint4* __restrict__ desc, const int n, const int nid_level,
                   int* __restrict__ word, double* __restrict__ weight, int* __restrict__ node_out) {
    const int gi = (blockIdx.x * blockDim.x + threadIdx.x) / kBowGroup;          // descriptor of this group
    if (gi >= n) return;                                                        // whole groups leave together
    const int lane = threadIdx.x & 31, gl = lane & (kBowGroup - 1);
    const unsigned gmask = 0xFFFFu << (lane & ~(kBowGroup - 1));
    const uint4 qa = __ldg(desc + 2 * (size_t)gi), qb = __ldg(desc + 2 * (size_t)gi + 1);
    int node = 0, level = 0, at_level = nid_level <= 0 ? 0 : -1;
    for (;;) {
        const int c0 = __ldg(v.child_off + node), nc = __ldg(v.child_off + node + 1) - c0;
        if (nc == 0) break;                                                     // leaf (the root always has children)
        ++level;
        unsigned best = 0xFFFFFFFFu;                                            // distance << 16 | child position: first minimum wins
        for (int j = gl; j < nc; j += kBowGroup) {
            const int id = __ldg(v.child_ids + c0 + j);
            const uint4 da = __ldg(v.desc + 2 * (size_t)id), db = __ldg(v.desc + 2 * (size_t)id + 1);
            const unsigned d = __popc(qa.x ^ da.x) + __popc(qa.y ^ da.y) + __popc(qa.z ^ da.z) + __popc(qa.w ^ da.w) +
                               __popc(qb.x ^ db.x) + __popc(qb.y ^ db.y) + __popc(qb.z ^ db.z) + __popc(qb.w ^ db.w);
            best = min(best, (d << 16) | (unsigned)j);
        }
#pragma unroll
        for (int o = kBowGroup / 2; o; o >>= 1) best = min(best, __shfl_xor_sync(gmask, best, o, kBowGroup));
        node = __ldg(v.child_ids + c0 + (int)(best & 0xFFFFu));
        if (level == nid_level) at_level = node;
    }
    if (gl == 0) {
        if (word) word[gi] = __ldg(v.word_of_node + node);
        if (weight) weight[gi] = __ldg(v.weight + node);
        if (node_out) node_out[gi] = at_level >= 0 ? at_level : node;            // shallower leaf: see include/mcs_b200.h
    }
}

cudaError_t launch_bow_descend(const VocabularyDev& v, const uint8_t* desc, int n, int levelsup, int* word, double* weight,
                               int* node, cudaStream_t st) {
    if (n <= 0) return cudaSuccess;
    const long long threads = (long long)n * kBowGroup;
    bow_descend_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(v, (const uint4*)desc, n, v.L - levelsup, word, weight, node);
    return cudaGetLastError();
}

template <int WORDS, bool MASKED>
__global__ void __launch_bounds__(256)
group_distance_kernel(const GroupQuery* __restrict__ queries, const int nq, const uint32_t* __restrict__ desc1,
                      const uint32_t* __restrict__ mask1, const uint32_t* __restrict__ desc2, const uint32_t* __restrict__ mask2,
                      const int* __restrict__ cand, int* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int qi = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (qi >= nq) return;
    const GroupQuery q = queries[qi];
    uint32_t qw[WORDS], qm[MASKED ? WORDS : 1];
#pragma unroll
    for (int k = 0; k < WORDS; ++k) {
        qw[k] = __ldg(desc1 + (size_t)q.feature * WORDS + k);
        if (MASKED) qm[k] = __ldg(mask1 + (size_t)q.feature * WORDS + k);
    }
    for (int j = lane; j < q.cand_count; j += 32) {
        const int id = __ldg(cand + q.cand_start + j);
        const uint32_t* dd = desc2 + (size_t)id * WORDS;
        unsigned dist = 0;
        if (MASKED) {
            const uint32_t* mm = mask2 + (size_t)id * WORDS;
#pragma unroll
            for (int k = 0; k < WORDS; ++k) {
                const uint32_t xw = qw[k] ^ __ldg(dd + k);
                dist += __popc(xw & qm[k]) + __popc(xw & __ldg(mm + k));
            }
            dist >>= 1;                                                         // DescriptorDistance64Masked (ref :2452-2474)
        } else {
#pragma unroll
            for (int k = 0; k < WORDS; ++k) dist += __popc(qw[k] ^ __ldg(dd + k));
        }
        out[q.out_off + j] = (int)dist;
    }
}

cudaError_t launch_group_distance(const GroupQuery* queries, int nq, const uint8_t* desc1, const uint8_t* mask1, const uint8_t* desc2,
                                  const uint8_t* mask2, const int* cand, int dim, int* out, cudaStream_t st) {
    if (nq <= 0) return cudaSuccess;
    const int blocks = (int)(((long long)nq * 32 + 255) / 256);
    const bool masked = mask1 && mask2;
#define MCS_GD(W, M) group_distance_kernel<W, M><<<blocks, 256, 0, st>>>(queries, nq, (const uint32_t*)desc1, (const uint32_t*)mask1, \
        (const uint32_t*)desc2, (const uint32_t*)mask2, cand, out)
    if (dim == 16) { if (masked) MCS_GD(4, true); else MCS_GD(4, false); }
    else if (dim == 32) { if (masked) MCS_GD(8, true); else MCS_GD(8, false); }
    else if (dim == 64) { if (masked) MCS_GD(16, true); else MCS_GD(16, false); }
    else return cudaErrorInvalidValue;
#undef MCS_GD
    return cudaGetLastError();
}

}  // namespace mcs
