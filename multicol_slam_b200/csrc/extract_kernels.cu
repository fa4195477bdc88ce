// extract_kernels.cu -- sm_100a kernels of the extractor half of the hot path.
//
//   K1 pyr_fast_kernel   : per pyramid level, fused  resize(level l-1 -> l)  +  5x5 box blur  +
//                          FAST-9/16 score  +  per-cell 3x3 NMS  +  mirror-mask filter  -> raw corners
//                          (ref src/mdBRIEFextractorOct.cpp:1158-1201, :863-949, :1301; OpenCV
//                          resize/boxFilter/FAST arithmetic of SURVEY.md Appendix A.1/A.3/A.5)
//   K2 octree_kernel     : DistributeOctTree, one CTA per (image, level)   (ref :569-861)
//   (K3, orientation + descriptor, lives in describe_kernel.cu)
//
// Everything is integer / exactly-rounded arithmetic; compile with -fmad=false so that no float or
// double expression is contracted (the CPU oracle is built with -ffp-contract=off).
#include "mcs_common.cuh"
#include "cam_model.cuh"
#include "kernels.h"

namespace mcs {

// ------------------------------------------------------------------------------------------------
// K1
// ------------------------------------------------------------------------------------------------
constexpr int kTileStride = 76;    // shared-memory row stride of the level tile (72 used)
constexpr int kScoreStride = 68;   // score tile: 66 used
constexpr int kMaxTileCorners = kTW * kTH / 4;

__device__ __forceinline__ bool has_arc9(uint32_t m) {
    m |= m << 16;
    uint32_t r = m & (m >> 1);
    r &= r >> 2;
    r &= r >> 4;
    r &= m >> 8;
    return (r & 0xFFFFu) != 0;
}

// FAST-9/16 corner test + score of OpenCV (cornerScore<16>): 0 if not a corner, else
// max(threshold, best 9-arc margin) - 1.   p -> centre pixel inside the shared-memory tile.
__device__ __forceinline__ int fast_score(const uint8_t* p, int t) {
    constexpr int S = kTileStride;
    const int v = p[0];
    const int lo = v - t, hi = v + t;
    {   // any 9-arc of the 16-ring contains >= 2 of the 4 compass pixels
        const int c0 = p[3 * S], c4 = p[3], c8 = p[-3 * S], c12 = p[-3];
        const int nb = (c0 > hi) + (c4 > hi) + (c8 > hi) + (c12 > hi);
        const int nd = (c0 < lo) + (c4 < lo) + (c8 < lo) + (c12 < lo);
        if (nb < 2 && nd < 2) return 0;
    }
    int r[16];
    r[0] = p[3 * S];       r[1] = p[3 * S + 1];   r[2] = p[2 * S + 2];   r[3] = p[S + 3];
    r[4] = p[3];           r[5] = p[-S + 3];      r[6] = p[-2 * S + 2];  r[7] = p[-3 * S + 1];
    r[8] = p[-3 * S];      r[9] = p[-3 * S - 1];  r[10] = p[-2 * S - 2]; r[11] = p[-S - 3];
    r[12] = p[-3];         r[13] = p[S - 3];      r[14] = p[2 * S - 2];  r[15] = p[3 * S - 1];
    uint32_t bm = 0, dm = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        bm |= (uint32_t)(r[k] > hi) << k;
        dm |= (uint32_t)(r[k] < lo) << k;
    }
    if (!has_arc9(bm) && !has_arc9(dm)) return 0;
    // Packed s16x2 lanes: lo = v - r (dark margin), hi = r - v (bright margin); one sliding-min tree over
    // the circular ring gives, per start k, the minimum over the 9-arc in both lanes (VIMNMX.S16x2).
    // NOTE: the scalar form max(best, max(mn9, -mx9)) is MISCOMPILED by ptxas 12.9 -O1+ for sm_100a
    // (the negation is dropped when folded into VIMNMX3) -- see tools/ptxas_vimnmx3_repro.cu.
    unsigned q[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) {
        const int d = v - r[k];
        q[k] = ((unsigned)d & 0xFFFFu) | ((unsigned)(-d) << 16);
    }
    unsigned q2[16], q4[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) q2[k] = __vmins2(q[k], q[(k + 1) & 15]);
#pragma unroll
    for (int k = 0; k < 16; ++k) q4[k] = __vmins2(q2[k], q2[(k + 2) & 15]);
    unsigned m = __vmins2(__vmins2(q4[0], q4[4]), q[8]);
#pragma unroll
    for (int k = 1; k < 16; ++k) m = __vmaxs2(m, __vmins2(__vmins2(q4[k], q4[(k + 4) & 15]), q[(k + 8) & 15]));
    const int dark = (int)(short)(m & 0xFFFFu), bright = (int)(short)(m >> 16);
    return max(t, max(dark, bright)) - 1;
}

__global__ void __launch_bounds__(256)
pyr_fast_kernel(const LevelGeom g, const int level, const int nlevels, const int fast_th,
                const uint8_t* __restrict__ src, const size_t src_img_bytes,
                uint8_t* __restrict__ dst, uint8_t* __restrict__ dst_blur,
                const uint8_t* __restrict__ mask0, const int mask_w, const size_t mask_bytes,
                const int* __restrict__ cam_of_image,
                uint32_t* __restrict__ raw, const size_t raw_img_stride, int* __restrict__ raw_count) {
    __shared__ __align__(16) uint8_t s_src[kSrcH * kSrcW];
    __shared__ __align__(16) uint8_t s_tile[kTileH * kTileStride];
    __shared__ __align__(16) uint16_t s_hsum[(kTH + 4) * kTW];
    __shared__ __align__(16) uint8_t s_score[(kTH + 2) * kScoreStride];
    __shared__ int16_t s_cellx[kTW + 2], s_celly[kTH + 2];
    __shared__ uint32_t s_list[kMaxTileCorners];
    __shared__ int s_n, s_base;

    const int tid = threadIdx.x;
    const int b = blockIdx.z;
    const int X0 = blockIdx.x * kTW, Y0 = blockIdx.y * kTH;
    const uint8_t* simg = src + (size_t)b * src_img_bytes;

    // output-space range that this tile needs (after reflection everything lies inside it)
    const int xa = max(X0 - kHalo, 0), xb = min(X0 + kTW + kHalo, g.w) - 1;
    const int ya = max(Y0 - kHalo, 0), yb = min(Y0 + kTH + kHalo, g.h) - 1;
    int sx_lo, sx_hi, sy_lo, sy_hi;
    if (level == 0) { sx_lo = xa; sx_hi = xb; sy_lo = ya; sy_hi = yb; }
    else {
        sx_lo = g.xofs[xa]; sx_hi = min(g.xofs[xb] + 1, g.sw - 1);
        sy_lo = min(max((int)g.yofs[ya], 0), g.sh - 1); sy_hi = min(max(g.yofs[yb] + 1, 0), g.sh - 1);
    }
    if (tid == 0) s_n = 0;
    // ---- stage the source region (rows are contiguous: coalesced byte loads, L2-resident source) ----
    {
        const int sw = sx_hi - sx_lo + 1, sh = sy_hi - sy_lo + 1;
        for (int i = tid; i < sw * sh; i += 256) {
            const int yy = i / sw, xx = i - yy * sw;
            s_src[yy * kSrcW + xx] = simg[(size_t)(sy_lo + yy) * g.spitch + sx_lo + xx];
        }
    }
    if (tid < kTW + 2) { const int x = X0 - 1 + tid; s_cellx[tid] = (x >= 0 && x < g.w) ? g.cellx[x] : (int16_t)-1; }
    if (tid >= 128 && tid < 128 + kTH + 2) { const int y = Y0 - 1 + tid - 128; s_celly[tid - 128] = (y >= 0 && y < g.h) ? g.celly[y] : (int16_t)-1; }
    __syncthreads();

    // ---- bilinear resize into the haloed level tile (OpenCV fixed-point arithmetic) ----
    for (int i = tid; i < kTileW * kTileH; i += 256) {
        const int ty = i / kTileW, tx = i - ty * kTileW;
        const int rx = reflect101(X0 - kHalo + tx, g.w), ry = reflect101(Y0 - kHalo + ty, g.h);
        int v = 0;
        if (rx >= xa && rx <= xb && ry >= ya && ry <= yb) {
            if (level == 0) {
                v = s_src[(ry - sy_lo) * kSrcW + rx - sx_lo];
            } else {
                const int sx = g.xofs[rx], sx1 = min(sx + 1, g.sw - 1);
                const int sy = min(max((int)g.yofs[ry], 0), g.sh - 1), sy1 = min(max(g.yofs[ry] + 1, 0), g.sh - 1);
                const int a0 = g.xa0[rx], a1 = g.xa1[rx], b0 = g.yb0[ry], b1 = g.yb1[ry];
                const uint8_t* r0 = s_src + (sy - sy_lo) * kSrcW - sx_lo;
                const uint8_t* r1 = s_src + (sy1 - sy_lo) * kSrcW - sx_lo;
                const int h0 = r0[sx] * a0 + r0[sx1] * a1;
                const int h1 = r1[sx] * a0 + r1[sx1] * a1;
                v = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
                v = min(max(v, 0), 255);
            }
        }
        s_tile[ty * kTileStride + tx] = (uint8_t)v;
    }
    __syncthreads();

    uint8_t* dimg = dst + (size_t)b * g.img_bytes;
    uint8_t* bimg = dst_blur + (size_t)b * g.img_bytes;
    // ---- store the unblurred tile; horizontal 5-sums for the blur ----
    for (int i = tid; i < kTH * (kTW / 4); i += 256) {
        const int y = i / (kTW / 4), x4 = (i - y * (kTW / 4)) * 4;
        if (Y0 + y < g.h && X0 + x4 < g.pitch) {
            const uint8_t* p = s_tile + (y + kHalo) * kTileStride + x4 + kHalo;
            const uint32_t wv = p[0] | (p[1] << 8) | (p[2] << 16) | ((uint32_t)p[3] << 24);
            *(uint32_t*)(dimg + (size_t)(Y0 + y) * g.pitch + X0 + x4) = wv;
        }
    }
    for (int i = tid; i < (kTH + 4) * kTW; i += 256) {
        const int y = i / kTW, x = i - y * kTW;
        const uint8_t* p = s_tile + (y + kHalo - 2) * kTileStride + x + kHalo;
        s_hsum[i] = (uint16_t)(p[-2] + p[-1] + p[0] + p[1] + p[2]);
    }
    // ---- FAST score on the tile + 1 ring (only inside FAST cell interiors) ----
    for (int i = tid; i < (kTH + 2) * (kTW + 2); i += 256) {
        const int y = i / (kTW + 2), x = i - y * (kTW + 2);
        int s = 0;
        if (s_cellx[x] >= 0 && s_celly[y] >= 0)
            s = fast_score(s_tile + (y + kHalo - 1) * kTileStride + x + kHalo - 1, fast_th);
        s_score[y * kScoreStride + x] = (uint8_t)s;
    }
    __syncthreads();

    // ---- blurred tile ----
    for (int i = tid; i < kTH * (kTW / 4); i += 256) {
        const int y = i / (kTW / 4), x4 = (i - y * (kTW / 4)) * 4;
        if (Y0 + y < g.h && X0 + x4 < g.pitch) {
            uint32_t wv = 0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint16_t* h = s_hsum + y * kTW + x4 + k;
                const int s = h[0] + h[kTW] + h[2 * kTW] + h[3 * kTW] + h[4 * kTW];
                wv |= (uint32_t)((s + 12) / 25) << (8 * k);
            }
            *(uint32_t*)(bimg + (size_t)(Y0 + y) * g.pitch + X0 + x4) = wv;
        }
    }
    // ---- per-cell 3x3 non-max suppression, mask filter, tile-local compaction ----
    const uint8_t* m0 = mask0 + (size_t)cam_of_image[b] * mask_bytes;
    for (int i = tid; i < kTH * kTW; i += 256) {
        const int y = i / kTW, x = i - y * kTW;
        const uint8_t* sc = s_score + (y + 1) * kScoreStride + x + 1;
        const int s = sc[0];
        if (s == 0) continue;
        const int cx = s_cellx[x + 1], cy = s_celly[y + 1];
        bool keep = true;
#pragma unroll
        for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
            for (int dx = -1; dx <= 1; ++dx) {
                if (dx == 0 && dy == 0) continue;
                const bool same = (s_cellx[x + 1 + dx] == cx) && (s_celly[y + 1 + dy] == cy);
                const int sn = same ? sc[dy * kScoreStride + dx] : 0;
                keep = keep && (s > sn);
            }
        if (!keep) continue;
        const int gx = X0 + x, gy = Y0 + y;
        if (m0[(size_t)g.my0[gy] * mask_w + g.mx0[gx]] == 0) continue;
        const int pos = atomicAdd(&s_n, 1);
        s_list[pos] = pack_corner(gx, gy, s);
    }
    __syncthreads();
    const int n = s_n;
    if (n == 0) return;
    if (tid == 0) s_base = atomicAdd(&raw_count[b * nlevels + level], n);
    __syncthreads();
    uint32_t* rlist = raw + (size_t)b * raw_img_stride + g.raw_off;
    for (int i = tid; i < n; i += 256) {
        const int pos = s_base + i;
        if (pos < g.raw_cap) rlist[pos] = s_list[i];
    }
}

// ------------------------------------------------------------------------------------------------
// K2  octree  (ref src/mdBRIEFextractorOct.cpp:569-861)
// ------------------------------------------------------------------------------------------------
// One CTA owns one (image, level).  The std::list<ExtractorNode> of the reference is kept as a dense
// array in LIST ORDER (index == position from the front); each pass rebuilds the array:
//   new list = [children created this pass, in reverse creation order] ++ [surviving nodes, old order]
// which is exactly what push_front of n1..n4 + erase(parent) produce.  Corners carry the list position
// of their node (node_of[]) and are re-labelled through a per-pass remap table.
//   pass A ("full"): every node with more than one corner is divided, in list order (ref :693-765).
//   pass B ("sorted"): the children created by the previous pass that hold >1 corner are divided
//        largest-first, ties by later creation first (deterministic stand-in for the pointer order of
//        the reference's sort, :782), stopping as soon as the list holds >= N nodes (ref :775-851).
// The best corner of a node is the first maximum of `response` in raw-list (= reference) order; raw lists
// are unordered on the device, so the order is carried by an analytic key (cell-row-major, then
// pixel-row-major inside the cell -- ref :892-949).
struct OctNode {
    short x0, y0, x1, y1;     // UL.x, UL.y, BR.x, BR.y  (cell-grid coordinates)
    int count;
    unsigned seq;             // creation sequence number | bit31: created in the last pass (pass-B candidate)
};
constexpr unsigned kNewFlag = 0x80000000u;

constexpr int kOctThreads = 512;
constexpr int kChunk = kMaxNodes / kOctThreads;   // list positions owned by one thread in the scans

// kOctThreads-wide exclusive scan; every thread of the CTA must call.
__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp, int& total) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();
    if (lane == 31) s_warp[wid] = inc;
    __syncthreads();
    if (wid == 0) {
        int w = lane < kOctThreads / 32 ? s_warp[lane] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        if (lane < kOctThreads / 32) s_warp[lane] = w;
    }
    __syncthreads();
    total = s_warp[kOctThreads / 32 - 1];
    const int r = inc - v + (wid ? s_warp[wid - 1] : 0);
    __syncthreads();
    return r;
}

__device__ __forceinline__ int quadrant(const OctNode& n, int x, int y) {
    // DivideNode (ref :575-619): halfX = ceil((UR.x-UL.x)/2);  n1 | n2 / n3 | n4
    const int mx = n.x0 + ((n.x1 - n.x0 + 1) >> 1);
    const int my = n.y0 + ((n.y1 - n.y0 + 1) >> 1);
    return (x < mx ? 0 : 1) + (y < my ? 0 : 2);
}
__device__ __forceinline__ int nonempty4(const int* c) { return (c[0] > 0) + (c[1] > 0) + (c[2] > 0) + (c[3] > 0); }

size_t octree_smem_bytes(int cap) { return (size_t)cap * (8 + 2 * sizeof(OctNode) + 16 + 2) + 64; }

__global__ void __launch_bounds__(kOctThreads)
octree_kernel(const PyramidGeom* __restrict__ geom, const int cap /* node capacity, <= kMaxNodes */,
              const uint32_t* __restrict__ raw, const size_t raw_img_stride,
              const int* __restrict__ raw_count, uint16_t* __restrict__ node_of_all,
              uint32_t* __restrict__ sel_xys /* [B][sel_total] packed corner */, int* __restrict__ sel_count /* [B][L] */,
              int* __restrict__ status) {
    extern __shared__ __align__(16) unsigned char oct_smem[];
    unsigned long long* s_key = (unsigned long long*)oct_smem;          // pass-B sort keys / scan scratch / best-corner keys
    OctNode* s_nodes0 = (OctNode*)(s_key + cap);
    OctNode* s_nodes1 = s_nodes0 + cap;
    int (*s_child)[4] = (int (*)[4])(s_nodes1 + cap);                   // corner count per child, then new position per child
    short* s_rank = (short*)(s_child + cap);                            // processing rank of a divided node, -1 = survives
    __shared__ int s_warp[kOctThreads / 32];
    __shared__ int s_nn, s_cut;
    __shared__ unsigned s_seq;

    const int level = blockIdx.x, b = blockIdx.y;
    const LevelGeom& g = geom->lv[level];
    const int L = geom->nlevels;
    const int tid = threadIdx.x;
    const int n_all = raw_count[b * L + level];
    const int n = min(n_all, g.raw_cap);
    const uint32_t* corners = raw + (size_t)b * raw_img_stride + g.raw_off;
    uint16_t* node_of = node_of_all + (size_t)b * raw_img_stride + g.raw_off;
    uint32_t* out = sel_xys + (size_t)b * geom->sel_total + g.sel_off;
    const int N = g.quota;

    if (n == 0) { if (tid == 0) sel_count[b * L + level] = 0; return; }
    if (n_all > g.raw_cap && tid == 0) atomicOr(status, 1);

    const int minB = kEdge - 3;
    // ---- roots (ref :640-679) ----
    const int nIni = g.nodes_ini;
    for (int i = tid; i < nIni; i += kOctThreads) s_child[i][0] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += kOctThreads) {
        const int x = corner_x(corners[i]) - minB;
        const int r = min((int)((double)(float)x / g.hX), nIni - 1);
        node_of[i] = (uint16_t)r;
        atomicAdd(&s_child[r][0], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int k = 0;
        for (int i = 0; i < nIni; ++i) {
            s_rank[i] = -1;
            if (s_child[i][0] == 0) continue;
            OctNode nd;
            nd.x0 = (short)(int)(g.hX * (double)i); nd.x1 = (short)(int)(g.hX * (double)(i + 1));
            nd.y0 = 0; nd.y1 = (short)(g.h - kEdge + 3 - minB);
            nd.count = s_child[i][0]; nd.seq = (unsigned)i;
            s_rank[i] = (short)k;
            s_nodes0[k++] = nd;
        }
        s_nn = k; s_seq = (unsigned)nIni;
    }
    __syncthreads();
    for (int i = tid; i < n; i += kOctThreads) node_of[i] = (uint16_t)s_rank[node_of[i]];
    __syncthreads();

    int cur = 0;
    bool finish = false, pass_b = false;
    while (!finish) {
        const int nn = s_nn;
        OctNode* nodes = cur ? s_nodes1 : s_nodes0;
        OctNode* nnodes = cur ? s_nodes0 : s_nodes1;
        // 1. corner counts per child of every node that may be divided in this pass
        for (int i = tid; i < nn * 4; i += kOctThreads) (&s_child[0][0])[i] = 0;
        __syncthreads();
        for (int i = tid; i < n; i += kOctThreads) {
            const int p = node_of[i];
            const OctNode nd = nodes[p];
            if (nd.count > 1 && (!pass_b || (nd.seq & kNewFlag))) {
                const uint32_t c = corners[i];
                atomicAdd(&s_child[p][quadrant(nd, corner_x(c) - minB, corner_y(c) - minB)], 1);
            }
        }
        __syncthreads();
        // 2. s_rank[p] = processing rank among the divided nodes (-1: survives), ndiv = how many
        int ndiv;
        if (!pass_b) {          // list order (ref :693-765)
            int flag[kChunk], sum = 0;
#pragma unroll
            for (int k = 0; k < kChunk; ++k) {
                const int p = tid * kChunk + k;
                flag[k] = (p < nn && nodes[p].count > 1) ? 1 : 0;
                sum += flag[k];
            }
            int base = block_exclusive_scan(sum, s_warp, ndiv);
#pragma unroll
            for (int k = 0; k < kChunk; ++k) {
                const int p = tid * kChunk + k;
                if (p < nn) s_rank[p] = flag[k] ? (short)base : (short)-1;
                base += flag[k];
            }
            __syncthreads();
        } else {                // largest first, later-created first among equals (ref :782-786)
            for (int p = tid; p < nn; p += kOctThreads) {
                const OctNode nd = nodes[p];
                s_key[p] = (nd.count > 1 && (nd.seq & kNewFlag))
                               ? (((unsigned long long)(unsigned)nd.count << 32) | (nd.seq & ~kNewFlag)) : 0ull;
            }
            __syncthreads();
            int mine = 0;
            for (int p = tid; p < nn; p += kOctThreads) {
                const unsigned long long kp = s_key[p];
                int r = -1;
                if (kp) { r = 0; for (int q = 0; q < nn; ++q) r += (s_key[q] > kp); ++mine; }
                s_rank[p] = (short)r;
            }
            int ncand;
            block_exclusive_scan(mine, s_warp, ncand);
            // growth of the list per candidate, in processing order; cut after the divide reaching N (ref :848-849)
            for (int p = tid; p < nn; p += kOctThreads) {
                const int r = s_rank[p];
                if (r >= 0) s_key[r] = (unsigned long long)(nonempty4(s_child[p]) - 1);
            }
            __syncthreads();
            if (tid == 0) {
                int size = nn, cut = ncand;
                for (int r = 0; r < ncand; ++r) {
                    size += (int)s_key[r];
                    if (size >= N) { cut = r + 1; break; }
                }
                s_cut = cut;
            }
            __syncthreads();
            ndiv = s_cut;
            for (int p = tid; p < nn; p += kOctThreads)
                if (s_rank[p] >= ndiv) s_rank[p] = -1;
            __syncthreads();
        }
        // 3. creation index of the first child of rank r  (children of rank r follow all children of ranks < r)
        for (int p = tid; p < nn; p += kOctThreads) {
            const int r = s_rank[p];
            if (r >= 0) s_key[r] = (unsigned long long)nonempty4(s_child[p]);
        }
        __syncthreads();
        int K;                  // number of nodes created by this pass
        {
            int loc[kChunk], sum = 0;
#pragma unroll
            for (int k = 0; k < kChunk; ++k) {
                const int r = tid * kChunk + k;
                loc[k] = r < ndiv ? (int)s_key[r] : 0;
                sum += loc[k];
            }
            int base = block_exclusive_scan(sum, s_warp, K);
#pragma unroll
            for (int k = 0; k < kChunk; ++k) {
                const int r = tid * kChunk + k;
                if (r < ndiv) s_key[r] = (unsigned long long)base;
                base += loc[k];
            }
            __syncthreads();
        }
        // 4. new list = [new children, reverse creation order] ++ [survivors, old order]
        int nsurv;
        {
            int loc[kChunk], sum = 0;
#pragma unroll
            for (int k = 0; k < kChunk; ++k) {
                const int p = tid * kChunk + k;
                loc[k] = (p < nn && s_rank[p] < 0) ? 1 : 0;
                sum += loc[k];
            }
            int base = block_exclusive_scan(sum, s_warp, nsurv);
            const unsigned seq0 = s_seq;
#pragma unroll
            for (int k = 0; k < kChunk; ++k) {
                const int p = tid * kChunk + k;
                if (p >= nn) continue;
                const OctNode nd = nodes[p];
                if (loc[k]) {
                    const int np = K + base;
                    OctNode c = nd; c.seq &= ~kNewFlag;
                    if (np < cap) nnodes[np] = c;
                    s_child[p][0] = s_child[p][1] = s_child[p][2] = s_child[p][3] = np;
                    ++base;
                } else {
                    int t = (int)s_key[s_rank[p]];
                    const int mx = nd.x0 + ((nd.x1 - nd.x0 + 1) >> 1);
                    const int my = nd.y0 + ((nd.y1 - nd.y0 + 1) >> 1);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const int cnt = s_child[p][q];
                        if (cnt == 0) { s_child[p][q] = 0; continue; }
                        OctNode c;
                        c.x0 = (q & 1) ? (short)mx : nd.x0; c.x1 = (q & 1) ? nd.x1 : (short)mx;
                        c.y0 = (q & 2) ? (short)my : nd.y0; c.y1 = (q & 2) ? nd.y1 : (short)my;
                        c.count = cnt; c.seq = (seq0 + (unsigned)t) | kNewFlag;
                        const int np = K - 1 - t;
                        if (np < cap) nnodes[np] = c;
                        s_child[p][q] = np;
                        ++t;
                    }
                }
            }
        }
        __syncthreads();
        const int new_nn = K + nsurv;
        // 5. relabel the corners
        for (int i = tid; i < n; i += kOctThreads) {
            const int p = node_of[i];
            int q = 0;
            if (s_rank[p] >= 0) {
                const uint32_t c = corners[i];
                q = quadrant(nodes[p], corner_x(c) - minB, corner_y(c) - minB);
            }
            node_of[i] = (uint16_t)s_child[p][q];
        }
        // 6. loop control (ref :767-775, :851-854)
        int n_to_expand;
        {
            int cnt = 0;
            for (int p = tid; p < min(K, cap); p += kOctThreads) cnt += (nnodes[p].count > 1);
            block_exclusive_scan(cnt, s_warp, n_to_expand);
        }
        if (tid == 0) { s_nn = min(new_nn, cap); s_seq += (unsigned)K; }
        cur ^= 1;
        if (new_nn > cap) { if (tid == 0) atomicOr(status, 2); finish = true; }
        if (new_nn >= N || new_nn == nn) finish = true;
        else if (!pass_b && new_nn + 3 * n_to_expand > N) pass_b = true;
        __syncthreads();
    }

    // ---- best corner per node: first maximum of response in reference raw order (ref :857-874) ----
    const int nn = s_nn;
    for (int p = tid; p < nn; p += kOctThreads) s_key[p] = 0ull;
    __syncthreads();
    const int cell_area = g.w_cell * g.h_cell;
    for (int i = tid; i < n; i += kOctThreads) {
        const uint32_t c = corners[i];
        const int x = corner_x(c) - kEdge, y = corner_y(c) - kEdge;       // relative to the first cell interior
        const int cj = x / g.w_cell, ci = y / g.h_cell;
        const unsigned order = (unsigned)((ci * g.n_cols + cj) * cell_area + (y - ci * g.h_cell) * g.w_cell + (x - cj * g.w_cell));
        const unsigned long long key = ((unsigned long long)corner_s(c) << 56) | ((unsigned long long)(0xFFFFFFu - order) << 32) | c;
        atomicMax(&s_key[node_of[i]], key);
    }
    __syncthreads();
    for (int p = tid; p < nn; p += kOctThreads)
        if (p < g.sel_cap) out[p] = (uint32_t)(s_key[p] & 0xFFFFFFFFull);
    if (tid == 0) {
        sel_count[b * L + level] = min(nn, g.sel_cap);
        if (nn > g.sel_cap) atomicOr(status, 4);
    }
}

// ------------------------------------------------------------------------------------------------
// host-callable launchers
// ------------------------------------------------------------------------------------------------
void launch_pyr_fast(const PyramidGeom& G, int level, int n_images, const uint8_t* src, size_t src_img_bytes,
                     uint8_t* dst, uint8_t* dst_blur, const uint8_t* mask0, int mask_w, size_t mask_bytes,
                     const int* cam_of_image, uint32_t* raw, int* raw_count, cudaStream_t st) {
    const LevelGeom& g = G.lv[level];
    dim3 grid(g.tiles_x, g.tiles_y, n_images);
    pyr_fast_kernel<<<grid, 256, 0, st>>>(g, level, G.nlevels, G.fast_threshold, src, src_img_bytes, dst, dst_blur, mask0,
                                          mask_w, mask_bytes, cam_of_image, raw, G.raw_total, raw_count);
}

cudaError_t launch_octree(const PyramidGeom& G, const PyramidGeom* G_dev, int n_images, const uint32_t* raw,
                          const int* raw_count, uint16_t* node_of, uint32_t* sel_xys, int* sel_count, int* status,
                          cudaStream_t st) {
    int cap = 0;
    for (int l = 0; l < G.nlevels; ++l) cap = max(cap, G.lv[l].quota + 8);
    cap = (cap + 31) & ~31;
    if (cap > kMaxNodes) return cudaErrorInvalidValue;
    const size_t smem = octree_smem_bytes(cap);
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(octree_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        configured = smem;
    }
    dim3 grid(G.nlevels, n_images);
    octree_kernel<<<grid, kOctThreads, smem, st>>>(G_dev, cap, raw, G.raw_total, raw_count, node_of, sel_xys, sel_count, status);
    return cudaSuccess;
}

}  // namespace mcs
