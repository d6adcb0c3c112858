// K5: cv2.groupRectangles(rects, min_neighbors, eps = 0.2) per frame, on the GPU.
// Replaces InferenceApp._run_nms (app/inference_app.py:169-217).  Algorithm of OpenCV
// objdetect/cascadedetect.cpp: cv::partition (union-find) on the SimilarRects predicate,
// per-class float32 mean + cvRound, drop classes with <= min_neighbors members, drop small
// classes nested in bigger ones; classes are numbered in order of their first member, and the
// merged box's "confidence" is the class size.
//
// No K x K adjacency bitmask is materialised: the predicate is evaluated on the fly and fed to
// a lock-free union-find whose roots always point to the smaller index (so the root of a class
// IS its first member), which keeps HBM traffic at O(K) instead of O(K^2 / 8).
#include "common.cuh"

namespace rodc {

struct GroupScratch {
    int32_t* seg;        // [F+1] first candidate of each frame
    int32_t* parent;     // [K]
    int32_t* cls;        // [K]   class index (within frame) of each candidate's root
    int32_t* sums;       // [K,5] per-class sums x,y,w,h and count (indexed seg[f] + class)
    int32_t* rrect;      // [K,4] per-class mean rect
    int32_t* ncls;       // [F]   classes per frame
    int32_t* tmp;        // [K,6] per-frame compacted output at seg[f]
    int32_t* fcount;     // [F]
    int32_t* foff;       // [F+1]
};

__host__ __device__ inline int64_t align16(int64_t v) { return (v + 15) / 16 * 16; }

static GroupScratch carve(void* scratch, int64_t K, int F) {
    char* p = reinterpret_cast<char*>(scratch);
    GroupScratch g;
    auto take = [&](int64_t n) { int32_t* r = reinterpret_cast<int32_t*>(p); p += align16(n * 4); return r; };
    g.seg = take(F + 1); g.parent = take(K); g.cls = take(K); g.sums = take(K * 5); g.rrect = take(K * 4);
    g.ncls = take(F); g.tmp = take(K * 6); g.fcount = take(F); g.foff = take(F + 1);
    return g;
}

// seg[f] = lower_bound(frame ids, f), f in [0, F]
__global__ void group_segments_kernel(const int32_t* __restrict__ cand, int64_t K, const int64_t* __restrict__ k_dev, int F,
                                      int32_t* __restrict__ seg) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f > F) return;
    K = eff_count(K, k_dev);
    int64_t lo = 0, hi = K;
    while (lo < hi) { int64_t mid = (lo + hi) >> 1; if (cand[mid * 5] < f) lo = mid + 1; else hi = mid; }
    seg[f] = (int32_t)lo;
}

__global__ void group_init_kernel(int32_t* __restrict__ parent, int32_t* __restrict__ sums, int64_t K, const int64_t* __restrict__ k_dev) {
    K = eff_count(K, k_dev);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < K; i += (int64_t)gridDim.x * blockDim.x) {
        parent[i] = (int32_t)i;
        #pragma unroll
        for (int k = 0; k < 5; ++k) sums[i * 5 + k] = 0;
    }
}

// Path-halving find: every store writes an ancestor (parents only ever decrease), so concurrent
// readers stay correct without synchronisation.
__device__ __forceinline__ int find_root(volatile int32_t* parent, int x) {
    int cur = x;
    while (true) {
        const int p = parent[cur];
        if (p == cur) return cur;
        const int gp = parent[p];
        if (gp != p) parent[cur] = gp;
        cur = gp;
    }
}

__device__ __forceinline__ int unite(int32_t* parent, int a, int b) {
    while (true) {
        a = find_root(parent, a); b = find_root(parent, b);
        if (a == b) return a;
        if (a < b) { int t = a; a = b; b = t; }          // a > b: the larger root links under the smaller
        if (atomicCAS(&parent[a], a, b) == a) return b;
    }
}

// SimilarRects (double arithmetic): all four edge deltas <= eps * (min(w) + min(h)) * 0.5
__device__ __forceinline__ bool similar(const int4 r1, const int4 r2, double eps) {
    const double delta = __dmul_rn(__dmul_rn(eps, (double)(min(r1.z, r2.z) + min(r1.w, r2.w))), 0.5);
    return (double)abs(r1.x - r2.x) <= delta && (double)abs(r1.y - r2.y) <= delta &&
           (double)abs(r1.x + r1.z - r2.x - r2.z) <= delta && (double)abs(r1.y + r1.w - r2.y - r2.w) <= delta;
}

__device__ __forceinline__ int4 load_xywh(const int32_t* cand, int64_t i) {
    const int32_t* c = cand + i * 5;
    return make_int4(c[1], c[2], c[3] - c[1], c[4] - c[2]);       // (x, y, w, h) as _run_nms builds them
}

// Pair scan over a 2-D triangle of 256 x 256 tiles of the (frame-sorted) candidate index space:
// block (bi, bj <= bi) tests candidates i of tile bi against candidates j < i of tile bj staged in
// shared memory.  Tiles that cannot share a frame exit at once, so the grid is sized from K alone
// (no host sync) while the K^2/2 predicate evaluations spread evenly over all SMs.  An integer
// prefilter (a superset of the exact predicate for eps <= 0.5) rejects most pairs before the
// double-precision SimilarRects test; unions are rare and go to the global forest with atomicCAS.
// (A one-CTA-per-frame variant with the forest in shared memory was measured 10x slower: the scan,
// not the atomics, is the cost, and it needs all SMs.)
constexpr int GT = 256;
__global__ void __launch_bounds__(GT) group_union_kernel(const int32_t* __restrict__ cand, int64_t K, const int64_t* __restrict__ k_dev,
                                                         const int32_t* __restrict__ seg, int32_t* __restrict__ parent, double eps) {
    __shared__ int4 tile[GT];
    __shared__ int tframe[GT];
    K = eff_count(K, k_dev);
    // linear block id -> (bi, bj) with bj <= bi
    const long long b = blockIdx.x;
    long long bi = (long long)((sqrt(8.0 * (double)b + 1.0) - 1.0) * 0.5);
    while (bi * (bi + 1) / 2 > b) --bi;
    while ((bi + 1) * (bi + 2) / 2 <= b) ++bi;
    const long long bj = b - bi * (bi + 1) / 2;
    const int64_t i0 = bi * GT, j0 = bj * GT;
    if (i0 >= K) return;                                    // grid sized for an upper bound of K
    // the i tile's first frame starts at seg[f_first]; j tiles that end before it hold other frames only
    const int f_first = cand[i0 * 5];
    if (j0 + GT <= seg[f_first]) return;
    const int64_t jl = j0 + threadIdx.x;
    int4 rj = make_int4(0, 0, 0, 0);
    int fj = -2;
    if (jl < K) { rj = load_xywh(cand, jl); fj = cand[jl * 5]; }
    tile[threadIdx.x] = rj; tframe[threadIdx.x] = fj;
    __syncthreads();
    const int64_t i = i0 + threadIdx.x;
    if (i >= K) return;
    const int4 ri = load_xywh(cand, i);
    const int fi = cand[i * 5];
    // |d| <= eps*(min w + min h)/2 <= (w_i + h_i)/4 for eps <= 0.5; no prefilter for larger eps
    const int reach = eps <= 0.5 ? ri.z + ri.w : (1 << 28);
    const int cnt = (int)min((int64_t)GT, i - j0);           // only j < i
    int my_root = (int)i;
    for (int q = 0; q < cnt; ++q) {
        const int4 r2 = tile[q];
        if (4 * abs(ri.x - r2.x) > reach || 4 * abs(ri.y - r2.y) > reach) continue;
        if (tframe[q] != fi) continue;
        if (!similar(ri, r2, eps)) continue;
        const int rjr = find_root(parent, (int)(j0 + q));
        if (rjr != my_root) my_root = unite(parent, my_root, rjr);
    }
}

constexpr int ELIG_CAP = 1536;      // suppressing classes of a frame kept in shared memory (36 KB); more -> global scan

// one block per frame: class numbering (order of first member = order of roots), sums, means,
// filters, ordered per-frame output into tmp.
__global__ void __launch_bounds__(1024) group_finalize_kernel(const int32_t* __restrict__ cand, const int32_t* __restrict__ seg,
                                                              int32_t* __restrict__ parent, int32_t* __restrict__ cls,
                                                              int32_t* __restrict__ sums, int32_t* __restrict__ rrect,
                                                              int32_t* __restrict__ tmp, int32_t* __restrict__ fcount,
                                                              int min_neighbors, double eps) {
    __shared__ int warp_sum[32];
    __shared__ int carry;
    const int f = blockIdx.x;
    const int s0 = seg[f], s1 = seg[f + 1];
    const int n = s1 - s0;
    if (n <= 0) { if (threadIdx.x == 0) fcount[f] = 0; return; }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    // 1. flatten + number the roots in index order (exclusive scan of is_root)
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        int root = -1, isroot = 0;
        if (i < n) { root = find_root(parent, s0 + i); isroot = (root == s0 + i); }
        const unsigned b = __ballot_sync(0xffffffffu, isroot);
        if (lane == 0) warp_sum[wid] = __popc(b);
        __syncthreads();
        if (threadIdx.x < 32) {
            int s = warp_sum[threadIdx.x];
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, s, o); if (threadIdx.x >= o) s += y; }
            warp_sum[threadIdx.x] = s;
        }
        __syncthreads();
        if (i < n) {
            parent[s0 + i] = root;                                     // flattened
            if (isroot) cls[s0 + i] = carry + (wid ? warp_sum[wid - 1] : 0) + __popc(b & ((1u << lane) - 1u));
        }
        __syncthreads();
        if (threadIdx.x == 0) carry += warp_sum[31];
        __syncthreads();
    }
    const int ncls = carry;
    // 2. per-class integer sums (exact, order independent)
    for (int i = threadIdx.x; i < n; i += 1024) {
        const int c = cls[parent[s0 + i]];
        const int4 r = load_xywh(cand, s0 + i);
        int32_t* sm = sums + (int64_t)(s0 + c) * 5;
        atomicAdd(sm + 0, r.x); atomicAdd(sm + 1, r.y); atomicAdd(sm + 2, r.z); atomicAdd(sm + 3, r.w); atomicAdd(sm + 4, 1);
    }
    __syncthreads();
    // 3. class means: s = 1.f / n; cvRound(float(sum) * s)
    for (int c = threadIdx.x; c < ncls; c += 1024) {
        const int32_t* sm = sums + (int64_t)(s0 + c) * 5;
        const float s = __fdiv_rn(1.f, (float)sm[4]);
        #pragma unroll
        for (int k = 0; k < 4; ++k) rrect[(int64_t)(s0 + c) * 4 + k] = __float2int_rn(__fmul_rn(__int2float_rn(sm[k]), s));
    }
    __syncthreads();
    // 4. filters + ordered output.  Only classes with more than min_neighbors members can suppress another one
    // (typically a tenth of all classes): their grown rectangles are collected in shared memory first, so the
    // nested-rectangle test scans that short list instead of every class of the frame.
    __shared__ int n_elig;
    __shared__ int elig[ELIG_CAP][6];                          // class, x0 - dx, y0 - dy, x1 + dx, y1 + dy, members
    if (threadIdx.x == 0) { carry = 0; n_elig = 0; }
    __syncthreads();
    for (int c = threadIdx.x; c < ncls; c += 1024) {
        const int n2 = sums[(int64_t)(s0 + c) * 5 + 4];
        if (n2 > min_neighbors) {
            const int slot = atomicAdd(&n_elig, 1);
            if (slot < ELIG_CAP) {
                const int32_t* r2 = rrect + (int64_t)(s0 + c) * 4;
                const int dx = __double2int_rn(__dmul_rn((double)r2[2], eps));
                const int dy = __double2int_rn(__dmul_rn((double)r2[3], eps));
                elig[slot][0] = c; elig[slot][1] = r2[0] - dx; elig[slot][2] = r2[1] - dy;
                elig[slot][3] = r2[0] + r2[2] + dx; elig[slot][4] = r2[1] + r2[3] + dy; elig[slot][5] = n2;
            }
        }
    }
    __syncthreads();
    const bool use_list = n_elig <= ELIG_CAP;
    const int n_list = n_elig;
    for (int base = 0; base < ncls; base += 1024) {
        const int c = base + threadIdx.x;
        int keep = 0, n1 = 0;
        int4 r1 = make_int4(0, 0, 0, 0);
        if (c < ncls) {
            n1 = sums[(int64_t)(s0 + c) * 5 + 4];
            const int32_t* rr = rrect + (int64_t)(s0 + c) * 4;
            r1 = make_int4(rr[0], rr[1], rr[2], rr[3]);
            if (n1 > min_neighbors) {
                keep = 1;
                if (use_list) {
                    for (int e = 0; e < n_list; ++e) {
                        const int n2 = elig[e][5];
                        if (elig[e][0] == c) continue;
                        if (r1.x >= elig[e][1] && r1.y >= elig[e][2] && r1.x + r1.z <= elig[e][3] && r1.y + r1.w <= elig[e][4] &&
                            (n2 > max(3, n1) || n1 < 3)) { keep = 0; break; }
                    }
                } else {
                    for (int j = 0; j < ncls; ++j) {
                        const int n2 = sums[(int64_t)(s0 + j) * 5 + 4];
                        if (j == c || n2 <= min_neighbors) continue;
                        const int32_t* r2 = rrect + (int64_t)(s0 + j) * 4;
                        const int dx = __double2int_rn(__dmul_rn((double)r2[2], eps));
                        const int dy = __double2int_rn(__dmul_rn((double)r2[3], eps));
                        if (r1.x >= r2[0] - dx && r1.y >= r2[1] - dy &&
                            r1.x + r1.z <= r2[0] + r2[2] + dx && r1.y + r1.w <= r2[1] + r2[3] + dy &&
                            (n2 > max(3, n1) || n1 < 3)) { keep = 0; break; }
                    }
                }
            }
        }
        const unsigned b = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_sum[wid] = __popc(b);
        __syncthreads();
        if (threadIdx.x < 32) {
            int s = warp_sum[threadIdx.x];
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, s, o); if (threadIdx.x >= o) s += y; }
            warp_sum[threadIdx.x] = s;
        }
        __syncthreads();
        if (keep) {
            const int pos = carry + (wid ? warp_sum[wid - 1] : 0) + __popc(b & ((1u << lane) - 1u));
            int32_t* o = tmp + (int64_t)(s0 + pos) * 6;
            o[0] = f; o[1] = r1.x; o[2] = r1.y; o[3] = r1.z; o[4] = r1.w; o[5] = n1;
        }
        __syncthreads();
        if (threadIdx.x == 0) carry += warp_sum[31];
        __syncthreads();
    }
    if (threadIdx.x == 0) fcount[f] = carry;
}

// single block: exclusive scan of the per-frame counts
__global__ void __launch_bounds__(1024) group_offsets_kernel(const int32_t* __restrict__ fcount, int F,
                                                             int32_t* __restrict__ foff, int64_t* __restrict__ out_count) {
    __shared__ int warp_sum[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < F; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < F ? fcount[i] : 0;
        int x = v;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, x, o); if ((threadIdx.x & 31) >= o) x += y; }
        if ((threadIdx.x & 31) == 31) warp_sum[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            int s = warp_sum[threadIdx.x];
            #pragma unroll
            for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, s, o); if (threadIdx.x >= o) s += y; }
            warp_sum[threadIdx.x] = s;
        }
        __syncthreads();
        const int incl = x + ((threadIdx.x >> 5) ? warp_sum[(threadIdx.x >> 5) - 1] : 0);
        if (i < F) foff[i] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) { foff[F] = carry; *out_count = carry; }
}

__global__ void group_emit_kernel(const int32_t* __restrict__ tmp, const int32_t* __restrict__ seg,
                                  const int32_t* __restrict__ fcount, const int32_t* __restrict__ foff,
                                  int32_t* __restrict__ out) {
    const int f = blockIdx.x;
    const int cnt = fcount[f];
    const int32_t* src = tmp + (int64_t)seg[f] * 6;
    int32_t* dst = out + (int64_t)foff[f] * 6;
    for (int i = threadIdx.x; i < cnt * 6; i += blockDim.x) dst[i] = src[i];
}

// min_neighbors <= 0: groupRectangles returns the input unchanged with weight 1
__global__ void group_passthrough_kernel(const int32_t* __restrict__ cand, int64_t K, const int64_t* __restrict__ k_dev,
                                         int32_t* __restrict__ out, int64_t* __restrict__ out_count) {
    K = eff_count(K, k_dev);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < K; i += (int64_t)gridDim.x * blockDim.x) {
        const int32_t* c = cand + i * 5;
        int32_t* o = out + i * 6;
        o[0] = c[0]; o[1] = c[1]; o[2] = c[2]; o[3] = c[3] - c[1]; o[4] = c[4] - c[2]; o[5] = 1;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *out_count = K;
}

}  // namespace rodc

using namespace rodc;

extern "C" int64_t rodc_group_scratch_bytes(int64_t k, int n_frames) {
    if (k < 0 || n_frames < 0) return -1;
    const int64_t K = k > 0 ? k : 1, F = n_frames > 0 ? n_frames : 1;
    return align16((F + 1) * 4) * 2 + align16(F * 4) * 2 + align16(K * 4) * 2 + align16(K * 20) + align16(K * 16) + align16(K * 24) + 256;
}

extern "C" int rodc_group_rectangles(const int32_t* cand, int64_t k, int n_frames, int min_neighbors, double eps,
                                     int32_t* out, int64_t* out_count, void* scratch, void* stream) {
    return rodc_group_rectangles_dyn(cand, k, nullptr, n_frames, min_neighbors, eps, out, out_count, scratch, stream);
}

extern "C" int rodc_group_rectangles_dyn(const int32_t* cand, int64_t k, const int64_t* k_dev, int n_frames, int min_neighbors,
                                         double eps, int32_t* out, int64_t* out_count, void* scratch, void* stream) {
    RODC_CHECK_ARG(out_count != nullptr, "rodc_group_rectangles: out_count is NULL");
    RODC_CHECK_ARG(k >= 0 && k < (1ll << 30), "rodc_group_rectangles: bad k");
    cudaStream_t st = as_stream(stream);
    if (k == 0) { RODC_CUDA(cudaMemsetAsync(out_count, 0, sizeof(int64_t), st)); return RODC_OK; }
    RODC_CHECK_ARG(cand && out && scratch && n_frames > 0, "rodc_group_rectangles: NULL argument");
    int blocks = (int)((k + 255) / 256);
    if (min_neighbors <= 0) {
        Prof prof(KC_GROUP, st, 0, 44.0 * (double)k, 0.0);
        group_passthrough_kernel<<<blocks > 1184 ? 1184 : blocks, 256, 0, st>>>(cand, k, k_dev, out, out_count);
        RODC_LAUNCH_CHECK();
        return RODC_OK;
    }
    GroupScratch g = carve(scratch, k, n_frames);
    // 20 K (candidates) + 24 K' (groups, K' <= K); the K^2 / 2 predicate evaluations are ALU work
    Prof prof(KC_GROUP, st, 0, 44.0 * (double)k, 0.0, 6);
    group_segments_kernel<<<(n_frames + 1 + 255) / 256, 256, 0, st>>>(cand, k, k_dev, n_frames, g.seg);
    RODC_LAUNCH_CHECK();
    group_init_kernel<<<blocks > 1184 ? 1184 : blocks, 256, 0, st>>>(g.parent, g.sums, k, k_dev);
    RODC_LAUNCH_CHECK();
    {
        const long long nb = (k + GT - 1) / GT;
        const long long tri = nb * (nb + 1) / 2;
        RODC_CHECK_ARG(tri < (1ll << 31), "rodc_group_rectangles: too many candidates for one launch (%lld)", (long long)k);
        group_union_kernel<<<(unsigned)tri, GT, 0, st>>>(cand, k, k_dev, g.seg, g.parent, eps);
    }
    RODC_LAUNCH_CHECK();
    group_finalize_kernel<<<n_frames, 1024, 0, st>>>(cand, g.seg, g.parent, g.cls, g.sums, g.rrect, g.tmp, g.fcount, min_neighbors, eps);
    RODC_LAUNCH_CHECK();
    group_offsets_kernel<<<1, 1024, 0, st>>>(g.fcount, n_frames, g.foff, out_count);
    RODC_LAUNCH_CHECK();
    group_emit_kernel<<<n_frames, 128, 0, st>>>(g.tmp, g.seg, g.fcount, g.foff, out);
    RODC_LAUNCH_CHECK();
    return RODC_OK;
}
