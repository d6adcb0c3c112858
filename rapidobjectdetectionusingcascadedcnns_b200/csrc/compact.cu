// K3: score-threshold rejection + stable survivor compaction.
// Replaces the per-row Python filter of the cascade loop (app/inference_cascade_app.py:219-247):
//   keep window w  iff  (double)softmax[w,1] > threshold   (strict; float32 probability promoted
//   to float64 against the Python float threshold, NumPy-1.14 scalar rules), order preserved.
// Survivors carry: metadata row (frame id + *_norm box), bottleneck row (gathered into the next
// stage's concat buffer) and the accumulated confidence (LAST_STEP / AVG / MULT, :226-232).
//
// Three small launches, no host sync: (1) warp-ballot counts per 1024-window tile, (2) one-block
// exclusive scan of the tile counts (+ total), (3) ballot/popc ranks -> scatter of ids, metadata
// and confidences, then (4) a row gather of the bottlenecks with 16-byte accesses.
#include "common.cuh"

namespace rodc {

constexpr int CT = 1024;     // windows per tile (one block of 1024 threads, one window per thread)

// thr == NaN selects the single-net rule of app/inference_app.py:140-149: keep iff np.argmax(p) != 0, i.e.
// p[1] > p[0] (a tie goes to index 0 = background) -- not the same as p[1] > 0.5 when fp32 rounding gives
// p[1] == 0.5 > p[0].
__device__ __forceinline__ bool keep_pred(const float* probs, int64_t w, int64_t n, double thr) {
    if (w >= n) return false;
    if (thr != thr) return probs[2 * w + 1] > probs[2 * w];
    return (double)probs[2 * w + 1] > thr;
}

__global__ void __launch_bounds__(CT) compact_count_kernel(const float* __restrict__ probs, int64_t n, const int64_t* __restrict__ n_dev,
                                                           double thr, int32_t* __restrict__ tile_count) {
    __shared__ int warp_cnt[CT / 32];
    n = eff_count(n, n_dev);
    const int64_t w = (int64_t)blockIdx.x * CT + threadIdx.x;
    const unsigned b = __ballot_sync(0xffffffffu, keep_pred(probs, w, n, thr));
    if ((threadIdx.x & 31) == 0) warp_cnt[threadIdx.x >> 5] = __popc(b);
    __syncthreads();
    if (threadIdx.x < 32) {
        int v = warp_cnt[threadIdx.x];
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) tile_count[blockIdx.x] = v;
    }
}

// single block: exclusive scan of tile_count[0..nt) in place -> tile_off, total -> count_out
__global__ void __launch_bounds__(1024) compact_scan_kernel(int32_t* __restrict__ tile_count, int nt, int64_t* __restrict__ count_out) {
    __shared__ int warp_sum[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    for (int base = 0; base < nt; base += 1024) {
        const int i = base + threadIdx.x;
        const int v = i < nt ? tile_count[i] : 0;
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
        const int wprefix = (threadIdx.x >> 5) ? warp_sum[(threadIdx.x >> 5) - 1] : 0;
        const int incl = x + wprefix;
        if (i < nt) tile_count[i] = carry + incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry += incl;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count_out = carry;
}

__global__ void __launch_bounds__(CT) compact_scatter_kernel(const float* __restrict__ probs, int64_t n, const int64_t* __restrict__ n_dev,
                                                             double thr, const int32_t* __restrict__ tile_off,
                                                             const int32_t* __restrict__ meta_in, int32_t* __restrict__ meta_out,
                                                             const float* __restrict__ conf_acc, float* __restrict__ conf_out, int conf_mode,
                                                             int32_t* __restrict__ keep_idx) {
    __shared__ int warp_cnt[CT / 32];
    n = eff_count(n, n_dev);
    const int64_t w = (int64_t)blockIdx.x * CT + threadIdx.x;
    const bool keep = keep_pred(probs, w, n, thr);
    const unsigned b = __ballot_sync(0xffffffffu, keep);
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) warp_cnt[wid] = __popc(b);
    __syncthreads();
    if (threadIdx.x < 32) {
        int s = warp_cnt[threadIdx.x];
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { int y = __shfl_up_sync(0xffffffffu, s, o); if (threadIdx.x >= o) s += y; }
        warp_cnt[threadIdx.x] = s;                   // inclusive
    }
    __syncthreads();
    if (!keep) return;
    const int pos = tile_off[blockIdx.x] + (wid ? warp_cnt[wid - 1] : 0) + __popc(b & ((1u << lane) - 1u));
    if (keep_idx) keep_idx[pos] = (int32_t)w;
    if (meta_in) {
        #pragma unroll
        for (int k = 0; k < 5; ++k) meta_out[(int64_t)pos * 5 + k] = meta_in[w * 5 + k];
    }
    if (conf_out) {
        const float p = probs[2 * w + 1];
        float c = p;
        if (conf_acc && conf_mode == 1) c = conf_acc[w] + p;
        else if (conf_acc && conf_mode == 2) c = conf_acc[w] * p;
        conf_out[pos] = c;
    }
}

// one warp per surviving row; count read from device memory (no host sync)
__global__ void __launch_bounds__(256) compact_gather_rows_kernel(const int32_t* __restrict__ keep_idx, const int64_t* __restrict__ count,
                                                                  const float* __restrict__ in, int ld_in, int width,
                                                                  float* __restrict__ out, int ld_out, int col_off) {
    const int64_t cnt = *count;
    const int lane = threadIdx.x & 31;
    const int64_t warps = (int64_t)gridDim.x * (blockDim.x >> 5);
    for (int64_t r = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); r < cnt; r += warps) {
        const float4* src = reinterpret_cast<const float4*>(in + (int64_t)keep_idx[r] * ld_in);
        float4* dst = reinterpret_cast<float4*>(out + r * ld_out + col_off);
        for (int k = lane; k < (width >> 2); k += 32) dst[k] = __ldg(src + k);
    }
}

// measurement only (profile level 2): the survivor count of this compaction for the byte bookkeeping
__global__ void compact_note_count_kernel(const int64_t* __restrict__ count, unsigned long long* __restrict__ slot) {
    *slot = (unsigned long long)*count;
}

}  // namespace rodc

using namespace rodc;

extern "C" int64_t rodc_compact_scratch_bytes(int64_t n) {
    if (n < 0) return -1;
    return ((n + CT - 1) / CT + 1) * (int64_t)sizeof(int32_t) + 64;
}

extern "C" int rodc_compact(const float* probs, int64_t n, double threshold,
                            const int32_t* meta_in, int32_t* meta_out,
                            const float* bott_in, int ld_in, int width, float* bott_out, int ld_out, int col_off,
                            const float* conf_acc_in, float* conf_out, int conf_mode,
                            int32_t* keep_idx_out, int64_t* count_out, void* scratch, void* stream) {
    return rodc_compact_dyn(probs, n, nullptr, threshold, meta_in, meta_out, bott_in, ld_in, width, bott_out, ld_out, col_off,
                            conf_acc_in, conf_out, conf_mode, keep_idx_out, count_out, scratch, stream);
}

extern "C" int rodc_compact_dyn(const float* probs, int64_t n, const int64_t* n_dev, double threshold,
                                const int32_t* meta_in, int32_t* meta_out,
                                const float* bott_in, int ld_in, int width, float* bott_out, int ld_out, int col_off,
                                const float* conf_acc_in, float* conf_out, int conf_mode,
                                int32_t* keep_idx_out, int64_t* count_out, void* scratch, void* stream) {
    RODC_CHECK_ARG(n_dev != count_out, "rodc_compact: the input count and the output count must be different buffers");
    RODC_CHECK_ARG(probs && count_out && scratch, "rodc_compact: NULL argument");
    RODC_CHECK_ARG(n >= 0 && n < (1ll << 31), "rodc_compact: bad n");
    RODC_CHECK_ARG((meta_in == nullptr) == (meta_out == nullptr), "rodc_compact: meta_in/meta_out must both be set or both NULL");
    RODC_CHECK_ARG(conf_mode >= 0 && conf_mode <= 2, "rodc_compact: conf_mode must be 0, 1 or 2");
    cudaStream_t st = as_stream(stream);
    if (n == 0) { RODC_CUDA(cudaMemsetAsync(count_out, 0, sizeof(int64_t), st)); return RODC_OK; }
    if (bott_in) {
        RODC_CHECK_ARG(bott_out && keep_idx_out, "rodc_compact: bottleneck gather needs bott_out and keep_idx_out");
        RODC_CHECK_ARG(width % 4 == 0 && ld_in % 4 == 0 && ld_out % 4 == 0 && col_off % 4 == 0,
                       "rodc_compact: bottleneck widths/strides must be multiples of 4 floats");
        RODC_CHECK_ARG(width > 0 && width <= ld_in && col_off >= 0 && col_off + width <= ld_out,
                       "rodc_compact: bottleneck row [%d, %d) does not fit the rows (ld_in %d, ld_out %d)",
                       col_off, col_off + width, ld_in, ld_out);
    }
    int32_t* tile = reinterpret_cast<int32_t*>(scratch);
    const int nt = (int)((n + CT - 1) / CT);
    // 8 n (score reads of the two passes) + per SURVIVOR 44 B (meta r + w, index, confidence) + 8 D_i (row
    // gather, r + w); the survivor count is data dependent and is noted on the device when the scope is recorded
    Prof prof(KC_COMPACT, st, width, 8.0 * (double)n, 0.0, bott_in ? 4 : 3);
    compact_count_kernel<<<nt, CT, 0, st>>>(probs, n, n_dev, threshold, tile);
    RODC_LAUNCH_CHECK();
    compact_scan_kernel<<<1, 1024, 0, st>>>(tile, nt, count_out);
    RODC_LAUNCH_CHECK();
    compact_scatter_kernel<<<nt, CT, 0, st>>>(probs, n, n_dev, threshold, tile, meta_in, meta_out, conf_acc_in, conf_out, conf_mode, keep_idx_out);
    RODC_LAUNCH_CHECK();
    if (bott_in) {
        int64_t blocks = (n + 7) / 8; if (blocks > 148 * 8) blocks = 148 * 8;
        compact_gather_rows_kernel<<<(unsigned)blocks, 256, 0, st>>>(keep_idx_out, count_out, bott_in, ld_in, width, bott_out, ld_out, col_off);
        RODC_LAUNCH_CHECK();
    }
    if (prof.recording()) {
        if (unsigned long long* slot = prof.area_slot(44.0 + (bott_in ? 8.0 * width : 0.0)))
            compact_note_count_kernel<<<1, 1, 0, st>>>(count_out, slot);
    }
    return RODC_OK;
}
