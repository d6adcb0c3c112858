// K1: integer-exact Pillow BILINEAR pyramid.  Every level is resampled from the ORIGINAL frame
// (data/image_info.py:201-250 via scipy.misc.imresize): horizontal pass -> u8 -> vertical pass
// -> u8, 22-bit fixed-point coefficients precomputed on the host (plan.cu).
//
// Data movement: the horizontal kernel stages one source row in shared memory ONCE and emits
// that row of every level (so the frame is read once per batch, not once per scale); the
// vertical kernel streams the horizontal rows with coalesced row-wise accesses.
#include "common.cuh"

namespace rodc {

#define PREC_BITS 22

__device__ __forceinline__ uint8_t clip8(int acc) {
    int v = acc >> PREC_BITS;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// grid: (H, n_frames); block: 256.  dynamic smem: W*3 bytes.
__global__ void __launch_bounds__(256) pyr_hpass_kernel(const uint8_t* __restrict__ frames, int H, int W,
                                                        const LevelInfo* __restrict__ levels, int n_levels,
                                                        const int32_t* __restrict__ tables,
                                                        uint8_t* __restrict__ scratch, int64_t scratch_bytes) {
    extern __shared__ __align__(16) uint8_t row[];
    const int y = blockIdx.x, f = blockIdx.y;
    const int row_bytes = W * 3;
    const uint8_t* src = frames + ((int64_t)f * H + y) * row_bytes;
    // vectorised staging when the row start is 16-byte aligned
    if ((((uintptr_t)src) & 15) == 0) {
        const int nv = row_bytes >> 4;
        const uint4* s4 = reinterpret_cast<const uint4*>(src);
        uint4* d4 = reinterpret_cast<uint4*>(row);
        for (int i = threadIdx.x; i < nv; i += blockDim.x) d4[i] = __ldg(s4 + i);
        for (int i = (nv << 4) + threadIdx.x; i < row_bytes; i += blockDim.x) row[i] = __ldg(src + i);
    } else {
        for (int i = threadIdx.x; i < row_bytes; i += blockDim.x) row[i] = __ldg(src + i);
    }
    __syncthreads();
    uint8_t* sc = scratch + (int64_t)f * scratch_bytes;
    for (int l = 0; l < n_levels; ++l) {
        const LevelInfo lv = levels[l];
        if (lv.identity) continue;
        const int32_t* bounds = tables + lv.hx_bounds;
        const int32_t* coef = tables + lv.hx_coef;
        uint8_t* out = sc + lv.scratch_off + (int64_t)y * lv.w * 3;
        const int nout = lv.w * 3;
        for (int i = threadIdx.x; i < nout; i += blockDim.x) {
            const int xx = i / 3, c = i - xx * 3;
            const int xmin = bounds[2 * xx], cnt = bounds[2 * xx + 1];
            const int32_t* k = coef + (int64_t)xx * lv.kx;
            const uint8_t* px = row + xmin * 3 + c;
            int acc = 1 << (PREC_BITS - 1);
            for (int t = 0; t < cnt; ++t) acc += (int)px[t * 3] * k[t];
            out[i] = clip8(acc);
        }
    }
}

// grid: (total level rows, n_frames); block 128.  One block = one output row of one level.
__global__ void __launch_bounds__(128) pyr_vpass_kernel(const uint8_t* __restrict__ frames, int H, int W,
                                                        const LevelInfo* __restrict__ levels, int n_levels,
                                                        const int32_t* __restrict__ tables,
                                                        const int32_t* __restrict__ row_start,   // [n_levels+1]
                                                        const uint8_t* __restrict__ scratch, int64_t scratch_bytes,
                                                        uint8_t* __restrict__ pyramid, int64_t pyramid_bytes) {
    const int r = blockIdx.x, f = blockIdx.y;
    int lo = 0, hi = n_levels;                       // level l with row_start[l] <= r < row_start[l+1]
    while (hi - lo > 1) { int mid = (lo + hi) >> 1; if (row_start[mid] <= r) lo = mid; else hi = mid; }
    const LevelInfo lv = levels[lo];
    const int yy = r - row_start[lo];
    const int nout = lv.w * 3;
    uint8_t* out = pyramid + (int64_t)f * pyramid_bytes + lv.byte_off + (int64_t)yy * nout;
    if (lv.identity) {
        const uint8_t* src = frames + ((int64_t)f * H + yy) * nout;
        for (int i = threadIdx.x; i < nout; i += blockDim.x) out[i] = __ldg(src + i);
        return;
    }
    const int32_t* bounds = tables + lv.vy_bounds;
    const int ymin = bounds[2 * yy], cnt = bounds[2 * yy + 1];
    const int32_t* k = tables + lv.vy_coef + (int64_t)yy * lv.ky;
    const uint8_t* src = scratch + (int64_t)f * scratch_bytes + lv.scratch_off + (int64_t)ymin * nout;
    for (int i = threadIdx.x; i < nout; i += blockDim.x) {
        int acc = 1 << (PREC_BITS - 1);
        for (int t = 0; t < cnt; ++t) acc += (int)__ldg(src + (int64_t)t * nout + i) * k[t];
        out[i] = clip8(acc);
    }
}

// Raw windows gather (API / parity path; the stage-0 kernel reads the pyramid in place).
__global__ void extract_windows_kernel(const uint8_t* __restrict__ pyramid, int64_t pyramid_bytes,
                                       const LevelInfo* __restrict__ levels, const int32_t* __restrict__ win,
                                       int64_t n_windows, int wsize, int64_t total_rows, uint8_t* __restrict__ out) {
    // one thread per (window, row): copies wsize*3 bytes
    const int row_bytes = wsize * 3;
    for (int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; r < total_rows; r += (int64_t)gridDim.x * blockDim.x) {
        const int64_t w = r / wsize; const int ry = (int)(r - w * wsize);
        const int64_t f = w / n_windows; const int64_t j = w - f * n_windows;
        const int32_t* wr = win + j * 8;
        const LevelInfo& lv = levels[wr[0]];
        const uint8_t* src = pyramid + f * pyramid_bytes + lv.byte_off + ((int64_t)(wr[2] + ry) * lv.w + wr[1]) * 3;
        uint8_t* dst = out + r * row_bytes;
        for (int i = 0; i < row_bytes; ++i) dst[i] = src[i];
    }
}

__global__ void window_meta_kernel(const int32_t* __restrict__ win, int64_t n_windows, int64_t total,
                                   int frame_id_base, int32_t* __restrict__ meta) {
    for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < total; w += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = w / n_windows; const int64_t j = w - f * n_windows;
        const int32_t* wr = win + j * 8;
        int32_t* m = meta + w * 5;
        m[0] = (int32_t)f + frame_id_base; m[1] = wr[4]; m[2] = wr[5]; m[3] = wr[6]; m[4] = wr[7];
    }
}

}  // namespace rodc

using namespace rodc;

extern "C" int rodc_pyramid(rodc_plan* p, const uint8_t* frames, int n_frames, uint8_t* pyramid,
                            uint8_t* scratch, void* stream) {
    RODC_CHECK_ARG(p && frames && pyramid && scratch, "rodc_pyramid: NULL argument");
    RODC_CHECK_ARG(n_frames >= 0, "rodc_pyramid: n_frames < 0");
    if (n_frames == 0 || p->levels.empty()) return RODC_OK;
    int rc = p->ensure_device(); if (rc) return rc;
    cudaStream_t st = as_stream(stream);
    const int nl = (int)p->levels.size();
    bool any_resample = false;
    for (auto& l : p->levels) any_resample |= !l.identity;
    if (any_resample) {
        size_t smem = (size_t)p->img_w * 3 + 16;
        if (smem > 48 * 1024) RODC_CUDA(cudaFuncSetAttribute(pyr_hpass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        dim3 grid(p->img_h, n_frames);
        pyr_hpass_kernel<<<grid, 256, smem, st>>>(frames, p->img_h, p->img_w, p->d_levels, nl, p->d_tables, scratch, p->scratch_bytes);
        RODC_LAUNCH_CHECK();
    }
    dim3 grid(p->total_rows, n_frames);
    pyr_vpass_kernel<<<grid, 128, 0, st>>>(frames, p->img_h, p->img_w, p->d_levels, nl, p->d_tables, p->d_row_start,
                                            scratch, p->scratch_bytes, pyramid, p->pyramid_bytes);
    RODC_LAUNCH_CHECK();
    return RODC_OK;
}

extern "C" int rodc_extract_windows(rodc_plan* p, const uint8_t* pyramid, int n_frames, uint8_t* windows, void* stream) {
    RODC_CHECK_ARG(p && pyramid && windows, "rodc_extract_windows: NULL argument");
    const int64_t N = (int64_t)p->win_level.size();
    if (n_frames <= 0 || N == 0) return RODC_OK;
    RODC_CHECK_ARG(p->fits, "rodc_extract_windows: a window leaves its level (the reference raises ValueError and skips the image)");
    int rc = p->ensure_device(); if (rc) return rc;
    const int64_t rows = N * n_frames * p->win;
    int blocks = (int)((rows + 255) / 256); if (blocks > 148 * 32) blocks = 148 * 32;
    extract_windows_kernel<<<blocks, 256, 0, as_stream(stream)>>>(pyramid, p->pyramid_bytes, p->d_levels, p->d_win, N, p->win, rows, windows);
    RODC_LAUNCH_CHECK();
    return RODC_OK;
}

extern "C" int rodc_window_meta(rodc_plan* p, int n_frames, int frame_id_base, int32_t* meta, void* stream) {
    RODC_CHECK_ARG(p && meta, "rodc_window_meta: NULL argument");
    const int64_t N = (int64_t)p->win_level.size();
    if (n_frames <= 0 || N == 0) return RODC_OK;
    int rc = p->ensure_device(); if (rc) return rc;
    const int64_t total = N * n_frames;
    int blocks = (int)((total + 255) / 256); if (blocks > 148 * 16) blocks = 148 * 16;
    window_meta_kernel<<<blocks, 256, 0, as_stream(stream)>>>(p->d_win, N, total, frame_id_base, meta);
    RODC_LAUNCH_CHECK();
    return RODC_OK;
}
