// K1: integer-exact Pillow BILINEAR pyramid.  Every level is resampled from the ORIGINAL frame
// (data/image_info.py:201-250 via scipy.misc.imresize): horizontal pass -> u8 -> vertical pass
// -> u8, 22-bit fixed-point coefficients precomputed on the host (plan.cu).
//
// Data movement: the horizontal kernel brings four source rows into shared memory ONCE with a TMA bulk copy,
// re-lays them planar per channel and emits those rows of every level (so the frame is read once per batch, not once per
// scale, and the coefficient table once per four rows) with packed dp4a arithmetic; the vertical
// kernel streams the horizontal rows with coalesced row-wise accesses.
#include "common.cuh"

namespace rodc {

#define PREC_BITS 22

__device__ __forceinline__ uint8_t clip8(int acc) {
    int v = acc >> PREC_BITS;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// Horizontal pass.  grid: (ceil(H / HROWS), n_frames); block: 256; dynamic smem: HROWS x 3 planar channel rows of
// the source.  Work items are the (level, output x) pairs of ALL levels, flattened (so narrow levels do not idle
// lanes).  The taps of an item are walked as aligned 4-pixel words: one LDS.32 per channel and word, three __dp4a
// per channel (the 22-bit Pillow coefficient is split into three u8 limbs, so the integer sum is exact:
// acc = a0 + (a1 << 8) + (a2 << 16)) -- ~1 instruction per multiply-add instead of 5.  A CTA handles HROWS source
// rows so that every coefficient word (the 158 KB table is the same for all rows and streams from L2) is fetched
// once per HROWS rows.
constexpr int HROWS = 4;
__global__ void __launch_bounds__(256) pyr_hpass_kernel(const uint8_t* __restrict__ frames, int H, int W,
                                                        const int32_t* __restrict__ items, int n_items,
                                                        const uint4* __restrict__ coef,
                                                        uint8_t* __restrict__ scratch, int64_t scratch_bytes) {
    extern __shared__ __align__(16) uint8_t rowsm[];
    __shared__ __align__(8) uint64_t bar;
    const int y0 = blockIdx.x * HROWS, f = blockIdx.y;
    const int wpad = ((W + 3) & ~3) + 4;                    // words past W only ever meet zero coefficients
    // 1. The HROWS source rows are ONE contiguous run of the frame: a single TMA bulk copy (cp.async.bulk, UBLKCP)
    //    brings them into shared memory -- each source pixel is fetched from HBM once per batch, by the copy engine,
    //    not by per-thread byte loads.  The run is widened to 16-byte boundaries (the copy's granularity): only
    //    16-byte words that hold a byte of the rows are read, so nothing outside the frames buffer is touched.
    uint8_t* raw = rowsm + (size_t)HROWS * 3 * wpad;        // [<= HROWS * W * 3 + 30], 16-byte aligned (wpad % 4 == 0, HROWS = 4)
    const int nrows = min(HROWS, H - y0);
    const uint8_t* src = frames + ((int64_t)f * H + y0) * W * 3;
    const int lead = (int)(reinterpret_cast<uintptr_t>(src) & 15);
    const uint32_t nbytes = (uint32_t)((lead + nrows * W * 3 + 15) & ~15);
    const uint32_t bar_a = (uint32_t)__cvta_generic_to_shared(&bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" :: "r"(bar_a));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar_a), "r"(nbytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     :: "r"((uint32_t)__cvta_generic_to_shared(raw)), "l"(reinterpret_cast<uint64_t>(src - lead)), "r"(nbytes), "r"(bar_a)
                     : "memory");
    }
    __syncthreads();                                        // barrier initialised before anybody polls it
    {
        uint32_t done;
        do {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(bar_a) : "memory");
        } while (!done);
    }
    // 2. interleaved RGB -> planar channel rows (the dp4a taps want four pixels of ONE channel per word)
    for (int r = 0; r < HROWS; ++r) {
        uint8_t* rs = rowsm + (size_t)r * 3 * wpad;
        const uint8_t* q0 = raw + lead + (size_t)min(r, nrows - 1) * W * 3;   // rows past H repeat row H-1 and are not stored
        for (int x = threadIdx.x; x < W; x += blockDim.x) {
            const uint8_t* q = q0 + x * 3;
            rs[x] = q[0]; rs[wpad + x] = q[1]; rs[2 * wpad + x] = q[2];
        }
        for (int x = W + threadIdx.x; x < wpad; x += blockDim.x) { rs[x] = 0; rs[wpad + x] = 0; rs[2 * wpad + x] = 0; }
    }
    __syncthreads();
    const uint32_t* p = reinterpret_cast<const uint32_t*>(rowsm);
    const int wq = wpad / 4;                                // words per channel row
    uint8_t* sc = scratch + (int64_t)f * scratch_bytes;
    for (int it = threadIdx.x; it < n_items; it += blockDim.x) {
        const int4 m = __ldg(reinterpret_cast<const int4*>(items) + it * 2);          // word0, n_words, coef word, out byte
        const int lw = __ldg(items + it * 8 + 4);
        const uint4* k = coef + m.z;
        uint32_t a[HROWS][3][3];
        #pragma unroll
        for (int r = 0; r < HROWS; ++r)
            #pragma unroll
            for (int ch = 0; ch < 3; ++ch) { a[r][ch][0] = 0u; a[r][ch][1] = 0u; a[r][ch][2] = 0u; }
        for (int wi = 0; wi < m.y; ++wi) {
            const uint4 c = __ldg(k + wi);
            #pragma unroll
            for (int r = 0; r < HROWS; ++r)
                #pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const uint32_t v = p[(r * 3 + ch) * wq + m.x + wi];
                    a[r][ch][0] = __dp4a(v, c.x, a[r][ch][0]); a[r][ch][1] = __dp4a(v, c.y, a[r][ch][1]); a[r][ch][2] = __dp4a(v, c.z, a[r][ch][2]);
                }
        }
        #pragma unroll
        for (int r = 0; r < HROWS; ++r) {
            if (y0 + r >= H) break;
            uint8_t* out = sc + m.w + (int64_t)(y0 + r) * lw * 3;
            #pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const int acc = (int)(a[r][ch][0] + (a[r][ch][1] << 8) + (a[r][ch][2] << 16)) + (1 << (PREC_BITS - 1));
                out[ch] = clip8(acc);
            }
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

__global__ void window_gather_rows_kernel(const int32_t* __restrict__ win, const LevelInfo* __restrict__ levels,
                                          int64_t n_windows, int64_t total, int64_t base_off, int64_t pyramid_bytes,
                                          int64_t* __restrict__ win_off, int32_t* __restrict__ win_pitch) {
    for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < total; w += (int64_t)gridDim.x * blockDim.x) {
        const int64_t f = w / n_windows; const int64_t j = w - f * n_windows;
        const int32_t* wr = win + j * 8;
        const LevelInfo& lv = levels[wr[0]];
        win_off[w] = base_off + f * pyramid_bytes + lv.byte_off + ((int64_t)wr[2] * lv.w + wr[1]) * 3;
        win_pitch[w] = lv.w * 3;
    }
}

// Level index of every stage-0 window (+ slot_base): the key into a rodc_scaled_slot table.
__global__ void window_levels_kernel(const int32_t* __restrict__ win, int64_t n_windows, int64_t total,
                                     int slot_base, int32_t* __restrict__ out) {
    for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < total; w += (int64_t)gridDim.x * blockDim.x)
        out[w] = win[(w % n_windows) * 8] + slot_base;
}

// out[i] = in[idx[i]] for rows of row_words 32-bit words (survivor bookkeeping next to rodc_compact).
__global__ void select_rows_kernel(const uint32_t* __restrict__ in, const int32_t* __restrict__ idx, int64_t n,
                                   int row_words, uint32_t* __restrict__ out) {
    const int64_t total = n * row_words;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / row_words; const int c = (int)(t - i * row_words);
        out[t] = __ldg(in + (int64_t)__ldg(idx + i) * row_words + c);
    }
}

// Resize-then-crop branch of the next-stage input (app/inference_cascade_app.py:374-394): the crop starts at
// (int(xmin_norm * next_scale), int(ymin_norm * next_scale)) of the frame resized to next_scale and is
// size x size; Window.raw raises ValueError when it leaves that image (data/rectangles.py:286-289) -> *err.
__global__ void window_scaled_rows_kernel(const int32_t* __restrict__ meta, const int32_t* __restrict__ slot, int64_t n,
                                          const rodc_scaled_slot* __restrict__ slots, int n_slots, int size,
                                          int64_t dense_rel_off, int64_t* __restrict__ win_off,
                                          int32_t* __restrict__ win_pitch, int32_t* __restrict__ err) {
    for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < n; w += (int64_t)gridDim.x * blockDim.x) {
        const int s = slot[w];
        if (s < 0 || s >= n_slots) { atomicOr(err, 2); win_off[w] = 0; win_pitch[w] = 0; continue; }
        const rodc_scaled_slot sl = slots[s];
        if (sl.individually) {                 // this window keeps the crop + cv2.resize input (materialised by the caller)
            win_off[w] = dense_rel_off + w * (int64_t)size * size * 3;
            win_pitch[w] = size * 3;
            continue;
        }
        const int32_t* m = meta + w * 5;
        const int nx = (int)((double)m[1] * sl.scale), ny = (int)((double)m[2] * sl.scale);
        if (nx + size > sl.width || ny + size > sl.height) { atomicOr(err, 1); win_off[w] = sl.base_off; win_pitch[w] = 0; continue; }
        win_off[w] = sl.base_off + (int64_t)(m[0] - sl.frame_base) * sl.frame_stride + ((int64_t)ny * sl.width + nx) * 3;
        win_pitch[w] = sl.width * 3;
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
    // planar rows + the raw rows of the bulk copy (widened to 16-byte boundaries on both sides)
    const size_t smem = (size_t)HROWS * 3 * ((((size_t)p->img_w + 3) & ~(size_t)3) + 4) + (size_t)HROWS * p->img_w * 3 + 32;
    if (any_resample && smem > 48 * 1024)
        RODC_CUDA(cudaFuncSetAttribute(pyr_hpass_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int64_t frame_bytes = (int64_t)p->img_h * p->img_w * 3;
    for (int f0 = 0; f0 < n_frames; f0 += 65535) {              // gridDim.y limit
        const int nf = n_frames - f0 < 65535 ? n_frames - f0 : 65535;
        const uint8_t* fr = frames + (int64_t)f0 * frame_bytes;
        uint8_t* sc = scratch + (int64_t)f0 * p->scratch_bytes;
        uint8_t* py = pyramid + (int64_t)f0 * p->pyramid_bytes;
        // SURVEY §8d books K1 as sum over levels of (3 W H source read + 3 w_k h_k level write) per frame: the h-pass
        // gets the source term, the v-pass the level term
        double lv_bytes = 0.0; int n_res = 0;
        for (auto& l : p->levels) { lv_bytes += 3.0 * l.w * l.h; n_res += l.identity ? 0 : 1; }
        if (any_resample) {
            dim3 grid((p->img_h + HROWS - 1) / HROWS, nf);
            Prof prof(KC_PYR_H, st, 0, (double)nf * n_res * (double)frame_bytes, 0.0);
            pyr_hpass_kernel<<<grid, 256, smem, st>>>(fr, p->img_h, p->img_w, p->d_hitems, p->n_hitems,
                                                       reinterpret_cast<const uint4*>(p->d_hcoef), sc, p->scratch_bytes);
            RODC_LAUNCH_CHECK();
        }
        dim3 grid(p->total_rows, nf);
        Prof prof(KC_PYR_V, st, 0, (double)nf * lv_bytes, 0.0);
        pyr_vpass_kernel<<<grid, 128, 0, st>>>(fr, p->img_h, p->img_w, p->d_levels, nl, p->d_tables, p->d_row_start,
                                                sc, p->scratch_bytes, py, p->pyramid_bytes);
        RODC_LAUNCH_CHECK();
    }
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
    Prof prof(KC_SOURCE, as_stream(stream), p->win, 2.0 * (double)N * n_frames * p->win * p->win * 3, 0.0);
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
    Prof prof(KC_WINTAB, as_stream(stream), 0, 52.0 * (double)total, 0.0);
    window_meta_kernel<<<blocks, 256, 0, as_stream(stream)>>>(p->d_win, N, total, frame_id_base, meta);
    RODC_LAUNCH_CHECK();
    return RODC_OK;
}

extern "C" int rodc_window_gather_rows(rodc_plan* p, int n_frames, int64_t pyramid_base_off, int64_t* win_off,
                                       int32_t* win_pitch, void* stream) {
    RODC_CHECK_ARG(p && win_off && win_pitch, "rodc_window_gather_rows: NULL argument");
    const int64_t N = (int64_t)p->win_level.size();
    if (n_frames <= 0 || N == 0) return RODC_OK;
    RODC_CHECK_ARG(p->fits, "rodc_window_gather_rows: a window leaves its level (the reference raises ValueError and skips the image)");
    int rc = p->ensure_device(); if (rc) return rc;
    const int64_t total = N * n_frames;
    int blocks = (int)((total + 255) / 256); if (blocks > 148 * 16) blocks = 148 * 16;
    Prof prof(KC_WINTAB, as_stream(stream), 0, 44.0 * (double)total, 0.0);
    window_gather_rows_kernel<<<blocks, 256, 0, as_stream(stream)>>>(p->d_win, p->d_levels, N, total, pyramid_base_off,
                                                                     p->pyramid_bytes, win_off, win_pitch);
    RODC_LAUNCH_CHECK();
    return RODC_OK;
}

extern "C" int rodc_window_levels(rodc_plan* p, int n_frames, int slot_base, int32_t* out, void* stream) {
    RODC_CHECK_ARG(p && out, "rodc_window_levels: NULL argument");
    const int64_t N = (int64_t)p->win_level.size();
    if (n_frames <= 0 || N == 0) return RODC_OK;
    int rc = p->ensure_device(); if (rc) return rc;
    const int64_t total = N * n_frames;
    int blocks = (int)((total + 255) / 256); if (blocks > 148 * 16) blocks = 148 * 16;
    Prof prof(KC_WINTAB, as_stream(stream), 0, 36.0 * (double)total, 0.0);
    window_levels_kernel<<<blocks, 256, 0, as_stream(stream)>>>(p->d_win, N, total, slot_base, out);
    RODC_LAUNCH_CHECK();
    return RODC_OK;
}

extern "C" int rodc_upload_frames(const void* const* host_frames, const int64_t* nbytes, const int64_t* dev_off, int n,
                                  uint8_t* dev_base, void* stream) {
    RODC_CHECK_ARG(host_frames && nbytes && dev_off && dev_base && n >= 0, "rodc_upload_frames: bad argument");
    cudaStream_t st = as_stream(stream);
    for (int i = 0; i < n; ++i) {
        RODC_CHECK_ARG(host_frames[i] && nbytes[i] >= 0 && dev_off[i] >= 0, "rodc_upload_frames: bad frame %d", i);
        RODC_CUDA(cudaMemcpyAsync(dev_base + dev_off[i], host_frames[i], (size_t)nbytes[i], cudaMemcpyHostToDevice, st));
    }
    return RODC_OK;
}

extern "C" int rodc_select_rows(const void* in, const int32_t* idx, int64_t n, int row_bytes, void* out, void* stream) {
    RODC_CHECK_ARG(in && idx && out, "rodc_select_rows: NULL argument");
    RODC_CHECK_ARG(row_bytes > 0 && (row_bytes & 3) == 0, "rodc_select_rows: row_bytes must be a positive multiple of 4 (got %d)", row_bytes);
    if (n <= 0) return RODC_OK;
    const int64_t total = n * (row_bytes / 4);
    int blocks = (int)((total + 255) / 256); if (blocks > 148 * 16) blocks = 148 * 16;
    Prof prof(KC_MISC, as_stream(stream), 0, (double)n * (2.0 * row_bytes + 4.0), 0.0);
    select_rows_kernel<<<blocks, 256, 0, as_stream(stream)>>>(static_cast<const uint32_t*>(in), idx, n, row_bytes / 4,
                                                              static_cast<uint32_t*>(out));
    RODC_LAUNCH_CHECK();
    return RODC_OK;
}

extern "C" int rodc_window_scaled_rows(const int32_t* meta, const int32_t* slot, int64_t n, const rodc_scaled_slot* slots,
                                       int n_slots, int size, int64_t dense_rel_off, int64_t* win_off, int32_t* win_pitch,
                                       int32_t* err, void* stream) {
    RODC_CHECK_ARG(meta && slot && slots && win_off && win_pitch && err, "rodc_window_scaled_rows: NULL argument");
    RODC_CHECK_ARG(size > 0 && n_slots > 0, "rodc_window_scaled_rows: bad size %d / %d slots", size, n_slots);
    if (n <= 0) return RODC_OK;
    int blocks = (int)((n + 255) / 256); if (blocks > 148 * 16) blocks = 148 * 16;
    Prof prof(KC_MISC, as_stream(stream), 0, (double)n * 36.0, 0.0);
    window_scaled_rows_kernel<<<blocks, 256, 0, as_stream(stream)>>>(meta, slot, n, slots, n_slots, size, dense_rel_off,
                                                                     win_off, win_pitch, err);
    RODC_LAUNCH_CHECK();
    return RODC_OK;
}
