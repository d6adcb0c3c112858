// K2a (+ fused K4): per-window conv3x3(3->C)+bias+ReLU -> maxpool3x3/1 SAME -> NHWC flatten.
//
// Graph: network/net.py:172-240 (conv_layer_0, pool_layer_0), network/net_builder.py:6-60.
// SAME padding is per window: zeros for the conv; the pool ignores padding, which for
// post-ReLU (>= 0) data equals zero padding.  Raw 0..255 pixels are fed without any
// normalisation (app/inference_cascade_app.py:200-202).
//
// Window sources (the u8 window never round-trips through HBM on the fast path):
//   PYRAMID  stage 0 reads its 12x12 window in place from the packed pyramid (K1 output)
//   RECROP   stages >= 1 crop the survivor's *_norm box from the ORIGINAL frame and apply
//            cv2.resize INTER_LINEAR (8U fixed point, 2x INTER_AREA shortcut) on the fly
//            (app/inference_cascade_app.py:363-373) -- this is K4 fused as the producer
//   DENSE    caller-materialised windows (run_inference_on_windows / run_inference_on_raw_data)
//
// Output: fp32 features (SIMT fc1 / tests) or the split-fp16 A operand of the tensor fc1:
//   f * 2^a_exp = hi + lo  (hi, lo fp16; 22 significant bits), planes [n, K] K-contiguous.
#include "stage.cuh"

namespace rodc {

// ------------------------------------------------------------------ cv2.resize arithmetic
// imgproc/resize.cpp: fx = (float)((dx+0.5)*scale_x - 0.5); sx = cvFloor(fx); fx -= sx;
// explicit _rn intrinsics keep nvcc from contracting mul+sub into an fma (CPU does not).
__device__ __forceinline__ void cv_tap(int d, int ssize, int dsize, bool xaxis, int& s, int& a0, int& a1) {
    const double inv = __ddiv_rn((double)dsize, (double)ssize);
    const double scale = __ddiv_rn(1.0, inv);
    const double t = __dsub_rn(__dmul_rn((double)d + 0.5, scale), 0.5);
    float f = __double2float_rn(t);
    int si = (int)floorf(f);
    f = __fsub_rn(f, (float)si);
    if (xaxis) {
        if (si < 0) { f = 0.f; si = 0; }
        if (si >= ssize - 1) { f = 0.f; si = ssize - 1; }
    }
    a0 = __float2int_rn(__fmul_rn(__fsub_rn(1.f, f), 2048.f));     // saturate_cast<short>(cbuf*2048)
    a1 = __float2int_rn(__fmul_rn(f, 2048.f));
    s = si;
}

struct Crop { const uint8_t* base; int pitch; int sw, sh; };      // base -> crop origin, pitch = W*3

__device__ __forceinline__ Crop make_crop(const SrcDev& src, int64_t w) {
    const int32_t* m = src.meta + w * 5;
    const int f = m[0];
    const int H = src.frame_hw[2 * f], W = src.frame_hw[2 * f + 1];
    int x0 = m[1], y0 = m[2], x1 = m[3], y1 = m[4];
    // the reference raises if the box leaves the frame (data/rectangles.py:302-305); the planner
    // never produces such boxes -- clamp so a bad table cannot read out of bounds.
    x0 = max(0, min(x0, W)); x1 = max(x0, min(x1, W)); y0 = max(0, min(y0, H)); y1 = max(y0, min(y1, H));
    Crop c;
    c.pitch = W * 3;
    c.base = src.frames_base + src.frame_off[f] + ((int64_t)y0 * W + x0) * 3;
    c.sw = x1 - x0; c.sh = y1 - y0;
    return c;
}

// One resized pixel (3 channels) of cv2.resize(crop, (S,S)).  tabs: xs,xa0,xa1,ys,yb0,yb1 [S each].
__device__ __forceinline__ void recrop_pixel(const Crop& c, int S, const int* tabs, int dy, int dx, int out[3]) {
    if (c.sw <= 0 || c.sh <= 0) { out[0] = out[1] = out[2] = 0; return; }
    if (c.sw == S && c.sh == S) {                                   // identity: copy
        const uint8_t* p = c.base + (int64_t)dy * c.pitch + dx * 3;
        out[0] = p[0]; out[1] = p[1]; out[2] = p[2];
        return;
    }
    if (c.sw == 2 * S && c.sh == 2 * S) {                           // INTER_AREA 2x2 fast path
        const uint8_t* p = c.base + (int64_t)(2 * dy) * c.pitch + (2 * dx) * 3;
        #pragma unroll
        for (int ch = 0; ch < 3; ++ch)
            out[ch] = ((int)p[ch] + (int)p[3 + ch] + (int)p[c.pitch + ch] + (int)p[c.pitch + 3 + ch] + 2) >> 2;
        return;
    }
    const int sx = tabs[dx], a0 = tabs[S + dx], a1 = tabs[2 * S + dx];
    const int sy = tabs[3 * S + dy], b0 = tabs[4 * S + dy], b1 = tabs[5 * S + dy];
    const int sx1 = min(sx + 1, c.sw - 1);
    const int y0 = min(max(sy, 0), c.sh - 1), y1 = min(max(sy + 1, 0), c.sh - 1);
    const uint8_t* r0 = c.base + (int64_t)y0 * c.pitch;
    const uint8_t* r1 = c.base + (int64_t)y1 * c.pitch;
    #pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
        const int T0 = (int)r0[sx * 3 + ch] * a0 + (int)r0[sx1 * 3 + ch] * a1;
        const int T1 = (int)r1[sx * 3 + ch] * a0 + (int)r1[sx1 * 3 + ch] * a1;
        int v = (((b0 * (T0 >> 4)) >> 16) + ((b1 * (T1 >> 4)) >> 16) + 2) >> 2;
        out[ch] = v < 0 ? 0 : (v > 255 ? 255 : v);
    }
}

__device__ __forceinline__ void fill_recrop_tables(const Crop& c, int S, int* tabs) {
    for (int t = threadIdx.x; t < 2 * S; t += blockDim.x) {
        int s, a0, a1;
        if (t < S) { cv_tap(t, c.sw > 0 ? c.sw : 1, S, true, s, a0, a1); tabs[t] = s; tabs[S + t] = a0; tabs[2 * S + t] = a1; }
        else { int d = t - S; cv_tap(d, c.sh > 0 ? c.sh : 1, S, false, s, a0, a1); tabs[3 * S + d] = s; tabs[4 * S + d] = a0; tabs[5 * S + d] = a1; }
    }
}

// Standalone K4 (API / parity): out [n,S,S,3] u8.  grid = n windows, block 128.
__global__ void __launch_bounds__(128) recrop_kernel(SrcDev src, int64_t n, int S, uint8_t* __restrict__ out) {
    extern __shared__ int tabs[];
    const int64_t w = blockIdx.x;
    if (w >= n) return;
    const Crop c = make_crop(src, w);
    fill_recrop_tables(c, S, tabs);
    __syncthreads();
    uint8_t* o = out + w * (int64_t)S * S * 3;
    for (int i = threadIdx.x; i < S * S; i += blockDim.x) {
        int px[3];
        recrop_pixel(c, S, tabs, i / S, i % S, px);
        o[i * 3 + 0] = (uint8_t)px[0]; o[i * 3 + 1] = (uint8_t)px[1]; o[i * 3 + 2] = (uint8_t)px[2];
    }
}

// ------------------------------------------------------------------ feature kernel
// C = 32 channels; thread <-> (channel pair, 4-pixel row group); weights of the pair live in
// registers for the whole window.  R = rows per strip (conv rows r0-1..r0+R kept in smem).
constexpr int FEAT_THREADS = 256;

template <int OUT_SPLIT>
__global__ void __launch_bounds__(FEAT_THREADS)
features_kernel(SrcDev src, int64_t first, int64_t n, int S, int R,
                const float* __restrict__ conv_w, const float* __restrict__ conv_b,
                float a_scale, float* __restrict__ out_f32, __half* __restrict__ out_hi, __half* __restrict__ out_lo) {
    constexpr int C = 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int P = S + 2;
    float* inp = reinterpret_cast<float*>(smem_raw);                 // [P][P][3] zero-padded input
    float* cbuf = inp + ((P * P * 3 + 3) & ~3);                      // [R+2][P][C] conv+ReLU, zero borders
    int* tabs = reinterpret_cast<int*>(cbuf + (R + 2) * P * C);      // [6][S] recrop tables

    const int64_t wrel = blockIdx.x;
    if (wrel >= n) return;
    const int64_t w = first + wrel;
    const int tid = threadIdx.x;

    // ---- 1. stage the u8 window as zero-padded floats
    for (int i = tid; i < P * P * 3; i += FEAT_THREADS) inp[i] = 0.f;
    for (int i = tid; i < (R + 2) * P * C; i += FEAT_THREADS) cbuf[i] = 0.f;
    Crop crop;
    if (src.kind == RODC_SRC_RECROP) { crop = make_crop(src, w); fill_recrop_tables(crop, S, tabs); }
    __syncthreads();
    if (src.kind == RODC_SRC_DENSE) {
        const uint8_t* p = src.dense + w * (int64_t)S * S * 3;
        for (int i = tid; i < S * S * 3; i += FEAT_THREADS) {
            const int y = i / (S * 3), r = i - y * S * 3;
            inp[((y + 1) * P + 1) * 3 + r] = (float)p[i];
        }
    } else if (src.kind == RODC_SRC_PYRAMID) {
        const int64_t f = w / src.n_windows, j = w - f * src.n_windows;
        const int32_t* wr = src.win + j * 8;
        const LevelInfo& lv = src.levels[wr[0]];
        const uint8_t* p = src.pyramid + f * src.pyramid_bytes + lv.byte_off + ((int64_t)wr[2] * lv.w + wr[1]) * 3;
        const int pitch = lv.w * 3;
        for (int i = tid; i < S * S * 3; i += FEAT_THREADS) {
            const int y = i / (S * 3), r = i - y * S * 3;
            inp[((y + 1) * P + 1) * 3 + r] = (float)p[(int64_t)y * pitch + r];
        }
    } else {
        for (int i = tid; i < S * S; i += FEAT_THREADS) {
            const int y = i / S, x = i - y * S;
            int px[3];
            recrop_pixel(crop, S, tabs, y, x, px);
            float* d = inp + ((y + 1) * P + (x + 1)) * 3;
            d[0] = (float)px[0]; d[1] = (float)px[1]; d[2] = (float)px[2];
        }
    }

    // ---- 2. weights of this thread's channel pair -> registers
    const int cp = tid & 15;                 // channels 2cp, 2cp+1
    float wreg[27][2];
    #pragma unroll
    for (int t = 0; t < 27; ++t) {
        wreg[t][0] = __ldg(conv_w + t * C + 2 * cp);
        wreg[t][1] = __ldg(conv_w + t * C + 2 * cp + 1);
    }
    const float b0 = __ldg(conv_b + 2 * cp), b1 = __ldg(conv_b + 2 * cp + 1);
    __syncthreads();

    const int XG = S >> 2;                   // 4-pixel groups per row (S is a multiple of 4)
    float* of = OUT_SPLIT ? nullptr : out_f32 + wrel * (int64_t)S * S * C;
    __half* oh = OUT_SPLIT ? out_hi + wrel * (int64_t)S * S * C : nullptr;
    __half* ol = OUT_SPLIT ? out_lo + wrel * (int64_t)S * S * C : nullptr;

    for (int r0 = 0; r0 < S; r0 += R) {
        // ---- 3a. conv + bias + ReLU for rows r0-1 .. r0+R into cbuf rows 0 .. R+1
        const int items = (R + 2) * XG * 16;
        for (int it = tid; it < items; it += FEAT_THREADS) {
            const int rest = it >> 4;
            const int xg = rest % XG, ry = rest / XG;
            const int y = r0 - 1 + ry;
            float2* dst = reinterpret_cast<float2*>(cbuf + ((ry * P) + xg * 4 + 1) * C + 2 * cp);
            if (y < 0 || y >= S) {
                #pragma unroll
                for (int px = 0; px < 4; ++px) dst[px * (C / 2)] = make_float2(0.f, 0.f);
                continue;
            }
            float acc[4][2];
            #pragma unroll
            for (int px = 0; px < 4; ++px) { acc[px][0] = b0; acc[px][1] = b1; }
            #pragma unroll
            for (int ky = 0; ky < 3; ++ky) {
                const float* row = inp + ((y + ky) * P + xg * 4) * 3;
                float v[18];
                #pragma unroll
                for (int q = 0; q < 18; ++q) v[q] = row[q];
                #pragma unroll
                for (int kx = 0; kx < 3; ++kx)
                    #pragma unroll
                    for (int ci = 0; ci < 3; ++ci) {
                        const int t = (ky * 3 + kx) * 3 + ci;
                        #pragma unroll
                        for (int px = 0; px < 4; ++px) {
                            acc[px][0] = fmaf(v[(px + kx) * 3 + ci], wreg[t][0], acc[px][0]);
                            acc[px][1] = fmaf(v[(px + kx) * 3 + ci], wreg[t][1], acc[px][1]);
                        }
                    }
            }
            #pragma unroll
            for (int px = 0; px < 4; ++px)
                dst[px * (C / 2)] = make_float2(fmaxf(acc[px][0], 0.f), fmaxf(acc[px][1], 0.f));
        }
        __syncthreads();
        // ---- 3b. maxpool 3x3 + store (rows r0 .. r0+R-1)
        const int rows = min(R, S - r0);
        const int pitems = rows * S * 16;
        for (int it = tid; it < pitems; it += FEAT_THREADS) {
            const int rest = it >> 4;
            const int x = rest % S, r = rest / S;
            const float2* c0 = reinterpret_cast<const float2*>(cbuf + ((r * P) + x) * C + 2 * cp);
            float2 m = make_float2(0.f, 0.f);
            #pragma unroll
            for (int dy = 0; dy < 3; ++dy)
                #pragma unroll
                for (int dx = 0; dx < 3; ++dx) {
                    const float2 q = c0[(dy * P + dx) * (C / 2)];
                    m.x = fmaxf(m.x, q.x); m.y = fmaxf(m.y, q.y);
                }
            const int64_t k = ((int64_t)(r0 + r) * S + x) * C + 2 * cp;
            if (OUT_SPLIT) {
                const float sx = m.x * a_scale, sy = m.y * a_scale;          // exact: power of two
                const __half hx = __float2half_rn(sx), hy = __float2half_rn(sy);
                const __half lx = __float2half_rn(sx - __half2float(hx)), ly = __float2half_rn(sy - __half2float(hy));
                *reinterpret_cast<__half2*>(oh + k) = __halves2half2(hx, hy);
                *reinterpret_cast<__half2*>(ol + k) = __halves2half2(lx, ly);
            } else {
                *reinterpret_cast<float2*>(of + k) = m;
            }
        }
        __syncthreads();
    }
}

static size_t feat_smem_bytes(int S, int R) {
    const int P = S + 2;
    return (size_t)((P * P * 3 + 3) & ~3) * 4 + (size_t)(R + 2) * P * 32 * 4 + (size_t)6 * S * 4;
}

static int pick_rows(int S) {
    // largest multiple-of-4 strip that keeps two CTAs per SM (<= 112 KB each)
    for (int R = S; R > 4; R -= 4)
        if (feat_smem_bytes(S, R) <= 112 * 1024) return R;
    return 4;
}

int resolve_source(const rodc_window_source* src, int in_size, SrcDev* out) {
    RODC_CHECK_ARG(src != nullptr, "window source is NULL");
    SrcDev d; memset(&d, 0, sizeof(d));
    d.kind = src->kind;
    if (src->kind == RODC_SRC_DENSE) {
        RODC_CHECK_ARG(src->dense != nullptr, "dense window source without data");
        d.dense = src->dense;
    } else if (src->kind == RODC_SRC_PYRAMID) {
        RODC_CHECK_ARG(src->plan && src->pyramid, "pyramid window source needs plan and pyramid");
        rodc_plan* p = const_cast<rodc_plan*>(src->plan);
        RODC_CHECK_ARG(p->win == in_size, "plan window %d != stage input %d", p->win, in_size);
        RODC_CHECK_ARG(p->fits, "a window leaves its level (the reference raises ValueError and skips the image)");
        int rc = p->ensure_device(); if (rc) return rc;
        d.pyramid = src->pyramid; d.pyramid_bytes = p->pyramid_bytes; d.levels = p->d_levels; d.win = p->d_win;
        d.n_windows = (int64_t)p->win_level.size();
    } else if (src->kind == RODC_SRC_RECROP) {
        RODC_CHECK_ARG(src->frames_base && src->frame_off && src->frame_hw && src->meta, "recrop window source incomplete");
        d.frames_base = src->frames_base; d.frame_off = src->frame_off; d.frame_hw = src->frame_hw; d.meta = src->meta;
    } else {
        set_error("unknown window source kind %d", src->kind);
        return RODC_ERR_ARG;
    }
    *out = d;
    return RODC_OK;
}

int launch_features(const rodc_stage* s, const SrcDev& src, int64_t first, int64_t n,
                    float* out_f32, __half* out_hi, __half* out_lo, cudaStream_t st) {
    if (n <= 0) return RODC_OK;
    RODC_CHECK_ARG(s->channels == 32, "feature kernel is built for 32 conv channels (got %d)", s->channels);
    RODC_CHECK_ARG((s->in_size & 3) == 0, "stage input size must be a multiple of 4 (got %d)", s->in_size);
    RODC_CHECK_ARG(n < (1ll << 31), "too many windows in one launch");
    const int S = s->in_size, R = pick_rows(S);
    const size_t smem = feat_smem_bytes(S, R);
    const float a_scale = ldexpf(1.f, s->a_exp);
    if (out_f32) {
        RODC_CUDA(cudaFuncSetAttribute(features_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        features_kernel<0><<<(unsigned)n, FEAT_THREADS, smem, st>>>(src, first, n, S, R, s->conv_w, s->conv_b, a_scale, out_f32, nullptr, nullptr);
    } else {
        RODC_CUDA(cudaFuncSetAttribute(features_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        features_kernel<1><<<(unsigned)n, FEAT_THREADS, smem, st>>>(src, first, n, S, R, s->conv_w, s->conv_b, a_scale, nullptr, out_hi, out_lo);
    }
    RODC_LAUNCH_CHECK();
    return RODC_OK;
}

}  // namespace rodc

using namespace rodc;

extern "C" int rodc_recrop(const uint8_t* frames_base, const int64_t* frame_off, const int32_t* frame_hw,
                           const int32_t* meta, int64_t n, int size, uint8_t* out, void* stream) {
    RODC_CHECK_ARG(frames_base && frame_off && frame_hw && meta && out, "rodc_recrop: NULL argument");
    RODC_CHECK_ARG(size > 0 && size <= 1024, "rodc_recrop: bad size %d", size);
    if (n <= 0) return RODC_OK;
    SrcDev d; memset(&d, 0, sizeof(d));
    d.kind = RODC_SRC_RECROP; d.frames_base = frames_base; d.frame_off = frame_off; d.frame_hw = frame_hw; d.meta = meta;
    recrop_kernel<<<(unsigned)n, 128, (size_t)6 * size * sizeof(int), as_stream(stream)>>>(d, n, size, out);
    RODC_LAUNCH_CHECK();
    return RODC_OK;
}
