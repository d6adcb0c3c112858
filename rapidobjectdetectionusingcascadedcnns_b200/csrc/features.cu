// K2a + K4: per-window conv3x3(3->C)+bias+ReLU -> maxpool3x3/1 SAME -> NHWC flatten.
//
// Graph: network/net.py:172-240 (conv_layer_0, pool_layer_0), network/net_builder.py:6-60.
// SAME padding is per window: zeros for the conv; the pool ignores padding, which for
// post-ReLU (>= 0) data equals zero padding.  Raw 0..255 pixels are fed without any
// normalisation (app/inference_cascade_app.py:200-202).
//
// Window sources.  rodc_stage_forward materialises every non-dense source per chunk (source_windows_kernel /
// recrop_windows_kernel) and the streaming feature kernels read dense u8 windows; the generic feature kernels
// still read every source in place:
//   PYRAMID  stage-0 windows inside the packed pyramid (K1 output)
//   GATHER   per-window (byte offset, row pitch) into any buffer (mixed frame shapes, resize-then-crop branch)
//   RECROP   stages >= 1: the survivor's *_norm box of the ORIGINAL frame through cv2.resize INTER_LINEAR
//            (8U fixed point, 2x INTER_AREA shortcut), app/inference_cascade_app.py:363-373 (K4)
//   DENSE    materialised windows (run_inference_on_windows / run_inference_on_raw_data / the workspace copy)
//
// Output: fp32 features (SIMT fc1 / tests) or the split-fp16 A operand of the tensor fc1:
//   f * 2^a_exp = hi + lo  (hi, lo fp16; 22 significant bits), planes [n, K] K-contiguous.
#include "stage.cuh"
#include <stdlib.h>
#include <string.h>

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

// The u8 input of a stage for ANY window source (what the reference holds as windows_raw and feeds to X,
// app/inference_cascade_app.py:200-202): out [n,S,S,3].  API / parity path; the stage kernels read the
// sources in place.
__global__ void __launch_bounds__(128) source_windows_kernel(SrcDev src, int64_t first, int64_t n, int S,
                                                             uint8_t* __restrict__ out, const int64_t* __restrict__ n_dev, int64_t n_off) {
    extern __shared__ int tabs[];
    n = eff_count(n, n_dev, n_off);
    for (int64_t wrel = blockIdx.x; wrel < n; wrel += gridDim.x) {
        const int64_t w = first + wrel;
        uint8_t* o = out + wrel * (int64_t)S * S * 3;
        if (src.kind == RODC_SRC_RECROP) {
            __syncthreads();                                   // tables of the previous window are free
            const Crop c = make_crop(src, w);
            fill_recrop_tables(c, S, tabs);
            __syncthreads();
            for (int i = threadIdx.x; i < S * S; i += blockDim.x) {
                int px[3];
                recrop_pixel(c, S, tabs, i / S, i % S, px);
                o[i * 3 + 0] = (uint8_t)px[0]; o[i * 3 + 1] = (uint8_t)px[1]; o[i * 3 + 2] = (uint8_t)px[2];
            }
            continue;
        }
        const uint8_t* p; int pitch;
        if (src.kind == RODC_SRC_DENSE) { p = src.dense + w * (int64_t)S * S * 3; pitch = S * 3; }
        else if (src.kind == RODC_SRC_GATHER) { p = src.gather_base + src.win_off[w]; pitch = src.win_pitch[w]; }
        else {
            const int64_t f = w / src.n_windows, j = w - f * src.n_windows;
            const int32_t* wr = src.win + j * 8;
            const LevelInfo& lv = src.levels[wr[0]];
            p = src.pyramid + f * src.pyramid_bytes + lv.byte_off + ((int64_t)wr[2] * lv.w + wr[1]) * 3;
            pitch = lv.w * 3;
        }
        if ((S & 3) == 0 && (reinterpret_cast<uintptr_t>(o) & 3) == 0) {
            // one 32-bit store per thread and step: the four source bytes come from the (at most two) aligned words
            // that hold them, so no word without a byte of the window row is ever read (432 byte loads + 432 byte
            // stores per 12-px window before: the copy ran at 0.1 of the HBM rate and was 3 % of a config-3 step)
            const int rw = S * 3 / 4;                              // words per window row
            uint32_t* ow = reinterpret_cast<uint32_t*>(o);
            for (int i = threadIdx.x; i < S * rw; i += blockDim.x) {
                const int y = i / rw, wx = i - y * rw;
                const uintptr_t ua = reinterpret_cast<uintptr_t>(p + (int64_t)y * pitch + 4 * wx);
                const uint32_t* q = reinterpret_cast<const uint32_t*>(ua & ~(uintptr_t)3);
                const int sh = (int)(ua & 3) * 8;
                const uint32_t w0 = __ldg(q);
                const uint32_t w1 = sh ? __ldg(q + 1) : 0u;
                ow[i] = __funnelshift_r(w0, w1, sh);
            }
        } else {
            for (int i = threadIdx.x; i < S * S * 3; i += blockDim.x) {
                const int y = i / (S * 3), b = i - y * S * 3;
                o[i] = p[(int64_t)y * pitch + b];
            }
        }
    }
}

// Stage-0 window copy, throughput form (DENSE / GATHER / PYRAMID sources, S % 4 == 0): a CTA takes COPY_G windows at
// a time; the first COPY_G threads resolve their base pointers (window table, level table) ONCE into shared memory,
// then every thread moves the same 32-bit word of all COPY_G windows -- COPY_G independent load pairs in flight per
// thread instead of one, and no per-thread table walk.  Same aligned-word rule as above: no 4-byte word without a
// byte of the window row is read.  (One window per CTA iteration ran at 0.15 of the HBM rate: latency bound.)
constexpr int COPY_G = 8;
__global__ void __launch_bounds__(128) window_copy_kernel(SrcDev src, int64_t first, int64_t n, int S,
                                                          uint32_t* __restrict__ out, const int64_t* __restrict__ n_dev, int64_t n_off) {
    __shared__ const uint8_t* base[COPY_G];
    __shared__ int pitch[COPY_G];
    n = eff_count(n, n_dev, n_off);
    const int rw = S * 3 / 4, wpw = S * rw;                    // words per window row / per window
    for (int64_t w0 = (int64_t)blockIdx.x * COPY_G; w0 < n; w0 += (int64_t)gridDim.x * COPY_G) {
        __syncthreads();                                        // the previous group's pointers are free
        if (threadIdx.x < COPY_G && w0 + threadIdx.x < n) {
            const int64_t w = first + w0 + threadIdx.x;
            const uint8_t* p; int pt;
            if (src.kind == RODC_SRC_DENSE) { p = src.dense + w * (int64_t)S * S * 3; pt = S * 3; }
            else if (src.kind == RODC_SRC_GATHER) { p = src.gather_base + src.win_off[w]; pt = src.win_pitch[w]; }
            else {
                const int64_t f = w / src.n_windows, j = w - f * src.n_windows;
                const int32_t* wr = src.win + j * 8;
                const LevelInfo& lv = src.levels[wr[0]];
                p = src.pyramid + f * src.pyramid_bytes + lv.byte_off + ((int64_t)wr[2] * lv.w + wr[1]) * 3;
                pt = lv.w * 3;
            }
            base[threadIdx.x] = p; pitch[threadIdx.x] = pt;
        }
        __syncthreads();
        const int g = (int)min((int64_t)COPY_G, n - w0);
        for (int i = threadIdx.x; i < wpw; i += blockDim.x) {
            const int y = i / rw, wx = i - y * rw;
            uint32_t v[COPY_G];
            #pragma unroll
            for (int k = 0; k < COPY_G; ++k) {
                if (k < g) {
                    const uintptr_t ua = reinterpret_cast<uintptr_t>(base[k] + (int64_t)y * pitch[k] + 4 * wx);
                    const uint32_t* q = reinterpret_cast<const uint32_t*>(ua & ~(uintptr_t)3);
                    const int sh = (int)(ua & 3) * 8;
                    const uint32_t a0 = __ldg(q);
                    const uint32_t a1 = sh ? __ldg(q + 1) : 0u;
                    v[k] = __funnelshift_r(a0, a1, sh);
                }
            }
            #pragma unroll
            for (int k = 0; k < COPY_G; ++k)
                if (k < g) out[(w0 + k) * wpw + i] = v[k];
        }
    }
}

// K4 for the default cascade sizes: cv2.resize(window.raw_norm, (S, S)) of every survivor into dense u8 windows
// (app/inference_cascade_app.py:363-373).  A thread owns 4 adjacent output pixels of a row (its x taps live in
// registers, 12 output bytes leave as three 32-bit stores) and walks down the rows; the tap tables of a window are
// built once by the first 2 S threads.  Same arithmetic as recrop_pixel, without the per-pixel divisions and table
// reads that made the generic kernel issue bound (147 -> ~60 instructions per pixel).
// Bytes p[0 .. len) (len = 3 or 6) as lo = bytes 0..3, hi = bytes 4..5, fetched with one or two aligned 8-byte
// loads instead of `len` byte loads (the scattered byte loads saturated the L1 request path).  The aligned word
// that holds a valid byte never leaves that byte's 256-byte allocation granule, so nothing outside the frames
// buffer is touched.
__device__ __forceinline__ void load_taps(const uint8_t* p, int len, uint32_t& lo, uint32_t& hi) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p);
    const unsigned long long* q = reinterpret_cast<const unsigned long long*>(a & ~(uintptr_t)7);
    const int off = (int)(a & 7);
    const unsigned long long w0 = __ldg(q);
    const unsigned long long w1 = off + len > 8 ? __ldg(q + 1) : 0ull;
    const int sh = off * 8;
    const unsigned long long v = sh ? (w0 >> sh) | (w1 << (64 - sh)) : w0;
    lo = (uint32_t)v; hi = (uint32_t)(v >> 32);
}

template <int S, int RPP>
__global__ void __launch_bounds__((S / 4) * RPP) recrop_windows_kernel(SrcDev src, int64_t first, int64_t n,
                                                                       uint8_t* __restrict__ out, const int64_t* __restrict__ n_dev, int64_t n_off) {
    constexpr int UN = S / 4, THREADS = UN * RPP;
    __shared__ int tabs[6 * S];
    n = eff_count(n, n_dev, n_off);
    const int u = threadIdx.x % UN, r0 = threadIdx.x / UN;
    for (int64_t wrel = blockIdx.x; wrel < n; wrel += gridDim.x) {
        const Crop c = make_crop(src, first + wrel);
        uint32_t* o = reinterpret_cast<uint32_t*>(out + wrel * (int64_t)S * S * 3);
        const bool empty = c.sw <= 0 || c.sh <= 0;
        const bool same = c.sw == S && c.sh == S, half = c.sw == 2 * S && c.sh == 2 * S;
        if (empty || same || half) {                            // uniform per window: zero / copy / INTER_AREA 2x2
            for (int r = r0; r < S; r += RPP) {
                uint32_t w[3] = {0u, 0u, 0u};
                if (!empty) {
                    #pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int x = 4 * u + i;
                        #pragma unroll
                        for (int ch = 0; ch < 3; ++ch) {
                            uint32_t v;
                            if (same) v = c.base[(int64_t)r * c.pitch + x * 3 + ch];
                            else {
                                const uint8_t* p = c.base + (int64_t)(2 * r) * c.pitch + (2 * x) * 3 + ch;
                                v = ((uint32_t)p[0] + p[3] + p[c.pitch] + p[c.pitch + 3] + 2u) >> 2;
                            }
                            const int b = i * 3 + ch;
                            w[b >> 2] |= v << (8 * (b & 3));
                        }
                    }
                }
                uint32_t* d = o + (r * S + 4 * u) * 3 / 4;
                d[0] = w[0]; d[1] = w[1]; d[2] = w[2];
            }
            continue;
        }
        __syncthreads();                                        // the previous window's tables are free
        for (int t = threadIdx.x; t < 2 * S; t += THREADS) {
            int sidx, a0, a1;
            if (t < S) { cv_tap(t, c.sw, S, true, sidx, a0, a1); tabs[t] = sidx; tabs[S + t] = a0; tabs[2 * S + t] = a1; }
            else { const int d = t - S; cv_tap(d, c.sh, S, false, sidx, a0, a1); tabs[3 * S + d] = sidx; tabs[4 * S + d] = a0; tabs[5 * S + d] = a1; }
        }
        __syncthreads();
        int xo[4], xlen[4];                                     // byte offset of the first x tap, bytes to fetch
        uint32_t xw[4];                                         // the two x weights of a pixel as 16-bit halves (<= 2048 each)
        #pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int x = 4 * u + i, sx = tabs[x];
            xo[i] = sx * 3;
            xlen[i] = sx + 1 <= c.sw - 1 ? 6 : 3;               // right border: the second tap is the same pixel (weight 0)
            xw[i] = (uint32_t)tabs[S + x] | ((uint32_t)tabs[2 * S + x] << 16);
        }
        for (int r = r0; r < S; r += RPP) {
            const int sy = tabs[3 * S + r];
            // (b * t) >> 16 == umulhi(b << 16, t) for b <= 2048, t < 2^16: one IMAD.HI instead of a multiply and a shift
            const uint32_t b0s = (uint32_t)tabs[4 * S + r] << 16, b1s = (uint32_t)tabs[5 * S + r] << 16;
            const int y0 = min(max(sy, 0), c.sh - 1), y1 = min(max(sy + 1, 0), c.sh - 1);
            const uint8_t* p0 = c.base + (int64_t)y0 * c.pitch;
            const uint8_t* p1 = c.base + (int64_t)y1 * c.pitch;
            uint32_t w[3] = {0u, 0u, 0u};
            #pragma unroll
            for (int i = 0; i < 4; ++i) {
                uint32_t l0, h0, l1, h1;                        // the two taps of both rows: 6 adjacent bytes each
                load_taps(p0 + xo[i], xlen[i], l0, h0);
                load_taps(p1 + xo[i], xlen[i], l1, h1);
                // xlen == 3: the bytes of the "second tap" are undefined, but its weight is 0 (cv_tap clamps f to 0)
                #pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    // bytes (ch, 3 + ch) of the six = this channel of the two taps -> the low half of a register, then the
                    // horizontal interpolation t0 * a0 + t1 * a1 is ONE dp2a (16-bit weights x 8-bit pixels, exact)
                    const uint32_t sel = ch == 0 ? 0x4430u : (ch == 1 ? 0x4441u : 0x4452u);
                    const uint32_t T0 = __dp2a_lo(xw[i], __byte_perm(l0, h0, sel), 0u);
                    const uint32_t T1 = __dp2a_lo(xw[i], __byte_perm(l1, h1, sel), 0u);
                    // weights are >= 0 and sum to 2048 per axis, so the result is already inside 0..255
                    const uint32_t v = ((__umulhi(b0s, T0 >> 4) + __umulhi(b1s, T1 >> 4) + 2u) >> 2) & 255u;
                    const int b = i * 3 + ch;
                    w[b >> 2] |= v << (8 * (b & 3));
                }
            }
            uint32_t* d = o + (r * S + 4 * u) * 3 / 4;
            d[0] = w[0]; d[1] = w[1]; d[2] = w[2];
        }
    }
}


constexpr int FEAT_THREADS = 256;

// ------------------------------------------------------------------ feature kernel, generic sizes (smem pooling)
// The 3x3x3 -> 32 convolution is a [pixels x 48] x [48 x 32] GEMM per window: K = 3 filter rows x
// (4 pixels x 4 channels), where the 4th pixel / 4th channel carry zero weights so that the A row of
// a pixel is ONE contiguous 32-byte run of the fp16 input staged as [P][P][4] -- the im2col matrix is
// never built.  u8 pixels are exact in fp16; the fp32 weights are split w * 2^e = hi + lo (fp16), so
// each product is exact and the conv costs 24 mma.sync.m16n8k16 per 16 pixels instead of 864 FFMA per
// pixel-row group.  (Legacy warp-level MMA is the right tool here: N = 32, K = 48 tiles are far too
// small to feed tcgen05, and the conv is < 6 % of the stage FLOPs.)
// Pooling is register blocked (each thread walks 4 rows x 4 pixels x 2 channels with a sliding window
// of horizontal maxima), the CTA loops over windows so the weight fragments are loaded once.
constexpr int CS = 40;            // floats per pixel in the conv buffer (32 + pad: conflict-free float2 stores)

__device__ __forceinline__ void mma16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                         uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// ST / RT: compile-time window side and strip height (0 = take the run-time arguments); the three
// reference sizes 12 / 24 / 48 get fully specialised code (constant divisions, unrolled pooling).
template <int OUT_SPLIT, int ST, int RT>
__global__ void __launch_bounds__(FEAT_THREADS, 2)
features_mma_kernel(SrcDev src, int64_t first, int64_t n, int S_rt, int R_rt,
                    const uint32_t* __restrict__ conv_frag, const float* __restrict__ conv_b, float conv_scale,
                    float a_scale, float* __restrict__ out_f32, __half* __restrict__ out_hi, __half* __restrict__ out_lo,
                    const int64_t* __restrict__ n_dev, int64_t n_off, int perm_side) {
    n = eff_count(n, n_dev, n_off);
    constexpr int C = 32;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int S = ST ? ST : S_rt;
    const int R = RT ? RT : R_rt;
    const int P = S + 2;
    const int inh_halves = (P * P * 4 + 16 + 7) & ~7;
    __half* inh = reinterpret_cast<__half*>(smem_raw);                         // [P][P][4] fp16, zero border
    float* cbuf = reinterpret_cast<float*>(smem_raw + (size_t)inh_halves * 2); // [R+2][P][CS]
    int* tabs = reinterpret_cast<int*>(cbuf + (R + 2) * P * CS);               // [6][S]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;

    // weight fragments of the whole conv (hi and lo planes) + biases of this thread's channels
    uint32_t bf[3][4][2][2];
    #pragma unroll
    for (int ky = 0; ky < 3; ++ky)
        #pragma unroll
        for (int nt = 0; nt < 4; ++nt)
            #pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint2 v = __ldg(reinterpret_cast<const uint2*>(conv_frag) + ((ky * 4 + nt) * 2 + h) * 32 + lane);
                bf[ky][nt][h][0] = v.x; bf[ky][nt][h][1] = v.y;
            }
    float2 bias2[4];
    #pragma unroll
    for (int nt = 0; nt < 4; ++nt) bias2[nt] = make_float2(__ldg(conv_b + 8 * nt + 2 * t), __ldg(conv_b + 8 * nt + 2 * t + 1));

    for (int i = tid; i < (R + 2) * P * CS; i += FEAT_THREADS) cbuf[i] = 0.f;   // borders stay zero forever

    for (int64_t wrel = blockIdx.x; wrel < n; wrel += gridDim.x) {
        const int64_t w = first + wrel;
        __syncthreads();                                      // previous window fully consumed
        // ---- 1. stage the window as zero-padded fp16 [P][P][4]
        {
            uint4* z = reinterpret_cast<uint4*>(inh);
            for (int i = tid; i < inh_halves / 8; i += FEAT_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
        }
        Crop crop;
        if (src.kind == RODC_SRC_RECROP) { crop = make_crop(src, w); fill_recrop_tables(crop, S, tabs); }
        __syncthreads();
        if (src.kind == RODC_SRC_RECROP) {
            for (int i = tid; i < S * S; i += FEAT_THREADS) {
                const int y = i / S, x = i - y * S;
                int px[3];
                recrop_pixel(crop, S, tabs, y, x, px);
                __half2* d = reinterpret_cast<__half2*>(inh + ((y + 1) * P + (x + 1)) * 4);
                d[0] = __halves2half2(__int2half_rn(px[0]), __int2half_rn(px[1]));
                d[1] = __halves2half2(__int2half_rn(px[2]), __int2half_rn(0));
            }
        } else {
            const uint8_t* p; int pitch;
            if (src.kind == RODC_SRC_DENSE) { p = src.dense + w * (int64_t)S * S * 3; pitch = S * 3; }
            else if (src.kind == RODC_SRC_GATHER) { p = src.gather_base + src.win_off[w]; pitch = src.win_pitch[w]; }
            else {
                const int64_t f = w / src.n_windows, j = w - f * src.n_windows;
                const int32_t* wr = src.win + j * 8;
                const LevelInfo& lv = src.levels[wr[0]];
                p = src.pyramid + f * src.pyramid_bytes + lv.byte_off + ((int64_t)wr[2] * lv.w + wr[1]) * 3;
                pitch = lv.w * 3;
            }
            for (int i = tid; i < S * S; i += FEAT_THREADS) {
                const int y = i / S, x = i - y * S;
                const uint8_t* q = p + (int64_t)y * pitch + x * 3;
                __half2* d = reinterpret_cast<__half2*>(inh + ((y + 1) * P + (x + 1)) * 4);
                d[0] = __halves2half2(__int2half_rn((int)q[0]), __int2half_rn((int)q[1]));
                d[1] = __halves2half2(__int2half_rn((int)q[2]), __int2half_rn(0));
            }
        }
        __syncthreads();

        float* of = OUT_SPLIT ? nullptr : out_f32 + wrel * (int64_t)S * S * C;
        __half* oh = OUT_SPLIT ? out_hi + wrel * (int64_t)S * S * C : nullptr;
        __half* ol = OUT_SPLIT ? out_lo + wrel * (int64_t)S * S * C : nullptr;

        for (int r0 = 0; r0 < S; r0 += R) {
            // ---- 2. conv + bias + ReLU of rows r0-1 .. r0+R  ->  cbuf rows 0 .. R+1 (mma.sync)
            const int npix = (R + 2) * S;
            for (int tile = warp; tile * 16 < npix; tile += FEAT_THREADS / 32) {
                int pa = tile * 16 + g, pb = pa + 8;
                const bool va = pa < npix, vb = pb < npix;
                pa = va ? pa : 0; pb = vb ? pb : 0;
                const int rya = pa / S, xa = pa - rya * S, ryb = pb / S, xb = pb - ryb * S;
                const int ya = r0 - 1 + rya, yb = r0 - 1 + ryb;
                const bool ina = va && ya >= 0 && ya < S, inb = vb && yb >= 0 && yb < S;
                const int yca = min(max(ya, 0), S - 1), ycb = min(max(yb, 0), S - 1);
                const uint32_t* rowa = reinterpret_cast<const uint32_t*>(inh + ((yca * P) + xa) * 4) + t;
                const uint32_t* rowb = reinterpret_cast<const uint32_t*>(inh + ((ycb * P) + xb) * 4) + t;
                float acc[4][4];
                #pragma unroll
                for (int nt = 0; nt < 4; ++nt) { acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f; }
                #pragma unroll
                for (int ky = 0; ky < 3; ++ky) {
                    const uint32_t a0 = rowa[ky * P * 2], a2 = rowa[ky * P * 2 + 4];
                    const uint32_t a1 = rowb[ky * P * 2], a3 = rowb[ky * P * 2 + 4];
                    #pragma unroll
                    for (int nt = 0; nt < 4; ++nt) {
                        mma16816(acc[nt], a0, a1, a2, a3, bf[ky][nt][1][0], bf[ky][nt][1][1]);     // lo weights first
                        mma16816(acc[nt], a0, a1, a2, a3, bf[ky][nt][0][0], bf[ky][nt][0][1]);
                    }
                }
                float* da = cbuf + ((rya * P) + xa + 1) * CS + 2 * t;
                float* db = cbuf + ((ryb * P) + xb + 1) * CS + 2 * t;
                #pragma unroll
                for (int nt = 0; nt < 4; ++nt) {
                    if (va) *reinterpret_cast<float2*>(da + 8 * nt) = ina
                        ? make_float2(fmaxf(fmaf(acc[nt][0], conv_scale, bias2[nt].x), 0.f), fmaxf(fmaf(acc[nt][1], conv_scale, bias2[nt].y), 0.f))
                        : make_float2(0.f, 0.f);
                    if (vb) *reinterpret_cast<float2*>(db + 8 * nt) = inb
                        ? make_float2(fmaxf(fmaf(acc[nt][2], conv_scale, bias2[nt].x), 0.f), fmaxf(fmaf(acc[nt][3], conv_scale, bias2[nt].y), 0.f))
                        : make_float2(0.f, 0.f);
                }
            }
            __syncthreads();
            // ---- 3. maxpool 3x3 (register blocked: 4 rows x 4 pixels x 2 channels per item) + store
            const int rows = min(R, S - r0);
            const int XG = S >> 2, SEG = (rows + 3) >> 2;
            const int items = SEG * XG * 16;
            for (int it = tid; it < items; it += FEAT_THREADS) {
                const int cp = it & 15;
                const int rest = it >> 4;
                const int xg = rest % XG, seg = rest / XG;
                const int rbeg = seg * 4, nrow = min(4, rows - rbeg);
                // horizontal 3-maxima of the six conv rows this item touches (all in registers)
                float2 h[6][4];
                #pragma unroll
                for (int rr = 0; rr < 6; ++rr) {
                    if (rr < nrow + 2) {
                        const float2* c0 = reinterpret_cast<const float2*>(cbuf + (((rbeg + rr) * P) + xg * 4) * CS + 2 * cp);
                        float2 v[6];
                        #pragma unroll
                        for (int q = 0; q < 6; ++q) v[q] = c0[q * (CS / 2)];
                        #pragma unroll
                        for (int px = 0; px < 4; ++px) {
                            h[rr][px].x = fmaxf(fmaxf(v[px].x, v[px + 1].x), v[px + 2].x);
                            h[rr][px].y = fmaxf(fmaxf(v[px].y, v[px + 1].y), v[px + 2].y);
                        }
                    }
                }
                #pragma unroll
                for (int r = 0; r < 4; ++r) {
                    if (r < nrow) {
                        const int64_t kbase = ((int64_t)(r0 + rbeg + r) * S + xg * 4) * C + 2 * cp;
                        #pragma unroll
                        for (int px = 0; px < 4; ++px) {
                            float2 m;
                            m.x = fmaxf(fmaxf(h[r][px].x, h[r + 1][px].x), h[r + 2][px].x);
                            m.y = fmaxf(fmaxf(h[r][px].y, h[r + 1][px].y), h[r + 2][px].y);
                            const int64_t k = OUT_SPLIT ? kperm(perm_side, kbase + px * C) : kbase + px * C;
                            if (OUT_SPLIT) {
                                const float sx = m.x * a_scale, sy = m.y * a_scale;            // exact: power of two
                                const __half2 h2 = __floats2half2_rn(sx, sy);
                                const float2 hf = __half22float2(h2);
                                *reinterpret_cast<__half2*>(oh + k) = h2;
                                *reinterpret_cast<__half2*>(ol + k) = __floats2half2_rn(sx - hf.x, sy - hf.y);
                            } else {
                                *reinterpret_cast<float2*>(of + k) = m;
                            }
                        }
                    }
                }
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------ feature kernel, 12 / 24 / 48 (streaming, barrier free)
// Same arithmetic as the generic kernel, but the conv output never touches shared memory: warp (cg, part) owns 16
// channels x one quarter of the window rows and streams down its rows.  M row r of an MMA tile maps to
// pixel 16j + 2 (r % 8) + r / 8, so every thread holds two horizontally adjacent pixels and the
// horizontal 3-max needs two lane shuffles per channel and row (pixel x = 2 TR g + i: a thread owns 2 TR
// adjacent pixels); the vertical 3-max is a sliding window of row registers.  Pixels are staged as
// (r, g, b, 1) * 2^(a_exp - c_exp): the constant channel carries the bias through the MMA and the
// accumulator IS the feature scaled for the fp16 split; the ReLU is applied once after the pooling.  The two n-tiles of a warp use permuted weight columns so that thread t owns
// channels 4t..4t+3 of its group: stores are 8-byte, four lanes fill one 32-byte sector.  One
// __syncthreads per window (input staging is double buffered).  This is also the producer the fused
// feature->fc1 kernel needs (round 2): only the final store changes.
// A warp row is VW = S x WPW pixels wide: one 48-px window, two 24-px windows side by side (lane groups 0-3 / 4-7:
// every lane and every MMA row is used, the horizontal max simply does not look across the window boundary) or one
// 12-px window (lane groups 6, 7 idle; packing four of them was measured slower: the extra per-lane state spills
// inside the row loop).
// Warp layout per window size (8 warps): S = 48 -> 1 window x 4 row parts x 2 channel groups;
// S = 24 -> 2 warp slots x 2 windows x 2 parts x 2 groups; S = 12 -> 4 windows x 1 part x 2 groups, so small
// windows are not dominated by the one-row halo that every row part recomputes.
template <int S> struct StreamCfg {
    static constexpr int NQ = S >= 48 ? 4 : (S >= 24 ? 2 : 1);       // row parts per window
    static constexpr int WPW = S == 24 ? 2 : 1;                       // windows side by side in one warp
    static constexpr int WPC = (4 / NQ) * WPW;                        // windows per CTA iteration
    static constexpr int CTAS_PER_SM = S == 12 ? 3 : 2;               // the 12-px kernel needs < 86 registers
    // halves per staged fp16 window [P][P][4] (+ read slack); two windows of one warp are placed 16 banks apart so
    // that the A-fragment loads of the two lane halves do not collide
    static constexpr int INH0 = ((S + 2) * (S + 2) * 4 + 64 + 7) & ~7;
    static constexpr int INH = WPW == 2 ? INH0 + ((16 - (INH0 / 2) % 32 + 32) % 32) * 2 : INH0;
};

template <int OUT_SPLIT, int S, int OCC = StreamCfg<S>::CTAS_PER_SM>
__global__ void __launch_bounds__(FEAT_THREADS, OCC)
features_stream_kernel(const uint8_t* __restrict__ dense, int64_t first, int64_t n,
                       const uint32_t* __restrict__ conv_frag, float pix_scale, float inv_a_scale,
                       float* __restrict__ out_f32, __half* __restrict__ out_hi, __half* __restrict__ out_lo,
                       const int64_t* __restrict__ n_dev, int64_t n_off, int probe_wrap) {
    n = eff_count(n, n_dev, n_off);
    constexpr int NQ = StreamCfg<S>::NQ, WPW = StreamCfg<S>::WPW, WPC = StreamCfg<S>::WPC, RQ = S / NQ;
    constexpr int VW = S * WPW;                                           // pixels of a warp row
    constexpr int C = 32, P = S + 2, TR = (VW + 15) / 16, NP = 2 * TR;   // NP pixels per thread and row
    constexpr bool ALL_IN = 8 * NP == VW;                                 // every lane group owns pixels
    static_assert(S % NP == 0 && VW <= 48, "a lane group lies inside one window or is idle");
    constexpr int INH = StreamCfg<S>::INH;                               // halves per input buffer (+ read slack)
    constexpr int WIN_BYTES = S * S * 3, UPW = S * S / 4, UPI = WPC * UPW;   // 12-byte units (4 pixels) per window / iteration
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __half* inh0 = reinterpret_cast<__half*>(smem_raw);                  // [2 buffers][WPC windows][INH]
    uint8_t* raw0 = smem_raw + (size_t)2 * WPC * INH * 2;                // [2 buffers][WPC windows][WIN_BYTES] u8 as fetched

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, t = lane & 3;
    const int cg = warp & 1, part = (warp >> 1) % NQ;
    const bool lane_in = ALL_IN || NP * g < VW;                           // S = 12 leaves lane groups 6, 7 idle
    const int wsub = lane_in ? (NP * g) / S : 0;                          // this lane's window inside the warp
    const int slot = ((warp >> 1) / NQ) * WPW + wsub;                    // ... and inside the CTA iteration
    const int ya = part * RQ;
    const int xbase = lane_in ? NP * g - wsub * S : 0;                    // this thread's pixels: xbase .. xbase + NP - 1

    // B fragments of this warp's two n-tiles (permuted columns, bias folded into the spare K slot)
    uint32_t bf[3][2][2][2];
    #pragma unroll
    for (int ky = 0; ky < 3; ++ky)
        #pragma unroll
        for (int sub = 0; sub < 2; ++sub)
            #pragma unroll
            for (int h = 0; h < 2; ++h) {
                const uint2 v = __ldg(reinterpret_cast<const uint2*>(conv_frag) + (((ky * 4 + (cg * 2 + sub)) * 2 + h) * 32 + lane));
                bf[ky][sub][h][0] = v.x; bf[ky][sub][h][1] = v.y;
            }
    const __half one_h = __float2half_rn(pix_scale);

    // Input staging.  The u8 windows of iteration i+1 are fetched with cp.async into a raw buffer while
    // iteration i computes (a window is 6.9 KB of the 295 KB it produces, but its load latency was 1/3 of the
    // kernel when it sat between two barriers); every thread converts exactly the bytes it fetched itself, so the
    // only barrier per iteration is the one that publishes the fp16 image.  The zero border (SAME padding) of
    // both fp16 buffers is written once: interior stores never touch it.
    {
        uint4* z = reinterpret_cast<uint4*>(inh0);
        for (int i = tid; i < 2 * WPC * INH / 8; i += FEAT_THREADS) z[i] = make_uint4(0u, 0u, 0u, 0u);
    }
    auto prefetch = [&](int64_t wrel0, int buf) {
        for (int u = tid; u < UPI; u += FEAT_THREADS) {
            const int sl = u / UPW, q = u - sl * UPW;
            if (wrel0 + sl < n) {
                const uint8_t* src = dense + (first + wrel0 + sl) * (int64_t)WIN_BYTES + q * 12;
                const uint32_t dst = (uint32_t)__cvta_generic_to_shared(raw0 + ((size_t)buf * WPC + sl) * WIN_BYTES + q * 12);
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(dst), "l"(src) : "memory");
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(dst + 4), "l"(src + 4) : "memory");
                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" :: "r"(dst + 8), "l"(src + 8) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
    prefetch((int64_t)blockIdx.x * WPC, 0);
    __syncthreads();                                           // zero fill visible before the first interior stores

    int iter = 0;
    for (int64_t wrel0 = (int64_t)blockIdx.x * WPC; wrel0 < n; wrel0 += (int64_t)gridDim.x * WPC, ++iter) {
        __half* inhb = inh0 + (size_t)(iter & 1) * WPC * INH;
        asm volatile("cp.async.wait_group 0;" ::: "memory");   // this thread's bytes of the current windows have landed
        const int64_t next0 = wrel0 + (int64_t)gridDim.x * WPC;
        if (next0 < n) prefetch(next0, (iter + 1) & 1);
        // ---- fp16 [P][P][4] = (r, g, b, 1) * 2^p from the thread's own 4-pixel units
        for (int u = tid; u < UPI; u += FEAT_THREADS) {
            const int sl = u / UPW, q = u - sl * UPW;
            if (wrel0 + sl >= n) continue;
            const uint32_t* rw = reinterpret_cast<const uint32_t*>(raw0 + ((size_t)(iter & 1) * WPC + sl) * WIN_BYTES + q * 12);
            const uint32_t w0 = rw[0], w1 = rw[1], w2 = rw[2];
            const int p0 = 4 * q, y = p0 / S, x = p0 - y * S;
            uint2* d = reinterpret_cast<uint2*>(inhb + (size_t)sl * INH + ((y + 1) * P + (x + 1)) * 4);
            const uint32_t px[4][3] = {{w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u},
                                       {w0 >> 24, w1 & 255u, (w1 >> 8) & 255u},
                                       {(w1 >> 16) & 255u, w1 >> 24, w2 & 255u},
                                       {(w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24}};
            #pragma unroll
            for (int i = 0; i < 4; ++i) {
                const __half2 rg = __floats2half2_rn((float)px[i][0] * pix_scale, (float)px[i][1] * pix_scale);
                const __half2 b1 = __halves2half2(__float2half_rn((float)px[i][2] * pix_scale), one_h);
                uint2 o;
                o.x = *reinterpret_cast<const uint32_t*>(&rg); o.y = *reinterpret_cast<const uint32_t*>(&b1);
                d[i] = o;
            }
        }
        __syncthreads();                                       // windows staged (and everyone left iteration iter - 1)

        const int64_t wrel = wrel0 + slot;
        if (wrel0 + (slot - wsub) >= n) continue;              // tail: all windows of this warp are beyond n
        const bool live = lane_in && wrel < n;                 // idle lanes / lanes of a missing window compute along, never store
        const __half* inh = inhb + (size_t)slot * INH;

        // thread-constant output pointers (row offset added per row, pixel offsets are immediates)
        // ONE running output pointer per thread (element offset of output row ya - 2 + rr; the lo plane sits at
        // the kernel-uniform distance out_lo - out_hi), opaque so that it is carried instead of being rebuilt from %tid
        // fp32 output (tests, SIMT fc1): the reference's NHWC order.  Split output: the permuted order of stage.cuh
        // kperm -- row = [channel half][x'][16], x' = i G + group, so one store instruction of the warp (fixed pixel i,
        // all lane groups) covers G x 32 contiguous bytes
        constexpr int G = S / NP;                              // lane groups per window
        constexpr int PIX = OUT_SPLIT ? G * 16 : C;            // element stride between a thread's pixels
#ifdef RODC_PROBES
        const int64_t wout = probe_wrap ? wrel % probe_wrap : wrel;     // probe: A stays inside probe_wrap rows (L2 resident)
#else
        const int64_t wout = wrel;
#endif
        int64_t oofs = OUT_SPLIT
            ? wout * (int64_t)S * S * C + (int64_t)(ya - 2) * S * C + cg * S * 16 + (xbase / NP) * 16 + 4 * t
            : wout * (int64_t)S * S * C + (int64_t)xbase * C + 16 * cg + 4 * t + (int64_t)(ya - 2) * S * C;
        asm volatile("" : "+l"(oofs));
        const int64_t lo_delta = OUT_SPLIT ? out_lo - out_hi : 0;
        const uint32_t* abase = reinterpret_cast<const uint32_t*>(inh + xbase * 4) + t;
        // lane / warp role flags in ONE register the compiler cannot rematerialise from %tid (S2R is a long-latency
        // read and was re-issued in every row): bit 0 left neighbour group, 1 right neighbour group, 2 first row
        // part, 3 last row part, 4 the lane's window exists (stores allowed)
        int flags = (lane_in && xbase > 0 ? 1 : 0) | (lane_in && xbase + NP < S ? 2 : 0) | (part == 0 ? 4 : 0) | (part == NQ - 1 ? 8 : 0) | (live ? 16 : 0);
        asm volatile("" : "+r"(flags));
        const bool has_left = flags & 1, has_right = flags & 2;

        // sliding window of horizontal maxima of (conv + bias) * 2^a_exp, BEFORE the ReLU: the ReLU commutes
        // with the max and is applied once at the output as max(., 0); zeros stand for SAME padding.
        float hA[NP][4], hB[NP][4];                            // max(h[y-2], h[y-1]) and h[y-1]
        #pragma unroll
        for (int i = 0; i < NP; ++i)
            #pragma unroll
            for (int c = 0; c < 4; ++c) { hA[i][c] = 0.f; hB[i][c] = 0.f; }

        // Row loop in part-relative terms (rr = 0 .. RQ + 1 <-> image row ya - 1 + rr) with running pointers, so
        // that nothing derived from %tid is needed inside it (the compiler re-read %tid every row otherwise).
        const bool first_part = flags & 4, last_part = flags & 8, owns = flags & 16;
        const uint32_t* arow = abase + (ya - 1) * P * 2;      // row ya - 1 (never dereferenced when that is row -1)
        #pragma unroll 1
        for (int rr = 0; rr <= RQ + 1; ++rr, arow += P * 2, oofs += S * C) {
            float hc[NP][4];
            if (!((rr == 0 && first_part) || (rr == RQ + 1 && last_part))) {
                float v[NP][4];
                #pragma unroll
                for (int j = 0; j < TR; ++j) {
                    // fragment row g -> pixel xbase + 2 j, row g + 8 -> pixel xbase + 2 j + 1
                    const uint32_t* base = arow + 2 * j * 2;
                    float acc[2][4];
                    #pragma unroll
                    for (int sub = 0; sub < 2; ++sub) { acc[sub][0] = acc[sub][1] = acc[sub][2] = acc[sub][3] = 0.f; }
                    #pragma unroll
                    for (int ky = 0; ky < 3; ++ky) {
                        const uint32_t a0 = base[ky * P * 2], a1 = base[ky * P * 2 + 2];
                        const uint32_t a2 = base[ky * P * 2 + 4], a3 = base[ky * P * 2 + 6];
                        #pragma unroll
                        for (int sub = 0; sub < 2; ++sub) {
                            mma16816(acc[sub], a0, a1, a2, a3, bf[ky][sub][1][0], bf[ky][sub][1][1]);   // lo weights first
                            mma16816(acc[sub], a0, a1, a2, a3, bf[ky][sub][0][0], bf[ky][sub][0][1]);
                        }
                    }
                    #pragma unroll
                    for (int sub = 0; sub < 2; ++sub) {
                        v[2 * j][2 * sub + 0] = acc[sub][0];
                        v[2 * j][2 * sub + 1] = acc[sub][1];
                        v[2 * j + 1][2 * sub + 0] = acc[sub][2];
                        v[2 * j + 1][2 * sub + 1] = acc[sub][3];
                    }
                    // values of idle lane groups are never looked at: their left neighbour is the last group of
                    // the row, whose has_right flag is clear
                }
                // horizontal 3-max: only the first / last pixel of a thread need a neighbour lane
                #pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float ln = __shfl_up_sync(0xffffffffu, v[NP - 1][c], 4);               // last pixel of lane group g-1
                    float rn = __shfl_down_sync(0xffffffffu, v[0][c], 4);                  // first pixel of lane group g+1
                    ln = has_left ? ln : 0.f;
                    rn = has_right ? rn : 0.f;
                    #pragma unroll
                    for (int i = 0; i < NP; ++i) {
                        const float l = i > 0 ? v[i > 0 ? i - 1 : 0][c] : ln;
                        const float r = i < NP - 1 ? v[i < NP - 1 ? i + 1 : i][c] : rn;
                        hc[i][c] = fmaxf(fmaxf(l, v[i][c]), r);
                    }
                }
            } else {
                #pragma unroll
                for (int i = 0; i < NP; ++i)
                    #pragma unroll
                    for (int c = 0; c < 4; ++c) hc[i][c] = 0.f;
            }
            // vertical: pooled row yy-1 = relu(max(h[yy-2], h[yy-1], h[yy]))
            if (rr >= 2 && owns) {                             // output row ya - 2 + rr, rr = 2 .. RQ + 1
                #pragma unroll
                for (int i = 0; i < NP; ++i) {
                    {
                        float m[4];
                        #pragma unroll
                        for (int c = 0; c < 4; ++c) m[c] = fmaxf(fmaxf(hA[i][c], hc[i][c]), 0.f);
                        if (OUT_SPLIT) {
                            const __half2 h01 = __floats2half2_rn(m[0], m[1]), h23 = __floats2half2_rn(m[2], m[3]);
                            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                            const __half2 l01 = __floats2half2_rn(m[0] - f01.x, m[1] - f01.y), l23 = __floats2half2_rn(m[2] - f23.x, m[3] - f23.y);
                            uint2 uh, ul;
                            uh.x = *reinterpret_cast<const uint32_t*>(&h01); uh.y = *reinterpret_cast<const uint32_t*>(&h23);
                            ul.x = *reinterpret_cast<const uint32_t*>(&l01); ul.y = *reinterpret_cast<const uint32_t*>(&l23);
                            *reinterpret_cast<uint2*>(out_hi + oofs + i * PIX) = uh;
                            *reinterpret_cast<uint2*>(out_hi + oofs + lo_delta + i * PIX) = ul;
                        } else {
                            *reinterpret_cast<float4*>(out_f32 + oofs + i * PIX) =
                                make_float4(m[0] * inv_a_scale, m[1] * inv_a_scale, m[2] * inv_a_scale, m[3] * inv_a_scale);
                        }
                    }
                }
            }
            #pragma unroll
            for (int i = 0; i < NP; ++i)
                #pragma unroll
                for (int c = 0; c < 4; ++c) {
                    hA[i][c] = fmaxf(hB[i][c], hc[i][c]);
                    hB[i][c] = hc[i][c];
                }
        }
    }
}

template <int S>
static size_t feat3_smem_bytes() {
    return (size_t)2 * StreamCfg<S>::WPC * StreamCfg<S>::INH * 2 + (size_t)2 * StreamCfg<S>::WPC * S * S * 3;
}

static size_t feat2_smem_bytes(int S, int R) {
    const int P = S + 2;
    const size_t inh = (size_t)((P * P * 4 + 16 + 7) & ~7) * 2;
    return inh + (size_t)(R + 2) * P * CS * 4 + (size_t)6 * S * 4;
}

static int pick_rows2(int S) {
    for (int R = S; R > 4; R -= 4)
        if (feat2_smem_bytes(S, R) <= 100 * 1024) return R;
    return 4;
}

// ------------------------------------------------------------------ general architecture (correctness path)
// Several conv layers, other filter counts, conv / pool strides (network/net.py:189-221 with the cf keys
// conv_filter_sizes, conv_filter_size, conv_stride, pooling_size, pooling_stride; TF SAME padding:
// out = ceil(in / s), pad_total = max((out - 1) s + k - in, 0), pad_before = pad_total / 2).  One CTA per
// window walks the layers with fp32 FMAs on ping-pong activation maps in a global scratch slice.  The
// reference's default net never takes this path; it exists so that the stage-model interface is complete.
struct GenNet {
    int n_conv, ksize, conv_stride, pool_size, pool_stride;
    int chan[8];
    const float* w[8];
    const float* b[8];
};

__device__ __forceinline__ int same_out(int in, int s) { return (in + s - 1) / s; }
__device__ __forceinline__ int same_pad_before(int in, int k, int s) {
    const int total = max((same_out(in, s) - 1) * s + k - in, 0);
    return total / 2;
}

template <int OUT_SPLIT>
__global__ void __launch_bounds__(FEAT_THREADS)
features_general_kernel(SrcDev src, int64_t first, int64_t n, int S, GenNet net, float* __restrict__ scratch, int64_t max_act,
                        float a_scale, float* __restrict__ out_f32, __half* __restrict__ out_hi, __half* __restrict__ out_lo,
                        int64_t K, const int64_t* __restrict__ n_dev, int64_t n_off) {
    extern __shared__ int gtabs[];
    n = eff_count(n, n_dev, n_off);
    float* bufA = scratch + (int64_t)blockIdx.x * 2 * max_act;
    float* bufB = bufA + max_act;
    const int tid = threadIdx.x;
    for (int64_t wrel = blockIdx.x; wrel < n; wrel += gridDim.x) {
        const int64_t w = first + wrel;
        __syncthreads();
        // ---- input window as float [S][S][3]
        Crop crop;
        if (src.kind == RODC_SRC_RECROP) { crop = make_crop(src, w); fill_recrop_tables(crop, S, gtabs); }
        __syncthreads();
        for (int i = tid; i < S * S; i += FEAT_THREADS) {
            const int y = i / S, x = i - y * S;
            float p0, p1, p2;
            if (src.kind == RODC_SRC_RECROP) {
                int px[3];
                recrop_pixel(crop, S, gtabs, y, x, px);
                p0 = (float)px[0]; p1 = (float)px[1]; p2 = (float)px[2];
            } else {
                const uint8_t* p; int pitch;
                if (src.kind == RODC_SRC_DENSE) { p = src.dense + w * (int64_t)S * S * 3; pitch = S * 3; }
                else if (src.kind == RODC_SRC_GATHER) { p = src.gather_base + src.win_off[w]; pitch = src.win_pitch[w]; }
                else {
                    const int64_t f = w / src.n_windows, j = w - f * src.n_windows;
                    const int32_t* wr = src.win + j * 8;
                    const LevelInfo& lv = src.levels[wr[0]];
                    p = src.pyramid + f * src.pyramid_bytes + lv.byte_off + ((int64_t)wr[2] * lv.w + wr[1]) * 3;
                    pitch = lv.w * 3;
                }
                const uint8_t* q = p + (int64_t)y * pitch + x * 3;
                p0 = (float)q[0]; p1 = (float)q[1]; p2 = (float)q[2];
            }
            bufA[i * 3 + 0] = p0; bufA[i * 3 + 1] = p1; bufA[i * 3 + 2] = p2;
        }
        __syncthreads();
        int H = S, cin = 3;
        for (int l = 0; l < net.n_conv; ++l) {
            const int cout = net.chan[l], k = net.ksize, cs = net.conv_stride;
            const int Ho = same_out(H, cs), pb = same_pad_before(H, k, cs);
            // conv + bias + ReLU: bufA [H][H][cin] -> bufB [Ho][Ho][cout]
            for (int i = tid; i < Ho * Ho * cout; i += FEAT_THREADS) {
                const int co = i % cout, xo = (i / cout) % Ho, yo = i / (cout * Ho);
                float acc = __ldg(net.b[l] + co);
                for (int ky = 0; ky < k; ++ky) {
                    const int yi = yo * cs + ky - pb;
                    if (yi < 0 || yi >= H) continue;
                    for (int kx = 0; kx < k; ++kx) {
                        const int xi = xo * cs + kx - pb;
                        if (xi < 0 || xi >= H) continue;
                        const float* in = bufA + ((int64_t)yi * H + xi) * cin;
                        const float* wt = net.w[l] + ((int64_t)(ky * k + kx) * cin) * cout + co;
                        for (int ci = 0; ci < cin; ++ci) acc = fmaf(in[ci], __ldg(wt + (int64_t)ci * cout), acc);
                    }
                }
                bufB[i] = fmaxf(acc, 0.f);
            }
            __syncthreads();
            // max pool SAME (padding excluded): bufB [Ho][Ho][cout] -> bufA [Hp][Hp][cout]
            const int ps = net.pool_size, pst = net.pool_stride;
            const int Hp = same_out(Ho, pst), ppb = same_pad_before(Ho, ps, pst);
            for (int i = tid; i < Hp * Hp * cout; i += FEAT_THREADS) {
                const int co = i % cout, xo = (i / cout) % Hp, yo = i / (cout * Hp);
                float m = -INFINITY;
                for (int ky = 0; ky < ps; ++ky) {
                    const int yi = yo * pst + ky - ppb;
                    if (yi < 0 || yi >= Ho) continue;
                    for (int kx = 0; kx < ps; ++kx) {
                        const int xi = xo * pst + kx - ppb;
                        if (xi < 0 || xi >= Ho) continue;
                        m = fmaxf(m, bufB[((int64_t)yi * Ho + xi) * cout + co]);
                    }
                }
                bufA[i] = m;
            }
            __syncthreads();
            H = Hp; cin = cout;
        }
        // ---- NHWC flatten -> output row
        for (int64_t i = tid; i < K; i += FEAT_THREADS) {
            const float v = bufA[i];
            if (OUT_SPLIT) {
                const float sv = v * a_scale;
                const __half h = __float2half_rn(sv);
                out_hi[wrel * K + i] = h;
                out_lo[wrel * K + i] = __float2half_rn(sv - __half2float(h));
            } else {
                out_f32[wrel * K + i] = v;
            }
        }
    }
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
    } else if (src->kind == RODC_SRC_GATHER) {
        RODC_CHECK_ARG(src->gather_base && src->win_off && src->win_pitch, "gather window source incomplete");
        d.gather_base = src->gather_base; d.win_off = src->win_off; d.win_pitch = src->win_pitch;
    } else {
        set_error("unknown window source kind %d", src->kind);
        return RODC_ERR_ARG;
    }
    *out = d;
    return RODC_OK;
}

// measurement only (profile level 2): sum of the survivors' crop areas, the data-dependent term of K4's
// algorithmic bytes 3 (w_n h_n + S^2) per survivor
__global__ void __launch_bounds__(256) crop_area_kernel(SrcDev src, int64_t first, int64_t n, unsigned long long* __restrict__ acc) {
    unsigned long long a = 0;
    for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < n; w += (int64_t)gridDim.x * blockDim.x) {
        const Crop c = make_crop(src, first + w);
        a += (unsigned long long)max(c.sw, 0) * (unsigned long long)max(c.sh, 0);
    }
    for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((threadIdx.x & 31) == 0 && a) atomicAdd(acc, a);
}

int launch_source_windows(const SrcDev& src, int64_t first, int64_t n, int size, uint8_t* out, cudaStream_t st,
                          const int64_t* n_dev, int64_t n_off) {
    if (n <= 0) return RODC_OK;
    const unsigned grid = (unsigned)(n < 148 * 64 ? n : 148 * 64);
    const bool rc_src = src.kind == RODC_SRC_RECROP;
    // stage-0 copy: window bytes read + written; K4: 3 S^2 written per survivor + 3 w_n h_n read (added below)
    Prof prof(rc_src ? KC_RECROP : KC_SOURCE, st, size, (rc_src ? 1.0 : 2.0) * (double)n * size * size * 3, 0.0);
    if (rc_src && prof.recording()) {
        if (unsigned long long* acc = prof.area_slot())
            crop_area_kernel<<<(unsigned)((n + 255) / 256 < 592 ? (n + 255) / 256 : 592), 256, 0, st>>>(src, first, n, acc);
    }
    const bool aligned = (reinterpret_cast<uintptr_t>(out) & 3) == 0;
    // (A variant that first stages the source rows in shared memory with coalesced 16-byte loads was measured SLOWER:
    // 254 vs 203 us per 18 944 windows at S = 48, 256 vs 128 us per 35 k at S = 24, profiles/r2b_ncu_small_kernels.json.
    // With the aligned 8-byte tap loads this kernel is bound by its integer ALU work -- sm 84 %, l1tex 65 % -- not by
    // the L1 request path any more, and the staging pass only adds instructions.)
    if (aligned && rc_src && size == 48) recrop_windows_kernel<48, 16><<<grid, 192, 0, st>>>(src, first, n, out, n_dev, n_off);
    else if (aligned && rc_src && size == 24) recrop_windows_kernel<24, 24><<<grid, 144, 0, st>>>(src, first, n, out, n_dev, n_off);
    else if (aligned && !rc_src && (size & 3) == 0) {
        const int64_t groups = (n + COPY_G - 1) / COPY_G;
        window_copy_kernel<<<(unsigned)(groups < 148 * 16 ? groups : 148 * 16), 128, 0, st>>>(src, first, n, size,
                                                                                             reinterpret_cast<uint32_t*>(out), n_dev, n_off);
    }
    else source_windows_kernel<<<grid, 128, (size_t)6 * size * sizeof(int), st>>>(src, first, n, size, out, n_dev, n_off);
    RODC_LAUNCH_CHECK();
    return RODC_OK;
}

int launch_features(const rodc_stage* s, const SrcDev& src, int64_t first, int64_t n,
                    float* out_f32, __half* out_hi, __half* out_lo, cudaStream_t st, float* general_scratch,
                    const int64_t* n_dev, int64_t n_off) {
    if (n <= 0) return RODC_OK;
    if (s->general) {
        RODC_CHECK_ARG(general_scratch != nullptr, "general-architecture stage needs activation scratch");
        RODC_CHECK_ARG(n < (1ll << 31), "too many windows in one launch");
        GenNet net;
        net.n_conv = s->n_conv; net.ksize = s->ksize; net.conv_stride = s->conv_stride;
        net.pool_size = s->pool_size; net.pool_stride = s->pool_stride;
        for (int l = 0; l < 8; ++l) { net.chan[l] = s->chan[l]; net.w[l] = s->gconv_w[l]; net.b[l] = s->gconv_b[l]; }
        const unsigned grid = (unsigned)(n < GENERAL_CTAS ? n : GENERAL_CTAS);
        const size_t smem = (size_t)6 * s->in_size * sizeof(int);
        Prof prof(KC_FEAT, st, s->in_size, (double)n * ((double)s->in_size * s->in_size * 3 + 4.0 * (double)s->K), 0.0);
        if (out_f32) features_general_kernel<0><<<grid, FEAT_THREADS, smem, st>>>(src, first, n, s->in_size, net, general_scratch, s->max_act,
                                                                                  1.f, out_f32, nullptr, nullptr, s->K, n_dev, n_off);
        else features_general_kernel<1><<<grid, FEAT_THREADS, smem, st>>>(src, first, n, s->in_size, net, general_scratch, s->max_act,
                                                                          ldexpf(1.f, s->a_exp), nullptr, out_hi, out_lo, s->K, n_dev, n_off);
        RODC_LAUNCH_CHECK();
        return RODC_OK;
    }
    RODC_CHECK_ARG(s->channels == 32, "feature kernel is built for 32 conv channels (got %d)", s->channels);
    RODC_CHECK_ARG((s->in_size & 3) == 0, "stage input size must be a multiple of 4 (got %d)", s->in_size);
    RODC_CHECK_ARG(n < (1ll << 31), "too many windows in one launch");
    const int S = s->in_size;
    // bytes: the u8 window in, 4 B per feature out (the materialised fc1 operand -- design traffic, not part of
    // SURVEY §8d's algorithmic bytes of K2, whose bound is the tensor pipe); flops: the 3x3x3 -> 32 conv
    Prof prof(KC_FEAT, st, S, (double)n * ((double)S * S * 3 + 4.0 * (double)s->K), 2.0 * 27.0 * (double)n * (double)s->K);
    const float a_scale = ldexpf(1.f, s->a_exp);
    const float conv_scale = ldexpf(1.f, -s->c_exp);
    // persistent-style grid: two CTAs per SM, each loops over windows
    const unsigned grid = (unsigned)(n < 148 * 2 ? n : 148 * 2);
    // RODC_FEATURES_GENERIC=1 forces the generic kernel (cross-check of the specialised one)
    static int use_v2 = -1;
    if (use_v2 < 0) { const char* e = getenv("RODC_FEATURES_GENERIC"); use_v2 = (e && e[0] == '1') ? 1 : 0; }
    // the streaming kernels read materialised u8 windows (rodc_stage_forward copies every other source into its
    // workspace first); other sources and sizes take the generic kernel
    if (!use_v2 && (S == 12 || S == 24 || S == 48) && src.kind == RODC_SRC_DENSE && (reinterpret_cast<uintptr_t>(src.dense) & 3) == 0) {
#define RODC_LAUNCH_FEAT3_OCC(SPLIT, ST_, OCC_)                                                                            \
        do {                                                                                                              \
            const size_t smem = feat3_smem_bytes<ST_>();                                                                  \
            RODC_CUDA(cudaFuncSetAttribute(features_stream_kernel<SPLIT, ST_, OCC_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
            const unsigned g3 = (unsigned)(n < 148 * OCC_ ? n : 148 * OCC_);                                                \
            features_stream_kernel<SPLIT, ST_, OCC_><<<g3, FEAT_THREADS, smem, st>>>(src.dense, first, n, s->conv_frag_perm, \
                                                                                 ldexpf(1.f, s->a_exp - s->c_exp),          \
                                                                                 ldexpf(1.f, -s->a_exp), out_f32, out_hi, out_lo, n_dev, n_off, \
                                                                                 SPLIT ? probe_a_wrap(ST_) : 0); \
        } while (0)
#ifdef RODC_PROBES
        // probe build: RODC_PROBE_FEAT_OCC3=1 runs the 24 / 48-px kernels with three CTAs per SM (80 registers, ~400 B
        // of spills per thread) instead of two (128 registers, none): an occupancy experiment, same results
        static int occ3 = -1;
        if (occ3 < 0) { const char* e = getenv("RODC_PROBE_FEAT_OCC3"); occ3 = (e && e[0] == '1') ? 1 : 0; }
#define RODC_LAUNCH_FEAT3(SPLIT, ST_)                                                                                      \
        do { if (occ3 && ST_ != 12) RODC_LAUNCH_FEAT3_OCC(SPLIT, ST_, 3); else RODC_LAUNCH_FEAT3_OCC(SPLIT, ST_, StreamCfg<ST_>::CTAS_PER_SM); } while (0)
#else
#define RODC_LAUNCH_FEAT3(SPLIT, ST_) RODC_LAUNCH_FEAT3_OCC(SPLIT, ST_, StreamCfg<ST_>::CTAS_PER_SM)
#endif
        if (out_f32) { if (S == 12) RODC_LAUNCH_FEAT3(0, 12); else if (S == 24) RODC_LAUNCH_FEAT3(0, 24); else RODC_LAUNCH_FEAT3(0, 48); }
        else { if (S == 12) RODC_LAUNCH_FEAT3(1, 12); else if (S == 24) RODC_LAUNCH_FEAT3(1, 24); else RODC_LAUNCH_FEAT3(1, 48); }
#undef RODC_LAUNCH_FEAT3
#undef RODC_LAUNCH_FEAT3_OCC
        RODC_LAUNCH_CHECK();
        return RODC_OK;
    }
#define RODC_LAUNCH_FEAT(SPLIT, ST_, RT_, Rv)                                                                              \
    do {                                                                                                                  \
        const size_t smem = feat2_smem_bytes(S, Rv);                                                                      \
        RODC_CUDA(cudaFuncSetAttribute(features_mma_kernel<SPLIT, ST_, RT_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        features_mma_kernel<SPLIT, ST_, RT_><<<grid, FEAT_THREADS, smem, st>>>(src, first, n, S, Rv, s->conv_frag, s->conv_b, \
                                                                                conv_scale, a_scale, out_f32, out_hi, out_lo, n_dev, n_off, s->kperm_side); \
    } while (0)
    { const int R = pick_rows2(S); if (out_f32) RODC_LAUNCH_FEAT(0, 0, 0, R); else RODC_LAUNCH_FEAT(1, 0, 0, R); }
#undef RODC_LAUNCH_FEAT
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
    return launch_source_windows(d, 0, n, size, out, as_stream(stream));
}

extern "C" int rodc_source_windows(const rodc_window_source* src, int64_t first, int64_t n, int size, uint8_t* out,
                                   void* stream) {
    RODC_CHECK_ARG(src && out, "rodc_source_windows: NULL argument");
    RODC_CHECK_ARG(size > 0 && size <= 1024 && first >= 0, "rodc_source_windows: bad size %d / first %lld", size, (long long)first);
    if (n <= 0) return RODC_OK;
    SrcDev d;
    int rc = resolve_source(src, size, &d); if (rc) return rc;
    return launch_source_windows(d, first, n, size, out, as_stream(stream));
}
