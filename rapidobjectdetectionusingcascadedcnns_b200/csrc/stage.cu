// K2 host side: stage handle (weights prepared on the device) and the stage forward
// (features -> fc1 -> fc2/softmax), chunked over the caller's workspace.
// Replaces session.run of one frozen stage graph (app/inference_cascade_app.py:196-215).
#include "stage.cuh"
#include <math.h>
#include <string.h>
#include <stdlib.h>
#include <new>
#include <vector>

rodc_stage::~rodc_stage() {
    float* f[] = {conv_w, conv_b, fc1_w, fc1_b, fc2_w, fc2_b};
    for (float* p : f) if (p) cudaFree(p);
    if (fc1_hi) cudaFree(fc1_hi);
    if (fc1_lo) cudaFree(fc1_lo);
    if (conv_frag) cudaFree(conv_frag);
    for (int l = 0; l < 8; ++l) { if (gconv_w[l]) cudaFree(gconv_w[l]); if (gconv_b[l]) cudaFree(gconv_b[l]); }
    if (conv_frag_perm) cudaFree(conv_frag_perm);
}

namespace rodc {

// W[K, N] fp32 (N contiguous) -> hi/lo fp16 planes [N, K] (K contiguous), W * 2^w_exp = hi + lo.
// 32x32 smem transpose so both the read and the writes are coalesced.
__global__ void __launch_bounds__(256) split_transpose_kernel(const float* __restrict__ W, int64_t K, int N, float scale,
                                                              __half* __restrict__ hi, __half* __restrict__ lo, int perm_side) {
    __shared__ float tile[32][33];
    const int64_t k0 = (int64_t)blockIdx.x * 32;
    const int n0 = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;       // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int64_t k = k0 + r;
        tile[r][tx] = (k < K && n0 + tx < N) ? W[k * N + n0 + tx] : 0.f;
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int n = n0 + r; const int64_t k = k0 + tx;
        if (n < N && k < K) {
            const float v = tile[tx][r] * scale;
            const __half h = __float2half_rn(v);
            const int64_t kp = kperm(perm_side, k);             // the order the feature kernels write A in
            hi[(int64_t)n * K + kp] = h;
            lo[(int64_t)n * K + kp] = __float2half_rn(v - __half2float(h));
        }
    }
}

int prepare_fc1_split(rodc_stage* s, cudaStream_t st) {
    dim3 grid((unsigned)((s->K + 31) / 32), (unsigned)((s->fc1_size + 31) / 32));
    Prof prof(KC_MISC, st, 0, 8.0 * (double)s->K * s->fc1_size, 0.0);
    split_transpose_kernel<<<grid, 256, 0, st>>>(s->fc1_w, s->K, s->fc1_size, ldexpf(1.f, s->w_exp), s->fc1_hi, s->fc1_lo,
                                                 s->kperm_side);
    RODC_LAUNCH_CHECK();
    return RODC_OK;
}

int probe_a_wrap(int in_size) {
#ifdef RODC_PROBES
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("RODC_PROBE_A_WRAP");
        v = e ? atoi(e) : 0;
        if (v < 0 || (v % 256) != 0) v = 0;
        if (v) fprintf(stderr, "librodc: RODC_PROBE_A_WRAP=%d is active -- the 48-px stage produces WRONG results (probe build)\n", v);
    }
    return in_size == 48 ? v : 0;
#else
    (void)in_size;
    return 0;
#endif
}

static int upload(float** dst, const float* src, size_t n) {
    RODC_CUDA(cudaMalloc(dst, (n ? n : 1) * sizeof(float)));
    if (n) RODC_CUDA(cudaMemcpy(*dst, src, n * sizeof(float), cudaMemcpyHostToDevice));
    return RODC_OK;
}

// largest e with bound * 2^e <= limit
static int pow2_headroom(double bound, double limit) {
    if (!(bound > 0.0)) return 0;
    int e = (int)floor(log2(limit / bound));
    if (e > 30) e = 30;
    if (e < -30) e = -30;
    return e;
}

}  // namespace rodc

using namespace rodc;

extern "C" int rodc_stage_create(int in_size, int conv_channels, int fc1_size, int n_prev,
                                 const float* conv_w, const float* conv_b,
                                 const float* fc1_w, const float* fc1_b,
                                 const float* fc2_w, const float* fc2_b, rodc_stage** out) {
    RODC_CHECK_ARG(out && conv_w && conv_b && fc1_w && fc1_b && fc2_w && fc2_b, "rodc_stage_create: NULL argument");
    RODC_CHECK_ARG(in_size >= 4 && (in_size % 4) == 0 && in_size <= 256, "rodc_stage_create: unsupported input size %d", in_size);
    RODC_CHECK_ARG(conv_channels == 32, "rodc_stage_create: only 32 conv channels are built (got %d)", conv_channels);
    RODC_CHECK_ARG(fc1_size > 0 && fc1_size % 256 == 0, "rodc_stage_create: fc1_size must be a multiple of 256 (got %d)", fc1_size);
    RODC_CHECK_ARG(n_prev >= 0 && n_prev % 4 == 0, "rodc_stage_create: bad previous-bottleneck width %d", n_prev);
    rodc_stage* s = new (std::nothrow) rodc_stage();
    if (!s) { set_error("out of host memory"); return RODC_ERR_NOMEM; }
    s->in_size = in_size; s->channels = conv_channels; s->fc1_size = fc1_size; s->n_prev = n_prev;
    s->K = (int64_t)in_size * in_size * conv_channels;
    s->kperm_side = (in_size == 12 || in_size == 24 || in_size == 48) ? in_size : 0;
    int rc = cudaGetDevice(&s->device) == cudaSuccess ? RODC_OK : RODC_ERR_CUDA;
    if (rc) { set_error("no CUDA device"); delete s; return rc; }
    const size_t nconv = (size_t)27 * conv_channels, nfc1 = (size_t)s->K * fc1_size, nfc2 = (size_t)(fc1_size + n_prev) * 2;
    if ((rc = upload(&s->conv_w, conv_w, nconv)) || (rc = upload(&s->conv_b, conv_b, conv_channels)) ||
        (rc = upload(&s->fc1_w, fc1_w, nfc1)) || (rc = upload(&s->fc1_b, fc1_b, fc1_size)) ||
        (rc = upload(&s->fc2_w, fc2_w, nfc2)) || (rc = upload(&s->fc2_b, fc2_b, 2))) { delete s; return rc; }
    // scaling exponents for the fp16 split: bound the largest possible feature / weight
    double fbound = 0.0;
    for (int c = 0; c < conv_channels; ++c) {
        double l1 = 0.0;
        for (int t = 0; t < 27; ++t) l1 += fabs((double)conv_w[t * conv_channels + c]);
        double b = 255.0 * l1 + fabs((double)conv_b[c]);
        if (b > fbound) fbound = b;
    }
    double wmax = 0.0;
    for (size_t i = 0; i < nfc1; ++i) { double a = fabs((double)fc1_w[i]); if (a > wmax) wmax = a; }
    s->a_exp = pow2_headroom(fbound, 32768.0);
    s->w_exp = pow2_headroom(wmax, 16384.0);
    // conv weights as mma.sync.m16n8k16 B fragments (features.cu): K index j of filter row ky is
    // (kx = j / 4, ci = j % 4); kx == 3 and ci == 3 are zero padding.  w * 2^c_exp = hi + lo (fp16).
    {
        double cmax = 0.0;
        for (size_t i = 0; i < nconv; ++i) { double a = fabs((double)conv_w[i]); if (a > cmax) cmax = a; }
        for (int c = 0; c < conv_channels; ++c) { double a = fabs((double)conv_b[c]); if (a > cmax) cmax = a; }   // bias rides in a K slot
        s->c_exp = pow2_headroom(cmax, 16384.0);
        // the streaming feature kernel stages pixels as u8 * 2^(a_exp - c_exp) in fp16: keep that exponent
        // inside the fp16 normal range (always true for sane nets: it is about -7 .. -12)
        if (s->a_exp - s->c_exp < -14) s->c_exp = s->a_exp + 14;
        if (s->a_exp - s->c_exp > 7) s->c_exp = s->a_exp - 7;
        const float cs = ldexpf(1.f, s->c_exp);
        std::vector<uint32_t> frag((size_t)3 * 4 * 2 * 32 * 2, 0u), fragp(frag.size(), 0u);
        auto wval = [&](int ky, int j, int nn) -> float {
            const int kx = j >> 2, ci = j & 3;
            if (kx > 2 || ci > 2) return 0.f;
            return conv_w[((ky * 3 + kx) * 3 + ci) * conv_channels + nn] * cs;
        };
        auto halfbits = [](float v, int lo) -> uint32_t {
            const __half h = __float2half_rn(v);
            __half r = h;
            if (lo) r = __float2half_rn(v - __half2float(h));
            unsigned short b; memcpy(&b, &r, 2);
            return (uint32_t)b;
        };
        for (int ky = 0; ky < 3; ++ky)
            for (int nt = 0; nt < 4; ++nt)
                for (int h = 0; h < 2; ++h)
                    for (int lane = 0; lane < 32; ++lane) {
                        const int gg = lane >> 2, tt = lane & 3, nn = nt * 8 + gg;
                        for (int r = 0; r < 2; ++r) {
                            const int j0 = 2 * tt + 8 * r;
                            const uint32_t lo16 = halfbits(wval(ky, j0, nn), h), hi16 = halfbits(wval(ky, j0 + 1, nn), h);
                            frag[((((size_t)(ky * 4 + nt) * 2 + h) * 32) + lane) * 2 + r] = lo16 | (hi16 << 16);
                            // permuted variant (features v3): n-tile nt = 2*cg + sub, column gg -> channel
                            // 16*cg + 4*(gg/2) + 2*sub + gg%2, so lane t owns channels 4t..4t+3 of its group
                            const int np = 16 * (nt >> 1) + 4 * (gg >> 1) + 2 * (nt & 1) + (gg & 1);
                            // K slot (ky = 1, kx = 1, ci = 3) multiplies the constant-one channel of the centre pixel: bias
                            auto wvalb = [&](int jj) -> float { return (ky == 1 && jj == 7) ? conv_b[np] * cs : wval(ky, jj, np); };
                            const uint32_t plo = halfbits(wvalb(j0), h), phi = halfbits(wvalb(j0 + 1), h);
                            fragp[((((size_t)(ky * 4 + nt) * 2 + h) * 32) + lane) * 2 + r] = plo | (phi << 16);
                        }
                    }
        if (cudaMalloc(&s->conv_frag, frag.size() * 4) != cudaSuccess ||
            cudaMemcpy(s->conv_frag, frag.data(), frag.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
            set_error("upload of the conv fragments failed"); delete s; return RODC_ERR_NOMEM;
        }
        if (cudaMalloc(&s->conv_frag_perm, fragp.size() * 4) != cudaSuccess ||
            cudaMemcpy(s->conv_frag_perm, fragp.data(), fragp.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) {
            set_error("upload of the conv fragments failed"); delete s; return RODC_ERR_NOMEM;
        }
    }
    cudaError_t e1 = cudaMalloc(&s->fc1_hi, nfc1 * sizeof(__half));
    cudaError_t e2 = cudaMalloc(&s->fc1_lo, nfc1 * sizeof(__half));
    if (e1 != cudaSuccess || e2 != cudaSuccess) { set_error("cudaMalloc of the split fc1 weights failed"); delete s; return RODC_ERR_NOMEM; }
    if ((rc = prepare_fc1_split(s, 0))) { delete s; return rc; }
    if (cudaDeviceSynchronize() != cudaSuccess) { set_error("weight preparation failed: %s", cudaGetErrorString(cudaGetLastError())); delete s; return RODC_ERR_CUDA; }
    *out = s;
    return RODC_OK;
}

extern "C" int rodc_stage_create_general(int in_size, int n_conv, const int32_t* channels, int ksize, int conv_stride,
                                         int pool_size, int pool_stride,
                                         const float* const* conv_w, const float* const* conv_b,
                                         int fc1_size, int n_prev, const float* fc1_w, const float* fc1_b,
                                         const float* fc2_w, const float* fc2_b, rodc_stage** out) {
    RODC_CHECK_ARG(out && channels && conv_w && conv_b && fc1_w && fc1_b && fc2_w && fc2_b, "rodc_stage_create_general: NULL argument");
    RODC_CHECK_ARG(n_conv >= 1 && n_conv <= 8, "rodc_stage_create_general: 1..8 conv layers (got %d)", n_conv);
    RODC_CHECK_ARG(ksize >= 1 && ksize <= 11 && conv_stride >= 1 && pool_size >= 1 && pool_stride >= 1,
                   "rodc_stage_create_general: bad geometry");
    // the reference default takes the specialised kernels
    if (n_conv == 1 && channels[0] == 32 && ksize == 3 && conv_stride == 1 && pool_size == 3 && pool_stride == 1 &&
        in_size % 4 == 0 && fc1_size % 256 == 0)
        return rodc_stage_create(in_size, 32, fc1_size, n_prev, conv_w[0], conv_b[0], fc1_w, fc1_b, fc2_w, fc2_b, out);
    RODC_CHECK_ARG(in_size >= 1 && in_size <= 512, "rodc_stage_create_general: bad input size %d", in_size);
    RODC_CHECK_ARG(fc1_size > 0 && fc1_size % 128 == 0, "rodc_stage_create_general: fc1_size must be a multiple of 128");
    rodc_stage* s = new (std::nothrow) rodc_stage();
    if (!s) { set_error("out of host memory"); return RODC_ERR_NOMEM; }
    s->general = 1; s->in_size = in_size; s->fc1_size = fc1_size; s->n_prev = n_prev;
    s->n_conv = n_conv; s->ksize = ksize; s->conv_stride = conv_stride; s->pool_size = pool_size; s->pool_stride = pool_stride;
    if (cudaGetDevice(&s->device) != cudaSuccess) { set_error("no CUDA device"); delete s; return RODC_ERR_CUDA; }
    int H = in_size, cin = 3, rc = RODC_OK;
    int64_t max_act = (int64_t)in_size * in_size * 3;
    double bound = 255.0;                                    // bound of the activations, for the fp16-split exponent
    for (int l = 0; l < n_conv; ++l) {
        const int cout = channels[l];
        RODC_CHECK_ARG(cout >= 1 && cout <= 1024, "rodc_stage_create_general: bad channel count %d", cout);
        s->chan[l] = cout;
        const size_t nw = (size_t)ksize * ksize * cin * cout;
        if ((rc = upload(&s->gconv_w[l], conv_w[l], nw)) || (rc = upload(&s->gconv_b[l], conv_b[l], cout))) { delete s; return rc; }
        double worst = 0.0;
        for (int co = 0; co < cout; ++co) {
            double l1 = 0.0;
            for (size_t t = 0; t < (size_t)ksize * ksize * cin; ++t) l1 += fabs((double)conv_w[l][t * cout + co]);
            const double b = bound * l1 + fabs((double)conv_b[l][co]);
            if (b > worst) worst = b;
        }
        bound = worst;
        const int Ho = (H + conv_stride - 1) / conv_stride;
        if ((int64_t)Ho * Ho * cout > max_act) max_act = (int64_t)Ho * Ho * cout;
        H = (Ho + pool_stride - 1) / pool_stride;
        cin = cout;
    }
    s->channels = cin;
    s->K = (int64_t)H * H * cin;
    s->max_act = max_act;
    RODC_CHECK_ARG(s->K % 8 == 0, "rodc_stage_create_general: flattened size %lld must be a multiple of 8", (long long)s->K);
    const size_t nfc1 = (size_t)s->K * fc1_size, nfc2 = (size_t)(fc1_size + n_prev) * 2;
    if ((rc = upload(&s->fc1_w, fc1_w, nfc1)) || (rc = upload(&s->fc1_b, fc1_b, fc1_size)) ||
        (rc = upload(&s->fc2_w, fc2_w, nfc2)) || (rc = upload(&s->fc2_b, fc2_b, 2))) { delete s; return rc; }
    double wmax = 0.0;
    for (size_t i = 0; i < nfc1; ++i) { double a = fabs((double)fc1_w[i]); if (a > wmax) wmax = a; }
    s->a_exp = pow2_headroom(bound, 32768.0);
    s->w_exp = pow2_headroom(wmax, 16384.0);
    if (cudaMalloc(&s->fc1_hi, nfc1 * sizeof(__half)) != cudaSuccess || cudaMalloc(&s->fc1_lo, nfc1 * sizeof(__half)) != cudaSuccess) {
        set_error("cudaMalloc of the split fc1 weights failed"); delete s; return RODC_ERR_NOMEM;
    }
    if ((rc = prepare_fc1_split(s, 0))) { delete s; return rc; }
    if (cudaDeviceSynchronize() != cudaSuccess) { set_error("weight preparation failed"); delete s; return RODC_ERR_CUDA; }
    *out = s;
    return RODC_OK;
}

extern "C" void rodc_stage_destroy(rodc_stage* s) { delete s; }

// general-architecture stages keep two activation maps per persistent CTA at the END of the workspace
static int64_t general_scratch_bytes(const rodc_stage* s) {
    return s->general ? (int64_t)GENERAL_CTAS * 2 * s->max_act * 4 + 256 : 0;
}

// bytes of one materialised u8 window (non-dense sources are copied into the workspace before the feature kernel)
static int64_t window_bytes(const rodc_stage* s) { return ((int64_t)s->in_size * s->in_size * 3 + 15) / 16 * 16; }

// workspace = feature chunk: fp32 [m,K] (SIMT) or two fp16 planes [m,K] (tensor): 4*K bytes per window,
// followed by the chunk's u8 windows [m, S, S, 3]
static int64_t chunk_windows(const rodc_stage* s, int64_t workspace_bytes) {
    workspace_bytes -= general_scratch_bytes(s);
    if (workspace_bytes < 0) return 0;
    int64_t m = workspace_bytes / (s->K * 4 + window_bytes(s));
    if (m >= 128) m = m / 128 * 128;
    return m;
}

extern "C" int64_t rodc_stage_workspace_bytes(const rodc_stage* s, int64_t n, int impl) {
    (void)impl;
    if (!s || n < 0) return -1;
    // enough for two full waves of 128-window tiles on 148 SMs, capped at the problem size
    int64_t m = (n + 127) / 128 * 128;
    const int64_t cap = 148 * 128 * (s->K > 40000 ? 1 : (s->K > 10000 ? 2 : 8));
    if (m > cap) m = cap;
    if (m < 128) m = 128;
    return m * (s->K * 4 + window_bytes(s)) + general_scratch_bytes(s);
}

extern "C" int rodc_stage_forward(const rodc_stage* s, const rodc_window_source* src, int64_t first, int64_t n,
                                  float* bottleneck, int ld_bott, float* logits, float* probs,
                                  void* workspace, int64_t workspace_bytes, int impl, void* stream) {
    return rodc_stage_forward_dyn(s, src, first, n, nullptr, bottleneck, ld_bott, logits, probs, workspace, workspace_bytes,
                                  impl, stream);
}

extern "C" int rodc_stage_forward_dyn(const rodc_stage* s, const rodc_window_source* src, int64_t first, int64_t n,
                                      const int64_t* n_dev, float* bottleneck, int ld_bott, float* logits, float* probs,
                                      void* workspace, int64_t workspace_bytes, int impl, void* stream) {
    RODC_CHECK_ARG(s && src && bottleneck && probs && workspace, "rodc_stage_forward: NULL argument");
    RODC_CHECK_ARG(ld_bott >= s->fc1_size + s->n_prev, "rodc_stage_forward: ld_bott %d < %d", ld_bott, s->fc1_size + s->n_prev);
    RODC_CHECK_ARG(impl == RODC_IMPL_TENSOR || impl == RODC_IMPL_SIMT, "rodc_stage_forward: unknown impl %d", impl);
    RODC_CHECK_ARG(!n_dev || impl == RODC_IMPL_TENSOR, "rodc_stage_forward_dyn: device-side counts need the tensor fc1");
    if (n <= 0) return RODC_OK;
    SrcDev sd;
    int rc = resolve_source(src, s->in_size, &sd); if (rc) return rc;
    const int64_t chunk = chunk_windows(s, workspace_bytes);
    RODC_CHECK_ARG(chunk >= 1, "rodc_stage_forward: workspace too small (%lld bytes, need >= %lld)",
                   (long long)workspace_bytes, (long long)(s->K * 4));
    cudaStream_t st = as_stream(stream);
    float* gscratch = nullptr;
    if (s->general) {
        const int64_t off = (workspace_bytes - general_scratch_bytes(s) + 255) / 256 * 256;
        gscratch = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + off);
        // the tensor fc1 needs fc1_size == 512 and K % 64 == 0; anything else runs the fp32 SIMT GEMM
        if (impl == RODC_IMPL_TENSOR && (s->fc1_size != 512 || s->K % 64 != 0)) impl = RODC_IMPL_SIMT;
        RODC_CHECK_ARG(!n_dev || impl == RODC_IMPL_TENSOR, "rodc_stage_forward_dyn: this architecture runs the SIMT fc1, which takes no device-side count");
    }
    // The default-architecture feature kernels stream dense u8 windows (asynchronous prefetch); pyramid / re-crop /
    // gather sources are materialised per chunk first -- a window is 2 % of the bytes it expands to, and a
    // separate, fully occupied kernel hides the gather latency that stalled the fused producer.
    const bool materialise = !s->general && sd.kind != RODC_SRC_DENSE;
    uint8_t* dense_ws = reinterpret_cast<uint8_t*>(workspace) + chunk * s->K * 4;
    for (int64_t c0 = 0; c0 < n; c0 += chunk) {
        const int64_t m = (n - c0) < chunk ? (n - c0) : chunk;
        float* bott = bottleneck + c0 * ld_bott;
        SrcDev csrc = sd;
        int64_t cfirst = first + c0;
        if (materialise) {
            if ((rc = launch_source_windows(sd, first + c0, m, s->in_size, dense_ws, st, n_dev, c0))) return rc;
            memset(&csrc, 0, sizeof(csrc));
            csrc.kind = RODC_SRC_DENSE; csrc.dense = dense_ws;
            cfirst = 0;
        }
        if (impl == RODC_IMPL_SIMT) {
            float* feat = reinterpret_cast<float*>(workspace);
            if ((rc = launch_features(s, csrc, cfirst, m, feat, nullptr, nullptr, st, gscratch))) return rc;
            if ((rc = launch_fc1_simt(s, feat, m, bott, ld_bott, st))) return rc;
        } else {
            __half* hi = reinterpret_cast<__half*>(workspace);
            __half* lo = hi + m * s->K;
            if ((rc = launch_features(s, csrc, cfirst, m, nullptr, hi, lo, st, gscratch, n_dev, c0))) return rc;
            if ((rc = launch_fc1_tensor(s, hi, lo, m, bott, ld_bott, st, n_dev, c0))) return rc;
        }
    }
    return launch_fc2_softmax(s, bottleneck, ld_bott, n, logits, probs, st, n_dev);
}

extern "C" int rodc_stage_features_f32(const rodc_stage* s, const rodc_window_source* src, int64_t first,
                                       int64_t n, float* features, void* stream) {
    RODC_CHECK_ARG(s && src && features, "rodc_stage_features_f32: NULL argument");
    if (s->general) { set_error("rodc_stage_features_f32 is a test hook of the default architecture only"); return RODC_ERR_UNSUP; }
    SrcDev sd;
    int rc = resolve_source(src, s->in_size, &sd); if (rc) return rc;
    return launch_features(s, sd, first, n, features, nullptr, nullptr, as_stream(stream));
}
