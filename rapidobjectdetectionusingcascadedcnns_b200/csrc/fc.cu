// K2b/K2c: fully connected layers of the stage net (network/net.py:233-236,139-165).
//   fc1_simt   : fp32 CUDA-core GEMM + bias + ReLU -- the cross-check implementation of the
//                tensor-core fc1 (fc1_tc.cu); same inputs, independent arithmetic.
//   fc2_softmax: [own 512 | previous bottleneck] x W2 + b2 -> softmax ("final_softmax").
#include "stage.cuh"

namespace rodc {

// C[m, N] = relu(A[m, K] * W[K, N] + b); A row-major (K contiguous), W row-major (N contiguous).
// 128x128 tile, BK = 8, 256 threads, 8x8 micro-tile.
__global__ void __launch_bounds__(256) fc1_simt_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                       const float* __restrict__ bias, float* __restrict__ out,
                                                       int ld_out, int64_t m, int64_t K, int N) {
    __shared__ __align__(16) float As[8][128 + 4];
    __shared__ __align__(16) float Bs[8][128];
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.y * 128;
    const int n0 = blockIdx.x * 128;
    const int ty = tid >> 4, tx = tid & 15;
    // two-level blocked summation (like Eigen's GEMM and like the tensor path): `acc` sums one block
    // of 64 k, `tot` sums the blocks -- keeps the fp32 rounding error ~sqrt(K/64) instead of ~sqrt(K)
    float acc[8][8], tot[8][8];
    #pragma unroll
    for (int i = 0; i < 8; ++i)
        #pragma unroll
        for (int j = 0; j < 8; ++j) { acc[i][j] = 0.f; tot[i][j] = 0.f; }
    const int a_row = tid >> 1, a_kq = (tid & 1) * 4;
    const int b_k = tid >> 5, b_n = (tid & 31) * 4;
    const bool a_ok = (m0 + a_row) < m;
    const float* a_ptr = A + (m0 + a_row) * K + a_kq;
    const float* b_ptr = W + (int64_t)b_k * N + n0 + b_n;
    for (int64_t k0 = 0; k0 < K; k0 += 8) {
        float4 av = a_ok ? *reinterpret_cast<const float4*>(a_ptr + k0) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 bv = *reinterpret_cast<const float4*>(b_ptr + k0 * N);
        As[a_kq + 0][a_row] = av.x; As[a_kq + 1][a_row] = av.y; As[a_kq + 2][a_row] = av.z; As[a_kq + 3][a_row] = av.w;
        *reinterpret_cast<float4*>(&Bs[b_k][b_n]) = bv;
        __syncthreads();
        #pragma unroll
        for (int k = 0; k < 8; ++k) {
            float a[8], b[8];
            *reinterpret_cast<float4*>(a) = *reinterpret_cast<const float4*>(&As[k][ty * 8]);
            *reinterpret_cast<float4*>(a + 4) = *reinterpret_cast<const float4*>(&As[k][ty * 8 + 4]);
            *reinterpret_cast<float4*>(b) = *reinterpret_cast<const float4*>(&Bs[k][tx * 8]);
            *reinterpret_cast<float4*>(b + 4) = *reinterpret_cast<const float4*>(&Bs[k][tx * 8 + 4]);
            #pragma unroll
            for (int i = 0; i < 8; ++i)
                #pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
        if ((k0 & 63) == 56 || k0 + 8 >= K) {
            #pragma unroll
            for (int i = 0; i < 8; ++i)
                #pragma unroll
                for (int j = 0; j < 8; ++j) { tot[i][j] += acc[i][j]; acc[i][j] = 0.f; }
        }
    }
    #pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int64_t r = m0 + ty * 8 + i;
        if (r >= m) continue;
        float* o = out + r * ld_out + n0 + tx * 8;
        #pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaxf(tot[i][j] + __ldg(bias + n0 + tx * 8 + j), 0.f);
    }
}

// One warp per window.  fp64 accumulation (the fp64 oracle is the arbiter; TF's own fp32 order is
// unspecified), logits rounded to fp32, softmax in fp32 as exp(x - max) / sum.
__global__ void __launch_bounds__(256) fc2_softmax_kernel(const float* __restrict__ bott, int ld, int width, int64_t n,
                                                          const int64_t* __restrict__ n_dev,
                                                          const float* __restrict__ W2, const float* __restrict__ b2,
                                                          float* __restrict__ logits, float* __restrict__ probs) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= eff_count(n, n_dev)) return;
    const float* row = bott + w * ld;
    double s0 = 0.0, s1 = 0.0;
    if ((width & 127) == 0 && (ld & 3) == 0 && (reinterpret_cast<uintptr_t>(bott) & 15) == 0) {
        // 16-byte row loads (512 B per warp and instruction), all loads of a row in flight before the fp64 FMAs
        const float4* row4 = reinterpret_cast<const float4*>(row);
        const float4* w4 = reinterpret_cast<const float4*>(W2);          // [k][2] floats: two k per float4
        #pragma unroll 4
        for (int k4 = lane; k4 < width / 4; k4 += 32) {
            const float4 v = row4[k4];
            const float4 wa = __ldg(w4 + 2 * k4), wb = __ldg(w4 + 2 * k4 + 1);
            s0 = fma((double)v.x, (double)wa.x, s0); s1 = fma((double)v.x, (double)wa.y, s1);
            s0 = fma((double)v.y, (double)wa.z, s0); s1 = fma((double)v.y, (double)wa.w, s1);
            s0 = fma((double)v.z, (double)wb.x, s0); s1 = fma((double)v.z, (double)wb.y, s1);
            s0 = fma((double)v.w, (double)wb.z, s0); s1 = fma((double)v.w, (double)wb.w, s1);
        }
    } else {
        for (int k = lane; k < width; k += 32) {
            const double v = (double)row[k];
            const float2 wv = __ldg(reinterpret_cast<const float2*>(W2) + k);
            s0 = fma(v, (double)wv.x, s0);
            s1 = fma(v, (double)wv.y, s1);
        }
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    }
    if (lane == 0) {
        const float l0 = (float)(s0 + (double)__ldg(b2)), l1 = (float)(s1 + (double)__ldg(b2 + 1));
        const float mx = fmaxf(l0, l1);
        const float e0 = expf(l0 - mx), e1 = expf(l1 - mx);
        const float inv = 1.f / (e0 + e1);
        if (logits) { logits[2 * w] = l0; logits[2 * w + 1] = l1; }
        probs[2 * w] = e0 * inv; probs[2 * w + 1] = e1 * inv;
    }
}

int launch_fc1_simt(const rodc_stage* s, const float* feat, int64_t m, float* bott, int ld_bott, cudaStream_t st) {
    if (m <= 0) return RODC_OK;
    RODC_CHECK_ARG(s->fc1_size % 128 == 0, "fc1_size must be a multiple of 128 (got %d)", s->fc1_size);
    RODC_CHECK_ARG(s->K % 8 == 0, "K must be a multiple of 8");
    dim3 grid(s->fc1_size / 128, (unsigned)((m + 127) / 128));
    Prof prof(KC_FC1_SIMT, st, s->in_size, (double)m * (double)s->K * 4.0 + (double)s->K * s->fc1_size * 4.0 + (double)m * s->fc1_size * 4.0,
              2.0 * (double)m * (double)s->K * (double)s->fc1_size);
    fc1_simt_kernel<<<grid, 256, 0, st>>>(feat, s->fc1_w, s->fc1_b, bott, ld_bott, m, s->K, s->fc1_size);
    RODC_LAUNCH_CHECK();
    return RODC_OK;
}

int launch_fc2_softmax(const rodc_stage* s, const float* bott, int ld_bott, int64_t n, float* logits, float* probs, cudaStream_t st,
                       const int64_t* n_dev) {
    if (n <= 0) return RODC_OK;
    const int width = s->fc1_size + s->n_prev;
    const int wpb = 8;
    // K2c: reads the [own | prev] bottleneck row, writes logits + softmax (4 * 512 (i + 1) B / window)
    Prof prof(KC_FC2, st, s->in_size, (double)n * (4.0 * width + 16.0), (double)n * 4.0 * width);
    fc2_softmax_kernel<<<(unsigned)((n + wpb - 1) / wpb), wpb * 32, 0, st>>>(bott, ld_bott, width, n, n_dev, s->fc2_w, s->fc2_b, logits, probs);
    RODC_LAUNCH_CHECK();
    return RODC_OK;
}

}  // namespace rodc
