// Stage-net handle shared by the K2 kernels.
#pragma once
#include "common.cuh"

struct rodc_stage {
    int in_size = 0;        // S: input window side
    int channels = 0;       // conv output channels (32)
    int fc1_size = 0;       // 512
    int n_prev = 0;         // width of the previous bottleneck (512 * stage index)
    int64_t K = 0;          // S*S*channels
    int device = -1;
    // fp32 weights (reference layouts)
    float* conv_w = nullptr;    // [3,3,3,C] HWIO
    float* conv_b = nullptr;    // [C]
    float* fc1_w = nullptr;     // [K, fc1]      (SIMT path)
    float* fc1_b = nullptr;     // [fc1]
    float* fc2_w = nullptr;     // [fc1 + n_prev, 2]
    float* fc2_b = nullptr;     // [2]
    // split-fp16 fc1 weights for the tensor path: W * 2^w_exp = hi + lo, stored [fc1, K] (K-major)
    __half* fc1_hi = nullptr;
    __half* fc1_lo = nullptr;
    uint32_t* conv_frag = nullptr;  // mma.sync B fragments of the conv: [3 ky][4 nt][hi,lo][32 lanes][2] half2
    uint32_t* conv_frag_perm = nullptr;   // same, n-tile columns permuted so a thread owns 4 adjacent channels (v3)
    // K order of the SPLIT fc1 operand (A planes written by the feature kernels, B planes fc1_hi / fc1_lo): in_size when
    // the streaming feature kernel's store-friendly order is in use (see kperm below), 0 = the reference's NHWC order
    int kperm_side = 0;
    int c_exp = 0;              // conv weights are scaled by 2^c_exp before the fp16 split
    int a_exp = 0;              // features are scaled by 2^a_exp before the fp16 split
    int w_exp = 0;              // weights  are scaled by 2^w_exp before the fp16 split
    // general architecture (several conv layers / other strides, network/net.py:189-221): layer table
    int general = 0;
    int n_conv = 1, ksize = 3, conv_stride = 1, pool_size = 3, pool_stride = 1;
    int chan[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    float* gconv_w[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    float* gconv_b[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    int64_t max_act = 0;        // floats of the largest activation map (per ping-pong buffer)
    ~rodc_stage();
};

namespace rodc {

// K order of the split fc1 operand for the streaming feature kernels (S = 12 / 24 / 48).  fc1 is a sum over K, so the
// order is free as long as the A planes and the prepared weights agree.  The reference's NHWC flatten order
// (k = (y S + x) 32 + c) makes a warp of the feature kernel scatter every store instruction over eight 128-byte lines
// (a warp owns 16 of a pixel's 32 channels and its eight lane groups own pixels six apart): the stores alone were
// 70 % of that kernel's L1 wavefronts (profiles/r2a_ncu_kernels.json: l1tex 79 %, the top unit).  In the permuted order
// a row is stored as [channel half][x'][16 channels] with x' = (x mod NP) G + x / NP (NP pixels per thread, G lane
// groups per window), so the eight lane groups of one store instruction write 8 x 32 = 256 contiguous bytes.
__host__ __device__ inline int64_t kperm(int S, int64_t k) {
    if (S != 12 && S != 24 && S != 48) return k;
    const int NPX = S == 12 ? 2 : 6, G = S / NPX;
    const int c = (int)(k & 31);
    const int64_t px = k >> 5;
    const int x = (int)(px % S);
    const int64_t y = px / S;
    return (((y * 2 + (c >> 4)) * S) + (x % NPX) * G + x / NPX) * 16 + (c & 15);
}

// Resolved window source passed by value to the feature kernel.
struct SrcDev {
    int kind;
    const uint8_t* dense;
    const uint8_t* pyramid; int64_t pyramid_bytes; const LevelInfo* levels; const int32_t* win; int64_t n_windows;
    const uint8_t* frames_base; const int64_t* frame_off; const int32_t* frame_hw; const int32_t* meta;
    const uint8_t* gather_base; const int64_t* win_off; const int32_t* win_pitch;
};

int resolve_source(const rodc_window_source* src, int in_size, SrcDev* out);

// features.cu
// n_dev / n_off (optional): device-side count, effective n = clamp(*n_dev - n_off, 0, n) (common.cuh eff_count)
int launch_features(const rodc_stage* s, const SrcDev& src, int64_t first, int64_t n,
                    float* out_f32, __half* out_hi, __half* out_lo, cudaStream_t st, float* general_scratch = nullptr,
                    const int64_t* n_dev = nullptr, int64_t n_off = 0);
int launch_source_windows(const SrcDev& src, int64_t first, int64_t n, int size, uint8_t* out, cudaStream_t st,
                          const int64_t* n_dev = nullptr, int64_t n_off = 0);
constexpr int GENERAL_CTAS = 296;     // persistent CTAs of the general-architecture feature kernel
// fc.cu
int launch_fc1_simt(const rodc_stage* s, const float* feat, int64_t m, float* bott, int ld_bott, cudaStream_t st);
int launch_fc2_softmax(const rodc_stage* s, const float* bott, int ld_bott, int64_t n, float* logits, float* probs, cudaStream_t st,
                       const int64_t* n_dev = nullptr);
// fc1_tc.cu
int launch_fc1_tensor(const rodc_stage* s, const __half* a_hi, const __half* a_lo, int64_t m,
                      float* bott, int ld_bott, cudaStream_t st, const int64_t* m_dev = nullptr, int64_t m_off = 0);
int prepare_fc1_split(rodc_stage* s, cudaStream_t st);
// -DRODC_PROBES builds only (0 otherwise): RODC_PROBE_A_WRAP=<rows> makes the 48-px stage keep its whole A operand in
// `rows` rows (the feature kernel writes window w to row w mod rows, fc1 reads those rows for every M tile): the A planes
// stay L2 resident, results are WRONG -- the measured upper bound of what producer fusion / L2 blocking could return
int probe_a_wrap(int in_size);

}  // namespace rodc
