// Shared helpers for librodc (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdarg.h>
#include <vector>
#include "../../include/rodc.h"

namespace rodc {

void set_error(const char* fmt, ...);

#define RODC_CHECK_ARG(cond, ...)                                   \
    do { if (!(cond)) { rodc::set_error(__VA_ARGS__); return RODC_ERR_ARG; } } while (0)

#define RODC_CUDA(call)                                             \
    do { cudaError_t e__ = (call);                                  \
         if (e__ != cudaSuccess) {                                  \
             rodc::set_error("%s failed at %s:%d: %s", #call, __FILE__, __LINE__, cudaGetErrorString(e__)); \
             return RODC_ERR_CUDA; } } while (0)

#define RODC_LAUNCH_CHECK()  RODC_CUDA(cudaGetLastError())

static inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Device-side element counts.  A launch sized for an upper bound n reads the real count from device memory
// (written by rodc_compact of the previous stage), so a whole cascade can be enqueued -- or captured in a CUDA
// graph -- without the host ever learning a survivor count: effective n = clamp(*n_dev - off, 0, n).
#ifdef __CUDACC__
__device__ __forceinline__ int64_t eff_count(int64_t n, const int64_t* n_dev, int64_t off = 0) {
    if (!n_dev) return n;
    const int64_t v = *n_dev - off;
    return v < 0 ? 0 : (v < n ? v : n);
}
#endif

// Kernel classes of the launch accounting / measurement hook (prof.cu, rodc_profile_table).
enum KernelClass { KC_PYR_H = 0, KC_PYR_V, KC_WINTAB, KC_SOURCE, KC_RECROP, KC_FEAT, KC_FC1, KC_FC1_SIMT, KC_FC2,
                   KC_COMPACT, KC_GROUP, KC_MISC, KC_COUNT };

// Scope around the launch(es) of one kernel class: counts them (rodc_launch_count) and, when profiling is on,
// brackets them with cudaEvents on `st`.  `bytes` / `flops` are the algorithmic figures of SURVEY.md §8(d).
int profile_level();
class Prof {
public:
    Prof(int kc, cudaStream_t st, int sub = 0, double bytes = 0.0, double flops = 0.0, int launches = 1);
    ~Prof();
    // device accumulator (one u64) for a data-dependent count (K4: pixels of the crops, K3: survivors), NULL unless
    // this scope is being recorded; bytes_per_unit x the accumulated count is added to the scope's bytes at collect time
    unsigned long long* area_slot(double bytes_per_unit = 3.0);
    bool recording() const { return rec_ >= 0; }
private:
    cudaStream_t st_;
    int rec_;
};

struct LevelInfo {            // one pyramid level of a plan (host + device copies)
    double  scale;
    int32_t w, h;             // level size
    int32_t step_x, step_y, nx, ny;
    int64_t first_window;
    int64_t byte_off;         // offset of the level in a frame's packed pyramid
    int64_t scratch_off;      // offset of the level's horizontal-pass rows in a frame's scratch
    int32_t identity;         // scale == 1.0: the level is the frame itself
    // coefficient tables (offsets into the plan's int32 table buffer)
    int32_t kx, ky;           // ksize of the horizontal / vertical pass
    int64_t hx_bounds, hx_coef;   // [w][2] and [w][kx]
    int64_t vy_bounds, vy_coef;   // [h][2] and [h][ky]
};

}  // namespace rodc

struct rodc_plan {
    int img_w, img_h, win;
    double factor, min_window_length;
    std::vector<rodc::LevelInfo> levels;
    std::vector<int32_t> win_level, win_x, win_y, norm;       // norm: [N*4]
    std::vector<int32_t> tables;                               // bounds + coefficients of all levels
    // horizontal pass, dp4a form: one item per (level, output x): [word0, n_words, coef word, out byte, level w, 0,0,0]
    // and per source word 4 packed coefficient limbs (c0 | c1 | c2 | pad), see plan.cu
    std::vector<int32_t> hitems;
    std::vector<uint32_t> hcoef;
    int64_t pyramid_bytes = 0, scratch_bytes = 0;
    int fits = 1;
    // device copies (lazily uploaded on first GPU use; device of first use)
    int device = -1;
    rodc::LevelInfo* d_levels = nullptr;
    int32_t* d_tables = nullptr;
    int32_t* d_hitems = nullptr;
    uint32_t* d_hcoef = nullptr;
    int n_hitems = 0;
    int32_t* d_win = nullptr;          // [N, 8]: level, x, y, level_w, norm[4] -- see plan.cu
    int32_t* d_row_start = nullptr;    // [n_levels+1] prefix of level heights (vertical-pass grid)
    int total_rows = 0;
    int ensure_device();
    ~rodc_plan();
};
