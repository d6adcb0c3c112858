// Host planner: pyramid scales, level sizes, sliding-window grid and Pillow coefficient
// tables.  Restates data/rectangles.py:338-425 (scale loop + grid), :248-262 (*_norm) and
// Pillow's precompute_coeffs / normalize_coeffs_8bpc (the arithmetic behind
// scipy.misc.imresize -> Image.resize(BILINEAR), data/image_info.py:240) in double precision,
// in the reference's operation order.
#include "common.cuh"
#include <math.h>
#include <string.h>
#include <new>

namespace rodc {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof(g_err), fmt, ap); va_end(ap);
}
const char* last_error() { return g_err; }

// Pillow Resample.c: bilinear_filter + precompute_coeffs (box = whole image) + 8bpc normalisation.
static void pillow_coeffs(int in_size, int out_size, std::vector<int32_t>& bounds, std::vector<int32_t>& coef, int& ksize) {
    const int PRECISION_BITS = 32 - 8 - 2;
    double scale = (double)in_size / (double)out_size;
    double filterscale = scale < 1.0 ? 1.0 : scale;
    double support = 1.0 * filterscale;
    ksize = (int)ceil(support) * 2 + 1;
    bounds.assign((size_t)out_size * 2, 0);
    coef.assign((size_t)out_size * ksize, 0);
    std::vector<double> k(ksize);
    double ss = 1.0 / filterscale;
    for (int xx = 0; xx < out_size; ++xx) {
        double center = (xx + 0.5) * scale;
        double ww = 0.0;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        for (int x = 0; x < xmax; ++x) {
            double a = (x + xmin - center + 0.5) * ss;
            if (a < 0.0) a = -a;
            double w = a < 1.0 ? 1.0 - a : 0.0;
            k[x] = w; ww += w;
        }
        for (int x = 0; x < xmax; ++x) {
            if (ww != 0.0) k[x] /= ww;
            double v = k[x] * (double)(1 << PRECISION_BITS);
            coef[(size_t)xx * ksize + x] = k[x] < 0 ? (int)(-0.5 + v) : (int)(0.5 + v);
        }
        bounds[(size_t)xx * 2] = xmin;
        bounds[(size_t)xx * 2 + 1] = xmax;
    }
}

// Places one level in the packed pyramid / scratch of a frame and appends its Pillow coefficient tables.
static void place_level(rodc_plan* p, LevelInfo& lv, int64_t& pyr, int64_t& scr) {
    lv.byte_off = pyr;
    pyr += (((int64_t)lv.w * lv.h * 3) + 15) / 16 * 16;
    lv.scratch_off = scr;
    if (!lv.identity) {
        scr += (((int64_t)lv.w * p->img_h * 3) + 15) / 16 * 16;
        std::vector<int32_t> b, c; int ks = 0;
        rodc::pillow_coeffs(p->img_w, lv.w, b, c, ks);
        lv.kx = ks;
        lv.hx_bounds = (int64_t)p->tables.size(); p->tables.insert(p->tables.end(), b.begin(), b.end());
        lv.hx_coef = (int64_t)p->tables.size();   p->tables.insert(p->tables.end(), c.begin(), c.end());
        // dp4a form of the same coefficients: taps grouped in aligned source words of 4 pixels, every
        // 22-bit coefficient split into three u8 limbs (exact integer arithmetic: sum = a0 + a1<<8 + a2<<16)
        for (int xx = 0; xx < lv.w; ++xx) {
            const int xmin = b[2 * xx], cnt = b[2 * xx + 1];
            const int word0 = xmin >> 2, nw = ((xmin + cnt + 3) >> 2) - word0;
            p->hitems.push_back(word0); p->hitems.push_back(nw);
            p->hitems.push_back((int32_t)(p->hcoef.size() / 4));
            p->hitems.push_back((int32_t)(lv.scratch_off + (int64_t)xx * 3));   // + y * w * 3 in the kernel
            p->hitems.push_back(lv.w);
            p->hitems.push_back(0); p->hitems.push_back(0); p->hitems.push_back(0);
            for (int wi = 0; wi < nw; ++wi) {
                uint32_t limb[3] = {0u, 0u, 0u};
                for (int q = 0; q < 4; ++q) {
                    const int x = (word0 + wi) * 4 + q, tap = x - xmin;
                    const uint32_t cf = (tap >= 0 && tap < cnt) ? (uint32_t)c[(size_t)xx * ks + tap] : 0u;
                    for (int l = 0; l < 3; ++l) limb[l] |= ((cf >> (8 * l)) & 0xffu) << (8 * q);
                }
                p->hcoef.push_back(limb[0]); p->hcoef.push_back(limb[1]); p->hcoef.push_back(limb[2]); p->hcoef.push_back(0u);
            }
        }
        rodc::pillow_coeffs(p->img_h, lv.h, b, c, ks);
        lv.ky = ks;
        lv.vy_bounds = (int64_t)p->tables.size(); p->tables.insert(p->tables.end(), b.begin(), b.end());
        lv.vy_coef = (int64_t)p->tables.size();   p->tables.insert(p->tables.end(), c.begin(), c.end());
    }
}

}  // namespace rodc

using rodc::LevelInfo;

extern "C" int rodc_version(void) { return RODC_VERSION; }
extern "C" const char* rodc_last_error(void) { return rodc::last_error(); }

extern "C" int rodc_device_info(int* sm_count, int* cc_major, int* cc_minor) {
    int dev = 0;
    RODC_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    RODC_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return RODC_OK;
}

extern "C" int rodc_plan_create(int img_w, int img_h, int win, double factor, double min_window_length,
                                rodc_plan** out) {
    RODC_CHECK_ARG(out != nullptr, "rodc_plan_create: out is NULL");
    RODC_CHECK_ARG(img_w > 0 && img_h > 0 && win > 0, "rodc_plan_create: bad sizes %dx%d win %d", img_w, img_h, win);
    RODC_CHECK_ARG(factor > 1.0, "rodc_plan_create: window_scale_factor must be > 1 (got %g)", factor);
    RODC_CHECK_ARG(min_window_length > 0.0, "rodc_plan_create: min_window_length must be > 0");
    rodc_plan* p = new (std::nothrow) rodc_plan();
    if (!p) { rodc::set_error("out of host memory"); return RODC_ERR_NOMEM; }
    p->img_w = img_w; p->img_h = img_h; p->win = win; p->factor = factor; p->min_window_length = min_window_length;

    double scale = 1.0;
    double iw = (double)img_w, ih = (double)img_h;              // ints in the reference until the first division
    const double max_windows_per_row = 1.0 / min_window_length; // rectangles.py:359
    const double max_img_length = max_windows_per_row * (double)win;
    int64_t n = 0, pyr = 0, scr = 0;
    while (true) {
        if (iw < max_img_length && ih < max_img_length) {       // :366
            LevelInfo lv; memset(&lv, 0, sizeof(lv));
            lv.scale = scale;
            int sx = (int)(0.4 * win), sy = (int)(0.4 * win);
            int tx = (int)(0.1 * iw), ty = (int)(0.1 * ih);
            lv.step_x = sx < tx ? sx : tx; if (lv.step_x < 1) lv.step_x = 1;   // :386
            lv.step_y = sy < ty ? sy : ty; if (lv.step_y < 1) lv.step_y = 1;   // :387
            lv.first_window = n;
            lv.identity = (scale == 1.0);
            if (lv.identity) { lv.w = img_w; lv.h = img_h; }
            else { lv.w = (int)((double)img_w * scale); lv.h = (int)((double)img_h * scale); }   // imresize size rule
            int level_index = (int)p->levels.size();
            int nx = 0, ny = 0;
            for (int xmin = 0; !((double)(xmin + win) >= iw); xmin += lv.step_x) {      // :395 float compare
                int col = 0;
                for (int ymin = 0; !((double)(ymin + win) >= ih); ymin += lv.step_y) {  // :403
                    p->win_level.push_back(level_index);
                    p->win_x.push_back(xmin); p->win_y.push_back(ymin);
                    p->norm.push_back((int32_t)((double)xmin / scale));                 // :248-262 int() truncation
                    p->norm.push_back((int32_t)((double)ymin / scale));
                    p->norm.push_back((int32_t)((double)(xmin + win) / scale));
                    p->norm.push_back((int32_t)((double)(ymin + win) / scale));
                    if (xmin + win > lv.w || ymin + win > lv.h) p->fits = 0;            // Window.raw ValueError
                    ++n; ++col;
                }
                ny = col; ++nx;
            }
            lv.nx = nx; lv.ny = ny;
            rodc::place_level(p, lv, pyr, scr);
            p->levels.push_back(lv);
        }
        scale /= factor; ih /= factor; iw /= factor;            // :418-420 repeated division
        if (ih < (double)win || iw < (double)win) break;        // :423
    }
    if (scr >= (1ll << 31)) {
        rodc::set_error("rodc_plan_create: horizontal-pass rows of one frame exceed 2 GiB");
        delete p; return RODC_ERR_ARG;
    }
    p->pyramid_bytes = pyr > 0 ? pyr : 16;
    p->scratch_bytes = scr > 0 ? scr : 16;
    *out = p;
    return RODC_OK;
}

// Pyramid of explicitly given scales, no window grid: the frames `image.raw_scaled(True, next_scale)` of the
// reference's resize-then-crop branch (app/inference_cascade_app.py:374-394 -> data/image_info.py:201-250:
// imresize(original, ratio) -> size (int(W*ratio), int(H*ratio)); ratio == 1.0 is the original itself).
extern "C" int rodc_plan_create_scales(int img_w, int img_h, int n_scales, const double* scales, rodc_plan** out) {
    RODC_CHECK_ARG(out != nullptr && (scales != nullptr || n_scales == 0), "rodc_plan_create_scales: NULL argument");
    RODC_CHECK_ARG(img_w > 0 && img_h > 0 && n_scales >= 0, "rodc_plan_create_scales: bad sizes %dx%d, %d scales", img_w, img_h, n_scales);
    rodc_plan* p = new (std::nothrow) rodc_plan();
    if (!p) { rodc::set_error("out of host memory"); return RODC_ERR_NOMEM; }
    p->img_w = img_w; p->img_h = img_h; p->win = 0; p->factor = 0.0; p->min_window_length = 0.0;
    int64_t pyr = 0, scr = 0;
    for (int i = 0; i < n_scales; ++i) {
        LevelInfo lv; memset(&lv, 0, sizeof(lv));
        lv.scale = scales[i];
        lv.identity = (scales[i] == 1.0);
        if (lv.identity) { lv.w = img_w; lv.h = img_h; }
        else { lv.w = (int)((double)img_w * scales[i]); lv.h = (int)((double)img_h * scales[i]); }
        if (!(scales[i] > 0.0) || lv.w < 1 || lv.h < 1 || lv.w > (1 << 16) || lv.h > (1 << 16)) {
            rodc::set_error("rodc_plan_create_scales: scale %g gives a %dx%d level", scales[i], lv.w, lv.h);
            delete p; return RODC_ERR_ARG;
        }
        rodc::place_level(p, lv, pyr, scr);
        if (scr >= (1ll << 31)) {
            rodc::set_error("rodc_plan_create_scales: horizontal-pass rows of one frame exceed 2 GiB; split the scale list");
            delete p; return RODC_ERR_ARG;
        }
        p->levels.push_back(lv);
    }
    p->pyramid_bytes = pyr > 0 ? pyr : 16;
    p->scratch_bytes = scr > 0 ? scr : 16;
    *out = p;
    return RODC_OK;
}

rodc_plan::~rodc_plan() {
    if (d_levels) cudaFree(d_levels);
    if (d_tables) cudaFree(d_tables);
    if (d_hitems) cudaFree(d_hitems);
    if (d_hcoef) cudaFree(d_hcoef);
    if (d_win) cudaFree(d_win);
    if (d_row_start) cudaFree(d_row_start);
}

int rodc_plan::ensure_device() {
    int dev = 0;
    RODC_CUDA(cudaGetDevice(&dev));
    if (device == dev) return RODC_OK;
    RODC_CHECK_ARG(device < 0, "rodc_plan used on device %d but uploaded to device %d", dev, device);
    size_t nl = levels.size() ? levels.size() : 1;
    RODC_CUDA(cudaMalloc(&d_levels, nl * sizeof(LevelInfo)));
    if (levels.size()) RODC_CUDA(cudaMemcpy(d_levels, levels.data(), levels.size() * sizeof(LevelInfo), cudaMemcpyHostToDevice));
    size_t nt = tables.size() ? tables.size() : 1;
    RODC_CUDA(cudaMalloc(&d_tables, nt * sizeof(int32_t)));
    if (tables.size()) RODC_CUDA(cudaMemcpy(d_tables, tables.data(), tables.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    n_hitems = (int)(hitems.size() / 8);
    RODC_CUDA(cudaMalloc(&d_hitems, (hitems.size() ? hitems.size() : 8) * sizeof(int32_t)));
    if (hitems.size()) RODC_CUDA(cudaMemcpy(d_hitems, hitems.data(), hitems.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    RODC_CUDA(cudaMalloc(&d_hcoef, (hcoef.size() ? hcoef.size() : 4) * sizeof(uint32_t)));
    if (hcoef.size()) RODC_CUDA(cudaMemcpy(d_hcoef, hcoef.data(), hcoef.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
    // packed window rows: [N,8] = level, x, y, level_w, norm[4]
    size_t N = win_level.size();
    std::vector<int32_t> rows((N ? N : 1) * 8, 0);
    for (size_t i = 0; i < N; ++i) {
        rows[i * 8 + 0] = win_level[i];
        rows[i * 8 + 1] = win_x[i];
        rows[i * 8 + 2] = win_y[i];
        rows[i * 8 + 3] = levels[win_level[i]].w;
        for (int k = 0; k < 4; ++k) rows[i * 8 + 4 + k] = norm[i * 4 + k];
    }
    RODC_CUDA(cudaMalloc(&d_win, rows.size() * sizeof(int32_t)));
    RODC_CUDA(cudaMemcpy(d_win, rows.data(), rows.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    std::vector<int32_t> rs(levels.size() + 1, 0);
    for (size_t l = 0; l < levels.size(); ++l) rs[l + 1] = rs[l] + levels[l].h;
    total_rows = rs.back();
    RODC_CUDA(cudaMalloc(&d_row_start, rs.size() * sizeof(int32_t)));
    RODC_CUDA(cudaMemcpy(d_row_start, rs.data(), rs.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
    device = dev;
    return RODC_OK;
}

extern "C" void rodc_plan_destroy(rodc_plan* p) { delete p; }
extern "C" int rodc_plan_n_levels(const rodc_plan* p) { return p ? (int)p->levels.size() : -1; }
extern "C" int64_t rodc_plan_n_windows(const rodc_plan* p) { return p ? (int64_t)p->win_level.size() : -1; }
extern "C" int rodc_plan_fits(const rodc_plan* p) { return p ? p->fits : 0; }
extern "C" int64_t rodc_plan_pyramid_bytes(const rodc_plan* p) { return p ? p->pyramid_bytes : -1; }
extern "C" int64_t rodc_plan_scratch_bytes(const rodc_plan* p) { return p ? p->scratch_bytes : -1; }

extern "C" int rodc_plan_levels(const rodc_plan* p, double* scale, int32_t* width, int32_t* height,
                                int32_t* step_x, int32_t* step_y, int32_t* nx, int32_t* ny,
                                int64_t* first_window, int64_t* byte_offset) {
    RODC_CHECK_ARG(p != nullptr, "rodc_plan_levels: plan is NULL");
    for (size_t i = 0; i < p->levels.size(); ++i) {
        const LevelInfo& l = p->levels[i];
        if (scale) scale[i] = l.scale;
        if (width) width[i] = l.w;
        if (height) height[i] = l.h;
        if (step_x) step_x[i] = l.step_x;
        if (step_y) step_y[i] = l.step_y;
        if (nx) nx[i] = l.nx;
        if (ny) ny[i] = l.ny;
        if (first_window) first_window[i] = l.first_window;
        if (byte_offset) byte_offset[i] = l.byte_off;
    }
    return RODC_OK;
}

extern "C" int rodc_plan_windows(const rodc_plan* p, int32_t* level, int32_t* x, int32_t* y, int32_t* norm) {
    RODC_CHECK_ARG(p != nullptr, "rodc_plan_windows: plan is NULL");
    size_t N = p->win_level.size();
    if (level) memcpy(level, p->win_level.data(), N * sizeof(int32_t));
    if (x) memcpy(x, p->win_x.data(), N * sizeof(int32_t));
    if (y) memcpy(y, p->win_y.data(), N * sizeof(int32_t));
    if (norm) memcpy(norm, p->norm.data(), N * 4 * sizeof(int32_t));
    return RODC_OK;
}
