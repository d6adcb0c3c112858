// Launch accounting + measurement hook of librodc.
//
// Every kernel launch of the library goes through a `Prof` scope (common.cuh): it bumps the process-wide
// launch counter (rodc_launch_count -- bench.py's "gpu_launches" is the difference of two reads, i.e. a
// measurement, not a hand-kept tally) and, when profiling is enabled, brackets the launch(es) with cudaEvents
// on the stream they are issued on, tagged with a kernel class, a sub key (window side) and the algorithmic
// bytes / FLOPs of SURVEY.md §8(d).  Level 1 records fc1 only (the dominant kernel: cheap enough to stay on
// inside the timed region), level 2 records every class.
#include "common.cuh"
#include <atomic>
#include <mutex>
#include <string.h>
#include <vector>

namespace rodc {

static std::atomic<int64_t> g_launches{0};
static std::atomic<int> g_level{0};

struct ProfRec { cudaEvent_t a, b; int kc, sub, launches; double bytes, flops; int area_slot; double area_scale; };
static std::mutex g_mu;
static std::vector<ProfRec> g_recs;

// device accumulators for data-dependent byte counts (K4: sum of the survivors' crop areas)
constexpr int AREA_SLOTS = 4096;
static unsigned long long* g_area_dev = nullptr;
static int g_area_next = 0;

struct Row { int kc, sub; double ms, bytes, flops; int64_t launches, records; };
static std::vector<Row> g_table;            // filled by the last collect

int profile_level() { return g_level.load(std::memory_order_relaxed); }

static int level_of(int kc) { return kc == KC_FC1 || kc == KC_FC1_SIMT ? 1 : 2; }

Prof::Prof(int kc, cudaStream_t st, int sub, double bytes, double flops, int launches) : st_(st), rec_(-1) {
    g_launches.fetch_add(launches, std::memory_order_relaxed);
    if (profile_level() < level_of(kc)) return;
    cudaEvent_t a, b;
    if (cudaEventCreate(&a) != cudaSuccess) return;
    if (cudaEventCreate(&b) != cudaSuccess) { cudaEventDestroy(a); return; }
    cudaEventRecord(a, st);
    std::lock_guard<std::mutex> lk(g_mu);
    g_recs.push_back({a, b, kc, sub, launches, bytes, flops, -1, 0.0});
    rec_ = (int)g_recs.size() - 1;
}

Prof::~Prof() {
    if (rec_ < 0) return;
    std::lock_guard<std::mutex> lk(g_mu);
    if (rec_ < (int)g_recs.size()) cudaEventRecord(g_recs[rec_].b, st_);
}

unsigned long long* Prof::area_slot(double bytes_per_unit) {
    if (rec_ < 0) return nullptr;
    std::lock_guard<std::mutex> lk(g_mu);
    if (!g_area_dev) {
        if (cudaMalloc(&g_area_dev, AREA_SLOTS * sizeof(unsigned long long)) != cudaSuccess) { g_area_dev = nullptr; return nullptr; }
        cudaMemset(g_area_dev, 0, AREA_SLOTS * sizeof(unsigned long long));
    }
    if (g_area_next >= AREA_SLOTS) return nullptr;
    const int slot = g_area_next++;
    g_recs[rec_].area_slot = slot;
    g_recs[rec_].area_scale = bytes_per_unit;
    return g_area_dev + slot;
}

static const char* const KC_NAMES[KC_COUNT] = {
    "K1 pyr_hpass", "K1 pyr_vpass", "window tables", "stage-0 window copy", "K4 recrop", "K2a features",
    "K2b fc1 (tcgen05)", "K2b fc1 (simt)", "K2c fc2+softmax", "K3 compact", "K5 group_rectangles", "row select / misc"};

}  // namespace rodc

using namespace rodc;

extern "C" int64_t rodc_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

extern "C" const char* rodc_kernel_class_name(int kc) { return kc >= 0 && kc < KC_COUNT ? KC_NAMES[kc] : ""; }

extern "C" int rodc_profile_enable(int level) {
    g_level.store(level < 0 ? 0 : level, std::memory_order_relaxed);
    return RODC_OK;
}

extern "C" int rodc_profile_collect(double* fc1_ms, double* fc1_flops, int64_t* fc1_launches,
                                    double* dom_ms, double* dom_flops, int64_t* dom_launches) {
    std::vector<ProfRec> recs;
    std::vector<unsigned long long> area;
    {
        std::lock_guard<std::mutex> lk(g_mu);
        recs.swap(g_recs);
        if (g_area_dev && g_area_next > 0) {
            area.resize(g_area_next);
            RODC_CUDA(cudaDeviceSynchronize());
            RODC_CUDA(cudaMemcpy(area.data(), g_area_dev, area.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
            RODC_CUDA(cudaMemset(g_area_dev, 0, area.size() * sizeof(unsigned long long)));
            g_area_next = 0;
        }
    }
    double ms = 0.0, fl = 0.0, dms = 0.0, dfl = 0.0, maxfl = 0.0;
    int64_t n1 = 0, dn = 0;
    std::vector<float> t(recs.size(), 0.f);
    g_table.clear();
    for (size_t i = 0; i < recs.size(); ++i) {
        ProfRec& r = recs[i];
        RODC_CUDA(cudaEventSynchronize(r.b));
        RODC_CUDA(cudaEventElapsedTime(&t[i], r.a, r.b));
        cudaEventDestroy(r.a); cudaEventDestroy(r.b);
        if (r.area_slot >= 0 && r.area_slot < (int)area.size()) r.bytes += r.area_scale * (double)area[r.area_slot];
        if (r.kc == KC_FC1 || r.kc == KC_FC1_SIMT) {
            ms += t[i]; fl += r.flops; ++n1;
            if (r.flops > maxfl) maxfl = r.flops;
        }
        Row* row = nullptr;
        for (auto& x : g_table) if (x.kc == r.kc && x.sub == r.sub) { row = &x; break; }
        if (!row) { g_table.push_back({r.kc, r.sub, 0.0, 0.0, 0.0, 0, 0}); row = &g_table.back(); }
        row->ms += t[i]; row->bytes += r.bytes; row->flops += r.flops; row->launches += r.launches; row->records += 1;
    }
    for (size_t i = 0; i < recs.size(); ++i)
        if ((recs[i].kc == KC_FC1 || recs[i].kc == KC_FC1_SIMT) && recs[i].flops == maxfl) { dms += t[i]; dfl += recs[i].flops; ++dn; }
    if (fc1_ms) *fc1_ms = ms;
    if (fc1_flops) *fc1_flops = fl;
    if (fc1_launches) *fc1_launches = n1;
    if (dom_ms) *dom_ms = dms;
    if (dom_flops) *dom_flops = dfl;
    if (dom_launches) *dom_launches = dn;
    return RODC_OK;
}

extern "C" int rodc_profile_table(int max_rows, int32_t* kclass, int32_t* sub, double* ms, double* bytes, double* flops,
                                  int64_t* launches, int64_t* records) {
    int n = 0;
    for (const Row& r : g_table) {
        if (n >= max_rows) break;
        if (kclass) kclass[n] = r.kc;
        if (sub) sub[n] = r.sub;
        if (ms) ms[n] = r.ms;
        if (bytes) bytes[n] = r.bytes;
        if (flops) flops[n] = r.flops;
        if (launches) launches[n] = r.launches;
        if (records) records[n] = r.records;
        ++n;
    }
    return n;
}
