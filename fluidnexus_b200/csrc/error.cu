// error.cu -- thread-local last-error string + version entry points.
#include <stdarg.h>

#include "common.cuh"

namespace fnx {
unsigned long long g_launches = 0;
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
}  // namespace fnx

extern "C" {
int fnx_abi_version(void) { return FNX_ABI_VERSION; }
const char *fnx_last_error(void) { return fnx::g_err; }
const char *fnx_build_arch(void) { return "sm_100a"; }
unsigned long long fnx_launch_count(void) { return fnx::g_launches; }
}

// ---- section profiler ------------------------------------------------------------------------------------------
#include <vector>
namespace fnx {
struct Prof {
    bool enabled = false;
    unsigned mask = 0;
    struct Pair { cudaEvent_t a, b; int sec; };
    std::vector<Pair> pairs;   // recorded this session
    std::vector<Pair> pool;    // reusable
    cudaEvent_t open_ev[SEC_COUNT];
    bool open[SEC_COUNT] = {false};
    Pair cur[SEC_COUNT];
};
static Prof g_prof;
void prof_begin(int s, cudaStream_t st) {
    if (!g_prof.enabled || !((g_prof.mask >> s) & 1u)) return;
    Prof::Pair p;
    if (!g_prof.pool.empty()) { p = g_prof.pool.back(); g_prof.pool.pop_back(); }
    else { cudaEventCreate(&p.a); cudaEventCreate(&p.b); }
    p.sec = s;
    cudaEventRecord(p.a, st);
    g_prof.cur[s] = p;
    g_prof.open[s] = true;
}
void prof_end(int s, cudaStream_t st) {
    if (!g_prof.enabled || !g_prof.open[s]) return;
    cudaEventRecord(g_prof.cur[s].b, st);
    g_prof.pairs.push_back(g_prof.cur[s]);
    g_prof.open[s] = false;
}
}  // namespace fnx

extern "C" {
static const char *kSectionNames[fnx::SEC_COUNT] = {"preprocess", "depth_sort", "emit", "tile_sort", "pack", "blend_fwd",
                                                    "blend_bwd", "geom_bwd", "image_loss", "physics"};
int fnx_profile_sections(void) { return fnx::SEC_COUNT; }
const char *fnx_profile_section_name(int32_t s) { return (s >= 0 && s < fnx::SEC_COUNT) ? kSectionNames[s] : ""; }
int fnx_profile_enable(uint32_t section_mask) {
    fnx::g_prof.enabled = section_mask != 0;
    fnx::g_prof.mask = section_mask;
    return FNX_OK;
}
// Blocks until the recorded sections finished; adds up their durations per section and clears the record.
int fnx_profile_collect(float *total_ms /*[sections]*/, int32_t *launches /*[sections]*/) {
    for (int s = 0; s < fnx::SEC_COUNT; s++) { if (total_ms) total_ms[s] = 0.f; if (launches) launches[s] = 0; }
    for (auto &p : fnx::g_prof.pairs) {
        cudaError_t e = cudaEventSynchronize(p.b);
        float ms = 0.f;
        if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, p.a, p.b);
        if (e != cudaSuccess) { fnx::set_error("profile collect failed: %s", cudaGetErrorString(e)); return FNX_ERR_CUDA; }
        if (total_ms) total_ms[p.sec] += ms;
        if (launches) launches[p.sec] += 1;
        fnx::g_prof.pool.push_back(p);
    }
    fnx::g_prof.pairs.clear();
    return FNX_OK;
}
}
