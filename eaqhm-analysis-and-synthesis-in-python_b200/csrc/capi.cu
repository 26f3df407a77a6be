// Library identification and error text of the eaQHM B200 C ABI.
#include <stdarg.h>

#include "common.cuh"

static thread_local char g_err[512] = "";

void eaqhm_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" int eaqhm_version(void) { return EAQHM_VERSION; }
extern "C" const char* eaqhm_last_error(void) { return g_err; }
extern "C" int eaqhm_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ---- profiler: launch counts always, CUDA-event timing on request ---------------------------
#include <vector>
namespace {
struct Pending { int kc; cudaEvent_t e0, e1; };
bool g_prof_on = false;
long long g_counts[KC_COUNT] = {0};
double g_ms[KC_COUNT] = {0};
std::vector<Pending> g_pending;
std::vector<cudaEvent_t> g_pool;
cudaEvent_t g_cur = nullptr;
cudaEvent_t take_event() {
    if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
void drain() {
    for (auto& p : g_pending) {
        float ms = 0.f;
        if (cudaEventSynchronize(p.e1) == cudaSuccess && cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess)
            g_ms[p.kc] += ms;
        g_pool.push_back(p.e0); g_pool.push_back(p.e1);
    }
    g_pending.clear();
}
}  // namespace

void eaqhm_prof_begin(int kc, cudaStream_t st) {
    g_counts[kc]++;
    if (!g_prof_on) return;
    g_cur = take_event();
    cudaEventRecord(g_cur, st);
}
void eaqhm_prof_end(int kc, cudaStream_t st) {
    if (!g_prof_on) return;
    cudaEvent_t e1 = take_event();
    cudaEventRecord(e1, st);
    g_pending.push_back({kc, g_cur, e1});
}

static const char* kKernelNames[KC_COUNT] = {"fallback_flags", "active", "basis", "gram", "chol_solve", "accept",
                                             "run_moments", "a0_moments", "interp", "synth", "reduce", "var_pass",
                                             "snapshot"};

extern "C" int eaqhm_profile_enable(int on) { drain(); g_prof_on = on != 0; return EAQHM_OK; }
extern "C" int eaqhm_profile_reset(void) {
    drain();
    for (int i = 0; i < KC_COUNT; i++) { g_counts[i] = 0; g_ms[i] = 0.0; }
    return EAQHM_OK;
}
extern "C" int eaqhm_profile_classes(void) { return KC_COUNT; }
extern "C" const char* eaqhm_profile_name(int kc) { return (kc >= 0 && kc < KC_COUNT) ? kKernelNames[kc] : ""; }
extern "C" int eaqhm_profile_read(double* ms, int64_t* counts, int32_t n) {
    CHECK_ARG(ms && counts && n >= KC_COUNT);
    drain();
    for (int i = 0; i < KC_COUNT; i++) { ms[i] = g_ms[i]; counts[i] = g_counts[i]; }
    return EAQHM_OK;
}
