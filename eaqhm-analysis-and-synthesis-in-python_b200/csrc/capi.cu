// Library identification and error text of the eaQHM B200 C ABI.
#include <stdarg.h>
#include <stdlib.h>

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
#include <mutex>
#include <vector>
namespace {
std::mutex g_prof_mu;   // launches of different host threads share the accounting below
struct Pending { int kc; cudaEvent_t e0, e1; cudaStream_t st; };
bool g_prof_on = false;
long long g_counts[KC_COUNT] = {0};
double g_ms[KC_COUNT] = {0};
std::vector<Pending> g_pending;
std::vector<cudaEvent_t> g_pool;
cudaEvent_t take_event() {
    if (!g_pool.empty()) { cudaEvent_t e = g_pool.back(); g_pool.pop_back(); return e; }
    cudaEvent_t e; cudaEventCreate(&e); return e;
}
// experiment aid (tools only): EAQHM_TRACE=<file> appends one line per timed launch: class, stream, start and end in ms
// since the first launch of the drained batch (a timeline of the chunk pipeline's two streams)
void drain() {
    static const char* trace_path = getenv("EAQHM_TRACE");
    FILE* tf = (trace_path && !g_pending.empty()) ? fopen(trace_path, "a") : nullptr;
    for (auto& p : g_pending) {
        float ms = 0.f;
        if (cudaEventSynchronize(p.e1) == cudaSuccess && cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess)
            g_ms[p.kc] += ms;
        if (tf) {
            float t0 = 0.f;
            cudaEventElapsedTime(&t0, g_pending.front().e0, p.e0);
            fprintf(tf, "%d %p %.3f %.3f\n", p.kc, (void*)p.st, t0, t0 + ms);
        }
        g_pool.push_back(p.e0); g_pool.push_back(p.e1);
    }
    if (tf) { fprintf(tf, "#\n"); fclose(tf); }
    g_pending.clear();
}
}  // namespace

// begin/end bracket one launch on one host thread: the start event travels in a thread-local slot, the shared
// lists are touched under the mutex
static thread_local cudaEvent_t t_cur = nullptr;
void eaqhm_prof_begin(int kc, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    g_counts[kc]++;
    if (!g_prof_on) { t_cur = nullptr; return; }
    t_cur = take_event();
    cudaEventRecord(t_cur, st);
}
void eaqhm_prof_end(int kc, cudaStream_t st) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    if (!g_prof_on || t_cur == nullptr) return;
    cudaEvent_t e1 = take_event();
    cudaEventRecord(e1, st);
    g_pending.push_back({kc, t_cur, e1, st});
    t_cur = nullptr;
}

int eaqhm_sm_count() {
    static int nsm[kMaxDevices] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) return 148;
    if (nsm[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        nsm[dev] = n;
    }
    return nsm[dev];
}

static const char* kKernelNames[KC_COUNT] = {"fallback_flags", "active", "basis", "gram", "chol_solve", "accept",
                                             "run_moments", "a0_moments", "interp", "synth", "reduce", "var_pass",
                                             "snapshot", "chol_update", "swipe"};

extern "C" int eaqhm_profile_enable(int on) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    drain();
    g_prof_on = on != 0;
    return EAQHM_OK;
}
extern "C" int eaqhm_profile_reset(void) {
    std::lock_guard<std::mutex> lk(g_prof_mu);
    drain();
    for (int i = 0; i < KC_COUNT; i++) { g_counts[i] = 0; g_ms[i] = 0.0; }
    return EAQHM_OK;
}
extern "C" int eaqhm_profile_classes(void) { return KC_COUNT; }
extern "C" const char* eaqhm_profile_name(int kc) { return (kc >= 0 && kc < KC_COUNT) ? kKernelNames[kc] : ""; }
extern "C" int eaqhm_profile_read(double* ms, int64_t* counts, int32_t n) {
    CHECK_ARG(ms && counts && n >= KC_COUNT);
    std::lock_guard<std::mutex> lk(g_prof_mu);
    drain();
    for (int i = 0; i < KC_COUNT; i++) { ms[i] = g_ms[i]; counts[i] = g_counts[i]; }
    return EAQHM_OK;
}
