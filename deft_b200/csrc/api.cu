// Library-level entry points: version, error string, device check.
#include "common.cuh"
#include <string.h>

namespace deft {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}
static unsigned long long g_launches = 0;
void count_launch() { __atomic_fetch_add(&g_launches, 1ull, __ATOMIC_RELAXED); }
// Programmatic dependent launch: -1 = not decided yet (environment DEFT_PDL, default on), 0 / 1 = set.
static int g_pdl = -1;
bool pdl_enabled() {
    int v = __atomic_load_n(&g_pdl, __ATOMIC_RELAXED);
    if (v < 0) {
        const char* e = getenv("DEFT_PDL");
        v = e ? (atoi(e) != 0) : 1;
        __atomic_store_n(&g_pdl, v, __ATOMIC_RELAXED);
    }
    return v != 0;
}
static long long* g_trace = nullptr;
long long* debug_trace_ptr() { return __atomic_load_n(&g_trace, __ATOMIC_RELAXED); }
static unsigned long long g_fallbacks = 0;
static unsigned long long g_dcn_patch = 0;
void note_dcn_patch_launch() { __atomic_fetch_add(&g_dcn_patch, 1ull, __ATOMIC_RELAXED); }
static int g_dcn_patch_mode = -1;
int dcn_patch_mode() {
    int v = __atomic_load_n(&g_dcn_patch_mode, __ATOMIC_RELAXED);
    if (v < 0) {
        const char* e = getenv("DEFT_DCN_PATCH");
        v = e ? atoi(e) : 0;
        if (v < 0 || v > 2) v = 0;
        __atomic_store_n(&g_dcn_patch_mode, v, __ATOMIC_RELAXED);
    }
    return v;
}
static int g_engine_log = -1;
void note_simt_fallback(const char* op, const char* fmt, ...) {
    __atomic_fetch_add(&g_fallbacks, 1ull, __ATOMIC_RELAXED);
    int v = __atomic_load_n(&g_engine_log, __ATOMIC_RELAXED);
    if (v < 0) {
        const char* e = getenv("DEFT_DEBUG_ENGINE");
        v = e ? (atoi(e) != 0) : 0;
        __atomic_store_n(&g_engine_log, v, __ATOMIC_RELAXED);
    }
    if (!v) return;
    char buf[256];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    fprintf(stderr, "deft_b200: %s served by the fp32 SIMT engine (tcgen05 dispatcher declined: %s)\n", op, buf);
}
}  // namespace deft

extern "C" {
void deft_debug_set_trace(void* device_buffer) { __atomic_store_n(&deft::g_trace, reinterpret_cast<long long*>(device_buffer), __ATOMIC_RELAXED); }
void deft_debug_set_engine_log(int enabled) { __atomic_store_n(&deft::g_engine_log, enabled ? 1 : 0, __ATOMIC_RELAXED); }
void deft_debug_set_dcn_patch(int mode) { __atomic_store_n(&deft::g_dcn_patch_mode, (mode < 0 || mode > 2) ? 0 : mode, __ATOMIC_RELAXED); }
unsigned long long deft_dcn_patch_launch_count(void) { return __atomic_load_n(&deft::g_dcn_patch, __ATOMIC_RELAXED); }
unsigned long long deft_simt_fallback_count(void) { return __atomic_load_n(&deft::g_fallbacks, __ATOMIC_RELAXED); }
unsigned long long deft_launch_count(void) { return __atomic_load_n(&deft::g_launches, __ATOMIC_RELAXED); }
void deft_set_pdl(int enabled) {
    if (getenv("DEFT_PDL") != nullptr) return;           // an explicit environment setting wins
    __atomic_store_n(&deft::g_pdl, enabled ? 1 : 0, __ATOMIC_RELAXED);
}
int deft_get_pdl(void) { return deft::pdl_enabled() ? 1 : 0; }
int deft_version(void) { return 100; }
const char* deft_last_error(void) { return deft::g_err; }
int deft_device_ok(void) {
    int dev = 0;
    cudaDeviceProp p;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) {
        deft::set_error("no CUDA device");
        return 0;
    }
    if (p.major != 10) {
        deft::set_error("deft_b200 is built for sm_100a only; found sm_%d%d", p.major, p.minor);
        return 0;
    }
    return 1;
}
}
