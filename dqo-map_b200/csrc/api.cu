// Error reporting and ABI bookkeeping for libdqomap_b200.so.
#include "common.cuh"
#include <mutex>
#include <stdarg.h>
#include <stdio.h>

namespace dqo {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static long long g_launches = 0;
void note_launch(int n) { g_launches += n; }

// One non-blocking side stream per device for work that can overlap the caller's stream (fork / join with events).
static std::mutex g_side_mu;
static cudaStream_t g_side[64] = {};
cudaStream_t side_stream() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return nullptr;
    std::lock_guard<std::mutex> lock(g_side_mu);
    if (!g_side[dev]) {
        // highest priority: the forked kernels are small and latency-bound, they should slip in between the blocks of
        // the large kernel they overlap instead of queueing behind its whole grid
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);
        if (cudaStreamCreateWithPriority(&g_side[dev], cudaStreamNonBlocking, hi) != cudaSuccess) g_side[dev] = nullptr;
    }
    return g_side[dev];
}

static int g_profile = 0;
static cudaEvent_t g_ev[ST_COUNT];
static bool g_ev_made = false, g_ev_set[ST_COUNT];
void stage_mark(cudaStream_t stream, int stage) {
    if (!g_profile) return;
    if (!g_ev_made) {
        for (int i = 0; i < ST_COUNT; i++) cudaEventCreate(&g_ev[i]);
        g_ev_made = true;
    }
    cudaEventRecord(g_ev[stage], stream);
    g_ev_set[stage] = true;
}
} // namespace dqo

extern "C" long long dqo_launch_count(void) { return dqo::g_launches; }
extern "C" void dqo_profile_enable(int on) {
    dqo::g_profile = on;
    for (int i = 0; i < dqo::ST_COUNT; i++) dqo::g_ev_set[i] = false;
}
// ms_out[i] = time between stage mark i and the previous recorded mark (0 for the two BEGIN marks / missing marks);
// returns ST_COUNT
extern "C" int dqo_profile_read(float *ms_out, int n) {
    using namespace dqo;
    for (int i = 0; i < n; i++) ms_out[i] = 0.f;
    if (!g_ev_made) return ST_COUNT;
    int prev = g_ev_set[0] ? 0 : -1;
    for (int i = 1; i < ST_COUNT && i < n; i++) {
        if (!g_ev_set[i]) continue;
        if (i != ST_BEGIN_BWD && prev >= 0) {
            cudaEventSynchronize(g_ev[i]);
            float ms = 0.f;
            if (cudaEventElapsedTime(&ms, g_ev[prev], g_ev[i]) == cudaSuccess) ms_out[i] = ms;
        }
        prev = i;
    }
    return ST_COUNT;
}

extern "C" int dqo_abi_version(void) { return DQO_ABI_VERSION; }
extern "C" const char *dqo_last_error(void) { return dqo::g_err; }
