// Error reporting, ABI bookkeeping and the few host-side resources of libdqomap_b200.so.
//
// The library keeps no state that couples independent callers:
//   * the last-error string and the stage-profiling events are per calling thread;
//   * the helper ("side") stream and the four fork / join events a forward pass needs are owned by the pair
//     (device, caller's stream): two callers on different streams never share a side stream, so their forked work does
//     not serialise, and nothing is created or destroyed on the hot path after the first call on a stream;
//   * the launch counter is a relaxed atomic (instrumentation only).
#include "common.cuh"
#include <cstdlib>
#include <atomic>
#include <mutex>
#include <stdarg.h>
#include <stdio.h>
#include <unordered_map>
#include <nvtx3/nvToolsExt.h>

namespace dqo {
static thread_local char g_err[512] = "";
void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static std::atomic<long long> g_launches{0};
bool pdl_enabled() {
    static const bool on = [] {
        const char *e = getenv("DQO_PDL");
        return !(e && e[0] == '0');
    }();
    return on;
}
static thread_local bool g_pdl_now = false;
void pdl_scope(long long n_gaussians) { g_pdl_now = pdl_enabled() && n_gaussians >= DQO_PDL_MIN_GAUSSIANS; }
bool pdl_active() { return g_pdl_now; }
void note_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// NVTX ranges around the C-ABI entry points (visible in nsys / ncu --nvtx; a no-op without a tool attached)
void nvtx_push(const char *name) { nvtxRangePushA(name); }
void nvtx_pop() { nvtxRangePop(); }

struct FjKey {
    int dev;
    cudaStream_t stream;
    bool operator==(const FjKey &o) const { return dev == o.dev && stream == o.stream; }
};
struct FjHash {
    size_t operator()(const FjKey &k) const { return std::hash<void *>()((void *)k.stream) * 31u + (size_t)k.dev; }
};
static std::mutex g_fj_mu;
static std::unordered_map<FjKey, ForkJoin *, FjHash> g_fj;
#define DQO_MAX_FORK_JOIN 1024

ForkJoin *fork_join(cudaStream_t caller) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
    std::lock_guard<std::mutex> lock(g_fj_mu);
    const FjKey key{dev, caller};
    auto it = g_fj.find(key);
    if (it != g_fj.end()) return it->second;
    if (g_fj.size() >= DQO_MAX_FORK_JOIN) return nullptr; // callers fall back to their own stream (no overlap)
    ForkJoin *fj = new ForkJoin();
    // highest priority: the forked kernels are small and latency-bound, they should slip in between the blocks of the
    // large kernel they overlap instead of queueing behind its whole grid
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    bool ok = cudaStreamCreateWithPriority(&fj->side, cudaStreamNonBlocking, hi) == cudaSuccess;
    for (int i = 0; i < 8 && ok; i++) ok = cudaEventCreateWithFlags(&fj->ev[i], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) {
        cudaGetLastError();
        delete fj; // leaks at most a stream / a few events on a failing device; the caller runs unforked
        return nullptr;
    }
    g_fj[key] = fj;
    return fj;
}

static thread_local int g_profile = 0;
static thread_local cudaEvent_t g_ev[ST_COUNT];
static thread_local bool g_ev_made = false, g_ev_set[ST_COUNT];
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

extern "C" long long dqo_launch_count(void) { return dqo::g_launches.load(std::memory_order_relaxed); }
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
