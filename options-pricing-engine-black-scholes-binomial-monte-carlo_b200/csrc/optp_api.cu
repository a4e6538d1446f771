// Context, error reporting and memory helpers of the C ABI (include/optp_b200.h).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include "optp_common.cuh"

namespace optp {

void comm_release();  // comm.cu

static Context g_ctx;
static thread_local char g_err[512] = "";

static std::recursive_mutex g_api_mutex;  // recursive: optp_init may call optp_shutdown

Context& ctx() { return g_ctx; }

ApiLock::ApiLock() { g_api_mutex.lock(); }
ApiLock::~ApiLock() { g_api_mutex.unlock(); }

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int cuda_fail(cudaError_t e, const char* what) {
    snprintf(g_err, sizeof(g_err), "CUDA error %d (%s) in %s", (int)e, cudaGetErrorString(e), what);
    cudaGetLastError();  // clear the sticky flag for non-fatal errors
    return -(int)e;
}

int arena_reserve(Arena& a, size_t bytes) {
    if (bytes <= a.cap) return 0;
    if (a.p) {
        OPTP_CUDA(cudaStreamSynchronize(g_ctx.stream));
        OPTP_CUDA(cudaFree(a.p));
        a.p = nullptr;
        a.cap = 0;
    }
    size_t want = bytes + bytes / 8 + 256;
    OPTP_CUDA(cudaMalloc(&a.p, want));
    a.cap = want;
    return 0;
}

int host_arena_reserve(Arena& a, size_t bytes) {
    if (bytes <= a.cap) return 0;
    if (a.p) {
        OPTP_CUDA(cudaStreamSynchronize(g_ctx.stream));
        OPTP_CUDA(cudaFreeHost(a.p));
        a.p = nullptr;
        a.cap = 0;
    }
    size_t want = 2 * bytes + 4096;
    OPTP_CUDA(cudaHostAlloc(&a.p, want, cudaHostAllocDefault));
    a.cap = want;
    return 0;
}

void reduction_scratch(double** partials, unsigned int** ticket) {
    const unsigned int slot = g_ctx.reduce_slot++ % (unsigned int)kReduceSlots;
    *partials = g_ctx.partials + (size_t)slot * kMaxBlocks * kMaxVals;
    *ticket = g_ctx.ticket + slot;
}

int require_ready() {
    if (!g_ctx.ready) return set_error(OPTP_E_NOINIT, "optp_init() has not been called");
    return 0;
}

static void arena_free(Arena& a) {
    if (a.p) cudaFree(a.p);
    a.p = nullptr;
    a.cap = 0;
}

}  // namespace optp

using namespace optp;

extern "C" {

int optp_abi_version(void) { return OPTP_ABI_VERSION; }

const char* optp_last_error(void) { return g_err; }

int optp_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

static int init_resources(Context& c, int device) {
    OPTP_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    OPTP_CUDA(cudaGetDeviceProperties(&prop, device));
    // the library carries sm_100a SASS only (arch-specific, no forward-compatible PTX): any other part --
    // sm_103, sm_12x included -- would pass init and fail at the first launch with "no kernel image"
    if (prop.major != 10 || prop.minor != 0)
        return set_error(OPTP_E_ARG, "device %d is sm_%d%d; this library is built for sm_100a only",
                         device, prop.major, prop.minor);
    // grids are sized as n_sm x (<= 16 resident CTAs) and must fit the block-partial buffer
    if (prop.multiProcessorCount * 16 > kMaxBlocks)
        return set_error(OPTP_E_TOO_LARGE, "device %d has %d SMs; the reduction scratch is sized for %d",
                         device, prop.multiProcessorCount, kMaxBlocks / 16);
    c.device = device;
    c.n_sm = prop.multiProcessorCount;
    OPTP_CUDA(cudaStreamCreateWithFlags(&c.stream, cudaStreamNonBlocking));
    OPTP_CUDA(cudaMalloc(&c.partials, sizeof(double) * (size_t)kReduceSlots * kMaxBlocks * kMaxVals));
    OPTP_CUDA(cudaMalloc(&c.ticket, sizeof(unsigned int) * kReduceSlots));
    OPTP_CUDA(cudaMemset(c.ticket, 0, sizeof(unsigned int) * kReduceSlots));
    OPTP_CUDA(cudaMalloc(&c.stats_dev, sizeof(double) * kMaxPayoffs * OPTP_NSTATS));
    OPTP_CUDA(cudaHostAlloc(&c.stats_host, sizeof(double) * kMaxPayoffs * OPTP_NSTATS,
                            cudaHostAllocDefault));
    return 0;
}

static void release_resources(Context& c) {
    arena_free(c.in0);
    arena_free(c.in1);
    arena_free(c.out0);
    arena_free(c.out1);
    for (int i = 0; i < kMaxTables; ++i) {
        arena_free(c.tables[i].f32);
        arena_free(c.tables[i].f4);
        arena_free(c.tables[i].f64);
    }
    arena_free(c.work);
    arena_free(c.cta_scratch);
    if (c.pin_in.p) cudaFreeHost(c.pin_in.p);
    if (c.pin_out.p) cudaFreeHost(c.pin_out.p);
    c.pin_in = Arena();
    c.pin_out = Arena();
    if (c.partials) cudaFree(c.partials);
    if (c.ticket) cudaFree(c.ticket);
    if (c.stats_dev) cudaFree(c.stats_dev);
    if (c.crr_op_count) cudaFree(c.crr_op_count);
    if (c.stats_host) cudaFreeHost(c.stats_host);
    if (c.stream) cudaStreamDestroy(c.stream);
    if (c.stream2) cudaStreamDestroy(c.stream2);
    if (c.stream3) cudaStreamDestroy(c.stream3);
    if (c.stream4) cudaStreamDestroy(c.stream4);
    for (int i = 0; i < 3; ++i) {
        if (c.book_stream[i]) cudaStreamDestroy(c.book_stream[i]);
        if (c.book_event[i]) cudaEventDestroy(c.book_event[i]);
    }
    for (int i = 0; i < c.n_events; ++i) cudaEventDestroy(c.events[i]);
    cudaGetLastError();
    c = Context();
}

int optp_init(int device) {
    ApiLock lock;
    Context& c = g_ctx;
    if (c.ready && c.device == device) return 0;
    if (c.ready) optp_shutdown();
    int n = 0;
    OPTP_CUDA(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n)
        return set_error(OPTP_E_ARG, "device %d out of range (%d visible)", device, n);
    if (int e = init_resources(c, device)) {
        release_resources(c);  // a half-built context is torn down; the error text stays
        return e;
    }
    c.launches = 0;
    {
        const char* nb = getenv("OPTP_CRR_NO_BAND");
        c.crr_no_band = nb && nb[0] == '1';
        const char* bo = getenv("OPTP_CRR_BAND_ONLY");
        c.crr_band_only = bo && bo[0] == '1';
    }
    c.ready = true;
    return 0;
}

int optp_shutdown(void) {
    ApiLock lock;
    Context& c = g_ctx;
    if (!c.ready) return 0;
    cudaSetDevice(c.device);
    cudaStreamSynchronize(c.stream);
    if (c.stream2) cudaStreamSynchronize(c.stream2);
    if (c.stream3) cudaStreamSynchronize(c.stream3);
    if (c.stream4) cudaStreamSynchronize(c.stream4);
    for (int i = 0; i < 3; ++i)
        if (c.book_stream[i]) cudaStreamSynchronize(c.book_stream[i]);
    comm_release();
    release_resources(c);
    return 0;
}

uint64_t optp_launch_count(void) { return g_ctx.launches; }

int optp_set_option(const char* key, long value) {
    ApiLock lock;
    if (!key) return set_error(OPTP_E_ARG, "null option key");
    if (!strcmp(key, "crr_band")) {
        g_ctx.crr_no_band = (value == 0);
        return 0;
    }
    if (!strcmp(key, "crr_band_kernel")) {
        if (value != 1 && value != 2) return set_error(OPTP_E_ARG, "crr_band_kernel must be 1 or 2");
        g_ctx.crr_band_kernel = (int)value;
        return 0;
    }
    if (!strcmp(key, "crr_zone_group")) {
        if (value != 0 && value != 1 && value != 2 && value != 4)
            return set_error(OPTP_E_ARG, "crr_zone_group must be 0, 1, 2 or 4");
        g_ctx.crr_zone_group = (int)value;
        return 0;
    }
    if (!strcmp(key, "lv_direct")) {
        g_ctx.lv_direct = (value != 0);
        return 0;
    }
    if (!strcmp(key, "lv_window")) {
        g_ctx.lv_no_window = (value == 0);
        return 0;
    }
    if (!strcmp(key, "crr_euro_sum")) {
        g_ctx.crr_no_euro_sum = (value == 0);
        return 0;
    }
    if (!strcmp(key, "crr_count_ops")) {  // 1 = count (and reset the counters), 0 = production kernels
        g_ctx.crr_count_ops = (value != 0);
        if (g_ctx.crr_op_count) {
            OPTP_CUDA(cudaSetDevice(g_ctx.device));
            OPTP_CUDA(cudaDeviceSynchronize());
            OPTP_CUDA(cudaMemset(g_ctx.crr_op_count, 0, 2 * sizeof(unsigned long long)));
        }
        return 0;
    }
    if (!strcmp(key, "rng_mode")) {
        if (value < 0 || value >= RNG_MODES) return set_error(OPTP_E_ARG, "rng_mode must be 0, 1 or 2");
        g_ctx.rng_mode = (int)value;
        return 0;
    }
    if (!strcmp(key, "crr_band_only")) {
        g_ctx.crr_band_only = (value != 0);
        return 0;
    }
    return set_error(OPTP_E_ARG, "unknown option '%s'", key);
}

int optp_get_counter(const char* key, uint64_t* value_out) {
    ApiLock lock;
    if (!key || !value_out) return set_error(OPTP_E_ARG, "null pointer argument");
    const bool fp64 = !strcmp(key, "crr_fp64_lane_ops"), nodes = !strcmp(key, "crr_node_updates");
    if (!fp64 && !nodes) return set_error(OPTP_E_ARG, "unknown counter '%s'", key);
    *value_out = 0;
    if (!g_ctx.ready || !g_ctx.crr_op_count) return 0;
    OPTP_CUDA(cudaSetDevice(g_ctx.device));
    OPTP_CUDA(cudaDeviceSynchronize());
    unsigned long long v[2] = {0, 0};
    OPTP_CUDA(cudaMemcpy(v, g_ctx.crr_op_count, sizeof(v), cudaMemcpyDeviceToHost));
    *value_out = fp64 ? v[0] : v[1];
    return 0;
}

int optp_host_alloc(void** p, uint64_t bytes) {
    if (!p) return set_error(OPTP_E_ARG, "null pointer");
    OPTP_CUDA(cudaHostAlloc(p, bytes ? bytes : 1, cudaHostAllocDefault));
    return 0;
}

int optp_host_free(void* p) {
    if (p) OPTP_CUDA(cudaFreeHost(p));
    return 0;
}

}  // extern "C"
