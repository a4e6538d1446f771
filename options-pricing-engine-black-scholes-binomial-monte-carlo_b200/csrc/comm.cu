// Multi-GPU reduction of Monte Carlo statistics blocks behind the C ABI (include/optp_b200.h).
// The reference's only exchange step is the sum of per-worker sufficient statistics
// (monte_carlo.py:67-74 over the futures of :126-139); across GPUs that is ONE all-reduce of
// n_items x OPTP_NSTATS doubles over NVLink.  NCCL is resolved at run time (dlopen), so the library
// neither links against it nor needs it on a single-GPU box.
#include <dlfcn.h>
#include <string.h>
#include "optp_common.cuh"

namespace optp {

// the handful of NCCL declarations used here (nccl.h: stable since 2.0)
typedef struct ncclComm* ncclComm_t;
typedef struct {
    char internal[OPTP_COMM_ID_BYTES];
} ncclUniqueId;
typedef int ncclResult_t;
constexpr int kNcclSum = 0, kNcclFloat64 = 8;

struct Nccl {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
    int (*GetVersion)(int*) = nullptr;
};

static Nccl g_nccl;
static ncclComm_t g_comm = nullptr;
static int g_rank = 0, g_world = 1;

static int nccl_load() {
    if (g_nccl.handle) return 0;
    // the copy already mapped into the process (torch ships its own) wins; then the system one
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return set_error(OPTP_E_NOINIT, "libnccl.so.2 not found: %s", dlerror());
    Nccl n;
    n.handle = h;
    *(void**)&n.GetUniqueId = dlsym(h, "ncclGetUniqueId");
    *(void**)&n.CommInitRank = dlsym(h, "ncclCommInitRank");
    *(void**)&n.CommDestroy = dlsym(h, "ncclCommDestroy");
    *(void**)&n.AllReduce = dlsym(h, "ncclAllReduce");
    *(void**)&n.GetErrorString = dlsym(h, "ncclGetErrorString");
    *(void**)&n.GetVersion = dlsym(h, "ncclGetVersion");
    if (!n.GetUniqueId || !n.CommInitRank || !n.CommDestroy || !n.AllReduce || !n.GetErrorString)
        return set_error(OPTP_E_NOINIT, "libnccl.so.2 lacks a required symbol");
    g_nccl = n;
    return 0;
}

static int nccl_fail(ncclResult_t r, const char* what) {
    return set_error(-1000 - (int)r, "NCCL error %d (%s) in %s", (int)r, g_nccl.GetErrorString(r), what);
}

void comm_release() {  // called by optp_shutdown
    if (g_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(g_comm);
    g_comm = nullptr;
    g_rank = 0;
    g_world = 1;
}

}  // namespace optp

using namespace optp;

extern "C" {

int optp_comm_unique_id(char id_out[OPTP_COMM_ID_BYTES]) {
    ApiLock lock;
    if (!id_out) return set_error(OPTP_E_ARG, "null pointer argument");
    if (int e = nccl_load()) return e;
    ncclUniqueId id;
    const ncclResult_t r = g_nccl.GetUniqueId(&id);
    if (r != 0) return nccl_fail(r, "ncclGetUniqueId");
    memcpy(id_out, id.internal, OPTP_COMM_ID_BYTES);
    return 0;
}

int optp_comm_init(int rank, int world, const char id[OPTP_COMM_ID_BYTES]) {
    OPTP_ENTER();
    if (world < 1 || rank < 0 || rank >= world) return set_error(OPTP_E_ARG, "rank %d outside a world of %d", rank, world);
    comm_release();
    if (world == 1) return 0;
    if (!id) return set_error(OPTP_E_ARG, "null pointer argument");
    if (int e = nccl_load()) return e;
    OPTP_CUDA(cudaSetDevice(ctx().device));
    ncclUniqueId uid;
    memcpy(uid.internal, id, OPTP_COMM_ID_BYTES);
    const ncclResult_t r = g_nccl.CommInitRank(&g_comm, world, uid, rank);
    if (r != 0) {
        g_comm = nullptr;
        return nccl_fail(r, "ncclCommInitRank");
    }
    g_rank = rank;
    g_world = world;
    return 0;
}

int optp_comm_destroy(void) {
    ApiLock lock;
    comm_release();
    return 0;
}

int optp_comm_info(int* rank_out, int* world_out) {
    ApiLock lock;
    if (rank_out) *rank_out = g_rank;
    if (world_out) *world_out = g_world;
    return 0;
}

int optp_allreduce_stats_dev(double* stats_dev, uint64_t n_doubles, void* stream) {
    OPTP_ENTER();
    if (n_doubles == 0 || g_world == 1) return 0;
    if (!stats_dev) return set_error(OPTP_E_ARG, "null pointer argument");
    if (!g_comm) return set_error(OPTP_E_NOINIT, "optp_comm_init has not been called");
    const ncclResult_t r = g_nccl.AllReduce(stats_dev, stats_dev, (size_t)n_doubles, kNcclFloat64, kNcclSum, g_comm,
                                            (cudaStream_t)stream);
    if (r != 0) return nccl_fail(r, "ncclAllReduce");
    return 0;
}

int optp_allreduce_stats(double* stats, uint64_t n_doubles) {
    OPTP_ENTER();
    if (n_doubles == 0 || g_world == 1) return 0;
    if (!stats) return set_error(OPTP_E_ARG, "null pointer argument");
    Context& c = ctx();
    OPTP_CUDA(cudaSetDevice(c.device));
    const size_t bytes = (size_t)n_doubles * sizeof(double);
    if (int e = arena_reserve(c.in1, bytes)) return e;
    if (int e = host_arena_reserve(c.pin_in, bytes)) return e;
    memcpy(c.pin_in.p, stats, bytes);
    OPTP_CUDA(cudaMemcpyAsync(c.in1.p, c.pin_in.p, bytes, cudaMemcpyHostToDevice, c.stream));
    if (int e = optp_allreduce_stats_dev((double*)c.in1.p, n_doubles, c.stream)) return e;
    OPTP_CUDA(cudaMemcpyAsync(c.pin_in.p, c.in1.p, bytes, cudaMemcpyDeviceToHost, c.stream));
    OPTP_CUDA(cudaStreamSynchronize(c.stream));
    memcpy(stats, c.pin_in.p, bytes);
    return 0;
}

}  // extern "C"
