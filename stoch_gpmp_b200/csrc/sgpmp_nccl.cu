// Split-particle mode over NCCL, driven from C (SURVEY §8b `sgpmp_allreduce_stats(ncclComm, ...)`, §8e).
//
// One problem's S samples are divided over the ranks of a communicator.  Per iteration every rank runs ONE fused
// launch over its own samples (sample -> cost -> local softmax statistics, sgpmp_iterate_stats: nothing is
// materialised, the weighted eps-sum is regenerated from the counter-based stream), the (m, Z, A) blocks are
// exchanged with ONE ncclAllGather issued on the SAME stream, and ONE launch merges them by log-sum-exp in a fixed
// rank order and applies mu += step L (A/Z) — identically on every rank.  All n_iters iterations are enqueued by
// one call: no host arithmetic, no host synchronisation, no torch kernels in between (the reference arithmetic being
// split is planner.py:263-275).
//
// NCCL is resolved at run time with dlopen (the library ships inside the torch wheel and is already mapped into a
// process that imported torch); the few declarations needed are restated here so that libsgpmp.so has no link-time
// dependency on it.
#include <dlfcn.h>

#include "sgpmp_common.cuh"

namespace sgpmp {

typedef struct { char internal[128]; } nccl_unique_id_t;     // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef void* nccl_comm_t;
enum { NCCL_FLOAT32 = 7, NCCL_FLOAT64 = 8 };                 // ncclDataType_t

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(nccl_unique_id_t*) = nullptr;
    int (*CommInitRank)(nccl_comm_t*, int, nccl_unique_id_t, int) = nullptr;
    int (*CommDestroy)(nccl_comm_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, nccl_comm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load(const char* path) {
    if (g_nccl.handle) return SGPMP_OK;
    void* h = nullptr;
    if (path && *path) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);       // already mapped by torch
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { set_error("sgpmp_nccl: cannot load libnccl (%s)", dlerror()); return SGPMP_ERR_UNSUPPORTED; }
    NcclApi a;
    a.handle = h;
    a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    a.CommInitRank = (decltype(a.CommInitRank))dlsym(h, "ncclCommInitRank");
    a.CommDestroy = (decltype(a.CommDestroy))dlsym(h, "ncclCommDestroy");
    a.AllGather = (decltype(a.AllGather))dlsym(h, "ncclAllGather");
    a.GetErrorString = (decltype(a.GetErrorString))dlsym(h, "ncclGetErrorString");
    if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllGather) {
        set_error("sgpmp_nccl: libnccl lacks a required symbol");
        return SGPMP_ERR_UNSUPPORTED;
    }
    g_nccl = a;
    return SGPMP_OK;
}

#define SGPMP_NCCL_CHECK(call, what)                                                                              \
    do {                                                                                                          \
        int r_ = (call);                                                                                          \
        if (r_ != 0) {                                                                                            \
            set_error("%s: NCCL error %d (%s)", what, r_, g_nccl.GetErrorString ? g_nccl.GetErrorString(r_) : "?"); \
            return SGPMP_ERR_CUDA;                                                                                \
        }                                                                                                         \
    } while (0)

}  // namespace sgpmp

using namespace sgpmp;

extern "C" int sgpmp_nccl_load(const char* path) { return nccl_load(path); }

extern "C" int sgpmp_comm_unique_id(void* id128) {
    SGPMP_REQUIRE(id128, "sgpmp_comm_unique_id: null pointer");
    int rc = nccl_load(nullptr);
    if (rc != SGPMP_OK) return rc;
    SGPMP_NCCL_CHECK(g_nccl.GetUniqueId((nccl_unique_id_t*)id128), "sgpmp_comm_unique_id");
    return SGPMP_OK;
}

extern "C" int sgpmp_comm_init(const void* id128, int32_t rank, int32_t n_ranks, void** comm) {
    SGPMP_REQUIRE(id128 && comm && n_ranks >= 1 && rank >= 0 && rank < n_ranks, "sgpmp_comm_init: invalid argument");
    int rc = nccl_load(nullptr);
    if (rc != SGPMP_OK) return rc;
    nccl_unique_id_t id;
    memcpy(&id, id128, sizeof(id));
    nccl_comm_t c = nullptr;
    SGPMP_NCCL_CHECK(g_nccl.CommInitRank(&c, n_ranks, id, rank), "sgpmp_comm_init");
    *comm = c;
    return SGPMP_OK;
}

extern "C" int sgpmp_comm_destroy(void* comm) {
    if (!comm || !g_nccl.handle) return SGPMP_OK;
    SGPMP_NCCL_CHECK(g_nccl.CommDestroy((nccl_comm_t)comm), "sgpmp_comm_destroy");
    return SGPMP_OK;
}

// all_gather of the per-rank statistics blocks [B*NP][M+2] -> stats_all [n_ranks][B*NP][M+2] (the exchange step; the
// reduction itself — the log-sum-exp merge — happens inside sgpmp_merge_apply_stats, in a fixed rank order)
extern "C" int sgpmp_allreduce_stats(void* comm, const sgpmp_shape_t* shape, const void* stats_local, void* stats_all, void* stream) {
    SGPMP_REQUIRE(shape_ok(shape), "sgpmp_allreduce_stats: invalid shape");
    SGPMP_REQUIRE(comm && stats_local && stats_all, "sgpmp_allreduce_stats: null pointer");
    SGPMP_REQUIRE(g_nccl.handle, "sgpmp_allreduce_stats: NCCL is not loaded (sgpmp_comm_init first)");
    const size_t count = (size_t)shape->B * shape->G * shape->K * ((size_t)shape->T * 2 * shape->n_dof + 2);
    SGPMP_NCCL_CHECK(g_nccl.AllGather(stats_local, stats_all, count, shape->dtype == SGPMP_F32 ? NCCL_FLOAT32 : NCCL_FLOAT64,
                                      (nccl_comm_t)comm, (cudaStream_t)stream), "sgpmp_allreduce_stats");
    return SGPMP_OK;
}

// n_iters split-particle iterations enqueued on `stream`.  shape_local: this rank's slice (S = S_local, sample_gid0 =
// rank * S_local).  comm may be null when n_ranks == 1 (the exchange is then the identity: stats_all = stats_local).
extern "C" int sgpmp_iterate_split_particles(const sgpmp_shape_t* shape_local, const sgpmp_cost_desc_t* desc, const double* tables,
                                             double step_size, int32_t n_iters, uint64_t seed, uint32_t draw0, void* means,
                                             void* means_pre, void* comm, int32_t n_ranks, void* stats_local, void* stats_all,
                                             void* costs, void* grad, void* stream) {
    SGPMP_REQUIRE(shape_ok(shape_local), "sgpmp_iterate_split_particles: invalid shape");
    SGPMP_REQUIRE(desc && tables && means && stats_local && n_iters >= 1 && n_ranks >= 1, "sgpmp_iterate_split_particles: invalid argument");
    SGPMP_REQUIRE(n_ranks == 1 || (comm && stats_all && g_nccl.handle), "sgpmp_iterate_split_particles: n_ranks > 1 needs a communicator and stats_all");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t esz = shape_local->dtype == SGPMP_F32 ? 4 : 8;
    const size_t n_part = (size_t)shape_local->B * shape_local->G * shape_local->K;
    const size_t M = (size_t)shape_local->T * 2 * shape_local->n_dof;
    for (int it = 0; it < n_iters; ++it) {
        const bool last = (it == n_iters - 1);
        if (last && means_pre) {
            if (cudaMemcpyAsync(means_pre, means, n_part * M * esz, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
                set_error("sgpmp_iterate_split_particles: means_pre copy failed");
                return SGPMP_ERR_CUDA;
            }
        }
        int rc = iterate_stats_launch(*shape_local, *desc, tables, seed, draw0 + (uint32_t)it, means, last ? costs : nullptr, stats_local, st);
        if (rc != SGPMP_OK) return rc;
        const void* merged = stats_local;
        if (n_ranks > 1) {
            SGPMP_NCCL_CHECK(g_nccl.AllGather(stats_local, stats_all, n_part * (M + 2), shape_local->dtype == SGPMP_F32 ? NCCL_FLOAT32 : NCCL_FLOAT64,
                                              (nccl_comm_t)comm, st), "sgpmp_iterate_split_particles");
            count_launch();
            merged = stats_all;
        }
        rc = merge_apply_stats_launch(*shape_local, tables, step_size, merged, n_ranks, means, last ? grad : nullptr, st);
        if (rc != SGPMP_OK) return rc;
    }
    return SGPMP_OK;
}
