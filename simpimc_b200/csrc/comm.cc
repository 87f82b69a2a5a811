// Multi-GPU data plane of a slice-sharded path behind the C ABI (include/simpimc_b200.h, section
// "slice sharding over several GPUs"): the ring halo of one slice of positions, the all-reduce of
// the shard partial sums and the ring rotation of the slices, issued with NCCL on the CONTEXT'S
// stream so that they order with the kernels around them (and can be captured into the same CUDA
// graph).  The reference never splits a path (its MPI ranks are independent walkers,
// src/framework/framework_class.h:44-53); what makes the split exact is that the pair action at
// level 0 couples slice b only with b + 1 (src/actions/pair_action/pair_action_class.h:282-288)
// while rho_k(b) and the k sums are slice-local (src/data_structures/species_class.h:391-395).
//
// Host-only translation unit written against the public C ABI plus three accessors
// (internal.h).  NCCL is bound at run time (dlopen of libnccl.so.2: the copy the process already
// holds, e.g. PyTorch's, else the system one), so the library loads on hosts without NCCL and
// single-GPU users never touch it; every entry point here fails loudly when it is missing.
#include <dlfcn.h>

#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/simpimc_b200.h"
#include "internal.h"

namespace {

// ---- the slice of nccl.h this file needs (ABI-stable since NCCL 2.0) --------------------------
typedef struct ncclComm *ncclComm_t;
typedef struct {
    char internal[128];
} ncclUniqueId;
enum { ncclSuccess = 0 };
enum { ncclSum = 0 };
enum { ncclFloat64 = 8 };

struct NcclApi {
    void *handle = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int *) = nullptr;
    std::string error;
};

NcclApi &Nccl() {
    static NcclApi api;
    if (api.handle || !api.error.empty()) return api;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // the copy already in the process
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        api.error = std::string("NCCL is not available (dlopen libnccl.so.2: ") + dlerror() + ")";
        return api;
    }
#define PIMC_NCCL_SYM(field, name)                                         \
    *reinterpret_cast<void **>(&api.field) = dlsym(h, name);               \
    if (!api.field) {                                                      \
        api.error = std::string("NCCL symbol missing: ") + name;           \
        return api;                                                        \
    }
    PIMC_NCCL_SYM(GetUniqueId, "ncclGetUniqueId")
    PIMC_NCCL_SYM(CommInitRank, "ncclCommInitRank")
    PIMC_NCCL_SYM(CommDestroy, "ncclCommDestroy")
    PIMC_NCCL_SYM(AllReduce, "ncclAllReduce")
    PIMC_NCCL_SYM(Send, "ncclSend")
    PIMC_NCCL_SYM(Recv, "ncclRecv")
    PIMC_NCCL_SYM(GroupStart, "ncclGroupStart")
    PIMC_NCCL_SYM(GroupEnd, "ncclGroupEnd")
    PIMC_NCCL_SYM(GetErrorString, "ncclGetErrorString")
    PIMC_NCCL_SYM(GetVersion, "ncclGetVersion")
#undef PIMC_NCCL_SYM
    api.handle = h;
    return api;
}

int Fail(int code, const std::string &msg) { return pimc_internal_fail(code, msg.c_str()); }

#define PIMC_NCCL(expr)                                                                                        \
    do {                                                                                                       \
        int r__ = (expr);                                                                                      \
        if (r__ != ncclSuccess) return Fail(PIMC_ERR_CUDA, std::string(#expr) + ": " + Nccl().GetErrorString(r__)); \
    } while (0)
#define PIMC_CU(expr)                                                                                          \
    do {                                                                                                       \
        cudaError_t e__ = (expr);                                                                              \
        if (e__ != cudaSuccess) return Fail(PIMC_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
    } while (0)

}  // namespace

struct pimc_comm {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1, device = 0;
    double *send = nullptr, *recv = nullptr;  // ring buffers, grown on demand
    size_t cap = 0;
    int64_t bytes_sent = 0;  // payload this rank handed to ncclSend / ncclAllReduce (measurement)
};

namespace {
int EnsureRing(pimc_comm *cm, size_t n_doubles) {
    if (cm->cap >= n_doubles) return PIMC_OK;
    if (cm->send) cudaFree(cm->send);
    if (cm->recv) cudaFree(cm->recv);
    cm->send = cm->recv = nullptr;
    cm->cap = 0;
    PIMC_CU(cudaMalloc((void **)&cm->send, n_doubles * sizeof(double)));
    PIMC_CU(cudaMalloc((void **)&cm->recv, n_doubles * sizeof(double)));
    cm->cap = n_doubles;
    return PIMC_OK;
}

/// send -> previous rank, recv <- next rank (the beta-periodic ring of slice shards).
int RingShift(pimc_comm *cm, cudaStream_t stream, size_t n_doubles) {
    NcclApi &nc = Nccl();
    const int prev = (cm->rank + cm->world - 1) % cm->world, next = (cm->rank + 1) % cm->world;
    PIMC_NCCL(nc.GroupStart());
    PIMC_NCCL(nc.Send(cm->send, n_doubles, ncclFloat64, prev, cm->comm, stream));
    PIMC_NCCL(nc.Recv(cm->recv, n_doubles, ncclFloat64, next, cm->comm, stream));
    PIMC_NCCL(nc.GroupEnd());
    cm->bytes_sent += (int64_t)(n_doubles * sizeof(double));
    return PIMC_OK;
}
}  // namespace

extern "C" {

int pimc_comm_unique_id(void *id128) {
    if (!id128) return Fail(PIMC_ERR_INVALID, "null id buffer");
    NcclApi &nc = Nccl();
    if (!nc.handle) return Fail(PIMC_ERR_UNSUPPORTED, nc.error);
    ncclUniqueId id;
    PIMC_NCCL(nc.GetUniqueId(&id));
    std::memcpy(id128, &id, sizeof(id));
    return PIMC_OK;
}

int pimc_comm_init(pimc_ctx *ctx, const void *id128, int32_t rank, int32_t world, pimc_comm **out) {
    if (!ctx || !out) return Fail(PIMC_ERR_INVALID, "null context or output");
    if (world < 1 || rank < 0 || rank >= world) return Fail(PIMC_ERR_INVALID, "bad rank / world size");
    pimc_comm *cm = new pimc_comm;
    cm->rank = rank;
    cm->world = world;
    cm->device = pimc_internal_device(ctx);
    if (world > 1) {
        if (!id128) {
            delete cm;
            return Fail(PIMC_ERR_INVALID, "null NCCL unique id");
        }
        NcclApi &nc = Nccl();
        if (!nc.handle) {
            delete cm;
            return Fail(PIMC_ERR_UNSUPPORTED, nc.error);
        }
        cudaError_t e = cudaSetDevice(cm->device);
        if (e != cudaSuccess) {
            delete cm;
            return Fail(PIMC_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e));
        }
        ncclUniqueId id;
        std::memcpy(&id, id128, sizeof(id));
        int r = nc.CommInitRank(&cm->comm, world, id, rank);
        if (r != ncclSuccess) {
            std::string msg = std::string("ncclCommInitRank: ") + nc.GetErrorString(r);
            delete cm;
            return Fail(PIMC_ERR_CUDA, msg);
        }
    }
    *out = cm;
    return PIMC_OK;
}

int pimc_comm_destroy(pimc_comm *cm) {
    if (!cm) return PIMC_OK;
    cudaSetDevice(cm->device);
    if (cm->send) cudaFree(cm->send);
    if (cm->recv) cudaFree(cm->recv);
    if (cm->comm) Nccl().CommDestroy(cm->comm);
    delete cm;
    return PIMC_OK;
}

int64_t pimc_comm_bytes_sent(pimc_comm *cm) { return cm ? cm->bytes_sent : 0; }

int pimc_halo_exchange(pimc_ctx *ctx, pimc_comm *cm, int32_t species) {
    if (!ctx || !cm) return Fail(PIMC_ERR_INVALID, "null context or communicator");
    if (cm->world == 1) return PIMC_OK;  // an unsharded path has no halo
    const int N = pimc_internal_n_part(ctx, species);
    if (N < 0) return Fail(PIMC_ERR_INVALID, "species index out of range");
    PIMC_CU(cudaSetDevice(cm->device));
    cudaStream_t stream = (cudaStream_t)pimc_ctx_stream(ctx);
    const size_t n = (size_t)pimc_internal_n_clones(ctx) * N * 3;
    int rc = EnsureRing(cm, n);
    if (rc != PIMC_OK) return rc;
    if ((rc = pimc_halo_pack(ctx, species, cm->send)) != PIMC_OK) return rc;
    if ((rc = RingShift(cm, stream, n)) != PIMC_OK) return rc;
    return pimc_halo_unpack(ctx, species, cm->recv);
}

int pimc_allreduce_sum(pimc_ctx *ctx, pimc_comm *cm, double *d_buf, int64_t n) {
    if (!ctx || !cm || !d_buf || n < 0) return Fail(PIMC_ERR_INVALID, "bad argument");
    if (cm->world == 1 || n == 0) return PIMC_OK;
    PIMC_CU(cudaSetDevice(cm->device));
    cudaStream_t stream = (cudaStream_t)pimc_ctx_stream(ctx);
    PIMC_NCCL(Nccl().AllReduce(d_buf, d_buf, (size_t)n, ncclFloat64, ncclSum, cm->comm, stream));
    cm->bytes_sent += n * (int64_t)sizeof(double);
    return PIMC_OK;
}

int pimc_rotate(pimc_ctx *ctx, pimc_comm *cm, int32_t shift) {
    if (!ctx || !cm) return Fail(PIMC_ERR_INVALID, "null context or communicator");
    if (cm->world == 1) return PIMC_OK;
    PIMC_CU(cudaSetDevice(cm->device));
    cudaStream_t stream = (cudaStream_t)pimc_ctx_stream(ctx);
    const int n_species = pimc_internal_n_species(ctx), C = pimc_internal_n_clones(ctx);
    int rc;
    for (int s = 0; s < n_species; ++s) {
        const size_t n = (size_t)C * pimc_internal_n_part(ctx, s) * 3 * (size_t)shift;
        if ((rc = EnsureRing(cm, n)) != PIMC_OK) return rc;
        if ((rc = pimc_rotate_pack(ctx, s, shift, cm->send)) != PIMC_OK) return rc;
        if ((rc = RingShift(cm, stream, n)) != PIMC_OK) return rc;
        if ((rc = pimc_rotate_apply(ctx, s, shift, cm->recv)) != PIMC_OK) return rc;
        if ((rc = pimc_halo_exchange(ctx, cm, s)) != PIMC_OK) return rc;
    }
    for (int s = 0; s < n_species; ++s)
        if ((rc = pimc_rhok_rebuild(ctx, s)) != PIMC_OK) return rc;
    return PIMC_OK;
}

int pimc_sharded_evaluate(pimc_ctx *ctx, pimc_comm *cm, int32_t which, pimc_action *const *actions, int32_t n_actions, double *d_out) {
    if (!ctx || !cm || !actions || !d_out || n_actions < 1) return Fail(PIMC_ERR_INVALID, "bad argument");
    const int C = pimc_internal_n_clones(ctx);
    // the actions' whole-path kernels side by side (side streams forked from the context's and joined back), then ONE all-reduce
    const int rc = pimc_internal_evaluate_many(ctx, which, actions, n_actions, d_out);
    if (rc != PIMC_OK) return rc;
    return pimc_allreduce_sum(ctx, cm, d_out, (int64_t)n_actions * C);
}

}  // extern "C"
