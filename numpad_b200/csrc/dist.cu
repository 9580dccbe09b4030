// Multi-GPU plumbing of the row-sharded Newton linear solve (SURVEY section 8e;
// replaces the MPI calls of the reference's admpisolve.py:146-188 -- the partial-y
// exchange of MpiJacobian.matvec and the scalar Allreduce of every _dot / _norm2).
//
// One process per GPU.  The library owns a NCCL communicator of its own (created from
// a unique id that the host distributes through torch.distributed), so that the C++
// solve phase can issue its collectives on the caller's stream without returning to
// Python: a halo exchange is `pack boundary values -> ncclAllGather` in front of the
// local SpMV, an inner product is `multi-dot kernel -> ncclAllReduce` on the device
// scalars.  NCCL is resolved at run time (dlopen of the copy torch has already
// loaded): the library has no link-time dependency on it and loads on a CPU-only box.
#include <dlfcn.h>
#include <string.h>

#include "common.cuh"
#include "dist.cuh"

namespace {

typedef struct { char internal[128]; } nccl_uid;
typedef int (*fn_get_uid)(nccl_uid*);
typedef int (*fn_init_rank)(void**, int, nccl_uid, int);
typedef int (*fn_destroy)(void*);
typedef int (*fn_allreduce)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_allgather)(const void*, void*, size_t, int, void*, cudaStream_t);
typedef int (*fn_bcast)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*fn_group)(void);
typedef const char* (*fn_errstr)(int);

struct nccl_api {
    void* lib = nullptr;
    fn_get_uid get_uid = nullptr;
    fn_init_rank init_rank = nullptr;
    fn_destroy destroy = nullptr;
    fn_allreduce allreduce = nullptr;
    fn_allgather allgather = nullptr;
    fn_bcast bcast = nullptr;
    fn_group group_start = nullptr, group_end = nullptr;
    fn_errstr errstr = nullptr;
} g_nccl;

constexpr int NCCL_DOUBLE = 8;     // ncclFloat64
constexpr int NCCL_SUM = 0;

int load_nccl(const char* path) {
    if (g_nccl.lib) return NPB_OK;
    void* h = nullptr;
    if (path && path[0]) h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) {
        npb_set_error("dist: cannot load NCCL (%s)", dlerror());
        return NPB_ERR_ARG;
    }
#define NPB_SYM(field, type, name)                                           \
    g_nccl.field = (type)dlsym(h, name);                                     \
    if (!g_nccl.field) { npb_set_error("dist: NCCL lacks %s", name); return NPB_ERR_ARG; }
    NPB_SYM(get_uid, fn_get_uid, "ncclGetUniqueId")
    NPB_SYM(init_rank, fn_init_rank, "ncclCommInitRank")
    NPB_SYM(destroy, fn_destroy, "ncclCommDestroy")
    NPB_SYM(allreduce, fn_allreduce, "ncclAllReduce")
    NPB_SYM(allgather, fn_allgather, "ncclAllGather")
    NPB_SYM(bcast, fn_bcast, "ncclBroadcast")
    NPB_SYM(group_start, fn_group, "ncclGroupStart")
    NPB_SYM(group_end, fn_group, "ncclGroupEnd")
    NPB_SYM(errstr, fn_errstr, "ncclGetErrorString")
#undef NPB_SYM
    g_nccl.lib = h;
    return NPB_OK;
}

#define NPB_NCCL(call)                                                       \
    do {                                                                     \
        int r_ = (call);                                                     \
        if (r_ != 0) {                                                       \
            npb_set_error("%s:%d NCCL: %s", __FILE__, __LINE__, g_nccl.errstr(r_)); \
            return NPB_ERR_CUDA;                                             \
        }                                                                    \
    } while (0)

__global__ void __launch_bounds__(256)
k_pack(int n, const int32_t* __restrict__ idx, const double* __restrict__ x,
       double* __restrict__ out) {
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (i < n) out[i] = x[idx[i]];
}

}  // namespace

// ---- internal API (dist.cuh) ------------------------------------------------------
int npb_dist_halo_exchange(const npb_halo& h, const double* x, cudaStream_t st) {
    // boundary values of x -> this rank's slot of the gather buffer -> every rank
    const npb_dist* d = (const npb_dist*)h.dist;
    double* mine = h.recv + (int64_t)d->rank * h.maxb;
    if (h.n_send > 0) {
        k_pack<<<npb_blocks(h.n_send, 256), 256, 0, st>>>(h.n_send, h.send_idx, x, mine);
        NPB_COUNT_LAUNCH();
    }
    if (h.maxb > 0)
        NPB_NCCL(g_nccl.allgather(mine, h.recv, (size_t)h.maxb, NCCL_DOUBLE, d->comm, st));
    return NPB_OK;
}

int npb_dist_allreduce_sum(const npb_dist* d, double* buf, int count, cudaStream_t st) {
    if (!d || d->world <= 1 || count <= 0) return NPB_OK;
    NPB_NCCL(g_nccl.allreduce(buf, buf, (size_t)count, NCCL_DOUBLE, NCCL_SUM, d->comm, st));
    return NPB_OK;
}

int npb_dist_allgatherv(const npb_dist* d, const double* mine, double* full,
                        const int64_t* offsets /*[world + 1], host*/, cudaStream_t st) {
    // rank r contributes full[offsets[r] : offsets[r+1]]; `mine` may alias that slice
    NPB_NCCL(g_nccl.group_start());
    for (int r = 0; r < d->world; ++r) {
        const size_t cnt = (size_t)(offsets[r + 1] - offsets[r]);
        if (cnt == 0) continue;
        int rc = g_nccl.bcast(r == d->rank ? (const void*)mine : (const void*)(full + offsets[r]),
                              full + offsets[r], cnt, NCCL_DOUBLE, r, d->comm, st);
        if (rc != 0) { g_nccl.group_end(); npb_set_error("dist: NCCL broadcast: %s", g_nccl.errstr(rc)); return NPB_ERR_CUDA; }
    }
    NPB_NCCL(g_nccl.group_end());
    return NPB_OK;
}

// ---- C-ABI ------------------------------------------------------------------------
extern "C" {

// 128-byte NCCL unique id (rank 0 creates it, the host broadcasts it)
int npb_dist_unique_id(const char* nccl_path, char* id128) {
    int rc = load_nccl(nccl_path);
    if (rc) return rc;
    nccl_uid u;
    NPB_NCCL(g_nccl.get_uid(&u));
    memcpy(id128, u.internal, 128);
    return NPB_OK;
}

// collective over all ranks: create the communicator; *handle is an npb_dist*
int npb_dist_init(const char* nccl_path, const char* id128, int rank, int world, void** handle) {
    int rc = load_nccl(nccl_path);
    if (rc) return rc;
    if (world < 1 || rank < 0 || rank >= world || !handle) return NPB_ERR_ARG;
    nccl_uid u;
    memcpy(u.internal, id128, 128);
    npb_dist* d = new npb_dist();
    d->rank = rank;
    d->world = world;
    int r = g_nccl.init_rank(&d->comm, world, u, rank);
    if (r != 0) {
        npb_set_error("dist: ncclCommInitRank: %s", g_nccl.errstr(r));
        delete d;
        return NPB_ERR_CUDA;
    }
    *handle = d;
    return NPB_OK;
}

int npb_dist_free(void* handle) {
    npb_dist* d = (npb_dist*)handle;
    if (!d) return NPB_OK;
    if (d->comm && g_nccl.destroy) g_nccl.destroy(d->comm);
    delete d;
    return NPB_OK;
}

int npb_dist_rank(void* handle) { return handle ? ((npb_dist*)handle)->rank : 0; }
int npb_dist_world(void* handle) { return handle ? ((npb_dist*)handle)->world : 1; }

// npb_allreduce_fn for npb_krylov_ops: ctx = the npb_dist handle
int npb_dist_allreduce_cb(void* ctx, double* buf, int count, void* stream) {
    return npb_dist_allreduce_sum((const npb_dist*)ctx, buf, count, (cudaStream_t)stream);
}

}  // extern "C"
