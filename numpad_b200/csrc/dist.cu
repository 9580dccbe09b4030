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

// ---- direct-store exchanges over NVLink peer memory --------------------------------
// Every rank stores its values straight into the slot it owns in each peer's heap, fences
// at system scope and bumps an arrival counter on the peer; then it waits until the
// counters of all peers have reached this exchange's sequence number.  Two parities of
// the data slots: a rank can be at most one exchange ahead of a peer that still reads
// the previous one (it cannot start exchange k+2 before every peer has published k+1,
// i.e. has finished reading k).
struct peer_ptrs { char* base[NPB_MAX_PEERS]; };

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void wait_peers(const unsigned long long* flags, int world, int rank,
                                           unsigned long long seq) {
    // threads 0 .. world-1 each watch one peer's arrival counter (in LOCAL memory)
    const int q = threadIdx.x;
    if (q < world && q != rank) {
        unsigned long long spins = 0;
        while (ld_acquire_sys(flags + q) < seq)
            if (++spins > (1ull << 31)) __trap();      // a dead peer must not hang the GPU
    }
    __syncthreads();
}

// one CTA: out slot of every peer <- x[idx[0..n)], then wait for the peers' slots.  The
// exchange number is read from (and bumped in) device memory: *dseq + 1
__global__ void __launch_bounds__(1024)
k_push_halo(int n, const int32_t* __restrict__ idx, const double* __restrict__ x, peer_ptrs pp,
            int64_t data_off, int64_t flag_off, int64_t slot, int64_t pstride, int world, int rank,
            unsigned long long* dseq, int wait) {
    const unsigned long long seq = *reinterpret_cast<volatile unsigned long long*>(dseq) + 1;
    slot += (int64_t)(seq & 1ull) * pstride;
    for (int i = threadIdx.x; i < n; i += 1024) {
        const double v = x[idx[i]];
        for (int q = 0; q < world; ++q)
            if (q != rank) reinterpret_cast<double*>(pp.base[q] + data_off)[slot + i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < world && threadIdx.x != rank)
        st_release_sys(reinterpret_cast<unsigned long long*>(pp.base[threadIdx.x] + flag_off) + rank, seq);
    if (wait)
        wait_peers(reinterpret_cast<const unsigned long long*>(pp.base[rank] + flag_off), world, rank, seq);
    if (threadIdx.x == 0) *dseq = seq;          // (every thread read it before the barrier above)
}

// one CTA: all-gather of a replicated vector `full` (LOCAL memory, fixed address) whose slice
// [off, off + n) this rank has just written: the slice goes into the exchange's parity half of
// every peer's heap region, then the peers' slices are copied from this rank's region into
// `full`.  Exchange number in device memory (*dseq + 1), total = length of the vector.
__global__ void __launch_bounds__(1024)
k_push_slice(int64_t n, peer_ptrs pp, int64_t data_off, int64_t flag_off, int64_t off, int64_t total,
             int world, int rank, unsigned long long* dseq, double* full) {
    const unsigned long long seq = *reinterpret_cast<volatile unsigned long long*>(dseq) + 1;
    const int64_t par = (int64_t)(seq & 1ull) * total;
    for (int64_t i = threadIdx.x; i < n; i += 1024) {
        const double v = full[off + i];
        for (int q = 0; q < world; ++q)
            if (q != rank) reinterpret_cast<double*>(pp.base[q] + data_off)[par + off + i] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < world && threadIdx.x != rank)
        st_release_sys(reinterpret_cast<unsigned long long*>(pp.base[threadIdx.x] + flag_off) + rank, seq);
    wait_peers(reinterpret_cast<const unsigned long long*>(pp.base[rank] + flag_off), world, rank, seq);
    const double* got = reinterpret_cast<const double*>(pp.base[rank] + data_off) + par;
    for (int64_t i = threadIdx.x; i < total; i += 1024)
        if (i < off || i >= off + n) full[i] = __ldcg(got + i);
    if (threadIdx.x == 0) *dseq = seq;
}

// buf[0..count) <- sum over ranks (fixed rank order: identical bits on every rank)
__global__ void __launch_bounds__(NPB_RED_SLOT)
k_allreduce_p2p(double* __restrict__ buf, int count, peer_ptrs pp, int64_t data_off,
                int64_t flag_off, int64_t pstride, int world, int rank, unsigned long long* dseq) {
    const int t = threadIdx.x;
    const unsigned long long seq = *reinterpret_cast<volatile unsigned long long*>(dseq) + 1;
    const int64_t parity_off = (int64_t)(seq & 1ull) * pstride;
    if (t < count) {
        const double v = buf[t];
        for (int q = 0; q < world; ++q)
            reinterpret_cast<double*>(pp.base[q] + data_off)[parity_off + (int64_t)rank * NPB_RED_SLOT + t] = v;
    }
    __threadfence_system();
    __syncthreads();
    if (t < world && t != rank)
        st_release_sys(reinterpret_cast<unsigned long long*>(pp.base[t] + flag_off) + rank, seq);
    wait_peers(reinterpret_cast<const unsigned long long*>(pp.base[rank] + flag_off), world, rank, seq);
    if (t < count) {
        const volatile double* mine = reinterpret_cast<const volatile double*>(pp.base[rank] + data_off) + parity_off;
        double s = 0.0;
        for (int q = 0; q < world; ++q) s += mine[(int64_t)q * NPB_RED_SLOT + t];
        buf[t] = s;
    }
    if (t == 0) *dseq = seq;
}

static inline peer_ptrs make_peers(const npb_dist* d) {
    peer_ptrs pp;
    for (int q = 0; q < NPB_MAX_PEERS; ++q) pp.base[q] = d->heap[q];
    return pp;
}

}  // namespace

// ---- internal API (dist.cuh) ------------------------------------------------------
int npb_dist_halo_exchange(const npb_halo& h, const double* x, cudaStream_t st, const double** xh,
                           npb_wait* wait) {
    if (wait) wait->flags = nullptr;
    // boundary values of x -> this rank's slot of the gather buffer -> every rank
    const npb_dist* d = (const npb_dist*)h.dist;
    if (h.p2p_data >= 0 && d->heap[d->rank]) {
        // h.seq: DEVICE counter of this operator's exchanges (the kernels read / bump it)
        const int64_t pstride = (int64_t)d->world * h.maxb;
        if (h.maxb > 0) {
            k_push_halo<<<1, 1024, 0, st>>>(h.n_send, h.send_idx, x, make_peers(d), h.p2p_data,
                                            h.p2p_flags, (int64_t)d->rank * h.maxb, pstride,
                                            d->world, d->rank, h.seq, wait ? 0 : 1);
            NPB_COUNT_LAUNCH();
            if (wait) {
                wait->flags = reinterpret_cast<const unsigned long long*>(d->heap[d->rank] + h.p2p_flags);
                wait->seq = 0;
                wait->world = d->world;
                wait->rank = d->rank;
                wait->dseq = h.seq;
                wait->pstride = pstride;
            }
        }
        // base of parity 0: the consuming kernel adds the parity of the exchange (npb_halo_begin)
        *xh = reinterpret_cast<const double*>(d->heap[d->rank] + h.p2p_data);
        return NPB_OK;
    }
    *xh = h.recv;
    double* mine = h.recv + (int64_t)d->rank * h.maxb;
    if (h.n_send > 0) {
        k_pack<<<npb_blocks(h.n_send, 256), 256, 0, st>>>(h.n_send, h.send_idx, x, mine);
        NPB_COUNT_LAUNCH();
    }
    if (h.maxb > 0)
        NPB_NCCL(g_nccl.allgather(mine, h.recv, (size_t)h.maxb, NCCL_DOUBLE, d->comm, st));
    return NPB_OK;
}

int npb_dist_allreduce_sum(const npb_dist* dc, double* buf, int count, cudaStream_t st) {
    npb_dist* d = const_cast<npb_dist*>(dc);
    if (!d || d->world <= 1 || count <= 0) return NPB_OK;
    if (d->red_data >= 0 && count <= NPB_RED_SLOT) {
        k_allreduce_p2p<<<1, NPB_RED_SLOT, 0, st>>>(buf, count, make_peers(d), d->red_data, d->red_flags,
                                                   (int64_t)d->world * NPB_RED_SLOT, d->world, d->rank,
                                                   reinterpret_cast<unsigned long long*>(d->heap[d->rank] + d->red_dseq));
        NPB_COUNT_LAUNCH();
        return NPB_OK;
    }
    NPB_NCCL(g_nccl.allreduce(buf, buf, (size_t)count, NCCL_DOUBLE, NCCL_SUM, d->comm, st));
    return NPB_OK;
}

double* npb_dist_heap_ptr(const npb_dist* d, int64_t off) {
    return (d && d->heap[d->rank]) ? reinterpret_cast<double*>(d->heap[d->rank] + off) : nullptr;
}

// direct-store all-gather of the replicated vector `full` (local memory): this rank's slice
// [offsets[rank], offsets[rank+1]) is already written in it; the heap region at data_off
// (2 parities x offsets[world] doubles) carries the peers' slices; dseq = device counter
int npb_dist_push_slice(const npb_dist* d, int64_t data_off, int64_t flag_off,
                        const int64_t* offsets, unsigned long long* dseq, double* full,
                        cudaStream_t st) {
    k_push_slice<<<1, 1024, 0, st>>>(offsets[d->rank + 1] - offsets[d->rank], make_peers(d), data_off,
                                     flag_off, offsets[d->rank], offsets[d->world], d->world, d->rank,
                                     dseq, full);
    NPB_COUNT_LAUNCH();
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
    cudaDeviceSynchronize();
    for (int q = 0; q < d->world && q < NPB_MAX_PEERS; ++q) {
        if (q != d->rank && d->ipc_base[q]) cudaIpcCloseMemHandle(d->ipc_base[q]);
    }     // (this rank's own heap belongs to the caller)
    if (d->comm && g_nccl.destroy) g_nccl.destroy(d->comm);
    delete d;
    return NPB_OK;
}

// ---- symmetric heap (direct-store exchanges).  The library allocates no device memory:
// create: the CALLER hands in `bytes` of device memory (`mem`, e.g. a torch tensor it keeps
// alive until npb_dist_free); it is zeroed and described by 72 bytes -- the CUDA IPC handle of
// the allocation that contains it (64) + the offset of `mem` inside that allocation (8).
// open: collective, all_handles = world x 72 bytes in rank order; alloc: bump allocation, the
// same call sequence on every rank gives the same offsets everywhere.  Without a heap every
// exchange goes through NCCL.
typedef int (*cu_get_range_fn)(unsigned long long* base, size_t* size, unsigned long long ptr);
int npb_dist_heap_create(void* handle, void* mem, int64_t bytes, char* ipc72) {
    npb_dist* d = (npb_dist*)handle;
    if (!d || d->world > NPB_MAX_PEERS || bytes <= 0 || !mem) return NPB_ERR_ARG;
    // base of the cudaMalloc allocation that contains `mem` (an allocator may hand out interior
    // pointers): driver entry point resolved at run time, no link-time dependency on libcuda
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qr;
    NPB_CUDA(cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &qr));
    if (!fn) { npb_set_error("heap_create: cuMemGetAddressRange unavailable"); return NPB_ERR_CUDA; }
    unsigned long long base = 0;
    size_t size = 0;
    if (((cu_get_range_fn)fn)(&base, &size, (unsigned long long)(uintptr_t)mem) != 0) {
        npb_set_error("heap_create: cuMemGetAddressRange failed");
        return NPB_ERR_CUDA;
    }
    const int64_t off = (int64_t)((unsigned long long)(uintptr_t)mem - base);
    if (off < 0 || (size_t)(off + bytes) > size) return NPB_ERR_ARG;
    NPB_CUDA(cudaMemset(mem, 0, (size_t)bytes));
    cudaIpcMemHandle_t hnd;
    NPB_CUDA(cudaIpcGetMemHandle(&hnd, (void*)(uintptr_t)base));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(ipc72, &hnd, 64);
    memcpy(ipc72 + 64, &off, 8);
    d->heap[d->rank] = (char*)mem;
    d->heap_bytes = bytes;
    d->heap_used = 0;
    return NPB_OK;
}

int npb_dist_heap_open(void* handle, const char* all_handles) {
    npb_dist* d = (npb_dist*)handle;
    if (!d || !d->heap[d->rank]) return NPB_ERR_ARG;
    for (int q = 0; q < d->world; ++q) {
        if (q == d->rank) continue;
        cudaIpcMemHandle_t hnd;
        int64_t off = 0;
        memcpy(&hnd, all_handles + (size_t)q * 72, 64);
        memcpy(&off, all_handles + (size_t)q * 72 + 64, 8);
        void* p = nullptr;
        NPB_CUDA(cudaIpcOpenMemHandle(&p, hnd, cudaIpcMemLazyEnablePeerAccess));
        d->ipc_base[q] = p;
        d->heap[q] = (char*)p + off;
    }
    // all-reduce slots: 2 parities x world x NPB_RED_SLOT doubles + world counters
    d->red_data = 0;
    d->red_flags = (int64_t)2 * d->world * NPB_RED_SLOT * 8;
    d->red_dseq = d->red_flags + 128;          // (the flags take world * 8 <= 64 bytes)
    d->heap_used = d->red_flags + 256;
    return NPB_OK;
}

// forget every allocation but the all-reduce slots and zero the freed part.  The caller
// brackets this with barriers: no peer may still push into the old layout, none may push
// into the new one before the memset has run.
int npb_dist_heap_reset(void* handle, void* stream) {
    npb_dist* d = (npb_dist*)handle;
    if (!d || !d->heap[d->rank]) return NPB_ERR_ARG;
    const int64_t base = d->red_flags + 256;
    NPB_CUDA(cudaMemsetAsync(d->heap[d->rank] + base, 0, (size_t)(d->heap_used - base), (cudaStream_t)stream));
    NPB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    d->heap_used = base;
    return NPB_OK;
}

// -> offset of `bytes` (rounded up to 256) in the heap, or < 0 when it is full
int64_t npb_dist_heap_alloc(void* handle, int64_t bytes) {
    npb_dist* d = (npb_dist*)handle;
    if (!d || !d->heap[d->rank] || bytes < 0) return -1;
    const int64_t b = (bytes + 255) / 256 * 256;
    if (d->heap_used + b > d->heap_bytes) return -1;
    const int64_t off = d->heap_used;
    d->heap_used += b;
    return off;
}

int npb_dist_rank(void* handle) { return handle ? ((npb_dist*)handle)->rank : 0; }
int npb_dist_world(void* handle) { return handle ? ((npb_dist*)handle)->world : 1; }

// npb_allreduce_fn for npb_krylov_ops: ctx = the npb_dist handle
int npb_dist_allreduce_cb(void* ctx, double* buf, int count, void* stream) {
    return npb_dist_allreduce_sum((const npb_dist*)ctx, buf, count, (cudaStream_t)stream);
}

}  // extern "C"
