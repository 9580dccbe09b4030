// Internal declarations of the multi-GPU layer (dist.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

constexpr int NPB_MAX_PEERS = 8;       // direct-store exchanges: one node, <= 8 GPUs
constexpr int NPB_RED_SLOT = 128;      // doubles per rank in the all-reduce slots

struct npb_dist {            // one per process: the library's NCCL communicator
    void* comm = nullptr;
    int rank = 0, world = 1;
    // symmetric heap for direct-store exchanges over NVLink peer memory: CALLER-OWNED device
    // memory with the same layout on every rank, opened on the peers through CUDA IPC
    char* heap[NPB_MAX_PEERS] = {};        // heap[q] = base of rank q's heap in this process
    void* ipc_base[NPB_MAX_PEERS] = {};    // mappings opened here (closed by npb_dist_free)
    int64_t heap_bytes = 0, heap_used = 0;
    int64_t red_data = -1, red_flags = -1; // offsets of the all-reduce slots
    int64_t red_dseq = -1;                 // offset of the all-reduce exchange counter (device)
};

// Halo of one row-sharded operator.  Local column numbering: [0, n_own) = the columns this
// rank owns, n_own + q * maxb + k = the k-th boundary value of rank q.  Every rank packs the
// owned entries the others read (send_idx, n_send <= maxb of them) into its slot of `recv`
// (world * maxb doubles) and one all-gather fills the other slots.
struct npb_halo {
    const void* dist;            // npb_dist*; NULL: operator is not sharded
    int64_t n_own;               // owned columns (local indices below this need no halo)
    const int32_t* send_idx;     // [n_send] local indices of the boundary values to publish
    int32_t n_send, maxb;
    double* recv;                // [world * maxb]
    // direct-store path (npb_dist_heap_*): offsets into the symmetric heap of 2 x world x maxb
    // doubles (two parities) and world 64-bit arrival counters; seq = [host] exchange counter
    int64_t p2p_data, p2p_flags;  // < 0: use the NCCL all-gather into `recv`
    unsigned long long* seq;
};

// what the consumer of a direct-store exchange still has to wait for: the arrival counters
// flags[q] (q != rank) of this rank reaching seq.  flags == NULL: nothing (already waited).
// The sequence number of an exchange lives ON THE DEVICE (dseq, bumped by the publishing
// kernel): neither the number nor the parity of the slots it selects is a kernel argument, so
// the launch sequence of a solve phase is identical from call to call and can be replayed as a
// CUDA graph.
struct npb_wait {
    const unsigned long long* flags;
    unsigned long long seq;              // filled in by npb_halo_begin from *dseq
    int world, rank;
    const unsigned long long* dseq;      // exchange counter of the operator (already bumped)
    long long pstride;                   // doubles between the two parities of the halo slots
};

__device__ __forceinline__ unsigned long long npb_ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
}

// prologue of a kernel that reads a halo: threads 0 .. world-1 each watch one peer's arrival
// counter (in LOCAL memory), then the CTA synchronises.  Call from ALL threads of the CTA.
__device__ __forceinline__ void npb_wait_peers(const npb_wait& w) {
    if (w.flags) {
        const int q = threadIdx.x;
        if (q < w.world && q != w.rank) {
            unsigned long long spins = 0;
            while (npb_ld_acquire_sys(w.flags + q) < w.seq)
                if (++spins > (1ull << 31)) __trap();      // a dead peer must not hang the GPU
        }
        __syncthreads();
    }
}

__device__ __forceinline__ unsigned long long npb_ld_volatile(const unsigned long long* p) {
    return *reinterpret_cast<const volatile unsigned long long*>(p);
}

// prologue of a kernel that reads a halo published by direct stores: pick the parity of this
// exchange (xh is handed in as the base of parity 0), then wait for the peers
__device__ __forceinline__ void npb_halo_begin(npb_wait& w, const double*& xh) {
    if (w.flags) {
        w.seq = npb_ld_volatile(w.dseq);
        xh += (long long)(w.seq & 1ull) * w.pstride;
    }
    npb_wait_peers(w);
}

// publishes / fetches the boundary values of x; *xh = where the SpMV reads them.  With
// `wait` != NULL a direct-store exchange only PUBLISHES and leaves the wait for the peers'
// values to the consumer kernel (overlaps the NVLink latency with that kernel's launch).
int npb_dist_halo_exchange(const npb_halo& h, const double* x, cudaStream_t st, const double** xh,
                           npb_wait* wait);
int npb_dist_allreduce_sum(const npb_dist* d, double* buf, int count, cudaStream_t st);
int npb_dist_allgatherv(const npb_dist* d, const double* mine, double* full,
                        const int64_t* offsets, cudaStream_t st);
double* npb_dist_heap_ptr(const npb_dist* d, int64_t off);
int npb_dist_push_slice(const npb_dist* d, int64_t data_off, int64_t flag_off,
                        const int64_t* offsets, unsigned long long* dseq, double* full,
                        cudaStream_t st);
