// Internal declarations of the multi-GPU layer (dist.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct npb_dist {            // one per process: the library's NCCL communicator
    void* comm = nullptr;
    int rank = 0, world = 1;
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
};

int npb_dist_halo_exchange(const npb_halo& h, const double* x, cudaStream_t st);
int npb_dist_allreduce_sum(const npb_dist* d, double* buf, int count, cudaStream_t st);
int npb_dist_allgatherv(const npb_dist* d, const double* mine, double* full,
                        const int64_t* offsets, cudaStream_t st);
