// COO -> canonical CSR (sort by (row, col), sum duplicates, optionally drop
// exact-zero sums).  Replaces scipy's coo_tocsr + sum_duplicates behind
// csr_jac.tocsr (reference adstate.py:268-272) for genuinely general blocks,
// and is the universal fallback of the chain-rule kernels (non-monotone
// relabels, general products, rows too long for the row-parallel kernels).
// The sort itself is CUB's device radix sort on 64-bit (row,col) keys; it is
// off the Newton hot path of the cavity workload.
#include <cub/cub.cuh>

#include "common.cuh"
#include "scan.cuh"

namespace {

constexpr int TPB = 256;

// key = row << col_bits | col with col_bits = bits needed for the largest
// column seen by the caller: the radix sort then runs over row_bits + col_bits
// instead of 64 bits (one 8-bit pass less for every 8 bits saved)
__global__ void __launch_bounds__(TPB)
k_make_keys(int64_t nnz, const int32_t* __restrict__ rows,
            const int32_t* __restrict__ cols, int col_bits, uint64_t* __restrict__ keys) {
    int64_t k = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (k < nnz)
        keys[k] = ((uint64_t)(uint32_t)rows[k] << col_bits) | (uint32_t)cols[k];
}

// heads of equal-key runs sum their run (in input order, like coo_tocsr),
// decide whether the entry survives and count it for its row
__global__ void __launch_bounds__(TPB)
k_reduce_runs(int64_t nnz, const uint64_t* __restrict__ keys,
              double* __restrict__ vals, int prune, int col_bits,
              int32_t* __restrict__ flag, int32_t* __restrict__ row_cnt) {
    int64_t k = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (k >= nnz) return;
    const uint64_t key = keys[k];
    int32_t f = 0;
    if (k == 0 || keys[k - 1] != key) {
        double s = vals[k];
        for (int64_t t = k + 1; t < nnz && keys[t] == key; ++t)
            s = __dadd_rn(s, vals[t]);
        vals[k] = s;
        f = (!prune || s != 0.0);
        if (f) atomicAdd(row_cnt + (int32_t)(key >> col_bits), 1);
    }
    flag[k] = f;
}

__global__ void __launch_bounds__(TPB)
k_compact(int64_t nnz, const uint64_t* __restrict__ keys,
          const double* __restrict__ vals, const int32_t* __restrict__ flag,
          const int32_t* __restrict__ pos, int col_bits, int32_t* __restrict__ oidx,
          double* __restrict__ out) {
    int64_t k = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (k >= nnz || !flag[k]) return;
    const int32_t o = pos[k];
    oidx[o] = (int32_t)(keys[k] & ((1ull << col_bits) - 1ull));
    out[o] = vals[k];
}

__global__ void __launch_bounds__(TPB)
k_max_cnt(int64_t nrows, const int32_t* __restrict__ cnt, int64_t* __restrict__ stats) {
    int64_t r = (int64_t)blockIdx.x * TPB + threadIdx.x;
    int32_t c = (r < nrows) ? cnt[r] : 0;
    for (int d = 16; d; d >>= 1) c = max(c, __shfl_xor_sync(0xffffffffu, c, d));
    if ((threadIdx.x & 31) == 0 && c > 0)
        atomicMax((unsigned long long*)(stats + 1), (unsigned long long)c);
}

struct Layout {
    size_t keys_in, keys_out, vals, flag, pos, row_cnt, scan_ws, cub, total;
    size_t cub_bytes;
};

inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

Layout make_layout(int64_t nrows, int64_t nnz) {
    Layout L;
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off += align256(bytes); return o; };
    L.keys_in = take(8 * (size_t)nnz);
    L.keys_out = take(8 * (size_t)nnz);
    L.vals = take(8 * (size_t)nnz);
    L.flag = take(4 * (size_t)nnz);
    L.pos = take(4 * ((size_t)nnz + 1));
    L.row_cnt = take(4 * ((size_t)nrows + 1));
    int64_t m = nnz > nrows ? nnz : nrows;
    L.scan_ws = take(8 * (size_t)npb_scan_ws_elems(m));
    size_t cb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cb, (const uint64_t*)nullptr,
                                    (uint64_t*)nullptr, (const double*)nullptr,
                                    (double*)nullptr, nnz > 0 ? nnz : 1);
    L.cub_bytes = cb;
    L.cub = take(cb);
    L.total = off;
    return L;
}

}  // namespace

#define ST(s) ((cudaStream_t)(s))

extern "C" {

int64_t npb_coo_to_csr_workspace(int64_t nrows, int64_t nnz) {
    return (int64_t)make_layout(nrows, nnz).total;
}

// Phase 1: sort + reduce inside the workspace; indptr_out[nrows+1] is final,
// stats[0] = nnz of the result, stats[1] = longest row.
static int bits_for(int64_t n) {
    int b = 1;
    while (b < 32 && ((int64_t)1 << b) < n) ++b;
    return b;
}

int npb_coo_to_csr_symbolic(int64_t nrows, int64_t ncols, int64_t nnz, const int32_t* rows,
                            const int32_t* cols, const double* vals, int prune,
                            void* ws, int64_t ws_bytes, int32_t* indptr_out,
                            int64_t* stats, void* stream) {
    const int col_bits = bits_for(ncols);
    Layout L = make_layout(nrows, nnz);
    if ((int64_t)L.total > ws_bytes) {
        npb_set_error("coo_to_csr: workspace %lld < %lld", (long long)ws_bytes,
                      (long long)L.total);
        return NPB_ERR_WORKSPACE;
    }
    char* base = (char*)ws;
    uint64_t* keys_in = (uint64_t*)(base + L.keys_in);
    uint64_t* keys_out = (uint64_t*)(base + L.keys_out);
    double* svals = (double*)(base + L.vals);
    int32_t* flag = (int32_t*)(base + L.flag);
    int32_t* pos = (int32_t*)(base + L.pos);
    int32_t* row_cnt = (int32_t*)(base + L.row_cnt);
    int64_t* scan_ws = (int64_t*)(base + L.scan_ws);
    cudaStream_t st = ST(stream);

    NPB_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(int64_t), st));
    NPB_CUDA(cudaMemsetAsync(row_cnt, 0, 4 * ((size_t)nrows + 1), st));
    if (nnz > 0) {
        k_make_keys<<<npb_blocks(nnz, TPB), TPB, 0, st>>>(nnz, rows, cols, col_bits, keys_in);
        NPB_COUNT_LAUNCH();
        const int row_bits = bits_for(nrows);
        size_t cb = L.cub_bytes;
        NPB_CUDA(cub::DeviceRadixSort::SortPairs(base + L.cub, cb, keys_in, keys_out,
                                                 vals, svals, nnz, 0, col_bits + row_bits, st));
        NPB_COUNT_LAUNCH();
        k_reduce_runs<<<npb_blocks(nnz, TPB), TPB, 0, st>>>(nnz, keys_out, svals, prune, col_bits, flag, row_cnt);
        NPB_COUNT_LAUNCH();
        NPB_CHECK_LAUNCH("coo_reduce_runs");
        int rc = npb_scan_launch(flag, pos, nnz, scan_ws, stats, st);
        if (rc) return rc;
        k_max_cnt<<<npb_blocks(nrows, TPB), TPB, 0, st>>>(nrows, row_cnt, stats);
        NPB_COUNT_LAUNCH();
    }
    // indptr from the per-row survivor counts (total goes to a scratch cell so
    // stats[0] keeps the value written by the flag scan)
    int64_t* scratch_total = scan_ws + npb_scan_ws_elems(nnz > nrows ? nnz : nrows) - 1;
    int rc = npb_scan_launch(row_cnt, indptr_out, nrows, scan_ws, scratch_total, st);
    if (rc) return rc;
    NPB_CHECK_LAUNCH("coo_to_csr_symbolic");
    return NPB_OK;
}

// Phase 2: compact the surviving entries into the caller's arrays.
int npb_coo_to_csr_numeric(int64_t nrows, int64_t ncols, int64_t nnz, void* ws,
                           int32_t* indices_out, double* data_out, void* stream) {
    if (nnz <= 0) return NPB_OK;
    Layout L = make_layout(nrows, nnz);
    char* base = (char*)ws;
    k_compact<<<npb_blocks(nnz, TPB), TPB, 0, ST(stream)>>>(
        nnz, (const uint64_t*)(base + L.keys_out), (const double*)(base + L.vals),
        (const int32_t*)(base + L.flag), (const int32_t*)(base + L.pos), bits_for(ncols),
        indices_out, data_out);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("coo_compact");
    return NPB_OK;
}

}  // extern "C"
