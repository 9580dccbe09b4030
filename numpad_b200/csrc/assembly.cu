// Chain-rule accumulation kernels: the device replacements of the scipy
// _sparsetools calls behind the reference's _multiply_ops / _add_ops
// (adstate.py:282-297) and of csr_jac/dia_jac.tocsr (adstate.py:244-272).
//
// Everything here is HBM-bound integer/fp64 streaming work.  Matrices are
// canonical CSR (sorted, duplicate-free rows); local Jacobian blocks are never
// materialised as CSR when they are
//   * a scalar        -> folded into an npb_scales chain (no traffic),
//   * a diagonal      -> column / row scaling that shares the pattern,
//   * a selection     -> `map[row] = column or -1` (4 B per row): a product
//                        with it is a column relabel (adjoint order) or a row
//                        gather (tangent order).
// Two-pass kernels (`*_symbolic` counts rows + scans, `*_numeric` fills) keep
// the caller in charge of every allocation.
#include "common.cuh"
#include "scan.cuh"

namespace {

constexpr int TPB = 256;

// scipy's csr_matmat drops entries whose value is exactly 0.0 (scipy >= 1.1x:
// `if (sums[head] != 0)`), so every product that can create zeros reports how
// many it produced; the host compacts the result only when that count is > 0
// (it almost never is).  Must be reached by all 32 lanes of a warp.
__device__ __forceinline__ void count_zeros(bool is_zero, int64_t* zero_count) {
    const unsigned m = __ballot_sync(0xffffffffu, is_zero);
    if (m && (threadIdx.x & 31) == 0 && zero_count)
        atomicAdd((unsigned long long*)zero_count, (unsigned long long)__popc(m));
}

// ---------------------------------------------------------------- elementwise
__global__ void __launch_bounds__(TPB)
k_apply_scales(int64_t nnz, const double* __restrict__ in, npb_scales sc,
               double* __restrict__ out) {
    int64_t k = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (k < nnz) out[k] = npb_apply(sc, in[k]);
}

// C = A * diag(d): data_out[k] = chain(data[k]) * d[col[k]]   (csr x dia_jac)
__global__ void __launch_bounds__(TPB)
k_scale_cols(int64_t nnz, const int32_t* __restrict__ idx,
             const double* __restrict__ in, npb_scales sc,
             const double* __restrict__ d, double* __restrict__ out,
             int64_t* __restrict__ zero_count) {
    int64_t k = (int64_t)blockIdx.x * TPB + threadIdx.x;
    bool z = false;
    if (k < nnz) {
        const double v = __dmul_rn(npb_apply(sc, in[k]), __ldg(d + idx[k]));
        out[k] = v;
        z = (v == 0.0);
    }
    count_zeros(z, zero_count);
}

// C = diag(d) * A: data_out[k] = d[row(k)] * chain(data[k])   (dia_jac x csr)
// one 8-lane group per row keeps the accesses of a warp inside 4 rows
__global__ void __launch_bounds__(TPB)
k_scale_rows(int64_t nrows, const int32_t* __restrict__ ptr,
             const double* __restrict__ in, npb_scales sc,
             const double* __restrict__ d, double* __restrict__ out,
             int64_t* __restrict__ zero_count) {
    const int lane = threadIdx.x & 7;
    int64_t r = ((int64_t)blockIdx.x * TPB + threadIdx.x) >> 3;
    int nz = 0;
    if (r < nrows) {
        const double dr = d[r];
        for (int32_t k = ptr[r] + lane, e = ptr[r + 1]; k < e; k += 8) {
            const double v = __dmul_rn(dr, npb_apply(sc, in[k]));
            out[k] = v;
            nz += (v == 0.0);
        }
    }
    for (int dd = 16; dd; dd >>= 1) nz += __shfl_xor_sync(0xffffffffu, nz, dd);
    if (nz && (threadIdx.x & 31) == 0 && zero_count)
        atomicAdd((unsigned long long*)zero_count, (unsigned long long)nz);
}

// C = A * S with S a selection whose map is strictly increasing on the kept
// columns and keeps every column: same row extents, relabelled columns.
__global__ void __launch_bounds__(TPB)
k_relabel_same(int64_t nnz, const int32_t* __restrict__ idx,
               const double* __restrict__ in, npb_scales sc,
               const int32_t* __restrict__ map, const double* __restrict__ w,
               int32_t* __restrict__ idx_out, double* __restrict__ out,
               int64_t* __restrict__ zero_count) {
    int64_t k = (int64_t)blockIdx.x * TPB + threadIdx.x;
    bool z = false;
    if (k < nnz) {
        const int32_t c = idx[k];
        idx_out[k] = __ldg(map + c);
        if (w) {
            const double v = __dmul_rn(npb_apply(sc, in[k]), __ldg(w + c));
            out[k] = v;
            z = (v == 0.0);
        }
    }
    if (w) count_zeros(z, zero_count);
}

// ------------------------------------------------------ relabel with dropping
__global__ void __launch_bounds__(TPB)
k_relabel_count(int64_t nrows, const int32_t* __restrict__ ptr,
                const int32_t* __restrict__ idx, const int32_t* __restrict__ map,
                int32_t* __restrict__ cnt, int64_t* __restrict__ stats) {
    int64_t r = (int64_t)blockIdx.x * TPB + threadIdx.x;
    int32_t c = 0;
    if (r < nrows) {
        for (int32_t k = ptr[r], e = ptr[r + 1]; k < e; ++k)
            c += (__ldg(map + idx[k]) >= 0);
        cnt[r] = c;
    }
    // block max -> one atomic per warp
    for (int d = 16; d; d >>= 1) c = max(c, __shfl_xor_sync(0xffffffffu, c, d));
    if ((threadIdx.x & 31) == 0 && c > 0)
        atomicMax((unsigned long long*)(stats + 1), (unsigned long long)c);
}

__global__ void __launch_bounds__(TPB)
k_relabel_fill(int64_t nrows, const int32_t* __restrict__ ptr,
               const int32_t* __restrict__ idx, const double* __restrict__ in,
               npb_scales sc, const int32_t* __restrict__ map,
               const double* __restrict__ w, const int32_t* __restrict__ optr,
               int32_t* __restrict__ oidx, double* __restrict__ out,
               int64_t* __restrict__ zero_count) {
    int64_t r = (int64_t)blockIdx.x * TPB + threadIdx.x;
    int nz = 0;
    if (r < nrows) {
        int32_t o = optr[r];
        // optr may come from a RECORDED structure (assembly plan replay) whose operand has
        // deviated: stay inside the row's slots and leave no slot undefined (the deviation
        // itself is detected by the checked fills upstream; this only keeps later kernels
        // of the doomed replay on valid indices)
        const int32_t oe = optr[r + 1];
        for (int32_t k = ptr[r], e = ptr[r + 1]; k < e; ++k) {
            const int32_t c = idx[k];
            const int32_t m = __ldg(map + c);
            if (m >= 0) {
                double v = npb_apply(sc, in[k]);
                if (w) v = __dmul_rn(v, __ldg(w + c));
                if (o < oe) { oidx[o] = m; out[o] = v; }
                nz += (v == 0.0);
                ++o;
            }
        }
        for (; o < oe; ++o) { oidx[o] = 0; out[o] = 0.0; }
    }
    for (int dd = 16; dd; dd >>= 1) nz += __shfl_xor_sync(0xffffffffu, nz, dd);
    if (nz && (threadIdx.x & 31) == 0 && zero_count)
        atomicAdd((unsigned long long*)zero_count, (unsigned long long)nz);
}

// ------------------------------------------------- row gather (tangent order)
__global__ void __launch_bounds__(TPB)
k_gather_count(int64_t nrows, const int32_t* __restrict__ map,
               const int32_t* __restrict__ bptr, int32_t* __restrict__ cnt,
               int64_t* __restrict__ stats) {
    int64_t r = (int64_t)blockIdx.x * TPB + threadIdx.x;
    int32_t c = 0;
    if (r < nrows) {
        const int32_t m = map[r];
        if (m >= 0) c = bptr[m + 1] - bptr[m];
        cnt[r] = c;
    }
    for (int d = 16; d; d >>= 1) c = max(c, __shfl_xor_sync(0xffffffffu, c, d));
    if ((threadIdx.x & 31) == 0 && c > 0)
        atomicMax((unsigned long long*)(stats + 1), (unsigned long long)c);
}

__global__ void __launch_bounds__(TPB)
k_gather_fill(int64_t nrows, const int32_t* __restrict__ map,
              const double* __restrict__ w, const int32_t* __restrict__ bptr,
              const int32_t* __restrict__ bidx, const double* __restrict__ bdat,
              npb_scales sc, const int32_t* __restrict__ optr,
              int32_t* __restrict__ oidx, double* __restrict__ out,
              int64_t* __restrict__ zero_count) {
    const int lane = threadIdx.x & 7;
    int64_t r = ((int64_t)blockIdx.x * TPB + threadIdx.x) >> 3;
    int nz = 0;
    const int32_t m = (r < nrows) ? map[r] : -1;
    if (r < nrows) {
        // (bounded to the row's slots and no slot left undefined: optr may be a recorded
        // structure, see k_relabel_fill)
        const int32_t o = optr[r], slots = optr[r + 1] - o;
        const int32_t b0 = m >= 0 ? bptr[m] : 0;
        const int32_t len = m >= 0 ? min(bptr[m + 1] - b0, slots) : 0;
        const double wr = w ? w[r] : 1.0;
        for (int32_t k = lane; k < len; k += 8) {
            oidx[o + k] = bidx[b0 + k];
            double v = npb_apply(sc, bdat[b0 + k]);
            if (w) v = __dmul_rn(wr, v);
            out[o + k] = v;
            nz += (v == 0.0);
        }
        for (int32_t k = len + lane; k < slots; k += 8) { oidx[o + k] = 0; out[o + k] = 0.0; }
    }
    for (int dd = 16; dd; dd >>= 1) nz += __shfl_xor_sync(0xffffffffu, nz, dd);
    if (nz && (threadIdx.x & 31) == 0 && zero_count)
        atomicAdd((unsigned long long*)zero_count, (unsigned long long)nz);
}

// ------------------------------------------------- drop explicit zero entries
__global__ void __launch_bounds__(TPB)
k_prune_count(int64_t nrows, const int32_t* __restrict__ ptr,
              const double* __restrict__ dat, npb_scales sc,
              int32_t* __restrict__ cnt, int64_t* __restrict__ stats) {
    int64_t r = (int64_t)blockIdx.x * TPB + threadIdx.x;
    int32_t c = 0;
    if (r < nrows) {
        for (int32_t k = ptr[r], e = ptr[r + 1]; k < e; ++k) c += (npb_apply(sc, dat[k]) != 0.0);
        cnt[r] = c;
    }
    for (int d = 16; d; d >>= 1) c = max(c, __shfl_xor_sync(0xffffffffu, c, d));
    if ((threadIdx.x & 31) == 0 && c > 0)
        atomicMax((unsigned long long*)(stats + 1), (unsigned long long)c);
}

__global__ void __launch_bounds__(TPB)
k_prune_fill(int64_t nrows, const int32_t* __restrict__ ptr,
             const int32_t* __restrict__ idx, const double* __restrict__ dat,
             npb_scales sc, const int32_t* __restrict__ optr,
             int32_t* __restrict__ oidx, double* __restrict__ out,
             int64_t* __restrict__ mismatch) {
    int64_t r = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (r >= nrows) return;
    int32_t o = optr[r];
    // mismatch != NULL: optr comes from a RECORDED structure (assembly plan replay): never
    // write past the row's slots, report a row whose non-zero count differs
    const int32_t oe = mismatch ? optr[r + 1] : INT32_MAX;
    for (int32_t k = ptr[r], e = ptr[r + 1]; k < e; ++k) {
        const double v = npb_apply(sc, dat[k]);
        if (v != 0.0) {
            if (o < oe) { oidx[o] = idx[k]; out[o] = v; }
            ++o;
        }
    }
    if (mismatch && o != oe) {
        atomicAdd((unsigned long long*)mismatch, 1ull);
        // the replay is void; leave defined (valid) entries for the kernels that still run
        for (; o < oe; ++o) { oidx[o] = 0; out[o] = 0.0; }
    }
}

// ---------------------------------------------------------------------- SpAdd
// C = A + B on canonical rows, entries whose sum is exactly 0.0 are dropped
// when prune != 0 -- the rule of scipy's csr_plus_csr.
template <bool FILL>
__global__ void __launch_bounds__(TPB)
k_spadd(int64_t nrows, const int32_t* __restrict__ ap,
        const int32_t* __restrict__ ai, const double* __restrict__ ad,
        npb_scales as, const int32_t* __restrict__ bp,
        const int32_t* __restrict__ bi, const double* __restrict__ bd,
        npb_scales bs, int prune, int32_t* __restrict__ cnt,
        int64_t* __restrict__ stats, const int32_t* __restrict__ optr,
        int32_t* __restrict__ oidx, double* __restrict__ out,
        int64_t* __restrict__ mismatch) {
    int64_t r = (int64_t)blockIdx.x * TPB + threadIdx.x;
    int32_t c = 0;
    if (r < nrows) {
        int32_t i = ap[r], ie = ap[r + 1], j = bp[r], je = bp[r + 1];
        int32_t o = FILL ? optr[r] : 0;
        // mismatch != NULL: optr is a RECORDED structure (assembly plan replay)
        const int32_t oe = (FILL && mismatch) ? optr[r + 1] : INT32_MAX;
        int32_t ca = (i < ie) ? ai[i] : INT32_MAX;
        int32_t cb = (j < je) ? bi[j] : INT32_MAX;
        while (i < ie || j < je) {
            int32_t col;
            double v;
            if (ca < cb) {
                col = ca; v = npb_apply(as, ad[i]);
                ++i; ca = (i < ie) ? ai[i] : INT32_MAX;
            } else if (cb < ca) {
                col = cb; v = npb_apply(bs, bd[j]);
                ++j; cb = (j < je) ? bi[j] : INT32_MAX;
            } else {
                col = ca;
                v = __dadd_rn(npb_apply(as, ad[i]), npb_apply(bs, bd[j]));
                ++i; ca = (i < ie) ? ai[i] : INT32_MAX;
                ++j; cb = (j < je) ? bi[j] : INT32_MAX;
            }
            if (!prune || v != 0.0) {
                if (FILL) {
                    if (o < oe) { oidx[o] = col; out[o] = v; }
                    ++o;
                }
                ++c;
            }
        }
        if (!FILL) cnt[r] = c;
        if (FILL && mismatch && o != oe) {
            atomicAdd((unsigned long long*)mismatch, 1ull);
            // the replay is void; leave defined (valid) entries for the kernels that still run
            for (; o < oe; ++o) { oidx[o] = 0; out[o] = 0.0; }
        }
    }
    if (!FILL) {
        for (int d = 16; d; d >>= 1) c = max(c, __shfl_xor_sync(0xffffffffu, c, d));
        if ((threadIdx.x & 31) == 0 && c > 0)
            atomicMax((unsigned long long*)(stats + 1), (unsigned long long)c);
    }
}

// ------------------------------------------------- block -> CSR conversions
__global__ void __launch_bounds__(TPB)
k_sel_count(int64_t nrows, const int32_t* __restrict__ map, int32_t* __restrict__ cnt) {
    int64_t r = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (r < nrows) cnt[r] = map[r] >= 0;
}

__global__ void __launch_bounds__(TPB)
k_sel_fill(int64_t nrows, const int32_t* __restrict__ map,
           const double* __restrict__ w, const int32_t* __restrict__ optr,
           int32_t* __restrict__ oidx, double* __restrict__ out) {
    int64_t r = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (r >= nrows) return;
    const int32_t m = map[r];
    const int32_t o = optr[r], oe = optr[r + 1];     // (bounded: optr may be a recorded structure)
    if (o < oe) {
        oidx[o] = m >= 0 ? m : 0;
        out[o] = m >= 0 ? (w ? w[r] : 1.0) : 0.0;
    }
}

__global__ void __launch_bounds__(TPB)
k_diag_csr(int64_t n, const double* __restrict__ d, double scale,
           int32_t* __restrict__ ptr, int32_t* __restrict__ idx,
           double* __restrict__ out) {
    int64_t r = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (r < n) {
        ptr[r] = (int32_t)r;
        idx[r] = (int32_t)r;
        out[r] = d ? d[r] : scale;
    }
    if (r == n) ptr[n] = (int32_t)n;
}

// COO row ids of a CSR pattern (8 lanes per row)
__global__ void __launch_bounds__(TPB)
k_expand_rows(int64_t nrows, const int32_t* __restrict__ ptr, int32_t* __restrict__ rows) {
    const int lane = threadIdx.x & 7;
    int64_t r = ((int64_t)blockIdx.x * TPB + threadIdx.x) >> 3;
    if (r >= nrows) return;
    for (int32_t k = ptr[r] + lane, e = ptr[r + 1]; k < e; k += 8) rows[k] = (int32_t)r;
}

// ------------------------------------------- general product, expansion half
// per A entry: how many B entries it meets
__global__ void __launch_bounds__(TPB)
k_spgemm_count(int64_t annz, const int32_t* __restrict__ aidx,
               const int32_t* __restrict__ bptr, int32_t* __restrict__ cnt) {
    int64_t k = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (k < annz) {
        const int32_t c = aidx[k];
        cnt[k] = bptr[c + 1] - bptr[c];
    }
}

// writes the COO triplets (row of A entry, col of B entry, a*b); `off` is the
// exclusive scan of the counts above.  8 lanes per A entry.
__global__ void __launch_bounds__(TPB)
k_spgemm_expand(int64_t annz, const int32_t* __restrict__ arow,
                const int32_t* __restrict__ aidx, const double* __restrict__ adat,
                npb_scales as, const int32_t* __restrict__ bptr,
                const int32_t* __restrict__ bidx, const double* __restrict__ bdat,
                npb_scales bs, const int32_t* __restrict__ off,
                int32_t* __restrict__ orow, int32_t* __restrict__ ocol,
                double* __restrict__ oval) {
    const int lane = threadIdx.x & 7;
    int64_t k = ((int64_t)blockIdx.x * TPB + threadIdx.x) >> 3;
    if (k >= annz) return;
    const int32_t c = aidx[k], r = arow[k];
    const double a = npb_apply(as, adat[k]);
    const int32_t b0 = bptr[c], len = bptr[c + 1] - b0, o = off[k];
    for (int32_t t = lane; t < len; t += 8) {
        orow[o + t] = r;
        ocol[o + t] = bidx[b0 + t];
        oval[o + t] = __dmul_rn(a, npb_apply(bs, bdat[b0 + t]));
    }
}

// max row length of a pattern
__global__ void __launch_bounds__(TPB)
k_max_row(int64_t nrows, const int32_t* __restrict__ ptr, int64_t* __restrict__ stats) {
    int64_t r = (int64_t)blockIdx.x * TPB + threadIdx.x;
    int32_t c = (r < nrows) ? ptr[r + 1] - ptr[r] : 0;
    for (int d = 16; d; d >>= 1) c = max(c, __shfl_xor_sync(0xffffffffu, c, d));
    if ((threadIdx.x & 31) == 0 && c > 0)
        atomicMax((unsigned long long*)(stats + 1), (unsigned long long)c);
}

}  // namespace

#define ST(s) ((cudaStream_t)(s))

extern "C" {

int npb_csr_apply_scales(int64_t nnz, const double* in, const npb_scales* sc,
                         double* out, void* stream) {
    if (nnz <= 0) return NPB_OK;
    k_apply_scales<<<npb_blocks(nnz, TPB), TPB, 0, ST(stream)>>>(nnz, in, *sc, out);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("csr_apply_scales");
    return NPB_OK;
}

int npb_csr_scale_cols(int64_t nnz, const int32_t* indices, const double* in,
                       const npb_scales* sc, const double* d, double* out,
                       int64_t* zero_count, void* stream) {
    if (zero_count) NPB_CUDA(cudaMemsetAsync(zero_count, 0, sizeof(int64_t), ST(stream)));
    if (nnz <= 0) return NPB_OK;
    k_scale_cols<<<npb_blocks(nnz, TPB), TPB, 0, ST(stream)>>>(nnz, indices, in, *sc, d, out, zero_count);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("csr_scale_cols");
    return NPB_OK;
}

int npb_csr_scale_rows(int64_t nrows, const int32_t* indptr, const double* in,
                       const npb_scales* sc, const double* d, double* out,
                       int64_t* zero_count, void* stream) {
    if (zero_count) NPB_CUDA(cudaMemsetAsync(zero_count, 0, sizeof(int64_t), ST(stream)));
    if (nrows <= 0) return NPB_OK;
    k_scale_rows<<<npb_blocks(nrows * 8, TPB), TPB, 0, ST(stream)>>>(nrows, indptr, in, *sc, d, out, zero_count);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("csr_scale_rows");
    return NPB_OK;
}

int npb_csr_relabel_same(int64_t nnz, const int32_t* indices, const double* in,
                         const npb_scales* sc, const int32_t* map,
                         const double* w, int32_t* indices_out, double* out,
                         int64_t* zero_count, void* stream) {
    if (zero_count) NPB_CUDA(cudaMemsetAsync(zero_count, 0, sizeof(int64_t), ST(stream)));
    if (nnz <= 0) return NPB_OK;
    if (w && !out) return NPB_ERR_ARG;
    k_relabel_same<<<npb_blocks(nnz, TPB), TPB, 0, ST(stream)>>>(nnz, indices, in, *sc, map, w, indices_out, out, zero_count);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("csr_relabel_same");
    return NPB_OK;
}

// stats[0] <- total nnz of the result, stats[1] <- its longest row
int npb_csr_relabel_symbolic(int64_t nrows, const int32_t* indptr,
                             const int32_t* indices, const int32_t* map,
                             int32_t* counts, int32_t* indptr_out,
                             int64_t* scan_ws, int64_t* stats, void* stream) {
    NPB_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(int64_t), ST(stream)));
    k_relabel_count<<<npb_blocks(nrows, TPB), TPB, 0, ST(stream)>>>(nrows, indptr, indices, map, counts, stats);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("csr_relabel_count");
    return npb_scan_launch(counts, indptr_out, nrows, scan_ws, stats, ST(stream));
}

int npb_csr_relabel_numeric(int64_t nrows, const int32_t* indptr,
                            const int32_t* indices, const double* in,
                            const npb_scales* sc, const int32_t* map,
                            const double* w, const int32_t* indptr_out,
                            int32_t* indices_out, double* out,
                            int64_t* zero_count, void* stream) {
    if (zero_count) NPB_CUDA(cudaMemsetAsync(zero_count, 0, sizeof(int64_t), ST(stream)));
    if (nrows <= 0) return NPB_OK;
    k_relabel_fill<<<npb_blocks(nrows, TPB), TPB, 0, ST(stream)>>>(nrows, indptr, indices, in, *sc, map, w, indptr_out, indices_out, out, zero_count);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("csr_relabel_fill");
    return NPB_OK;
}

int npb_csr_gather_rows_symbolic(int64_t nrows_out, const int32_t* map,
                                 const int32_t* b_indptr, int32_t* counts,
                                 int32_t* indptr_out, int64_t* scan_ws,
                                 int64_t* stats, void* stream) {
    NPB_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(int64_t), ST(stream)));
    k_gather_count<<<npb_blocks(nrows_out, TPB), TPB, 0, ST(stream)>>>(nrows_out, map, b_indptr, counts, stats);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("csr_gather_count");
    return npb_scan_launch(counts, indptr_out, nrows_out, scan_ws, stats, ST(stream));
}

int npb_csr_gather_rows_numeric(int64_t nrows_out, const int32_t* map,
                                const double* w, const int32_t* b_indptr,
                                const int32_t* b_indices, const double* b_data,
                                const npb_scales* sc, const int32_t* indptr_out,
                                int32_t* indices_out, double* out,
                                int64_t* zero_count, void* stream) {
    if (zero_count) NPB_CUDA(cudaMemsetAsync(zero_count, 0, sizeof(int64_t), ST(stream)));
    if (nrows_out <= 0) return NPB_OK;
    k_gather_fill<<<npb_blocks(nrows_out * 8, TPB), TPB, 0, ST(stream)>>>(nrows_out, map, w, b_indptr, b_indices, b_data, *sc, indptr_out, indices_out, out, zero_count);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("csr_gather_fill");
    return NPB_OK;
}

// drop the entries whose (scaled) value is exactly 0.0
int npb_csr_prune_symbolic(int64_t nrows, const int32_t* indptr, const double* data,
                           const npb_scales* sc, int32_t* counts, int32_t* indptr_out,
                           int64_t* scan_ws, int64_t* stats, void* stream) {
    NPB_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(int64_t), ST(stream)));
    k_prune_count<<<npb_blocks(nrows, TPB), TPB, 0, ST(stream)>>>(nrows, indptr, data, *sc, counts, stats);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("csr_prune_count");
    return npb_scan_launch(counts, indptr_out, nrows, scan_ws, stats, ST(stream));
}

int npb_csr_prune_numeric(int64_t nrows, const int32_t* indptr, const int32_t* indices,
                          const double* data, const npb_scales* sc,
                          const int32_t* indptr_out, int32_t* indices_out, double* out,
                          void* stream) {
    if (nrows <= 0) return NPB_OK;
    k_prune_fill<<<npb_blocks(nrows, TPB), TPB, 0, ST(stream)>>>(nrows, indptr, indices, data, *sc, indptr_out, indices_out, out, nullptr);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("csr_prune_fill");
    return NPB_OK;
}

// the same into a RECORDED structure (replay of an assembly plan): rows are written inside
// their recorded slots only and *mismatch (device, reset here) counts the rows whose number
// of non-zeros differs from the recording
int npb_csr_prune_numeric_checked(int64_t nrows, const int32_t* indptr, const int32_t* indices,
                                  const double* data, const npb_scales* sc,
                                  const int32_t* indptr_out, int32_t* indices_out, double* out,
                                  int64_t* mismatch, void* stream) {
    if (!mismatch) return NPB_ERR_ARG;
    NPB_CUDA(cudaMemsetAsync(mismatch, 0, sizeof(int64_t), ST(stream)));
    if (nrows <= 0) return NPB_OK;
    k_prune_fill<<<npb_blocks(nrows, TPB), TPB, 0, ST(stream)>>>(nrows, indptr, indices, data, *sc, indptr_out, indices_out, out, mismatch);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("csr_prune_fill_checked");
    return NPB_OK;
}

int npb_spadd_symbolic(int64_t nrows, const int32_t* a_indptr,
                       const int32_t* a_indices, const double* a_data,
                       const npb_scales* a_sc, const int32_t* b_indptr,
                       const int32_t* b_indices, const double* b_data,
                       const npb_scales* b_sc, int prune, int32_t* counts,
                       int32_t* indptr_out, int64_t* scan_ws, int64_t* stats,
                       void* stream) {
    NPB_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(int64_t), ST(stream)));
    k_spadd<false><<<npb_blocks(nrows, TPB), TPB, 0, ST(stream)>>>(
        nrows, a_indptr, a_indices, a_data, *a_sc, b_indptr, b_indices, b_data,
        *b_sc, prune, counts, stats, nullptr, nullptr, nullptr, nullptr);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("spadd_count");
    return npb_scan_launch(counts, indptr_out, nrows, scan_ws, stats, ST(stream));
}

int npb_spadd_numeric(int64_t nrows, const int32_t* a_indptr,
                      const int32_t* a_indices, const double* a_data,
                      const npb_scales* a_sc, const int32_t* b_indptr,
                      const int32_t* b_indices, const double* b_data,
                      const npb_scales* b_sc, int prune,
                      const int32_t* indptr_out, int32_t* indices_out,
                      double* out, void* stream) {
    if (nrows <= 0) return NPB_OK;
    k_spadd<true><<<npb_blocks(nrows, TPB), TPB, 0, ST(stream)>>>(
        nrows, a_indptr, a_indices, a_data, *a_sc, b_indptr, b_indices, b_data,
        *b_sc, prune, nullptr, nullptr, indptr_out, indices_out, out, nullptr);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("spadd_fill");
    return NPB_OK;
}

// fill of a RECORDED structure (indptr_out from the symbolic phase of an earlier sum with the
// same operand patterns, see _dev.AssemblyPlan): rows are written inside their recorded slots
// only and *mismatch (device, reset here) counts the rows whose number of entries differs --
// a sum that cancels exactly now but did not then, or the reverse
int npb_spadd_numeric_checked(int64_t nrows, const int32_t* a_indptr,
                              const int32_t* a_indices, const double* a_data,
                              const npb_scales* a_sc, const int32_t* b_indptr,
                              const int32_t* b_indices, const double* b_data,
                              const npb_scales* b_sc, int prune, const int32_t* indptr_out,
                              int32_t* indices_out, double* out, int64_t* mismatch,
                              void* stream) {
    if (!mismatch) return NPB_ERR_ARG;
    NPB_CUDA(cudaMemsetAsync(mismatch, 0, sizeof(int64_t), ST(stream)));
    if (nrows <= 0) return NPB_OK;
    k_spadd<true><<<npb_blocks(nrows, TPB), TPB, 0, ST(stream)>>>(
        nrows, a_indptr, a_indices, a_data, *a_sc, b_indptr, b_indices, b_data,
        *b_sc, prune, nullptr, nullptr, indptr_out, indices_out, out, mismatch);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("spadd_fill_checked");
    return NPB_OK;
}

int npb_sel_to_csr_symbolic(int64_t nrows, const int32_t* map, int32_t* counts,
                            int32_t* indptr_out, int64_t* scan_ws,
                            int64_t* stats, void* stream) {
    NPB_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(int64_t), ST(stream)));
    k_sel_count<<<npb_blocks(nrows, TPB), TPB, 0, ST(stream)>>>(nrows, map, counts);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("sel_count");
    return npb_scan_launch(counts, indptr_out, nrows, scan_ws, stats, ST(stream));
}

int npb_sel_to_csr_numeric(int64_t nrows, const int32_t* map, const double* w,
                           const int32_t* indptr_out, int32_t* indices_out,
                           double* out, void* stream) {
    if (nrows <= 0) return NPB_OK;
    k_sel_fill<<<npb_blocks(nrows, TPB), TPB, 0, ST(stream)>>>(nrows, map, w, indptr_out, indices_out, out);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("sel_fill");
    return NPB_OK;
}

// d == NULL: scale * identity
int npb_diag_to_csr(int64_t n, const double* d, double scale, int32_t* indptr,
                    int32_t* indices, double* data, void* stream) {
    k_diag_csr<<<npb_blocks(n + 1, TPB), TPB, 0, ST(stream)>>>(n, d, scale, indptr, indices, data);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("diag_to_csr");
    return NPB_OK;
}

int npb_csr_expand_rows(int64_t nrows, const int32_t* indptr, int32_t* rows,
                        void* stream) {
    if (nrows <= 0) return NPB_OK;
    k_expand_rows<<<npb_blocks(nrows * 8, TPB), TPB, 0, ST(stream)>>>(nrows, indptr, rows);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("csr_expand_rows");
    return NPB_OK;
}

int npb_csr_max_row(int64_t nrows, const int32_t* indptr, int64_t* stats,
                    void* stream) {
    NPB_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(int64_t), ST(stream)));
    if (nrows <= 0) return NPB_OK;
    k_max_row<<<npb_blocks(nrows, TPB), TPB, 0, ST(stream)>>>(nrows, indptr, stats);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("csr_max_row");
    return NPB_OK;
}

// general product C = A*B, expansion into COO (the caller then runs
// npb_coo_to_csr): offsets[a_nnz+1] <- exclusive scan of per-entry counts
int npb_spgemm_expand_symbolic(int64_t a_nnz, const int32_t* a_indices,
                               const int32_t* b_indptr, int32_t* counts,
                               int32_t* offsets, int64_t* scan_ws,
                               int64_t* stats, void* stream) {
    NPB_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(int64_t), ST(stream)));
    if (a_nnz > 0) {
        k_spgemm_count<<<npb_blocks(a_nnz, TPB), TPB, 0, ST(stream)>>>(a_nnz, a_indices, b_indptr, counts);
        NPB_COUNT_LAUNCH();
        NPB_CHECK_LAUNCH("spgemm_count");
    }
    return npb_scan_launch(counts, offsets, a_nnz, scan_ws, stats, ST(stream));
}

int npb_spgemm_expand_numeric(int64_t a_nnz, const int32_t* a_rows,
                              const int32_t* a_indices, const double* a_data,
                              const npb_scales* a_sc, const int32_t* b_indptr,
                              const int32_t* b_indices, const double* b_data,
                              const npb_scales* b_sc, const int32_t* offsets,
                              int32_t* out_rows, int32_t* out_cols,
                              double* out_vals, void* stream) {
    if (a_nnz <= 0) return NPB_OK;
    k_spgemm_expand<<<npb_blocks(a_nnz * 8, TPB), TPB, 0, ST(stream)>>>(
        a_nnz, a_rows, a_indices, a_data, *a_sc, b_indptr, b_indices, b_data,
        *b_sc, offsets, out_rows, out_cols, out_vals);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("spgemm_expand");
    return NPB_OK;
}

int64_t npb_scan_workspace_elems(int64_t n) { return npb_scan_ws_elems(n); }

}  // extern "C"
