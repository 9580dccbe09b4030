// Row-batched general SpGEMM  C = A B  with shared-memory staging.
//
// Role: the general `csr_matmat` of the chain rule (reference adstate.py:290-297 ->
// scipy _sparsetools csr_matmat) and the Galerkin products R (A P) of the multigrid setup.
// The first version expanded every partial product to HBM and ran a device-wide radix
// sort over all of them (coo.cu) -- ~8 passes over 16 B per product.  Here consecutive
// rows of A are batched so that a batch's partial products (<= CAP) fit in shared memory:
// one CTA expands them there, sorts them by (row, column) with a block radix sort (stable:
// duplicates stay in generation order, so sums are deterministic and ordered like the
// expansion), adds the runs and writes only the unique entries -- one pass over A, the
// referenced rows of B, and C.
//
// Algorithmic bytes: bytes(A) + (rows of B as referenced, mostly L2 hits) + bytes(C).
// Rows whose products do not fit (ub > CAP / 2) make the caller fall back to coo.cu.
#include <cub/block/block_radix_sort.cuh>
#include <cub/block/block_reduce.cuh>
#include <cub/block/block_scan.cuh>

#include "common.cuh"
#include "scan.cuh"

namespace {

constexpr int SG_THREADS = 256;
constexpr int SG_ITEMS = 8;
constexpr int SG_CAP = SG_THREADS * SG_ITEMS;       // partial products per batch
#ifndef SG_RADIX
#define SG_RADIX 4
#endif
// digits of SG_RADIX bits (6-bit digits need fewer passes over the ~18-20 significant key bits but a
// larger rank storage: measured 1.2-1.7x SLOWER on the Galerkin products of the 2048^2 cavity)
typedef cub::BlockRadixSort<unsigned long long, SG_THREADS, SG_ITEMS, double, SG_RADIX> sg_sort_t;

// ub[i] = max(1, sum over entries (i,k) of max(1, nnz(B_k))): products, entries and rows of a
// batch are all bounded by the sum of ub over its rows
__global__ void __launch_bounds__(256)
k_sg_ub(int64_t m, const int32_t* __restrict__ aptr, const int32_t* __restrict__ aidx,
        const int32_t* __restrict__ bptr, int32_t* __restrict__ ub, int64_t* __restrict__ stats) {
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    int32_t u = 0;
    if (i < m) {
        int64_t s = 0;
        for (int32_t e = aptr[i], e1 = aptr[i + 1]; e < e1; ++e) {
            const int32_t k = aidx[e];
            const int32_t l = bptr[k + 1] - bptr[k];
            s += l > 1 ? l : 1;
        }
        u = (int32_t)(s < 1 ? 1 : (s > 0x3fffffff ? 0x3fffffff : s));
        ub[i] = u;
    }
    // stats[1] = largest ub (the scan fills stats[0])
    for (int d = 16; d; d >>= 1) u = max(u, __shfl_xor_sync(0xffffffffu, u, d));
    if ((threadIdx.x & 31) == 0 && u > 0) atomicMax((long long*)(stats + 1), (long long)u);
}

__global__ void __launch_bounds__(256)
k_sg_max(int64_t m, const int32_t* __restrict__ cnt, int64_t* __restrict__ stats) {
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    int32_t u = i < m ? cnt[i] : 0;
    for (int d = 16; d; d >>= 1) u = max(u, __shfl_xor_sync(0xffffffffu, u, d));
    if ((threadIdx.x & 31) == 0 && u > 0) atomicMax((long long*)(stats + 1), (long long)u);
}

// batch b = rows [first r with ubptr[r] >= b T, first r with ubptr[r] >= (b+1) T)
__global__ void __launch_bounds__(256)
k_sg_batches(int64_t m, const int32_t* __restrict__ ubptr, int64_t T, int64_t nb,
             int32_t* __restrict__ brow) {
    int64_t b = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (b > nb) return;
    if (b == nb) { brow[b] = (int32_t)m; return; }
    const int64_t target = b * T;
    int64_t lo = 0, hi = m;
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        if ((int64_t)ubptr[mid] >= target) hi = mid; else lo = mid + 1;
    }
    brow[b] = (int32_t)lo;
}

struct sg_smem {
    union {
        typename sg_sort_t::TempStorage sort;
        typename cub::BlockScan<int, SG_THREADS>::TempStorage scan;
        typename cub::BlockReduce<int, SG_THREADS>::TempStorage red;
    } tmp;
    int cmin, cmax;
    unsigned long long key[SG_CAP + 1];
    double val[SG_CAP];
    int32_t rptr[SG_CAP + 1];
};

__global__ void __launch_bounds__(SG_THREADS)
k_sg_numeric(const int32_t* __restrict__ brow, const int32_t* __restrict__ aptr,
             const int32_t* __restrict__ aidx, const double* __restrict__ adat, npb_scales sa,
             const int32_t* __restrict__ bptr, const int32_t* __restrict__ bidx,
             const double* __restrict__ bdat, npb_scales sb, const int32_t* __restrict__ ubptr,
             int key_bits, int prune, int32_t* __restrict__ tidx, double* __restrict__ tval,
             int32_t* __restrict__ cnt, int32_t* __restrict__ tstart) {
    extern __shared__ __align__(16) unsigned char sg_raw[];
    sg_smem& S = *reinterpret_cast<sg_smem*>(sg_raw);
    const int tid = threadIdx.x;
    const int32_t r0 = brow[blockIdx.x], r1 = brow[blockIdx.x + 1];
    const int nr = r1 - r0;
    if (nr <= 0) return;
    for (int i = tid; i <= nr; i += SG_THREADS) S.rptr[i] = aptr[r0 + i];
    __syncthreads();
    const int32_t e0 = S.rptr[0], ne = S.rptr[nr] - e0;       // <= SG_CAP by construction
    // 1. products per entry, exclusive offsets (blocked arrangement: entry tid*ITEMS + j)
    int len[SG_ITEMS], off[SG_ITEMS];
#pragma unroll
    for (int j = 0; j < SG_ITEMS; ++j) {
        const int e = tid * SG_ITEMS + j;
        len[j] = 0;
        if (e < ne) { const int32_t k = aidx[e0 + e]; len[j] = bptr[k + 1] - bptr[k]; }
    }
    int total;
    cub::BlockScan<int, SG_THREADS>(S.tmp.scan).ExclusiveSum(len, off, total);
    __syncthreads();
    for (int p = total + tid; p < SG_CAP + 1; p += SG_THREADS) S.key[p] = ~0ull;
    // 2. expansion into shared memory: key = (row - r0) << 32 | column
    int cmin = 0x7fffffff, cmax = 0;
#pragma unroll
    for (int j = 0; j < SG_ITEMS; ++j) {
        const int e = tid * SG_ITEMS + j;
        if (e >= ne || len[j] == 0) continue;
        int lo = 0, hi = nr;                       // row of entry e: last i with rptr[i] - e0 <= e
        while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (S.rptr[mid] - e0 <= e) lo = mid; else hi = mid; }
        const unsigned long long rk = (unsigned long long)lo << 32;
        const int32_t k = aidx[e0 + e];
        const double a = npb_apply(sa, adat[e0 + e]);
        const int32_t b0 = bptr[k];
        for (int t = 0; t < len[j]; ++t) {
            const int32_t c = bidx[b0 + t];
            cmin = min(cmin, c);
            cmax = max(cmax, c);
            S.key[off[j] + t] = rk | (unsigned long long)(uint32_t)c;
            S.val[off[j] + t] = __dmul_rn(a, npb_apply(sb, bdat[b0 + t]));
        }
    }
    // 3. stable sort by (row, column).  Keys are compressed to (row << cb) | (column - cmin)
    // with cb = bits of the batch's column range: a batch of a stencil-like product spans a
    // few thousand columns, so the radix sort runs over ~20 bits instead of 44
    cmin = cub::BlockReduce<int, SG_THREADS>(S.tmp.red).Reduce(cmin, cub::Min());
    __syncthreads();
    cmax = cub::BlockReduce<int, SG_THREADS>(S.tmp.red).Reduce(cmax, cub::Max());
    if (tid == 0) { S.cmin = cmin; S.cmax = cmax; }
    __syncthreads();
    cmin = S.cmin;
    const unsigned span = total > 0 ? (unsigned)(S.cmax - cmin) : 0u;
    const int cb = span ? 32 - __clz(span) : 1;
    const int rbits = 32 - __clz((unsigned)nr);            // rows 0 .. nr-1, padding row nr
    const unsigned long long cmask = (1ull << cb) - 1ull;
    (void)key_bits;
    unsigned long long k8[SG_ITEMS];
    double v8[SG_ITEMS];
#pragma unroll
    for (int j = 0; j < SG_ITEMS; ++j) {
        const int p = tid * SG_ITEMS + j;
        const unsigned long long k = S.key[p];
        k8[j] = p < total ? (((k >> 32) << cb) | (unsigned long long)((uint32_t)k - (uint32_t)cmin))
                          : ((unsigned long long)nr << cb);
        v8[j] = S.val[p];
    }
    __syncthreads();
    sg_sort_t(S.tmp.sort).Sort(k8, v8, 0, cb + rbits);
    __syncthreads();
#pragma unroll
    for (int j = 0; j < SG_ITEMS; ++j) {
        // back to (row << 32) | column for the run sums below
        k8[j] = ((k8[j] >> cb) << 32) | (unsigned long long)((uint32_t)(k8[j] & cmask) + (uint32_t)cmin);
        S.key[tid * SG_ITEMS + j] = k8[j];
        S.val[tid * SG_ITEMS + j] = v8[j];
    }
    __syncthreads();
    // 4. heads of equal-key runs: sum the run in order, decide survival, rank the survivors
    double sum[SG_ITEMS];
    int keep[SG_ITEMS], nkeep = 0;
#pragma unroll
    for (int j = 0; j < SG_ITEMS; ++j) {
        const int p = tid * SG_ITEMS + j;
        keep[j] = 0;
        sum[j] = 0.0;
        if (p < total && (p == 0 || S.key[p - 1] != k8[j])) {
            double s = v8[j];
            for (int q = p + 1; q < total && S.key[q] == k8[j]; ++q) s = __dadd_rn(s, S.val[q]);
            sum[j] = s;
            keep[j] = (!prune || s != 0.0) ? 1 : 0;
            nkeep += keep[j];
        }
    }
    int base;
    cub::BlockScan<int, SG_THREADS>(S.tmp.scan).ExclusiveSum(nkeep, base);
    const int64_t t0 = ubptr[r0];
#pragma unroll
    for (int j = 0; j < SG_ITEMS; ++j) {
        if (!keep[j]) continue;
        const int32_t row = r0 + (int32_t)(k8[j] >> 32);
        tidx[t0 + base] = (int32_t)(uint32_t)k8[j];
        tval[t0 + base] = sum[j];
        // first kept entry of its row records where the row starts in the scratch arrays:
        // the smallest position among the row's survivors
        atomicMin(tstart + row, (int32_t)(t0 + base));
        atomicAdd(cnt + row, 1);
        ++base;
    }
}

// ---- one warp per row -----------------------------------------------------------------
// The Galerkin products of the multigrid setup (R (A P)) and most products of the chain rule have
// SHORT rows: 20..150 partial products.  A CTA-wide radix sort of 2048 of them (above) costs ~10
// sort passes per batch; here every warp owns one row at a time: the products of the row are
// expanded into the warp's slice of shared memory (lane e takes entry e of the A row: one round
// of dependent loads for all entries), the keys (column, position) are sorted in REGISTERS with a
// bitonic network of shuffles (CAP / 32 keys per lane; the position makes the keys unique, so the
// order of equal columns is the generation order, as in the stable radix sort), runs are summed in
// that order by the lane that holds their head, survivors ranked with warp ballots.  Same output contract as
// k_sg_numeric (tidx / tval at ubptr[row], cnt, tstart), bit-identical results.
// ascending bitonic sort of 32 * K unique keys held by a warp, element e = r * 32 + lane in register
// r: partners less than 32 apart are exchanged with shuffles, the others are registers of the same
// lane (a first version sorted in shared memory: ~0.3 us per compare-exchange round, 118 us for a
// 512-key row -- every round a dependent load / store pair with two-way bank conflicts)
template <int K>
__device__ __forceinline__ void warp_bitonic_sort(unsigned long long (&x)[K], int lane) {
#pragma unroll
    for (int k = 2; k <= 32 * K; k <<= 1) {
#pragma unroll
        for (int j = k >> 1; j > 0; j >>= 1) {
            if (j >= 32) {
                const int rj = j >> 5;
#pragma unroll
                for (int r = 0; r < K; ++r) {
                    const int rq = r ^ rj;
                    if (rq > r) {
                        const bool up = (((r * 32 + lane) & k) == 0);
                        const unsigned long long a = x[r], b = x[rq];
                        if ((a > b) == up) { x[r] = b; x[rq] = a; }
                    }
                }
            } else {
#pragma unroll
                for (int r = 0; r < K; ++r) {
                    const unsigned long long other = __shfl_xor_sync(0xffffffffu, x[r], j);
                    const bool up = (((r * 32 + lane) & k) == 0);
                    const bool lower = (lane & j) == 0;
                    const unsigned long long lo = x[r] < other ? x[r] : other;
                    const unsigned long long hi = x[r] < other ? other : x[r];
                    x[r] = (lower == up) ? lo : hi;
                }
            }
        }
    }
}

template <int CAP>
__global__ void __launch_bounds__(256)
k_sg_rows_warp(int64_t m, const int32_t* __restrict__ aptr, const int32_t* __restrict__ aidx,
               const double* __restrict__ adat, npb_scales sa, const int32_t* __restrict__ bptr,
               const int32_t* __restrict__ bidx, const double* __restrict__ bdat, npb_scales sb,
               const int32_t* __restrict__ ubptr, int prune, int32_t* __restrict__ tidx,
               double* __restrict__ tval, int32_t* __restrict__ cnt, int32_t* __restrict__ tstart) {
    extern __shared__ __align__(16) unsigned char sgw_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long* key = reinterpret_cast<unsigned long long*>(sgw_raw) + (size_t)warp * CAP;
    double* val = reinterpret_cast<double*>(sgw_raw + (size_t)8 * CAP * 8) + (size_t)warp * CAP;
    const int64_t nwarps = (int64_t)gridDim.x * 8;
    for (int64_t row = (int64_t)blockIdx.x * 8 + warp; row < m; row += nwarps) {
        const int32_t e0 = aptr[row], ne = aptr[row + 1] - e0;
        // 1. expansion
        int count = 0;
        for (int c0 = 0; c0 < ne; c0 += 32) {
            const int e = c0 + lane;
            int32_t b0 = 0, len = 0;
            double a = 0.0;
            if (e < ne) {
                const int32_t k = aidx[e0 + e];
                a = npb_apply(sa, adat[e0 + e]);
                b0 = bptr[k];
                len = bptr[k + 1] - b0;
            }
            int off = len;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const int t = __shfl_up_sync(0xffffffffu, off, d);
                if (lane >= d) off += t;
            }
            const int chunk_total = __shfl_sync(0xffffffffu, off, 31);
            off = count + off - len;
            for (int t = 0; t < len; ++t) {
                const int p = off + t;
                key[p] = ((unsigned long long)(uint32_t)bidx[b0 + t] << 16) | (unsigned long long)p;
                val[p] = __dmul_rn(a, npb_apply(sb, bdat[b0 + t]));
            }
            count += chunk_total;
        }
        __syncwarp();
        // 2. sort the keys in registers (the values stay where the expansion put them: the low 16
        // bits of a key are the position of its value)
        {
            constexpr int K = CAP / 32;
            unsigned long long x[K];
#pragma unroll
            for (int r = 0; r < K; ++r) {
                const int p = r * 32 + lane;
                x[r] = p < count ? key[p] : ~0ull;
            }
            warp_bitonic_sort<K>(x, lane);
#pragma unroll
            for (int r = 0; r < K; ++r) key[r * 32 + lane] = x[r];
        }
        __syncwarp();
        // 3. run sums by the lane that holds the head, survivors ranked in column order
        const int64_t t0 = ubptr[row];
        int kept = 0;
        for (int p0 = 0; p0 < count; p0 += 32) {
            const int p = p0 + lane;
            int keep = 0;
            double s = 0.0;
            uint32_t col = 0;
            if (p < count) {
                col = (uint32_t)(key[p] >> 16);
                if (p == 0 || (uint32_t)(key[p - 1] >> 16) != col) {
                    s = val[(int)(key[p] & 0xffffull)];
                    for (int q = p + 1; q < count && (uint32_t)(key[q] >> 16) == col; ++q)
                        s = __dadd_rn(s, val[(int)(key[q] & 0xffffull)]);
                    keep = (!prune || s != 0.0) ? 1 : 0;
                }
            }
            const unsigned mask = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const int r = kept + __popc(mask & ((1u << lane) - 1u));
                tidx[t0 + r] = (int32_t)col;
                tval[t0 + r] = s;
            }
            kept += __popc(mask);
        }
        if (lane == 0) { cnt[row] = kept; tstart[row] = (int32_t)t0; }
        __syncwarp();
    }
}

__global__ void __launch_bounds__(256)
k_sg_compact(int64_t m, const int32_t* __restrict__ cnt, const int32_t* __restrict__ tstart,
             const int32_t* __restrict__ cptr, const int32_t* __restrict__ tidx,
             const double* __restrict__ tval, int32_t* __restrict__ cidx, double* __restrict__ cdat) {
    // 8 lanes per row
    const int64_t r = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 3;
    const int lane = threadIdx.x & 7;
    if (r >= m) return;
    const int32_t n = cnt[r], s = tstart[r], d = cptr[r];
    for (int32_t k = lane; k < n; k += 8) { cidx[d + k] = tidx[s + k]; cdat[d + k] = tval[s + k]; }
}

__global__ void __launch_bounds__(256)
k_sg_init(int64_t m, int32_t* __restrict__ cnt, int32_t* __restrict__ tstart) {
    int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i < m) { cnt[i] = 0; tstart[i] = 0x7fffffff; }
}

// out row r = in row r with every column c replaced by map[c], re-sorted by the new column
// (insertion sort by one thread per row: for the short rows of a symmetric permutation
// P A P^T, whose rows are almost sorted already).  map must be injective on each row.
__global__ void __launch_bounds__(256)
k_relabel_sort_rows(int64_t m, const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                    const double* __restrict__ dat, const int32_t* __restrict__ map,
                    int32_t* __restrict__ oidx, double* __restrict__ odat) {
    int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (r >= m) return;
    const int32_t s = ptr[r], e = ptr[r + 1];
    for (int32_t k = s; k < e; ++k) {
        const int32_t c = map[idx[k]];
        const double v = dat[k];
        int32_t p = k;
        while (p > s && oidx[p - 1] > c) { oidx[p] = oidx[p - 1]; odat[p] = odat[p - 1]; --p; }
        oidx[p] = c;
        odat[p] = v;
    }
}

// row r of A S for a selection S given as a column map (map[c] < 0: column dropped) with
// optional weights w[c]: columns relabelled, the row re-sorted (stable insertion sort, one
// thread per row), equal columns summed in input order, exact zeros dropped when prune.
// Works in the scratch segment [ptr[r], ptr[r+1]) of tidx / tval; cnt[r] = entries kept.
__global__ void __launch_bounds__(256)
k_relabel_merge_rows(int64_t m, const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                     const double* __restrict__ dat, npb_scales sc, const int32_t* __restrict__ map,
                     const double* __restrict__ w, int prune, int32_t* __restrict__ tidx,
                     double* __restrict__ tval, int32_t* __restrict__ cnt,
                     int32_t* __restrict__ tstart) {
    int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (r >= m) return;
    const int32_t s = ptr[r], e = ptr[r + 1];
    int32_t n = 0;
    for (int32_t k = s; k < e; ++k) {
        const int32_t c0 = idx[k];
        const int32_t c = map[c0];
        if (c < 0) continue;
        double v = npb_apply(sc, dat[k]);
        if (w) v = __dmul_rn(v, w[c0]);
        int32_t p = s + n;
        while (p > s && tidx[p - 1] > c) { tidx[p] = tidx[p - 1]; tval[p] = tval[p - 1]; --p; }
        tidx[p] = c;
        tval[p] = v;
        ++n;
    }
    int32_t o = s;
    for (int32_t k = s; k < s + n;) {
        const int32_t c = tidx[k];
        double sum = tval[k];
        int32_t q = k + 1;
        for (; q < s + n && tidx[q] == c; ++q) sum = __dadd_rn(sum, tval[q]);
        if (!prune || sum != 0.0) { tidx[o] = c; tval[o] = sum; ++o; }
        k = q;
    }
    cnt[r] = o - s;
    tstart[r] = s;
}

}  // namespace

#define ST(s) ((cudaStream_t)(s))

extern "C" {

int npb_spgemm_rows_cap(void) { return SG_CAP; }

// tuning / test switch: 0 keeps rows of <= 512 partial products on the CTA-batched kernel too;
// returns the old value
static int npb_sg_warp_rows = 1;
int npb_spgemm_set_warp_rows(int on) {
    const int old = npb_sg_warp_rows;
    npb_sg_warp_rows = on ? 1 : 0;
    return old;
}

// phase 1: ub (m), ubptr (m + 1) = exclusive scan; stats[0] = total products bound,
// stats[1] = largest ub.  The caller checks stats[1] <= npb_spgemm_rows_cap() / 2.
int npb_spgemm_rows_bound(int64_t m, const int32_t* a_indptr, const int32_t* a_indices,
                          const int32_t* b_indptr, int32_t* ub, int32_t* ubptr,
                          int64_t* scan_ws, int64_t* stats, void* stream) {
    if (m <= 0) return NPB_ERR_ARG;
    NPB_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(int64_t), ST(stream)));
    k_sg_ub<<<npb_blocks(m, 256), 256, 0, ST(stream)>>>(m, a_indptr, a_indices, b_indptr, ub, stats);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("spgemm_rows_ub");
    return npb_scan_launch(ub, ubptr, m, scan_ws, stats, ST(stream));
}

// phase 2: batches of rows (brow: nb + 1 entries, nb = total / T + 1 with T = cap - max_ub),
// expansion + sort + run sums in shared memory; unique entries land in tidx / tval (`total`
// entries of scratch) and cnt / tstart (m) describe them per row.  key_bits = 32 + bits
// needed for the largest number of rows in a batch (<= cap).
int npb_spgemm_rows_numeric(int64_t m, int64_t total, int64_t max_ub, const int32_t* a_indptr,
                            const int32_t* a_indices, const double* a_data,
                            const npb_scales* a_scales, const int32_t* b_indptr,
                            const int32_t* b_indices, const double* b_data,
                            const npb_scales* b_scales, const int32_t* ubptr, int prune,
                            int32_t* brow, int32_t* tidx, double* tval, int32_t* cnt,
                            int32_t* tstart, void* stream) {
    cudaStream_t st = ST(stream);
    if (max_ub > SG_CAP / 2) { npb_set_error("spgemm_rows: row with %lld products", (long long)max_ub); return NPB_ERR_ARG; }
    if (npb_sg_warp_rows && max_ub > 16 && max_ub <= 512 && m > 0) {
        // short rows: one warp per row (no batches; cnt / tstart are written for every row).  Rows of
        // <= 16 products (D G of the cavity: 8) leave half a warp idle and stay on the batches.
        // Measured on the 21 products of the 2048^2 preconditioner setup (tools/profile_setup.py):
        // 55.4 -> 46.1 ms in total, the two largest 12.0 -> 9.0 and 12.0 -> 7.9 ms; the warp kernel is
        // bound by its four dependent rounds of global loads per row (row pointers, A entries, B row
        // pointers, B entries), not by the sort
        int64_t g = (m + 7) / 8;
        const int64_t gcap = (int64_t)NPB_NUM_SMS * 32;
        if (g > gcap) g = gcap;
#define NPB_SGW(CAP_) do { \
            const size_t shw = (size_t)8 * (CAP_) * 16; \
            static bool conf_ = false; \
            if (!conf_ && shw > 48 * 1024) { \
                NPB_CUDA(cudaFuncSetAttribute(k_sg_rows_warp<CAP_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shw)); \
                conf_ = true; \
            } \
            k_sg_rows_warp<CAP_><<<(unsigned)g, 256, shw, st>>>(m, a_indptr, a_indices, a_data, *a_scales, b_indptr, \
                b_indices, b_data, *b_scales, ubptr, prune, tidx, tval, cnt, tstart); \
        } while (0)
        if (max_ub <= 32) NPB_SGW(32);
        else if (max_ub <= 64) NPB_SGW(64);
        else if (max_ub <= 128) NPB_SGW(128);
        else if (max_ub <= 256) NPB_SGW(256);
        else NPB_SGW(512);
#undef NPB_SGW
        NPB_COUNT_LAUNCH();
        NPB_CHECK_LAUNCH("spgemm_rows_warp");
        return NPB_OK;
    }
    const int64_t T = SG_CAP - max_ub;
    const int64_t nb = total / T + 1;
    k_sg_batches<<<npb_blocks(nb + 1, 256), 256, 0, st>>>(m, ubptr, T, nb, brow);
    k_sg_init<<<npb_blocks(m, 256), 256, 0, st>>>(m, cnt, tstart);
    static bool configured = false;
    const size_t sh = sizeof(sg_smem) + 16;
    if (!configured) {
        NPB_CUDA(cudaFuncSetAttribute(k_sg_numeric, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
        configured = true;
    }
    int rb = 1;
    while ((1 << rb) < SG_CAP) ++rb;
    k_sg_numeric<<<(unsigned)nb, SG_THREADS, sh, st>>>(brow, a_indptr, a_indices, a_data, *a_scales,
                                                      b_indptr, b_indices, b_data, *b_scales, ubptr,
                                                      32 + rb + 1, prune, tidx, tval, cnt, tstart);
    NPB_COUNT_LAUNCH(); NPB_COUNT_LAUNCH(); NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("spgemm_rows_numeric");
    return NPB_OK;
}

// how many batches phase 2 uses (size of brow minus one)
int64_t npb_spgemm_rows_batches(int64_t total, int64_t max_ub) {
    const int64_t T = SG_CAP - max_ub;
    return T > 0 ? total / T + 1 : -1;
}

// c_indptr = exclusive scan of cnt; stats = {nnz(C), longest row}
int npb_spgemm_rows_indptr(int64_t m, const int32_t* cnt, int32_t* c_indptr, int64_t* scan_ws,
                           int64_t* stats, void* stream) {
    NPB_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(int64_t), ST(stream)));
    if (m > 0) {
        k_sg_max<<<npb_blocks(m, 256), 256, 0, ST(stream)>>>(m, cnt, stats);
        NPB_COUNT_LAUNCH();
    }
    return npb_scan_launch(cnt, c_indptr, m, scan_ws, stats, ST(stream));
}

// A times a selection with a non-monotone column map (aggregation maps, scatters): see
// k_relabel_merge_rows.  tidx / tval: nnz(A) entries of scratch; then npb_spgemm_rows_indptr
// on cnt and npb_spgemm_rows_compact.  For short rows (the sort is quadratic in the row length).
int npb_csr_relabel_merge_rows(int64_t m, const int32_t* indptr, const int32_t* indices,
                               const double* data, const npb_scales* scales, const int32_t* map,
                               const double* weights, int prune, int32_t* tidx, double* tval,
                               int32_t* cnt, int32_t* tstart, void* stream) {
    if (m <= 0) return NPB_OK;
    k_relabel_merge_rows<<<npb_blocks(m, 256), 256, 0, ST(stream)>>>(m, indptr, indices, data, *scales, map,
                                                                    weights, prune, tidx, tval, cnt, tstart);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("csr_relabel_merge_rows");
    return NPB_OK;
}

// column relabel by an injective map with per-row re-sort (see k_relabel_sort_rows); the row
// pointers are shared with the input
int npb_csr_relabel_sort_rows(int64_t m, const int32_t* indptr, const int32_t* indices,
                              const double* data, const int32_t* map, int32_t* indices_out,
                              double* data_out, void* stream) {
    if (m <= 0) return NPB_OK;
    k_relabel_sort_rows<<<npb_blocks(m, 256), 256, 0, ST(stream)>>>(m, indptr, indices, data, map,
                                                                   indices_out, data_out);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("csr_relabel_sort_rows");
    return NPB_OK;
}

// phase 3 (after npb_spgemm_rows_indptr and the allocation of C)
int npb_spgemm_rows_compact(int64_t m, const int32_t* cnt, const int32_t* tstart,
                            const int32_t* c_indptr, const int32_t* tidx, const double* tval,
                            int32_t* c_indices, double* c_data, void* stream) {
    if (m <= 0) return NPB_OK;
    k_sg_compact<<<npb_blocks(m * 8, 256), 256, 0, ST(stream)>>>(m, cnt, tstart, c_indptr, tidx, tval,
                                                                c_indices, c_data);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("spgemm_rows_compact");
    return NPB_OK;
}

}  // extern "C"
