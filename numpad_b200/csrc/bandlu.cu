// Banded LU with partial pivoting on the device + reverse Cuthill-McKee
// ordering on the host.
//
// Role: the robust direct solve for SMALL Newton systems (n * bandwidth^2
// cheap), e.g. the reference's own config 1 (test/cavity.py at N=50: n=7400,
// RCM bandwidth 150) whose late, convection-dominated Jacobians defeat every
// cheap preconditioner (hard part 1); it also serves as
// the coarsest-level solver of the multigrid hierarchy.  It stands where the
// reference calls SuperLU (adsolve.py:193-195, 78, 141); like SuperLU it is
// a pivoted sparse-direct method, restricted to a band.
//
// Storage is LAPACK's general band layout: A(i,j) at AB[kl+ku+i-j + j*ldab],
// ldab = 2*kl+ku+1 (kl extra superdiagonals receive the fill of row swaps).
// One CTA walks the columns; per column the pivot search, row swap, scaling
// and rank-1 update are block-parallel over <= kl*(kl+ku) entries that sit in
// L2.  This is latency-bound by construction (n dependent steps) and is not a
// bandwidth kernel; it is sized for n up to ~1e5.
#include <algorithm>
#include <queue>
#include <vector>

#include "common.cuh"

namespace {

constexpr int LU_T = 1024;

__global__ void __launch_bounds__(LU_T)
k_band_fill(int64_t n, const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
            const double* __restrict__ dat, const int32_t* __restrict__ inv,
            int kl, int ku, double* __restrict__ AB, int ldab) {
    const int lane = threadIdx.x & 7;
    int64_t r = ((int64_t)blockIdx.x * LU_T + threadIdx.x) >> 3;
    if (r >= n) return;
    const int pi = inv[r];
    for (int32_t k = ptr[r] + lane, e = ptr[r + 1]; k < e; k += 8) {
        const int pj = inv[idx[k]];
        if (pi - pj > kl || pj - pi > ku) continue;     // outside the band: caller's ordering is stale
        AB[(int64_t)(kl + ku + pi - pj) + (int64_t)pj * ldab] = dat[k];
    }
}

__global__ void __launch_bounds__(LU_T)
k_band_lu(int n, int kl, int ku, double* __restrict__ AB, int ldab,
          int32_t* __restrict__ ipiv, int32_t* __restrict__ info) {
    extern __shared__ double sm[];       // multipliers l[kl] | pivot row u[kl+ku+1] | reduce scratch
    double* sl = sm;
    double* su = sm + kl + 1;
    double* rv = su + (kl + ku + 2);
    int* ri = (int*)(rv + LU_T / 32);
    const int kv = kl + ku;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    int ju = 0;
    __shared__ int s_p;
    for (int j = 0; j < n; ++j) {
        const int km = min(kl, n - 1 - j);
        double* col = AB + (int64_t)j * ldab + kv;      // col[i] = A(j+i, j)
        // ---- pivot search over i = 0..km (first maximal |.|)
        double best = -1.0;
        int bi = 0;
        for (int i = tid; i <= km; i += LU_T) {
            double v = fabs(col[i]);
            if (v > best) { best = v; bi = i; }
        }
        for (int d = 16; d; d >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, best, d);
            int oi = __shfl_xor_sync(0xffffffffu, bi, d);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) { rv[warp] = best; ri[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            best = (lane < LU_T / 32) ? rv[lane] : -1.0;
            bi = (lane < LU_T / 32) ? ri[lane] : 0;
            for (int d = 16; d; d >>= 1) {
                double ov = __shfl_xor_sync(0xffffffffu, best, d);
                int oi = __shfl_xor_sync(0xffffffffu, bi, d);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (lane == 0) {
                s_p = bi;
                ipiv[j] = j + bi;
                if (best == 0.0 && *info == 0) *info = j + 1;   // exactly singular
            }
        }
        __syncthreads();
        const int p = s_p;
        ju = max(ju, min(j + ku + p, n - 1));
        const int nc = ju - j;                            // columns right of j touched
        // ---- swap rows j and j+p over columns j..ju, stage the pivot row
        for (int c = tid; c <= nc; c += LU_T) {
            double* a = AB + (int64_t)(j + c) * ldab + (kv - c);     // A(j, j+c)
            double top = a[0];
            if (p) { double bot = a[p]; a[p] = top; a[0] = bot; top = bot; }
            su[c] = top;
        }
        __syncthreads();
        const double piv = su[0];
        if (piv != 0.0) {
            const double rinv = 1.0 / piv;
            for (int i = 1 + tid; i <= km; i += LU_T) {
                double l = col[i] * rinv;
                col[i] = l;
                sl[i] = l;
            }
            __syncthreads();
            // ---- rank-1 update A(j+i, j+c) -= l[i] * u[c], i fastest (contiguous)
            const int tot = km * nc;
            for (int t = tid; t < tot; t += LU_T) {
                const int c = 1 + t / km, i = 1 + t % km;
                double* a = AB + (int64_t)(j + c) * ldab + (kv - c + i);
                *a = fma(-sl[i], su[c], *a);
            }
        }
        __syncthreads();
    }
}

// in-place solve A x = b with the factors above (dgbtrs, no transpose)
__global__ void __launch_bounds__(LU_T)
k_band_solve(int n, int kl, int ku, const double* __restrict__ AB, int ldab,
             const int32_t* __restrict__ ipiv, double* __restrict__ b) {
    const int kv = kl + ku;
    const int tid = threadIdx.x;
    __shared__ double s_bj;
    // L y = P b
    for (int j = 0; j < n; ++j) {
        const int lm = min(kl, n - 1 - j);
        if (tid == 0) {
            const int p = ipiv[j];
            double bj = b[p];
            if (p != j) { b[p] = b[j]; b[j] = bj; }
            s_bj = bj;
        }
        __syncthreads();
        const double bj = s_bj;
        const double* col = AB + (int64_t)j * ldab + kv;
        for (int i = 1 + tid; i <= lm; i += LU_T) b[j + i] = fma(-bj, col[i], b[j + i]);
        __syncthreads();
    }
    // U x = y  (kv superdiagonals)
    for (int j = n - 1; j >= 0; --j) {
        const double* col = AB + (int64_t)j * ldab + kv;
        if (tid == 0) { double v = b[j] / col[0]; b[j] = v; s_bj = v; }
        __syncthreads();
        const double bj = s_bj;
        const int lo = max(0, j - kv);
        for (int i = lo + tid; i < j; i += LU_T) b[i] = fma(-bj, col[i - j], b[i]);
        __syncthreads();
    }
}

}  // namespace

#define ST(s) ((cudaStream_t)(s))

extern "C" {

// [host] Reverse Cuthill-McKee on the symmetrised pattern of an n x n CSR
// matrix.  perm[k] = old index placed at position k.  bw[0], bw[1] = lower /
// upper bandwidth of the permuted matrix.
int npb_rcm_host(int64_t n, const int32_t* indptr, const int32_t* indices,
                 int32_t* perm, int32_t* bw) {
    if (n < 0) return NPB_ERR_ARG;
    std::vector<std::vector<int32_t>> adj((size_t)n);
    for (int64_t i = 0; i < n; ++i)
        for (int32_t k = indptr[i]; k < indptr[i + 1]; ++k) {
            int32_t j = indices[k];
            if (j == i) continue;
            adj[i].push_back(j);
            adj[j].push_back((int32_t)i);
        }
    std::vector<int32_t> deg((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        auto& a = adj[i];
        std::sort(a.begin(), a.end());
        a.erase(std::unique(a.begin(), a.end()), a.end());
        deg[i] = (int32_t)a.size();
    }
    for (int64_t i = 0; i < n; ++i)
        std::sort(adj[i].begin(), adj[i].end(), [&](int32_t x, int32_t y) {
            return deg[x] != deg[y] ? deg[x] < deg[y] : x < y;
        });
    std::vector<int32_t> order;
    order.reserve((size_t)n);
    std::vector<char> seen((size_t)n, 0);
    std::vector<int32_t> by_deg((size_t)n);
    for (int64_t i = 0; i < n; ++i) by_deg[i] = (int32_t)i;
    std::sort(by_deg.begin(), by_deg.end(), [&](int32_t x, int32_t y) {
        return deg[x] != deg[y] ? deg[x] < deg[y] : x < y;
    });
    auto bfs = [&](int32_t root, std::vector<int32_t>& out) {
        // level-structure BFS; returns the last visited node (far end)
        size_t head = out.size();
        out.push_back(root);
        seen[root] = 1;
        while (head < out.size()) {
            int32_t v = out[head++];
            for (int32_t w : adj[v])
                if (!seen[w]) { seen[w] = 1; out.push_back(w); }
        }
        return out.back();
    };
    for (int32_t s : by_deg) {
        if (seen[s]) continue;
        // pseudo-peripheral start: two exploratory sweeps inside the component
        std::vector<int32_t> tmp;
        int32_t far = bfs(s, tmp);
        for (int32_t v : tmp) seen[v] = 0;
        tmp.clear();
        int32_t far2 = bfs(far, tmp);
        for (int32_t v : tmp) seen[v] = 0;
        bfs(far2, order);
    }
    std::reverse(order.begin(), order.end());
    std::vector<int32_t> inv((size_t)n);
    for (int64_t k = 0; k < n; ++k) { perm[k] = order[k]; inv[order[k]] = (int32_t)k; }
    int32_t lo = 0, up = 0;
    for (int64_t i = 0; i < n; ++i)
        for (int32_t k = indptr[i]; k < indptr[i + 1]; ++k) {
            int32_t d = inv[i] - inv[indices[k]];
            if (d > lo) lo = d;
            if (-d > up) up = -d;
        }
    bw[0] = lo;
    bw[1] = up;
    return NPB_OK;
}

// AB (ldab x n, zero-initialised by the caller) <- P A P^T, inv[old] = new
int npb_band_fill(int64_t n, const int32_t* indptr, const int32_t* indices,
                  const double* data, const int32_t* inv, int kl, int ku, double* AB,
                  int ldab, void* stream) {
    if (ldab < 2 * kl + ku + 1) return NPB_ERR_ARG;
    if (n <= 0) return NPB_OK;
    k_band_fill<<<npb_blocks(n * 8, LU_T), LU_T, 0, ST(stream)>>>(n, indptr, indices, data, inv,
                                                               kl, ku, AB, ldab);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("band_fill");
    return NPB_OK;
}

// *info (device int32, zeroed here): 0 ok, j+1 = exact zero pivot in column j
int npb_band_lu_factor(int n, int kl, int ku, double* AB, int ldab, int32_t* ipiv,
                       int32_t* info, void* stream) {
    if (ldab < 2 * kl + ku + 1 || n < 0) return NPB_ERR_ARG;
    NPB_CUDA(cudaMemsetAsync(info, 0, sizeof(int32_t), ST(stream)));
    if (n == 0) return NPB_OK;
    size_t sh = sizeof(double) * ((size_t)(kl + 1) + (size_t)(kl + ku + 2) + LU_T / 32) +
                sizeof(int) * (LU_T / 32);
    if (sh > 200 * 1024) { npb_set_error("band_lu: bandwidth too large for shared memory"); return NPB_ERR_ARG; }
    if (sh > 48 * 1024)
        NPB_CUDA(cudaFuncSetAttribute(k_band_lu, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    k_band_lu<<<1, LU_T, sh, ST(stream)>>>(n, kl, ku, AB, ldab, ipiv, info);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("band_lu");
    return NPB_OK;
}

int npb_band_lu_solve(int n, int kl, int ku, const double* AB, int ldab,
                      const int32_t* ipiv, double* b, void* stream) {
    if (n <= 0) return NPB_OK;
    k_band_solve<<<1, LU_T, 0, ST(stream)>>>(n, kl, ku, AB, ldab, ipiv, b);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("band_solve");
    return NPB_OK;
}

}  // extern "C"
