// Algebraic multigrid (smoothed aggregation, V-cycle, damped-Jacobi smoother)
// and the least-squares-commutator block preconditioner built on it.
//
// Role (seam S3): the preconditioner of the FGMRES that replaces SuperLU
// `spsolve` in the reference Newton loop (adsolve.py:193-195).  The cavity
// Jacobian is a saddle-point matrix (N^2-1 zero diagonals):
// with V = rows with a diagonal entry, P = rows without,
//      J = [ A  G ]  (V rows)        S = -D A^-1 G  ~  -(DG) (D A G)^-1 (DG)
//          [ D  0 ]  (P rows)
// M^-1 r (block upper triangular):  p = -L~^-1 (D A G) L~^-1 r_P  with L = -D G ;
// u = A~^-1 (r_V - G p).   L~^-1 = k CG steps preconditioned by a V-cycle,
// A~^-1 = k stationary V-cycles.
// The setup (aggregation, Galerkin products) runs once per matrix through the
// chain-rule kernels of assembly.cu/coo.cu; everything below is the solve
// phase: SpMV-shaped, HBM-bound kernels launched from C++ (no Python in the
// Krylov loop).
#include <math.h>

#include <map>

#include "common.cuh"
#include "dist.cuh"
#include "scan.cuh"

int npb_spmv_launch(int64_t nrows, int64_t nnz, const int32_t* ptr, const int32_t* idx,
                    const double* dat, const double* x, const double* b, double* y,
                    int mode, cudaStream_t st);
int npb_axpby_launch(int64_t n, double a, const double* x, double b, double* y, cudaStream_t st);
int npb_spmv_ex(int64_t nrows, int64_t nnz, const int32_t* ptr, const int32_t* idx,
                const double* dat, const double* x, int mode, const double* b,
                const double* dinv, double omega, double* y, int variant, cudaStream_t st,
                const npb_halo* halo, const npb_wait* wait, const float* dat32);
int npb_dot_multi_launch(int64_t n, int k, const double* V, int64_t ld, const double* w,
                         double* out, void* scratch, cudaStream_t st);
int npb_spmv_dia_launch(int64_t nrows, int64_t ncols, int ndiag, int64_t ld,
                        const int32_t* offsets_host, const float* vals, const double* x, int mode,
                        const double* b, const double* dinv, double omega, double* y,
                        cudaStream_t st, double* dot_out, void* red, int ell_w, const int32_t* ell_idx,
                        const float* ell_val);
int npb_dia_cg_dir_launch(int64_t nrows, int ndiag, int64_t ld, const int32_t* offsets_host,
                          const float* vals, const double* z, const double* rz_new,
                          const double* rz_old, int first, double* p, double* q, double* pq_out,
                          void* red, cudaStream_t st);

extern "C" {
typedef int (*npb_apply_fn)(void* ctx, const double* x, double* y, void* stream);
typedef int (*npb_allreduce_fn)(void* ctx, double* buf, int count, void* stream);
struct npb_krylov_ops {
    npb_apply_fn matvec; void* matvec_ctx;
    npb_apply_fn precond; void* precond_ctx;
    npb_allreduce_fn allreduce; void* allreduce_ctx;
};
struct npb_csr_view { int64_t nrows, nnz; const int32_t* indptr; const int32_t* indices; const double* data; };
struct npb_krylov_info { int32_t iters, restarts, status, pad; double bnorm, rnorm0, rnorm; };
int npb_csr_matvec_cb(void* ctx, const double* x, double* y, void* stream);
int64_t npb_fgmres_workspace_bytes(int64_t n, int restart, int flexible);
int npb_fgmres(int64_t n, const npb_krylov_ops* ops, const double* b, double* x, int restart,
               int maxiter, double rtol, double atol, int flexible, void* work,
               int64_t work_bytes, npb_krylov_info* info, void* stream);
}

namespace {

constexpr int TPB = 256;

__global__ void __launch_bounds__(TPB)
k_jacobi0(int64_t n, const double* __restrict__ dinv, double omega,
          const double* __restrict__ b, double* __restrict__ x) {
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i < n) x[i] = omega * dinv[i] * b[i];
}

// y = Ainv b (dense row-major n x n), one warp per row
__global__ void __launch_bounds__(TPB)
k_dense_mv(int n, const double* __restrict__ M, const double* __restrict__ b,
           double* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int r = (blockIdx.x * TPB + threadIdx.x) >> 5;
    double s = 0.0;
    if (r < n)
        for (int c = lane; c < n; c += 32) s = fma(M[(int64_t)r * n + c], b[c], s);
    for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (r < n && lane == 0) y[r] = s;
}

__global__ void __launch_bounds__(TPB)
k_gather(int64_t n, const int32_t* __restrict__ idx, const double* __restrict__ src,
         double* __restrict__ dst) {
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i < n) dst[i] = src[idx[i]];
}

__global__ void __launch_bounds__(TPB)
k_scatter(int64_t n, const int32_t* __restrict__ idx, const double* __restrict__ src,
          double* __restrict__ dst) {
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i < n) dst[idx[i]] = src[i];
}

// CG with every scalar on the device (no host round trip inside the
// preconditioner):  p = z + (rz_new / rz_old) p   (first = 1: p = z)
__global__ void __launch_bounds__(TPB)
k_cg_dir(int64_t n, const double* __restrict__ rz_new, const double* __restrict__ rz_old,
         int first, const double* __restrict__ z, double* __restrict__ p) {
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i >= n) return;
    if (first) { p[i] = z[i]; return; }
    const double den = *rz_old;
    const double beta = den != 0.0 ? *rz_new / den : 0.0;
    p[i] = fma(beta, p[i], z[i]);
}

// alpha = rz / pq ;  x += alpha p ;  r = r_in - alpha q   (r_in may be r; r == NULL: the last
// step of a fixed number, whose residual nobody reads)
__global__ void __launch_bounds__(TPB)
k_cg_update(int64_t n, const double* __restrict__ rz, const double* __restrict__ pq,
            const double* __restrict__ p, const double* __restrict__ q,
            double* __restrict__ x, const double* r_in, double* r, int first) {
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i >= n) return;
    const double den = *pq;
    const double alpha = den != 0.0 ? *rz / den : 0.0;
    x[i] = first ? alpha * p[i] : fma(alpha, p[i], x[i]);
    if (r) r[i] = fma(-alpha, q[i], r_in[i]);
}

// y = d .* x
__global__ void __launch_bounds__(TPB)
k_diag_scale(int64_t n, const double* __restrict__ d, const double* __restrict__ x,
             double* __restrict__ y) {
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i < n) y[i] = d[i] * x[i];
}

// u = a - dinv .* g
__global__ void __launch_bounds__(TPB)
k_sub_scaled(int64_t n, const double* a, const double* __restrict__ dinv,
             const double* __restrict__ g, double* u) {
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i < n) u[i] = a[i] - dinv[i] * g[i];
}

// ------------------------------------------------------------------ setup
// diag[r] = A(r, r) (0 when absent)
__global__ void __launch_bounds__(TPB)
k_csr_diag(int64_t nrows, const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
           const double* __restrict__ dat, double* __restrict__ diag) {
    int64_t r = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (r >= nrows) return;
    double d = 0.0;
    for (int32_t k = ptr[r], e = ptr[r + 1]; k < e; ++k)
        if (idx[k] == r) d = dat[k];
    diag[r] = d;
}

// Strength-filtered copy of A for prolongator smoothing: off-diagonal entries
// with |a_ij| < theta * max_k |a_ik| are dropped and lumped onto the diagonal
// (row sums preserved).  FILL = false counts, FILL = true writes.
template <bool FILL>
__global__ void __launch_bounds__(TPB)
k_filter(int64_t nrows, const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
         const double* __restrict__ dat, double theta, int32_t* __restrict__ cnt,
         const int32_t* __restrict__ optr, int32_t* __restrict__ oidx, double* __restrict__ out) {
    int64_t r = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (r >= nrows) return;
    double mx = 0.0;
    for (int32_t k = ptr[r], e = ptr[r + 1]; k < e; ++k)
        if (idx[k] != r) mx = fmax(mx, fabs(dat[k]));
    const double thr = theta * mx;
    int32_t c = 0, o = FILL ? optr[r] : 0, dpos = -1;
    double lump = 0.0;
    for (int32_t k = ptr[r], e = ptr[r + 1]; k < e; ++k) {
        const bool diag = idx[k] == r;
        if (diag || fabs(dat[k]) >= thr) {
            if (FILL) { oidx[o] = idx[k]; out[o] = dat[k]; if (diag) dpos = o; ++o; }
            ++c;
        } else {
            lump += dat[k];
        }
    }
    if (FILL) { if (dpos >= 0) out[dpos] += lump; }
    else cnt[r] = c;
}

// Prolongator truncation: keep the entries of a row with |p| >= theta * max|p|
// and rescale them so that the row sum is preserved (constants stay in the
// range of P).  Limits the stencil growth of the Galerkin coarse operators.
template <bool FILL>
__global__ void __launch_bounds__(TPB)
k_truncate_rows(int64_t nrows, const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                const double* __restrict__ dat, double theta, int32_t* __restrict__ cnt,
                const int32_t* __restrict__ optr, int32_t* __restrict__ oidx,
                double* __restrict__ out) {
    int64_t r = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (r >= nrows) return;
    double mx = 0.0, sum_all = 0.0, sum_kept = 0.0;
    for (int32_t k = ptr[r], e = ptr[r + 1]; k < e; ++k) { mx = fmax(mx, fabs(dat[k])); sum_all += dat[k]; }
    const double thr = theta * mx;
    int32_t c = 0;
    for (int32_t k = ptr[r], e = ptr[r + 1]; k < e; ++k)
        if (fabs(dat[k]) >= thr && dat[k] != 0.0) { ++c; sum_kept += dat[k]; }
    if (!FILL) { cnt[r] = c; return; }
    // rescale only when the kept part carries most of the row sum (P of an
    // M-matrix-like operator is essentially non-negative)
    const double scale = (fabs(sum_kept) >= 0.5 * fabs(sum_all) && sum_kept != 0.0) ? sum_all / sum_kept : 1.0;
    int32_t o = optr[r];
    for (int32_t k = ptr[r], e = ptr[r + 1]; k < e; ++k)
        if (fabs(dat[k]) >= thr && dat[k] != 0.0) { oidx[o] = idx[k]; out[o] = dat[k] * scale; ++o; }
}

__device__ __forceinline__ uint32_t hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
    return x;
}

// One Luby round on the strength graph (|a_ij| >= theta * max_k |a_ik|):
// state 0 undecided, 1 root, 2 covered.  Reads `state`, writes `next`.
__global__ void __launch_bounds__(TPB)
k_mis_select(int64_t n, const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
             const double* __restrict__ dat, double theta, const int8_t* __restrict__ state,
             int8_t* __restrict__ next) {
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i >= n) return;
    int8_t s = state[i];
    if (s == 0) {
        double mx = 0.0;
        for (int32_t k = ptr[i], e = ptr[i + 1]; k < e; ++k)
            if (idx[k] != i) mx = fmax(mx, fabs(dat[k]));
        const double thr = theta * mx;
        const uint32_t pi = hash32((uint32_t)i);
        bool root = true;
        for (int32_t k = ptr[i], e = ptr[i + 1]; k < e && root; ++k) {
            const int32_t j = idx[k];
            if (j == i || fabs(dat[k]) < thr || dat[k] == 0.0 || state[j] != 0) continue;
            const uint32_t pj = hash32((uint32_t)j);
            if (pj > pi || (pj == pi && j > i)) root = false;
        }
        if (root) s = 1;
    }
    next[i] = s;
}

__global__ void __launch_bounds__(TPB)
k_mis_cover(int64_t n, const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
            const double* __restrict__ dat, double theta, int8_t* __restrict__ state,
            int32_t* __restrict__ undecided) {
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    bool und = false;
    if (i < n && state[i] == 0) {
        double mx = 0.0;
        for (int32_t k = ptr[i], e = ptr[i + 1]; k < e; ++k)
            if (idx[k] != i) mx = fmax(mx, fabs(dat[k]));
        const double thr = theta * mx;
        bool cov = false;
        for (int32_t k = ptr[i], e = ptr[i + 1]; k < e; ++k) {
            const int32_t j = idx[k];
            if (j != i && dat[k] != 0.0 && fabs(dat[k]) >= thr && state[j] == 1) { cov = true; break; }
        }
        // benign race: neighbours only move 0 -> 2 here, roots are already final
        if (cov) state[i] = 2; else und = true;
    }
    const unsigned m = __ballot_sync(0xffffffffu, und);
    if (m && (threadIdx.x & 31) == 0) atomicAdd(undecided, __popc(m));
}

__global__ void __launch_bounds__(TPB)
k_mis_flags(int64_t n, const int8_t* __restrict__ state, int force_roots, int32_t* __restrict__ flag) {
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i < n) flag[i] = (state[i] == 1) || (force_roots && state[i] == 0);
}

// agg[i] = id of own aggregate (roots) or of the most strongly connected root
__global__ void __launch_bounds__(TPB)
k_mis_join(int64_t n, const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
           const double* __restrict__ dat, const int32_t* __restrict__ flag,
           const int32_t* __restrict__ root_id, int32_t* __restrict__ agg) {
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i >= n) return;
    if (flag[i]) { agg[i] = root_id[i]; return; }
    double best = -1.0;
    int32_t bj = -1;
    for (int32_t k = ptr[i], e = ptr[i + 1]; k < e; ++k) {
        const int32_t j = idx[k];
        if (j != i && flag[j] && fabs(dat[k]) > best) { best = fabs(dat[k]); bj = j; }
    }
    agg[i] = bj >= 0 ? root_id[bj] : -1;    // -1 cannot happen for covered nodes
}

// ---- distance-2 maximal independent set (MIS-2) aggregation ----------------------------
// Roots are at least 3 strong edges apart, so aggregates (a root, its strong neighbours,
// and the nodes two strong edges away) hold ~7 nodes of a 5-point stencil instead of the
// ~2.75 of the distance-1 set above: operator complexity ~1.3 instead of ~2.5-3.
// key = (hash, index): unique priorities.  Luby round: propagate the maximum key of the
// undecided nodes over two strong edges; a node that sees its own key is a root; nodes
// within two strong edges of a root are excluded.
__device__ __forceinline__ unsigned long long mis_key(int64_t i) {
    return ((unsigned long long)hash32((uint32_t)i) << 32) | (unsigned long long)(uint32_t)i;
}

__device__ __forceinline__ double row_threshold(const int32_t* __restrict__ ptr,
                                                const int32_t* __restrict__ idx,
                                                const double* __restrict__ dat, int64_t i,
                                                double theta) {
    if (theta <= 0.0) return 0.0;
    double mx = 0.0;
    for (int32_t k = ptr[i], e = ptr[i + 1]; k < e; ++k)
        if (idx[k] != i) mx = fmax(mx, fabs(dat[k]));
    return theta * mx;
}

// out[i] = max(in'[i], max over strong neighbours j of in'[j]);  FIRST: in'[i] = undecided ? key : 0
template <bool FIRST>
__global__ void __launch_bounds__(TPB)
k_mis2_max(int64_t n, const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
           const double* __restrict__ dat, double theta, const int8_t* __restrict__ state,
           const unsigned long long* __restrict__ in, unsigned long long* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i >= n) return;
    const double thr = row_threshold(ptr, idx, dat, i, theta);
    unsigned long long m = FIRST ? (state[i] == 0 ? mis_key(i) : 0ull) : in[i];
    for (int32_t k = ptr[i], e = ptr[i + 1]; k < e; ++k) {
        const int32_t j = idx[k];
        if (j == i || dat[k] == 0.0 || fabs(dat[k]) < thr) continue;
        const unsigned long long v = FIRST ? (state[j] == 0 ? mis_key(j) : 0ull) : in[j];
        m = v > m ? v : m;
    }
    out[i] = m;
}

__global__ void __launch_bounds__(TPB)
k_mis2_root(int64_t n, const unsigned long long* __restrict__ m2, int8_t* __restrict__ state) {
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i < n && state[i] == 0 && m2[i] == mis_key(i)) state[i] = 1;
}

// near_out[i] = near_in'[i] or any strong neighbour j with near_in'[j];  FIRST: near_in' = root.
// LAST: undecided nodes that are near become excluded (state 2); counts the still undecided.
template <bool FIRST, bool LAST>
__global__ void __launch_bounds__(TPB)
k_mis2_near(int64_t n, const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
            const double* __restrict__ dat, double theta, int8_t* __restrict__ state,
            const int8_t* __restrict__ near_in, int8_t* __restrict__ near_out,
            int32_t* __restrict__ undecided) {
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    bool und = false;
    if (i < n) {
        const int8_t si = state[i];
        bool near = FIRST ? (si == 1) : (near_in[i] != 0);
        if (!near && (si == 0 || !LAST)) {
            const double thr = row_threshold(ptr, idx, dat, i, theta);
            for (int32_t k = ptr[i], e = ptr[i + 1]; k < e && !near; ++k) {
                const int32_t j = idx[k];
                if (j == i || dat[k] == 0.0 || fabs(dat[k]) < thr) continue;
                near = FIRST ? (state[j] == 1) : (near_in[j] != 0);
            }
        }
        if (!LAST) near_out[i] = near ? 1 : 0;
        else if (si == 0) { if (near) state[i] = 2; else und = true; }
    }
    if (LAST) {
        const unsigned m = __ballot_sync(0xffffffffu, und);
        if (m && (threadIdx.x & 31) == 0) atomicAdd(undecided, __popc(m));
    }
}

// label = index of the root a node belongs to (-1: none yet).  INIT: roots label themselves.
// Otherwise an unlabelled node takes the label of its most strongly connected labelled
// neighbour in the strength graph (snapshot semantics: reads `in`, writes `out`).
template <bool INIT>
__global__ void __launch_bounds__(TPB)
k_mis2_join(int64_t n, const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
            const double* __restrict__ dat, double theta, const int8_t* __restrict__ state,
            const int32_t* __restrict__ in, int32_t* __restrict__ out) {
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i >= n) return;
    if (INIT) { out[i] = state[i] == 1 ? (int32_t)i : -1; return; }
    int32_t lab = in[i];
    if (lab < 0) {
        const double thr = row_threshold(ptr, idx, dat, i, theta);
        double best = -1.0;
        for (int32_t k = ptr[i], e = ptr[i + 1]; k < e; ++k) {
            const int32_t j = idx[k];
            if (j == i || dat[k] == 0.0 || fabs(dat[k]) < thr) continue;
            const int32_t lj = in[j];
            if (lj >= 0 && fabs(dat[k]) > best) { best = fabs(dat[k]); lab = lj; }
        }
    }
    out[i] = lab;
}

// leftover nodes found their own aggregate; flag = "is a root"
__global__ void __launch_bounds__(TPB)
k_mis2_flags(int64_t n, int32_t* __restrict__ label, int32_t* __restrict__ flag) {
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i >= n) return;
    int32_t l = label[i];
    if (l < 0) { l = (int32_t)i; label[i] = l; }
    flag[i] = (l == (int32_t)i);
}

__global__ void __launch_bounds__(TPB)
k_mis2_ids(int64_t n, const int32_t* __restrict__ label, const int32_t* __restrict__ root_id,
           int32_t* __restrict__ agg) {
    int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (i < n) agg[i] = root_id[label[i]];
}

__global__ void __launch_bounds__(TPB)
k_csr_to_dense(int64_t n, const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
               const double* __restrict__ dat, double* __restrict__ M, int ld) {
    int64_t r = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (r >= n) return;
    for (int32_t k = ptr[r], e = ptr[r + 1]; k < e; ++k) M[r * ld + idx[k]] = dat[k];
}

// Gauss-Jordan inverse with partial pivoting, one CTA, W = [A | I] (n x 2n,
// row-major, in global memory / L2); on exit the right half is A^-1.
// Per column: block arg-max, row swap, pivot row scaled and staged in shared
// memory, then one warp per row eliminates with coalesced accesses.  Only the
// columns that can be non-trivial are touched: k+1..n-1 of the left half and
// 0..n-1 of the right half restricted to what earlier steps filled is not
// tracked -- the right half is swept fully (n <= 2048).
__global__ void __launch_bounds__(1024)
k_gauss_jordan(int n, double* __restrict__ W, int32_t* __restrict__ info) {
    extern __shared__ double prow[];      // 2n doubles: the scaled pivot row
    __shared__ double rv[32];
    __shared__ int ri[32];
    __shared__ int s_p;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ld = 2 * n;
    for (int k = 0; k < n; ++k) {
        double best = -1.0;
        int bi = k;
        for (int r = k + tid; r < n; r += 1024) {
            double v = fabs(W[(int64_t)r * ld + k]);
            if (v > best) { best = v; bi = r; }
        }
        for (int d = 16; d; d >>= 1) {
            double ov = __shfl_xor_sync(0xffffffffu, best, d);
            int oi = __shfl_xor_sync(0xffffffffu, bi, d);
            if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
        }
        if (lane == 0) { rv[warp] = best; ri[warp] = bi; }
        __syncthreads();
        if (warp == 0) {
            best = rv[lane]; bi = ri[lane];
            for (int d = 16; d; d >>= 1) {
                double ov = __shfl_xor_sync(0xffffffffu, best, d);
                int oi = __shfl_xor_sync(0xffffffffu, bi, d);
                if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
            }
            if (lane == 0) { s_p = bi; if (best == 0.0 && *info == 0) *info = k + 1; }
        }
        __syncthreads();
        const int p = s_p;
        double* rk = W + (int64_t)k * ld;
        double* rp = W + (int64_t)p * ld;
        const double piv = rp[k];
        __syncthreads();                  // everyone holds the pivot before row p is rewritten
        if (piv == 0.0) continue;
        const double rinv = 1.0 / piv;
        // swap rows k and p (columns >= k of the left half, all of the right
        // half), scale the new pivot row, stage it in shared memory
        for (int c = k + tid; c < ld; c += 1024) {
            const double top = rk[c], bot = rp[c];
            const double v = bot * rinv;
            if (p != k) rp[c] = top;
            rk[c] = v;
            prow[c] = v;
        }
        __syncthreads();
        // eliminate column k from every other row: one warp per row
        for (int r = warp; r < n; r += 32) {
            if (r == k) continue;
            double* rr = W + (int64_t)r * ld;
            const double f = rr[k];
            if (f == 0.0) continue;
            for (int c = k + 1 + lane; c < ld; c += 32) rr[c] = fma(-f, prow[c], rr[c]);
            if (lane == 0) rr[k] = 0.0;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(TPB)
k_copy_right_half(int n, const double* __restrict__ W, double* __restrict__ Minv) {
    int64_t t = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (t < (int64_t)n * n) {
        const int r = (int)(t / n), c = (int)(t % n);
        Minv[t] = W[(int64_t)r * 2 * n + n + c];
    }
}

__global__ void __launch_bounds__(TPB)
k_set_identity_right(int n, double* __restrict__ W) {
    int64_t r = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (r < n) W[r * 2 * n + n + r] = 1.0;
}

}  // namespace

#define ST(s) ((cudaStream_t)(s))

extern "C" {

struct npb_csr_mat {
    int64_t nrows, ncols, nnz;
    const int32_t* indptr;
    const int32_t* indices;
    const double* data;
    npb_halo halo;         // halo.dist != NULL: rows of a row-sharded operator (dist.cuh)
    const float* data32;   // != NULL: the values rounded to fp32 (preconditioner operators:
                           // 8 instead of 12 bytes per entry; sums, x and y stay fp64)
    // != NULL: diagonal-storage image (fp32, ndiag x nrows) of a stencil operator whose entries
    // sit on ndiag <= 8 constant offsets; preferred over CSR by spmv_mat (single GPU)
    const float* dia_vals;
    int32_t ndiag, dia_ld;     // dia_ld: floats between consecutive diagonals (>= nrows)
    int32_t dia_off[8];
    // hybrid image (ell_w > 0): the entries on none of the diagonals in an ELL part of that
    // width, slot k of row r at [k * dia_ld + r]
    const int32_t* ell_idx;
    const float* ell_val;
    int32_t ell_w, pad2;
};

// y = epilogue(M x): 0: Mx  1: b - Mx  2: b + Mx  3: x + omega dinv (b - Mx); a sharded
// operator first publishes / fetches the boundary values of x
// (dot_out: only with a diagonal-storage operator and mode 3, see npb_spmv_dia_launch)
static inline int spmv_mat(const npb_csr_mat& M, const double* x, const double* b, double* y,
                           int mode, cudaStream_t st, const double* dinv = nullptr,
                           double omega = 0.0, double* dot_out = nullptr, void* red = nullptr) {
    if (M.dia_vals && !M.halo.dist)
        return npb_spmv_dia_launch(M.nrows, M.ncols, M.ndiag, M.dia_ld, M.dia_off, M.dia_vals, x, mode, b,
                                   dinv, omega, y, st, dot_out, red, M.ell_w, M.ell_idx, M.ell_val);
    if (M.halo.dist) {
        npb_halo h = M.halo;
        const double* xh = nullptr;
        npb_wait w;
        int rc = npb_dist_halo_exchange(M.halo, x, st, &xh, &w);   // publish; the SpMV waits
        if (rc) return rc;
        h.recv = const_cast<double*>(xh);
        return npb_spmv_ex(M.nrows, M.nnz, M.indptr, M.indices, M.data, x, mode, b, dinv, omega, y,
                           0, st, &h, &w, M.data32);
    }
    return npb_spmv_ex(M.nrows, M.nnz, M.indptr, M.indices, M.data, x, mode, b, dinv, omega, y, 0,
                       st, nullptr, nullptr, M.data32);
}

struct npb_amg_level {
    npb_csr_mat A;         // level operator
    npb_csr_mat P;         // prolongation  (n x n_coarse)   [unused on the coarsest level]
    npb_csr_mat R;         // restriction   (n_coarse x n)
    const double* dinv;    // 1 / diag(A)
    double* x;             // work vectors of length n (x, x2 ping-pong, b, r)
    double* x2;
    double* b;
    double* r;
    // multi-GPU: the last row-sharded level (gather != 0) hands its restricted residual
    // to the replicated levels below: every rank computes its slice [offsets[rank],
    // offsets[rank+1]) of the coarse right-hand side and the slices are all-gathered
    int32_t gather, pad;
    const int64_t* offsets;     // [host] world + 1 entries
    // direct-store variant: the coarse right-hand side lives in the symmetric heap
    // (2 parities x offsets[world] doubles at g_data, world arrival counters at g_flags)
    int64_t g_data, g_flags;    // < 0: NCCL broadcasts into the next level's b
    unsigned long long* g_seq;  // [host] exchange counter
};

struct npb_amg {
    int32_t nlevels;       // >= 1
    int32_t nu_pre, nu_post;
    int32_t coarse_n;      // size of the dense coarse inverse (0: smooth only)
    double omega;
    const double* coarse_inv;   // dense row-major inverse of the coarsest operator
    npb_amg_level* levels;      // [host] array of nlevels descriptors
    const void* dist;           // npb_dist* when the fine levels are row-sharded, else NULL
    int32_t gamma, gamma_from;  // gamma > 1: levels >= gamma_from are visited gamma times per cycle
};

static int jacobi_sweep(const npb_amg_level& L, double omega, const double* b, const double* xin,
                        double* xout, cudaStream_t st) {
    // xout = xin + omega dinv (b - A xin): the SpMV kernel with the Jacobi epilogue
    return spmv_mat(L.A, xin, b, xout, 3, st, L.dinv, omega);
}

// One multigrid cycle on level l from a zero initial guess; the result is left in X[l] (level 0:
// in xfine).  Which of the two ping-pong buffers of a level holds its iterate is kept in the
// caller's arrays: the level descriptors are never modified, so the launch sequence of a cycle
// (buffers included) is the same on every call and a captured CUDA graph of it can be replayed.
// gamma > 1: levels from gamma_from down are visited gamma times (W-cycle on the coarse part of
// the hierarchy, where a visit costs a few launches): repeated coarse corrections, each from a
// zero coarse guess, between the pre- and the post-smoothing of the parent.
struct cycle_bufs {
    double* X[64];
    double* X2[64];
    double* CB[64];          // right-hand side of level l (l >= 1)
    const double* bfine;
    double* xfine;
    double* dot_out;         // != NULL: the last fine sweep also returns bfine . xfine (DIA level 0)
    void* red;
};

static int amg_cycle(const npb_amg* h, int l, cycle_bufs& w, cudaStream_t st) {
    const int nl = h->nlevels;
    npb_amg_level& L = h->levels[l];
    const int64_t n = L.A.nrows;
    const double* bl = (l == 0) ? w.bfine : w.CB[l];
    if (l == nl - 1 && h->coarse_n > 0) {
        double* xl = (l == 0) ? w.xfine : w.X[l];
        k_dense_mv<<<npb_blocks((int64_t)h->coarse_n * 32, TPB), TPB, 0, st>>>(h->coarse_n, h->coarse_inv, bl, xl);
        NPB_COUNT_LAUNCH();
        return NPB_OK;
    }
    // pre-smoothing from zero: x = omega dinv b, then nu_pre-1 sweeps
    double* cur = w.X[l];
    double* oth = w.X2[l];
    k_jacobi0<<<npb_blocks(n, TPB), TPB, 0, st>>>(n, L.dinv, h->omega, bl, cur);
    NPB_COUNT_LAUNCH();
    NPB_COUNT_BYTES(24.0 * n);
    const int extra = (l == nl - 1) ? (h->nu_pre + h->nu_post + 6) : (h->nu_pre - 1);
    for (int s = 0; s < extra; ++s) {
        jacobi_sweep(L, h->omega, bl, cur, oth, st);
        double* t = cur; cur = oth; oth = t;
    }
    w.X[l] = cur; w.X2[l] = oth;                      // the iterate is in X[l]
    if (l == nl - 1) {
        if (l == 0) NPB_CUDA(cudaMemcpyAsync(w.xfine, w.X[l], n * 8, cudaMemcpyDeviceToDevice, st));
        return NPB_OK;
    }
    const int reps = (h->gamma > 1 && l + 1 >= h->gamma_from && l + 2 < nl) ? h->gamma : 1;
    for (int rep = 0; rep < reps; ++rep) {
        // r = b - A x ; b_{l+1} = R r
        int rc = spmv_mat(L.A, w.X[l], bl, L.r, 1, st);
        if (rc) return rc;
        if (L.gather) {
            const npb_dist* d = (const npb_dist*)h->dist;
            if (L.g_data >= 0 && npb_dist_heap_ptr(d, 0)) {
                // own slice written in place, the peers' slices arrive through the heap (g_seq:
                // DEVICE exchange counter) and are copied into the level's right-hand side
                rc = spmv_mat(L.R, L.r, nullptr, w.CB[l + 1] + L.offsets[d->rank], 0, st);
                if (!rc) rc = npb_dist_push_slice(d, L.g_data, L.g_flags, L.offsets, L.g_seq, w.CB[l + 1], st);
            } else {
                double* mine = w.CB[l + 1] + L.offsets[d->rank];
                rc = spmv_mat(L.R, L.r, nullptr, mine, 0, st);
                if (!rc) rc = npb_dist_allgatherv(d, mine, w.CB[l + 1], L.offsets, st);
            }
        } else {
            rc = spmv_mat(L.R, L.r, nullptr, w.CB[l + 1], 0, st);
        }
        if (rc) return rc;
        if ((rc = amg_cycle(h, l + 1, w, st))) return rc;
        // x += P x_{l+1}
        if ((rc = spmv_mat(L.P, w.X[l + 1], w.X[l], w.X[l], 2, st))) return rc;
    }
    cur = w.X[l];
    oth = w.X2[l];
    for (int s = 0; s < h->nu_post; ++s) {
        double* out = (l == 0 && s == h->nu_post - 1) ? w.xfine : oth;
        if (out == w.xfine && w.dot_out) {
            int rc = spmv_mat(L.A, cur, bl, out, 3, st, L.dinv, h->omega, w.dot_out, w.red);
            if (rc) return rc;
        } else {
            jacobi_sweep(L, h->omega, bl, cur, out, st);
        }
        if (out == w.xfine) { cur = w.xfine; break; }
        double* t = cur; cur = oth; oth = t;
    }
    if (l == 0) {
        if (cur != w.xfine) NPB_CUDA(cudaMemcpyAsync(w.xfine, cur, n * 8, cudaMemcpyDeviceToDevice, st));
    } else {
        w.X[l] = cur; w.X2[l] = oth;
    }
    return NPB_OK;
}

// x <- one cycle (V, or W on the coarse levels) on b from a zero initial guess.  b and x have
// the fine size and must not alias the hierarchy's work vectors.
// the cycle can return b . x with its last sweep when that sweep is a Jacobi sweep of a
// diagonal-storage operator on one GPU (the preconditioned CG of the pressure solves)
static bool amg_cycle_can_dot(const npb_amg* h) {
    const npb_csr_mat& A0 = h->levels[0].A;
    return h->nlevels > 1 && h->nu_post >= 1 && A0.dia_vals && A0.ell_w == 0 && !A0.halo.dist && !h->dist;
}

static int amg_vcycle_dot(const npb_amg* h, const double* b, double* x, double* dot_out, void* red,
                          cudaStream_t st) {
    const int nl = h->nlevels;
    if (nl > 64) { npb_set_error("amg_vcycle: more than 64 levels"); return NPB_ERR_ARG; }
    cycle_bufs w;
    for (int l = 0; l < nl; ++l) { w.X[l] = h->levels[l].x; w.X2[l] = h->levels[l].x2; w.CB[l] = h->levels[l].b; }
    w.bfine = b;
    w.xfine = x;
    w.dot_out = dot_out;
    w.red = red;
    int rc = amg_cycle(h, 0, w, st);
    if (rc) return rc;
    NPB_CHECK_LAUNCH("amg_vcycle");
    return NPB_OK;
}

int npb_amg_vcycle(const npb_amg* h, const double* b, double* x, void* stream) {
    return amg_vcycle_dot(h, b, x, nullptr, nullptr, ST(stream));
}

// x <- k V-cycles on A x = b (stationary iteration, zero start); r = scratch (fine size)
static int amg_solve(const npb_amg* h, const npb_csr_mat& A, int k, const double* b, double* x,
                     double* r, double* dx, cudaStream_t st) {
    int rc = npb_amg_vcycle(h, b, x, (void*)st);
    for (int it = 1; it < k && !rc; ++it) {
        rc = spmv_mat(A, x, b, r, 1, st);
        if (!rc) rc = npb_amg_vcycle(h, r, dx, (void*)st);
        if (!rc) rc = npb_axpby_launch(A.nrows, 1.0, dx, 1.0, x, st);
    }
    return rc;
}

// x <- k steps of CG on A x = b preconditioned by one V-cycle (A SPD), zero
// start.  r, z, p, q: work vectors; scal: 4 device doubles; red: zeroed
// reduction scratch (npb_reduce_scratch_bytes).  b is not modified.
static int amg_pcg(const npb_amg* h, const npb_csr_mat& A, int k, const double* b, double* x,
                   double* r, double* z, double* p, double* q, double* scal, void* red,
                   cudaStream_t st) {
    const int64_t n = A.nrows;
    int rc;
    // one GPU, diagonal-storage operator: r . z comes out of the last sweep of the cycle, and
    // direction update, operator product and p . q are one kernel (q = L z + beta q)
    const bool fuse_rz = amg_cycle_can_dot(h);
    const bool fuse_dir = A.dia_vals && A.ell_w == 0 && !A.halo.dist && !h->dist && A.nrows == A.ncols;
    for (int i = 0; i < k; ++i) {
        double* rz_new = scal + (i & 1);
        double* rz_old = scal + ((i + 1) & 1);
        const double* rin = (i == 0) ? b : r;          // the first residual is b itself
        if ((rc = amg_vcycle_dot(h, rin, z, fuse_rz ? rz_new : nullptr, red, st))) return rc;
        if (!fuse_rz && (rc = npb_dot_multi_launch(n, 1, rin, n, z, rz_new, red, st))) return rc;
        if ((rc = npb_dist_allreduce_sum((const npb_dist*)h->dist, rz_new, 1, st))) return rc;
        if (fuse_dir) {
            if ((rc = npb_dia_cg_dir_launch(n, A.ndiag, A.dia_ld, A.dia_off, A.dia_vals, z, rz_new, rz_old,
                                            i == 0, p, q, scal + 2, red, st))) return rc;
        } else {
            k_cg_dir<<<npb_blocks(n, TPB), TPB, 0, st>>>(n, rz_new, rz_old, i == 0, z, p);
            NPB_COUNT_LAUNCH();
            if ((rc = spmv_mat(A, p, nullptr, q, 0, st))) return rc;
            if ((rc = npb_dot_multi_launch(n, 1, p, n, q, scal + 2, red, st))) return rc;
        }
        if ((rc = npb_dist_allreduce_sum((const npb_dist*)h->dist, scal + 2, 1, st))) return rc;
        k_cg_update<<<npb_blocks(n, TPB), TPB, 0, st>>>(n, rz_new, scal + 2, p, q, x, rin,
                                                       i == k - 1 ? nullptr : r, i == 0);
        NPB_COUNT_LAUNCH();
        NPB_COUNT_BYTES(8.0 * n * (i == k - 1 ? 3 : 6));
    }
    return NPB_OK;
}

// y = M x as an npb_apply_fn; ctx = npb_csr_mat (row-sharded or not)
int npb_mat_matvec_cb(void* ctx, const double* x, double* y, void* stream) {
    return spmv_mat(*(const npb_csr_mat*)ctx, x, nullptr, y, 0, ST(stream));
}

// only the halo exchange of a sharded operator (tools/shard_micro.py)
int npb_mat_halo_exchange(void* ctx, const double* x, void* stream) {
    const npb_csr_mat& M = *(const npb_csr_mat*)ctx;
    const double* xh = nullptr;
    return M.halo.dist ? npb_dist_halo_exchange(M.halo, x, ST(stream), &xh, nullptr) : NPB_OK;
}

// y = k V-cycles applied to x: usable directly as an npb_apply_fn
struct npb_amg_precond {
    const npb_amg* amg;
    npb_csr_mat A;
    int32_t cycles, pad;
    double* r;     // 2 work vectors of the fine size
    double* dx;
};

int npb_amg_apply_cb(void* ctx, const double* x, double* y, void* stream) {
    const npb_amg_precond* p = (const npb_amg_precond*)ctx;
    return amg_solve(p->amg, p->A, p->cycles, x, y, p->r, p->dx, ST(stream));
}

// least-squares-commutator block preconditioner (see file header)
struct npb_lsc {
    int64_t n, nv, np;
    const int32_t* vset;       // rows with a diagonal (size nv), increasing
    const int32_t* pset;       // rows without        (size np), increasing
    npb_csr_mat A, G, D;       // J[V,V], J[V,P], J[P,V]
    npb_csr_mat Lm;            // -D G
    const double* a_dinv;      // 1 / diag(A)
    const npb_amg* amg_a;
    const npb_amg* amg_l;
    int32_t cycles_a, cycles_l;
    double* wv[5];             // work vectors of size nv
    double* wp[8];             // work vectors of size np
    double* scal;              // 4 device doubles (CG scalars)
    void* red;                 // zeroed reduction scratch
    // robust mode for convection-dominated momentum blocks (a_gmres > 0): the
    // A-solve is a_gmres steps of Jacobi-preconditioned GMRES (cannot diverge,
    // unlike a Jacobi-smoothed V-cycle on a matrix without diagonal dominance)
    int32_t a_gmres;
    int32_t graph_id;          // > 0: the launch sequence may be replayed as a CUDA graph (unique id)
    void* a_work;              // npb_fgmres_workspace_bytes(nv, a_gmres, 0)
    int64_t a_work_bytes;
};

// the host structs above are declared a second time in include/numpad_b200.h (and mirrored
// with ctypes, tests/test_abi.py): a one-sided edit must not compile
static_assert(sizeof(npb_halo) == 64 && sizeof(npb_csr_mat) == 192, "npb_csr_mat layout");
static_assert(sizeof(npb_amg_level) == 656 && sizeof(npb_amg) == 56, "npb_amg layout");
static_assert(sizeof(npb_amg_precond) == 224 && sizeof(npb_lsc) == 984, "npb_lsc layout");

// y = d .* x as an npb_apply_fn (Jacobi preconditioner; ctx = npb_diag_ctx)
struct npb_diag_ctx { int64_t n; const double* d; };
typedef npb_diag_ctx diag_ctx;
int npb_diag_apply_cb(void* ctx, const double* x, double* y, void* stream);
static int diag_scale_cb(void* ctx, const double* x, double* y, void* stream) {
    return npb_diag_apply_cb(ctx, x, y, stream);
}
int npb_diag_apply_cb(void* ctx, const double* x, double* y, void* stream) {
    const diag_ctx* c = (const diag_ctx*)ctx;
    if (c->n > 0) {
        k_diag_scale<<<npb_blocks(c->n, TPB), TPB, 0, ST(stream)>>>(c->n, c->d, x, y);
        NPB_COUNT_LAUNCH();
    }
    return NPB_OK;
}

static int lsc_solve_a(const npb_lsc* c, const double* b, double* x, double* r, double* dx,
                       cudaStream_t st) {
    if (c->a_gmres <= 0) return amg_solve(c->amg_a, c->A, c->cycles_a, b, x, r, dx, st);
    npb_csr_view view{c->A.nrows, c->A.nnz, c->A.indptr, c->A.indices, c->A.data};
    diag_ctx dc{c->nv, c->a_dinv};
    npb_krylov_ops ops{npb_csr_matvec_cb, &view, diag_scale_cb, &dc, nullptr, nullptr};
    npb_krylov_info info;
    NPB_CUDA(cudaMemsetAsync(x, 0, (size_t)c->nv * 8, st));
    int rc = npb_fgmres(c->nv, &ops, b, x, c->a_gmres, c->a_gmres, 1e-3, 0.0, 0, c->a_work,
                        c->a_work_bytes, &info, (void*)st);
    return (rc == NPB_ERR_NOCONV) ? NPB_OK : rc;     // a fixed budget, not a tolerance
}


// the launch sequence between the gathers of r and the scatters into z: fixed buffers, device-
// resident scalars, no host decision -- capturable
static int lsc_core(const npb_lsc* c, cudaStream_t st) {
    // block upper-triangular form:  p = S~^-1 r_P ;  u = A~^-1 (r_V - G p)
    const int64_t np = c->np;
    double *rv = c->wv[0], *us = c->wv[1], *t1 = c->wv[2], *t2 = c->wv[3];
    double *rp = c->wp[0], *q = c->wp[1], *p1 = c->wp[2], *s1 = c->wp[3], *s2 = c->wp[4];
    double *s3 = c->wp[5], *s4 = c->wp[6];
    int rc;
    if (np > 0) {
        // p1 = L^-1 r_P            (L = -DG, so S^-1 ~ -L^-1 (D A G) L^-1)
        // cycles_l: CG steps of the first L-solve in the low 16 bits, of the second one in
        // the high 16 bits (0: the same)
        const int cl1 = c->cycles_l & 0xffff;
        const int cl2 = (c->cycles_l >> 16) ? (c->cycles_l >> 16) : cl1;
        if ((rc = amg_pcg(c->amg_l, c->Lm, cl1, rp, p1, s1, s2, s3, s4, c->scal, c->red, st))) return rc;
        // q = D A G p1
        if ((rc = spmv_mat(c->G, p1, nullptr, t1, 0, st))) return rc;
        if ((rc = spmv_mat(c->A, t1, nullptr, t2, 0, st))) return rc;
        if ((rc = spmv_mat(c->D, t2, nullptr, q, 0, st))) return rc;
        // p = -L^-1 q
        if ((rc = amg_pcg(c->amg_l, c->Lm, cl2, q, p1, s1, s2, s3, s4, c->scal, c->red, st))) return rc;
        if ((rc = npb_axpby_launch(np, 0.0, p1, -1.0, p1, st))) return rc;
        // r_V <- r_V - G p
        if ((rc = spmv_mat(c->G, p1, rv, rv, 1, st))) return rc;
    }
    // u = A~^-1 (r_V - G p)
    return lsc_solve_a(c, rv, us, t1, t2, st);
}

// CUDA graphs of lsc_core, one per preconditioner (npb_lsc::graph_id > 0, unique for the
// lifetime of the process; released by npb_lsc_graph_release).  The first application runs
// eagerly (lazy kernel attributes settle), the second is captured on an internal stream and
// instantiated, later ones are ONE cudaGraphLaunch instead of ~300 kernel launches: the
// coarse-level kernels of the two multigrid hierarchies are shorter than their launch cost.
struct lsc_graph {
    int calls = 0;
    cudaGraphExec_t exec = nullptr;
    unsigned long long launches = 0;
    double bytes = 0.0;
    bool failed = false;
};
static std::map<int, lsc_graph> g_lsc_graphs;
static cudaStream_t g_capture_stream = nullptr;
static int g_graphs_enabled = 1;

int npb_lsc_graphs_enable(int on) { const int old = g_graphs_enabled; g_graphs_enabled = on ? 1 : 0; return old; }

int npb_lsc_graph_release(int graph_id) {
    auto it = g_lsc_graphs.find(graph_id);
    if (it == g_lsc_graphs.end()) return NPB_OK;
    if (it->second.exec) cudaGraphExecDestroy(it->second.exec);
    g_lsc_graphs.erase(it);
    return NPB_OK;
}

int npb_lsc_apply_cb(void* ctx, const double* r, double* z, void* stream) {
    const npb_lsc* c = (const npb_lsc*)ctx;
    cudaStream_t st = ST(stream);
    const int64_t nv = c->nv, np = c->np;
    int rc = NPB_OK;
    k_gather<<<npb_blocks(nv, TPB), TPB, 0, st>>>(nv, c->vset, r, c->wv[0]);
    NPB_COUNT_LAUNCH();
    if (np > 0) {
        k_gather<<<npb_blocks(np, TPB), TPB, 0, st>>>(np, c->pset, r, c->wp[0]);
        NPB_COUNT_LAUNCH();
    }
    // (a row-sharded preconditioner gets a graph id only when EVERY exchange of it goes by direct
    // stores -- device-side exchange counters, no NCCL call inside the sequence; shard.py decides)
    const bool graphable = g_graphs_enabled && c->graph_id > 0 && c->a_gmres <= 0;
    lsc_graph* g = graphable ? &g_lsc_graphs[c->graph_id] : nullptr;
    if (g && !g->failed && g->exec) {
        if (cudaGraphLaunch(g->exec, st) != cudaSuccess) {
            npb_set_error("lsc_apply: cudaGraphLaunch failed: %s", cudaGetErrorString(cudaGetLastError()));
            return NPB_ERR_CUDA;
        }
        npb_launch_counter += g->launches;
        npb_bytes_counter += g->bytes;
    } else if (g && !g->failed && g->calls >= 1) {
        // capture on an internal stream (the caller's may be the legacy default stream, which
        // cannot be captured); nothing executes during capture
        if (!g_capture_stream) NPB_CUDA(cudaStreamCreateWithFlags(&g_capture_stream, cudaStreamNonBlocking));
        const unsigned long long l0 = npb_launch_counter;
        const double b0 = npb_bytes_counter;
        cudaGraph_t graph = nullptr;
        bool ok = cudaStreamBeginCapture(g_capture_stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        if (ok) {
            rc = lsc_core(c, g_capture_stream);
            ok = (cudaStreamEndCapture(g_capture_stream, &graph) == cudaSuccess) && rc == NPB_OK && graph;
        }
        if (ok) ok = cudaGraphInstantiate(&g->exec, graph, 0) == cudaSuccess;
        if (graph) cudaGraphDestroy(graph);
        g->launches = npb_launch_counter - l0;
        g->bytes = npb_bytes_counter - b0;
        npb_launch_counter = l0;
        npb_bytes_counter = b0;
        if (!ok) {                       // fall back to eager launches for this preconditioner
            cudaGetLastError();
            g->failed = true;
            g->exec = nullptr;
            if ((rc = lsc_core(c, st))) return rc;
        } else {
            if (cudaGraphLaunch(g->exec, st) != cudaSuccess) {
                npb_set_error("lsc_apply: cudaGraphLaunch failed: %s", cudaGetErrorString(cudaGetLastError()));
                return NPB_ERR_CUDA;
            }
            npb_launch_counter += g->launches;
            npb_bytes_counter += g->bytes;
        }
    } else {
        if ((rc = lsc_core(c, st))) return rc;
    }
    if (g) g->calls++;
    if (np > 0) {
        k_scatter<<<npb_blocks(np, TPB), TPB, 0, st>>>(np, c->pset, c->wp[2], z);
        NPB_COUNT_LAUNCH();
    }
    k_scatter<<<npb_blocks(nv, TPB), TPB, 0, st>>>(nv, c->vset, c->wv[1], z);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("lsc_apply");
    return NPB_OK;
}

// ------------------------------------------------------------- setup pieces
int npb_csr_diagonal(int64_t nrows, const int32_t* indptr, const int32_t* indices,
                     const double* data, double* diag, void* stream) {
    if (nrows <= 0) return NPB_OK;
    k_csr_diag<<<npb_blocks(nrows, TPB), TPB, 0, ST(stream)>>>(nrows, indptr, indices, data, diag);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("csr_diagonal");
    return NPB_OK;
}

// Aggregation by a maximal independent set of the strength graph.
// state, state2: int8[n] scratch; flag, root_id(n+1), agg: int32; counters:
// int32[2] device.  On return agg[i] in [0, *n_agg) and stats[0] = n_agg.
int npb_amg_aggregate(int64_t n, const int32_t* indptr, const int32_t* indices,
                      const double* data, double theta, int max_rounds, int8_t* state,
                      int8_t* state2, int32_t* flag, int32_t* root_id, int32_t* agg,
                      int32_t* undecided /*device int32*/, int64_t* scan_ws, int64_t* stats,
                      void* stream) {
    cudaStream_t st = ST(stream);
    NPB_CUDA(cudaMemsetAsync(state, 0, n, st));
    int8_t* cur = state;
    int8_t* nxt = state2;
    for (int round = 0; round < max_rounds; ++round) {
        k_mis_select<<<npb_blocks(n, TPB), TPB, 0, st>>>(n, indptr, indices, data, theta, cur, nxt);
        NPB_COUNT_LAUNCH();
        NPB_CUDA(cudaMemsetAsync(undecided, 0, sizeof(int32_t), st));
        k_mis_cover<<<npb_blocks(n, TPB), TPB, 0, st>>>(n, indptr, indices, data, theta, nxt, undecided);
        NPB_COUNT_LAUNCH();
        int8_t* t = cur; cur = nxt; nxt = t;
        int32_t und = 0;
        NPB_CUDA(cudaMemcpyAsync(&und, undecided, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        NPB_CUDA(cudaStreamSynchronize(st));
        if (und == 0) break;
    }
    k_mis_flags<<<npb_blocks(n, TPB), TPB, 0, st>>>(n, cur, 1, flag);
    NPB_COUNT_LAUNCH();
    int rc = npb_scan_launch(flag, root_id, n, scan_ws, stats, st);
    if (rc) return rc;
    k_mis_join<<<npb_blocks(n, TPB), TPB, 0, st>>>(n, indptr, indices, data, flag, root_id, agg);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("amg_aggregate");
    return NPB_OK;
}

// Aggregation by a distance-2 maximal independent set of the strength graph (see the
// kernels above).  keys: 2 n uint64 scratch; state, near: int8[n]; label, label2, flag,
// agg: int32[n]; root_id: int32[n + 1]; undecided: device int32.  stats[0] = n_agg.
int npb_amg_aggregate_mis2(int64_t n, const int32_t* indptr, const int32_t* indices,
                           const double* data, double theta, int max_rounds, int8_t* state,
                           int8_t* near, uint64_t* keys, int32_t* label, int32_t* label2,
                           int32_t* flag, int32_t* root_id, int32_t* agg, int32_t* undecided,
                           int64_t* scan_ws, int64_t* stats, void* stream) {
    cudaStream_t st = ST(stream);
    if (n <= 0) return NPB_ERR_ARG;
    const int nb = npb_blocks(n, TPB);
    unsigned long long* m1 = (unsigned long long*)keys;
    unsigned long long* m2 = m1 + n;
    NPB_CUDA(cudaMemsetAsync(state, 0, n, st));
    for (int round = 0; round < max_rounds; ++round) {
        k_mis2_max<true><<<nb, TPB, 0, st>>>(n, indptr, indices, data, theta, state, nullptr, m1);
        k_mis2_max<false><<<nb, TPB, 0, st>>>(n, indptr, indices, data, theta, state, m1, m2);
        k_mis2_root<<<nb, TPB, 0, st>>>(n, m2, state);
        NPB_CUDA(cudaMemsetAsync(undecided, 0, sizeof(int32_t), st));
        k_mis2_near<true, false><<<nb, TPB, 0, st>>>(n, indptr, indices, data, theta, state, nullptr, near, nullptr);
        k_mis2_near<false, true><<<nb, TPB, 0, st>>>(n, indptr, indices, data, theta, state, near, nullptr, undecided);
        NPB_COUNT_LAUNCH(); NPB_COUNT_LAUNCH(); NPB_COUNT_LAUNCH(); NPB_COUNT_LAUNCH(); NPB_COUNT_LAUNCH();
        int32_t und = 0;
        NPB_CUDA(cudaMemcpyAsync(&und, undecided, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        NPB_CUDA(cudaStreamSynchronize(st));
        if (und == 0) break;
    }
    k_mis2_join<true><<<nb, TPB, 0, st>>>(n, indptr, indices, data, theta, state, nullptr, label);
    k_mis2_join<false><<<nb, TPB, 0, st>>>(n, indptr, indices, data, theta, state, label, label2);
    k_mis2_join<false><<<nb, TPB, 0, st>>>(n, indptr, indices, data, theta, state, label2, label);
    k_mis2_join<false><<<nb, TPB, 0, st>>>(n, indptr, indices, data, theta, state, label, label2);
    k_mis2_flags<<<nb, TPB, 0, st>>>(n, label2, flag);
    NPB_COUNT_LAUNCH(); NPB_COUNT_LAUNCH(); NPB_COUNT_LAUNCH(); NPB_COUNT_LAUNCH(); NPB_COUNT_LAUNCH();
    int rc = npb_scan_launch(flag, root_id, n, scan_ws, stats, st);
    if (rc) return rc;
    k_mis2_ids<<<nb, TPB, 0, st>>>(n, label2, root_id, agg);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("amg_aggregate_mis2");
    return NPB_OK;
}

int npb_csr_filter_symbolic(int64_t nrows, const int32_t* indptr, const int32_t* indices,
                            const double* data, double theta, int32_t* counts,
                            int32_t* indptr_out, int64_t* scan_ws, int64_t* stats, void* stream) {
    NPB_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(int64_t), ST(stream)));
    k_filter<false><<<npb_blocks(nrows, TPB), TPB, 0, ST(stream)>>>(nrows, indptr, indices, data, theta, counts, nullptr, nullptr, nullptr);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("csr_filter_count");
    return npb_scan_launch(counts, indptr_out, nrows, scan_ws, stats, ST(stream));
}

int npb_csr_filter_numeric(int64_t nrows, const int32_t* indptr, const int32_t* indices,
                           const double* data, double theta, const int32_t* indptr_out,
                           int32_t* indices_out, double* out, void* stream) {
    if (nrows <= 0) return NPB_OK;
    k_filter<true><<<npb_blocks(nrows, TPB), TPB, 0, ST(stream)>>>(nrows, indptr, indices, data, theta, nullptr, indptr_out, indices_out, out);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("csr_filter_fill");
    return NPB_OK;
}

int npb_csr_truncate_symbolic(int64_t nrows, const int32_t* indptr, const int32_t* indices,
                              const double* data, double theta, int32_t* counts,
                              int32_t* indptr_out, int64_t* scan_ws, int64_t* stats, void* stream) {
    NPB_CUDA(cudaMemsetAsync(stats, 0, 2 * sizeof(int64_t), ST(stream)));
    k_truncate_rows<false><<<npb_blocks(nrows, TPB), TPB, 0, ST(stream)>>>(nrows, indptr, indices, data, theta, counts, nullptr, nullptr, nullptr);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("csr_truncate_count");
    return npb_scan_launch(counts, indptr_out, nrows, scan_ws, stats, ST(stream));
}

int npb_csr_truncate_numeric(int64_t nrows, const int32_t* indptr, const int32_t* indices,
                             const double* data, double theta, const int32_t* indptr_out,
                             int32_t* indices_out, double* out, void* stream) {
    if (nrows <= 0) return NPB_OK;
    k_truncate_rows<true><<<npb_blocks(nrows, TPB), TPB, 0, ST(stream)>>>(nrows, indptr, indices, data, theta, nullptr, indptr_out, indices_out, out);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("csr_truncate_fill");
    return NPB_OK;
}

// dense inverse of a small CSR matrix: W = scratch n x 2n, Minv = n x n out
int npb_dense_inverse(int n, const int32_t* indptr, const int32_t* indices, const double* data,
                      double* W, double* Minv, int32_t* info, void* stream) {
    cudaStream_t st = ST(stream);
    if (n <= 0) return NPB_OK;
    NPB_CUDA(cudaMemsetAsync(W, 0, sizeof(double) * 2 * (size_t)n * n, st));
    NPB_CUDA(cudaMemsetAsync(info, 0, sizeof(int32_t), st));
    k_csr_to_dense<<<npb_blocks(n, TPB), TPB, 0, st>>>(n, indptr, indices, data, W, 2 * n);
    k_set_identity_right<<<npb_blocks(n, TPB), TPB, 0, st>>>(n, W);
    {
        size_t sh = sizeof(double) * 2 * (size_t)n;
        if (sh > 48 * 1024)
            NPB_CUDA(cudaFuncSetAttribute(k_gauss_jordan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
        k_gauss_jordan<<<1, 1024, sh, st>>>(n, W, info);
    }
    k_copy_right_half<<<npb_blocks((int64_t)n * n, TPB), TPB, 0, st>>>(n, W, Minv);
    NPB_COUNT_LAUNCH(); NPB_COUNT_LAUNCH(); NPB_COUNT_LAUNCH(); NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("dense_inverse");
    return NPB_OK;
}

// CG building blocks with device-resident scalars (used by the multi-GPU
// preconditioner, which interleaves them with NCCL all-reduces of rz / pq)
int npb_cg_dir(int64_t n, const double* rz_new, const double* rz_old, int first,
               const double* z, double* p, void* stream) {
    if (n <= 0) return NPB_OK;
    k_cg_dir<<<npb_blocks(n, TPB), TPB, 0, ST(stream)>>>(n, rz_new, rz_old, first, z, p);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("cg_dir");
    return NPB_OK;
}

int npb_cg_update(int64_t n, const double* rz, const double* pq, const double* p,
                  const double* q, double* x, double* r, int first, void* stream) {
    if (n <= 0) return NPB_OK;
    k_cg_update<<<npb_blocks(n, TPB), TPB, 0, ST(stream)>>>(n, rz, pq, p, q, x, r, r, first);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("cg_update");
    return NPB_OK;
}

int npb_gather(int64_t n, const int32_t* idx, const double* src, double* dst, void* stream) {
    if (n <= 0) return NPB_OK;
    k_gather<<<npb_blocks(n, TPB), TPB, 0, ST(stream)>>>(n, idx, src, dst);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("gather");
    return NPB_OK;
}

}  // extern "C"
