// Block-tridiagonal direct solver over the level sets of the matrix graph.
//
// Seam replaced: `splinalg.spsolve(J, rhs)` (SuperLU, adsolve.py:193-195) for the systems on
// which preconditioned Krylov fails: convection-dominated saddle points (the reference's own
// march from rest, test/cavity.py:97-114) and the compressible-flow Jacobians at the scripts'
// time steps (test/euler_kec.py:223-257, test/navierstokes.py:262-296).
//
// A breadth-first search from a pseudo-peripheral node splits the unknowns into level sets;
// a level only couples to its two neighbours, so in level order the matrix is block
// tridiagonal, (L_k | D_k | U_k) per block row, with blocks of ~3N unknowns on an N x N grid.
// Block LU:  S_0 = D_0,  S_k = D_k - L_k S_{k-1}^{-1} U_{k-1};  the dense S_k^{-1} are kept
// (sum of m_k^2 doubles: 51 GB for the 1024^2 cavity, O(N^3)), so that a solve is two sweeps of
// dense matrix-vector products -- HBM-bound, no triangular dependency chains.  The kernels
// here are the sparse side: dense images of windows of the CSR matrix, window x dense products
// (L_k S^{-1} and its product with U_{k-1} are combinations of a few dense rows each, not
// GEMMs) and the dense matrix-vector product; the dense inversions are library calls.
#include <algorithm>
#include <vector>

#include "common.cuh"

#define ST(s) ((cudaStream_t)(s))

int npb_axpby_launch(int64_t n, double a, const double* x, double b, double* y, cudaStream_t st);

namespace {

constexpr int BT_T = 256;

// out[(r - r0) * ld + (c - c0)] += alpha * A[r, c] for r in [r0, r1), c in [c0, c1)
__global__ void __launch_bounds__(BT_T)
k_bt_window_to_dense(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                     const double* __restrict__ dat, int64_t r0, int64_t r1, int32_t c0, int32_t c1,
                     double alpha, double* __restrict__ out, int64_t ld) {
    const int lane = threadIdx.x & 7;
    const int64_t r = r0 + (((int64_t)blockIdx.x * BT_T + threadIdx.x) >> 3);
    if (r >= r1) return;
    for (int32_t k = ptr[r] + lane, e = ptr[r + 1]; k < e; k += 8) {
        const int32_t c = idx[k];
        if (c >= c0 && c < c1) out[(r - r0) * ld + (c - c0)] += alpha * dat[k];
    }
}

// C[r - r0][j] = beta * C[r - r0][j] + alpha * sum_{c in [c0, c1)} A[r, c] * B[c - c0][j]
// one CTA row-tile: blockIdx.y = row, blockIdx.x = tile of BT_T * 2 columns
constexpr int BT_MAXE = 64;
__global__ void __launch_bounds__(BT_T)
k_bt_window_spmm(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                 const double* __restrict__ dat, int64_t r0, int32_t c0, int32_t c1,
                 const double* __restrict__ B, int64_t ldb, int64_t ncols, double alpha,
                 double beta, double* __restrict__ C, int64_t ldc) {
    __shared__ int32_t sc[BT_MAXE];
    __shared__ double sv[BT_MAXE];
    __shared__ int cnt;
    const int64_t r = r0 + blockIdx.y;
    const int32_t b = ptr[r], e = ptr[r + 1];
    const int64_t j0 = (int64_t)blockIdx.x * (BT_T * 2) + threadIdx.x;
    double acc0 = 0.0, acc1 = 0.0;
    for (int32_t base = b; base < e; base += BT_MAXE) {
        __syncthreads();
        if (threadIdx.x == 0) {
            int n = 0;
            for (int32_t k = base; k < e && k < base + BT_MAXE; ++k) {
                const int32_t c = idx[k];
                if (c >= c0 && c < c1) { sc[n] = c - c0; sv[n] = dat[k]; ++n; }
            }
            cnt = n;
        }
        __syncthreads();
        const int n = cnt;
        for (int t = 0; t < n; ++t) {
            const double* brow = B + (int64_t)sc[t] * ldb;
            if (j0 < ncols) acc0 = fma(sv[t], brow[j0], acc0);
            if (j0 + BT_T < ncols) acc1 = fma(sv[t], brow[j0 + BT_T], acc1);
        }
    }
    double* crow = C + (int64_t)blockIdx.y * ldc;
    if (j0 < ncols) crow[j0] = (beta == 0.0 ? 0.0 : beta * crow[j0]) + alpha * acc0;
    if (j0 + BT_T < ncols) crow[j0 + BT_T] = (beta == 0.0 ? 0.0 : beta * crow[j0 + BT_T]) + alpha * acc1;
}

// y[r - r0] = beta * y0[r - r0] + alpha * sum_{c in [c0, c1)} A[r, c] * x[c - c0]   (y0 may be y)
__global__ void __launch_bounds__(BT_T)
k_bt_window_spmv(const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                 const double* __restrict__ dat, int64_t r0, int64_t r1, int32_t c0, int32_t c1,
                 const double* __restrict__ x, double alpha, double beta, const double* y0,
                 double* y) {
    const int lane = threadIdx.x & 7;
    const int64_t r = r0 + (((int64_t)blockIdx.x * BT_T + threadIdx.x) >> 3);
    double s = 0.0;
    if (r < r1)
        for (int32_t k = ptr[r] + lane, e = ptr[r + 1]; k < e; k += 8) {
            const int32_t c = idx[k];
            if (c >= c0 && c < c1) s = fma(dat[k], x[c - c0], s);
        }
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    if (r < r1 && lane == 0) y[r - r0] = (beta == 0.0 ? 0.0 : beta * y0[r - r0]) + alpha * s;
}

// y = beta * y0 + alpha * M x, M dense m x m row-major: one warp per row, 128-bit loads
__global__ void __launch_bounds__(BT_T)
k_bt_gemv(int m, const double* __restrict__ M, int64_t ld, const double* __restrict__ x,
          double alpha, double beta, const double* __restrict__ y0, double* __restrict__ y) {
    const int lane = threadIdx.x & 31;
    const int64_t r = ((int64_t)blockIdx.x * BT_T + threadIdx.x) >> 5;
    if (r >= m) return;
    const double* row = M + r * ld;
    double s0 = 0.0, s1 = 0.0;
    if ((((uintptr_t)row | (uintptr_t)x) & 15) == 0) {
        const double2* row2 = (const double2*)row;
        const double2* x2 = (const double2*)x;
        const int m2 = m >> 1;
        int j = lane;
        for (; j + 32 < m2; j += 64) {
            const double2 a = __ldcs(row2 + j), b = __ldcs(row2 + j + 32);
            const double2 u = x2[j], v = x2[j + 32];
            s0 = fma(a.x, u.x, s0); s0 = fma(a.y, u.y, s0);
            s1 = fma(b.x, v.x, s1); s1 = fma(b.y, v.y, s1);
        }
        for (; j < m2; j += 32) {
            const double2 a = __ldcs(row2 + j);
            const double2 u = x2[j];
            s0 = fma(a.x, u.x, s0); s0 = fma(a.y, u.y, s0);
        }
        if ((m & 1) && lane == 0) s1 = fma(row[m - 1], x[m - 1], s1);
    } else {
        for (int j = lane; j < m; j += 32) s0 = fma(row[j], x[j], s0);
    }
    double s = s0 + s1;
#pragma unroll
    for (int d = 16; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (lane == 0) y[r] = (beta == 0.0 ? 0.0 : beta * y0[r]) + alpha * s;
}

// y = beta * y0 + alpha * M^T x (column combination): thread per output column, rows strided
// over blockIdx.y with a shared-memory tile of partial sums
constexpr int GT_ROWS = 64;
__global__ void __launch_bounds__(BT_T)
k_bt_gemv_t_partial(int m, const double* __restrict__ M, int64_t ld, const double* __restrict__ x,
                    double* __restrict__ part /* gridDim.y x m */) {
    const int64_t j = (int64_t)blockIdx.x * BT_T + threadIdx.x;
    const int r0 = blockIdx.y * GT_ROWS;
    const int r1 = min(m, r0 + GT_ROWS);
    if (j >= m) return;
    double s = 0.0;
    for (int r = r0; r < r1; ++r) s = fma(__ldcs(M + (int64_t)r * ld + j), x[r], s);
    part[(int64_t)blockIdx.y * m + j] = s;
}
__global__ void __launch_bounds__(BT_T)
k_bt_gemv_t_final(int m, int nparts, const double* __restrict__ part, double alpha, double beta,
                  const double* __restrict__ y0, double* __restrict__ y) {
    const int64_t j = (int64_t)blockIdx.x * BT_T + threadIdx.x;
    if (j >= m) return;
    double s = 0.0;
    for (int p = 0; p < nparts; ++p) s += part[(int64_t)p * m + j];
    y[j] = (beta == 0.0 ? 0.0 : beta * y0[j]) + alpha * s;
}

}  // namespace

extern "C" {

// Level sets of the graph of A + A^T by breadth-first search from a pseudo-peripheral node
// (two exploratory sweeps per connected component).  All arrays on the HOST.  level[i] in
// [0, *nlevels); components are numbered one after the other.
int npb_bfs_levels_host(int64_t n, const int32_t* ptr, const int32_t* idx, const int32_t* ptr_t,
                        const int32_t* idx_t, int32_t* level, int32_t* nlevels) {
    if (n < 0) return NPB_ERR_ARG;
    std::vector<int32_t> mark((size_t)n, -1), order, frontier, next;
    for (int64_t i = 0; i < n; ++i) level[i] = -1;
    int32_t base = 0;
    // sweep from `root`, writing levels into `lev` starting at `b`; returns (last node, depth)
    auto sweep = [&](int32_t root, int32_t* lev, int32_t b, std::vector<int32_t>& visited,
                     int32_t& depth) -> int32_t {
        frontier.assign(1, root);
        lev[root] = b;
        visited.push_back(root);
        int32_t d = b, last = root;
        while (!frontier.empty()) {
            next.clear();
            for (int32_t v : frontier) {
                for (int pass = 0; pass < 2; ++pass) {
                    const int32_t* p = pass ? ptr_t : ptr;
                    const int32_t* ix = pass ? idx_t : idx;
                    if (!p) continue;
                    for (int32_t k = p[v]; k < p[v + 1]; ++k) {
                        const int32_t w = ix[k];
                        if (lev[w] < 0) { lev[w] = d + 1; next.push_back(w); visited.push_back(w); }
                    }
                }
            }
            if (!next.empty()) { ++d; last = next.back(); }
            frontier.swap(next);
        }
        depth = d - b + 1;
        return last;
    };
    std::vector<int32_t> visited;
    for (int64_t s = 0; s < n; ++s) {
        if (level[s] >= 0) continue;
        int32_t depth = 0, root = (int32_t)s;
        for (int trial = 0; trial < 2; ++trial) {
            visited.clear();
            const int32_t far = sweep(root, mark.data(), 0, visited, depth);
            for (int32_t v : visited) mark[v] = -1;
            root = far;
        }
        visited.clear();
        sweep(root, level, base, visited, depth);
        base += depth;
    }
    *nlevels = base;
    return NPB_OK;
}

// Two-sided level structure for the twisted factorisation: BFS balls around a pseudo-peripheral
// node A (levels 0 .. TA-1) and around the node B farthest from it (levels K-1 down to TA+1),
// and ONE middle level TA holding everything in neither ball.  With TA + TB <= dist(A, B) the two
// balls are disjoint and not adjacent, so the matrix is block tridiagonal in this order too, and
// both elimination chains run over growing BFS balls -- whose leading blocks are non-singular
// for saddle points as long as the seed has a non-zero diagonal (has_diag, may be NULL): a seed
// without one is replaced by a neighbour that has.  Connected graphs only (returns
// NPB_ERR_ARG otherwise: the caller uses the one-sided levels).  All arrays on the HOST.
int npb_bfs_levels_two_sided_host(int64_t n, const int32_t* ptr, const int32_t* idx,
                                  const int32_t* ptr_t, const int32_t* idx_t,
                                  const unsigned char* has_diag, int32_t* level,
                                  int32_t* nlevels, int32_t* middle) {
    if (n <= 0) return NPB_ERR_ARG;
    std::vector<int32_t> dist((size_t)n), frontier, next;
    auto bfs = [&](int32_t root, std::vector<int32_t>& d, int64_t& visited) -> int32_t {
        std::fill(d.begin(), d.end(), -1);
        frontier.assign(1, root);
        d[root] = 0;
        visited = 1;
        int32_t last = root, lev = 0;
        while (!frontier.empty()) {
            next.clear();
            for (int32_t v : frontier)
                for (int pass = 0; pass < 2; ++pass) {
                    const int32_t* p = pass ? ptr_t : ptr;
                    const int32_t* ix = pass ? idx_t : idx;
                    if (!p) continue;
                    for (int32_t k = p[v]; k < p[v + 1]; ++k) {
                        const int32_t w = ix[k];
                        if (d[w] < 0) { d[w] = lev + 1; next.push_back(w); ++visited; }
                    }
                }
            if (!next.empty()) { ++lev; last = next.back(); }
            frontier.swap(next);
        }
        return last;
    };
    auto with_diag = [&](int32_t v) -> int32_t {
        if (!has_diag || has_diag[v]) return v;
        for (int pass = 0; pass < 2; ++pass) {
            const int32_t* p = pass ? ptr_t : ptr;
            const int32_t* ix = pass ? idx_t : idx;
            if (!p) continue;
            for (int32_t k = p[v]; k < p[v + 1]; ++k)
                if (has_diag[ix[k]]) return ix[k];
        }
        return v;
    };
    int64_t vis = 0;
    int32_t a = bfs(0, dist, vis);
    if (vis != n) return NPB_ERR_ARG;                 // not connected
    a = bfs(a, dist, vis);                            // two exploratory sweeps
    a = with_diag(a);
    std::vector<int32_t> da((size_t)n), db((size_t)n);
    int32_t b = with_diag(bfs(a, da, vis));
    bfs(b, db, vis);
    const int32_t D = da[b];
    if (D < 4) return NPB_ERR_ARG;
    // ball sizes by radius
    std::vector<int64_t> ca((size_t)D + 2, 0), cb((size_t)D + 2, 0);
    for (int64_t i = 0; i < n; ++i) {
        if (da[i] <= D) ++ca[da[i] + 1];
        if (db[i] <= D) ++cb[db[i] + 1];
    }
    for (int32_t r = 1; r <= D + 1; ++r) { ca[r] += ca[r - 1]; cb[r] += cb[r - 1]; }   // c[r] = #{d < r}
    int32_t TA = 1;
    int64_t best = -1;
    for (int32_t t = 1; t <= D - 1; ++t) {            // TB = D - t >= 1
        const int64_t diff = ca[t] > cb[D - t] ? ca[t] - cb[D - t] : cb[D - t] - ca[t];
        if (best < 0 || diff < best) { best = diff; TA = t; }
    }
    const int32_t TB = D - TA;
    for (int64_t i = 0; i < n; ++i) {
        if (da[i] < TA) level[i] = da[i];
        else if (db[i] < TB) level[i] = TA + 1 + (TB - 1 - db[i]);
        else level[i] = TA;
    }
    *nlevels = TA + TB + 1;
    *middle = TA;
    return NPB_OK;
}

int npb_bt_window_to_dense(const int32_t* indptr, const int32_t* indices, const double* data,
                           int64_t r0, int64_t r1, int64_t c0, int64_t c1, double alpha,
                           double* out, int64_t ld, void* stream) {
    if (r1 <= r0) return NPB_OK;
    k_bt_window_to_dense<<<npb_blocks((r1 - r0) * 8, BT_T), BT_T, 0, ST(stream)>>>(
        indptr, indices, data, r0, r1, (int32_t)c0, (int32_t)c1, alpha, out, ld);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("bt_window_to_dense");
    return NPB_OK;
}

int npb_bt_window_spmm(const int32_t* indptr, const int32_t* indices, const double* data,
                       int64_t r0, int64_t r1, int64_t c0, int64_t c1, const double* B,
                       int64_t ldb, int64_t ncols, double alpha, double beta, double* C,
                       int64_t ldc, void* stream) {
    if (r1 <= r0 || ncols <= 0) return NPB_OK;
    if (r1 - r0 > 65535) { npb_set_error("bt_window_spmm: more than 65535 rows in a level"); return NPB_ERR_ARG; }
    dim3 grid((unsigned)((ncols + 2 * BT_T - 1) / (2 * BT_T)), (unsigned)(r1 - r0));
    k_bt_window_spmm<<<grid, BT_T, 0, ST(stream)>>>(indptr, indices, data, r0, (int32_t)c0,
                                                    (int32_t)c1, B, ldb, ncols, alpha, beta, C, ldc);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("bt_window_spmm");
    return NPB_OK;
}

int npb_bt_window_spmv(const int32_t* indptr, const int32_t* indices, const double* data,
                       int64_t r0, int64_t r1, int64_t c0, int64_t c1, const double* x,
                       double alpha, double beta, const double* y0, double* y, void* stream) {
    if (r1 <= r0) return NPB_OK;
    k_bt_window_spmv<<<npb_blocks((r1 - r0) * 8, BT_T), BT_T, 0, ST(stream)>>>(
        indptr, indices, data, r0, r1, (int32_t)c0, (int32_t)c1, x, alpha, beta, y0 ? y0 : y, y);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("bt_window_spmv");
    return NPB_OK;
}

// y = beta * y0 + alpha * op(M) x for a dense m x m row-major M (op = transpose when trans);
// y may alias y0; scratch: ceil(m / 64) * m doubles (transposed product only)
int64_t npb_bt_gemv_scratch_doubles(int64_t m) { return ((m + GT_ROWS - 1) / GT_ROWS) * m; }

int npb_bt_gemv(int64_t m, const double* M, int64_t ld, const double* x, double alpha,
                double beta, const double* y0, double* y, int trans, double* scratch,
                void* stream) {
    if (m <= 0) return NPB_OK;
    if (!trans) {
        k_bt_gemv<<<npb_blocks(m * 32, BT_T), BT_T, 0, ST(stream)>>>((int)m, M, ld, x, alpha, beta, y0, y);
        NPB_COUNT_LAUNCH();
    } else {
        const int np = (int)((m + GT_ROWS - 1) / GT_ROWS);
        dim3 grid((unsigned)npb_blocks(m, BT_T), (unsigned)np);
        k_bt_gemv_t_partial<<<grid, BT_T, 0, ST(stream)>>>((int)m, M, ld, x, scratch);
        k_bt_gemv_t_final<<<npb_blocks(m, BT_T), BT_T, 0, ST(stream)>>>((int)m, np, scratch, alpha, beta, y0, y);
        NPB_COUNT_LAUNCH(); NPB_COUNT_LAUNCH();
    }
    NPB_COUNT_BYTES(8.0 * (double)m * (double)m + 24.0 * (double)m);
    NPB_CHECK_LAUNCH("bt_gemv");
    return NPB_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------- solve phase in one call
// x = A^-1 b (trans = 0) or A^-T b (trans = 1) from the stored inverses: forward sweep
//   c_k = b_k - (lower block) z_{k-1},  z_k = S_k^-1 c_k,
// backward sweep  z_k -= S_k^-1 (upper block) z_{k+1}.  All arrays of the sweep live in
// buffers owned by the factorisation (bp: permuted right-hand side, z: permuted solution,
// c / scratch: work), so the launch sequence -- ~8 K small kernels for a 1024^2 cavity -- is
// the same on every call: after one eager call it is captured into a CUDA graph (graph_id > 0,
// one graph per direction) and replayed; the sweeps are then bound by the 2 x (sum m_k^2) x 8
// bytes of the inverses they stream, not by launches.
#include <map>

namespace {
struct bt_graph { int calls = 0; cudaGraphExec_t exec = nullptr; bool failed = false; unsigned long long launches = 0; double bytes = 0.0; };
std::map<long long, bt_graph> g_bt_graphs;
cudaStream_t g_bt_capture = nullptr;


// Twisted sweeps: the factorisation ran from BOTH ends towards level M (S_k for k < M from the
// top, for k > M from the bottom, S_M joins them; M = K - 1 is the plain one-chain form).  The
// two chains of each sweep are independent: with st2 != st they run on two streams (two
// branches of the captured graph), each with its own work buffers.
int bt_chain_fwd(int k0, int k1, int dir, const int64_t* off, const double* const* sinv,
                 const int32_t* ip, const int32_t* ix, const double* dat, const double* bp, double* z,
                 double* c, double* scratch, int trans, cudaStream_t st) {
    // k = k0, k0 + dir, ... (k1 exclusive): c_k = b_k - (block towards the chain's start) z_prev
    int rc = NPB_OK;
    for (int k = k0; k != k1 && !rc; k += dir) {
        const int64_t r0 = off[k], r1 = off[k + 1], m = r1 - r0;
        const double* src = bp + r0;
        if (k != k0) {
            const int p = k - dir;                    // previous level of this chain
            rc = npb_bt_window_spmv(ip, ix, dat, r0, r1, off[p], off[p + 1], z + off[p], -1.0, 1.0, bp + r0, c, st);
            src = c;
        }
        if (!rc) rc = npb_bt_gemv(m, sinv[k], m, src, 1.0, 0.0, nullptr, z + r0, trans, scratch, st);
    }
    return rc;
}

int bt_chain_bwd(int k0, int k1, int dir, const int64_t* off, const double* const* sinv,
                 const int32_t* ip, const int32_t* ix, const double* dat, double* z, double* c,
                 double* scratch, int trans, cudaStream_t st) {
    // k = k0, k0 + dir, ... (k1 exclusive), away from the middle: z_k -= S_k^-1 (block towards
    // the middle) z_{k - dir}
    int rc = NPB_OK;
    for (int k = k0; k != k1 && !rc; k += dir) {
        const int64_t r0 = off[k], r1 = off[k + 1], m = r1 - r0;
        const int p = k - dir;
        rc = npb_bt_window_spmv(ip, ix, dat, r0, r1, off[p], off[p + 1], z + off[p], 1.0, 0.0, nullptr, c, st);
        if (!rc) rc = npb_bt_gemv(m, sinv[k], m, c, -1.0, 1.0, z + r0, z + r0, trans, scratch, st);
    }
    return rc;
}

struct bt_fork { cudaStream_t st2; cudaEvent_t e_fork, e_join; };
bt_fork g_bt_fork = {};

int bt_sweeps(int K, int M, const int64_t* off, const double* const* sinv, const int32_t* ip,
              const int32_t* ix, const double* dat, const double* bp, double* z, double* c,
              double* c2, double* scratch, double* scratch2, int trans, cudaStream_t st,
              const bt_fork* fk) {
    int rc = NPB_OK;
    const bool two = fk && M < K - 1;
    cudaStream_t sb = two ? fk->st2 : st;
    // forward: chain A = levels 0 .. M-1 downwards, chain B = levels K-1 .. M+1 upwards
    if (two) { NPB_CUDA(cudaEventRecord(fk->e_fork, st)); NPB_CUDA(cudaStreamWaitEvent(sb, fk->e_fork, 0)); }
    rc = bt_chain_fwd(0, M, +1, off, sinv, ip, ix, dat, bp, z, c, scratch, trans, st);
    if (!rc) rc = bt_chain_fwd(K - 1, M, -1, off, sinv, ip, ix, dat, bp, z, two ? c2 : c, two ? scratch2 : scratch, trans, sb);
    if (two) { NPB_CUDA(cudaEventRecord(fk->e_join, sb)); NPB_CUDA(cudaStreamWaitEvent(st, fk->e_join, 0)); }
    if (rc) return rc;
    {   // middle level: c = b_M - (lower) z_{M-1} - (upper) z_{M+1};  z_M = S_M^-1 c
        const int64_t r0 = off[M], r1 = off[M + 1], m = r1 - r0;
        const double* src = bp + r0;
        if (M > 0) {
            rc = npb_bt_window_spmv(ip, ix, dat, r0, r1, off[M - 1], r0, z + off[M - 1], -1.0, 1.0, src, c, st);
            src = c;
        }
        if (!rc && M < K - 1) {
            rc = npb_bt_window_spmv(ip, ix, dat, r0, r1, r1, off[M + 2], z + r1, -1.0, 1.0, src, c, st);
            src = c;
        }
        if (!rc) rc = npb_bt_gemv(m, sinv[M], m, src, 1.0, 0.0, nullptr, z + r0, trans, scratch, st);
        if (rc) return rc;
    }
    // backward, away from the middle
    if (two) { NPB_CUDA(cudaEventRecord(fk->e_fork, st)); NPB_CUDA(cudaStreamWaitEvent(sb, fk->e_fork, 0)); }
    rc = bt_chain_bwd(M - 1, -1, -1, off, sinv, ip, ix, dat, z, c, scratch, trans, st);
    if (!rc) rc = bt_chain_bwd(M + 1, K, +1, off, sinv, ip, ix, dat, z, two ? c2 : c, two ? scratch2 : scratch, trans, sb);
    if (two) { NPB_CUDA(cudaEventRecord(fk->e_join, sb)); NPB_CUDA(cudaStreamWaitEvent(st, fk->e_join, 0)); }
    return rc;
}
}  // namespace

extern "C" {

int npb_bt_graph_release(int graph_id) {
    for (int t = 0; t < 2; ++t) {
        auto it = g_bt_graphs.find((long long)graph_id * 2 + t);
        if (it == g_bt_graphs.end()) continue;
        if (it->second.exec) cudaGraphExecDestroy(it->second.exec);
        g_bt_graphs.erase(it);
    }
    return NPB_OK;
}

// off: [host] K + 1 level offsets; sinv: [host] K device pointers (row-major m_k x m_k);
// (indptr, indices, data): the level-ordered matrix (its transpose when trans); M: the level
// the two chains of the factorisation meet at (K - 1: one chain); c, scratch: 2 x the longest
// level / 2 x npb_bt_gemv_scratch_doubles(longest level) doubles (one half per chain)
int npb_bt_solve(int K, int M, const int64_t* off, const double* const* sinv, const int32_t* indptr,
                 const int32_t* indices, const double* data, const double* bp, double* z,
                 double* c, double* scratch, int64_t mmax, int trans, int graph_id, void* stream) {
    cudaStream_t st = ST(stream);
    if (K <= 0) return NPB_OK;
    if (M < 0 || M >= K) return NPB_ERR_ARG;
    double* c2 = c + mmax;
    double* scratch2 = scratch + npb_bt_gemv_scratch_doubles(mmax);
    bt_graph* g = graph_id > 0 ? &g_bt_graphs[(long long)graph_id * 2 + (trans ? 1 : 0)] : nullptr;
    if (g && !g->failed && !g->exec && g->calls >= 1) {
        if (!g_bt_capture) {
            NPB_CUDA(cudaStreamCreateWithFlags(&g_bt_capture, cudaStreamNonBlocking));
            NPB_CUDA(cudaStreamCreateWithFlags(&g_bt_fork.st2, cudaStreamNonBlocking));
            NPB_CUDA(cudaEventCreateWithFlags(&g_bt_fork.e_fork, cudaEventDisableTiming));
            NPB_CUDA(cudaEventCreateWithFlags(&g_bt_fork.e_join, cudaEventDisableTiming));
        }
        const unsigned long long l0 = npb_launch_counter;
        const double b0 = npb_bytes_counter;
        cudaGraph_t graph = nullptr;
        bool ok = cudaStreamBeginCapture(g_bt_capture, cudaStreamCaptureModeThreadLocal) == cudaSuccess;
        int rc = NPB_OK;
        if (ok) {
            // the second stream joins the capture through the fork event: two graph branches
            rc = bt_sweeps(K, M, off, sinv, indptr, indices, data, bp, z, c, c2, scratch, scratch2, trans,
                           g_bt_capture, &g_bt_fork);
            ok = (cudaStreamEndCapture(g_bt_capture, &graph) == cudaSuccess) && rc == NPB_OK && graph;
        }
        if (ok) ok = cudaGraphInstantiate(&g->exec, graph, 0) == cudaSuccess;
        if (graph) cudaGraphDestroy(graph);
        g->launches = npb_launch_counter - l0;
        g->bytes = npb_bytes_counter - b0;
        npb_launch_counter = l0;
        npb_bytes_counter = b0;
        if (!ok) { cudaGetLastError(); g->failed = true; g->exec = nullptr; }
    }
    if (g) g->calls++;
    if (g && g->exec) {
        if (cudaGraphLaunch(g->exec, st) != cudaSuccess) {
            npb_set_error("bt_solve: cudaGraphLaunch failed: %s", cudaGetErrorString(cudaGetLastError()));
            return NPB_ERR_CUDA;
        }
        npb_launch_counter += g->launches;
        npb_bytes_counter += g->bytes;
        return NPB_OK;
    }
    return bt_sweeps(K, M, off, sinv, indptr, indices, data, bp, z, c, c2, scratch, scratch2, trans, st, nullptr);
}

}  // extern "C"
