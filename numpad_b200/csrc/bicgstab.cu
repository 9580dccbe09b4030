// BiCGStab (right-preconditioned, host-driven) and a block-Jacobi preconditioner.
//
// Seam: the same as gmres.cu -- `splinalg.spsolve(J, rhs)` of the reference Newton loop
// (adsolve.py:193-195) -- for Jacobians with a full diagonal whose coupling is dominated by
// small diagonal blocks (the 4 conserved variables of a cell in test/euler_kec.py /
// test/navierstokes.py at moderate CFL): short recurrences, 8 work vectors instead of a Krylov
// basis.  Operator, preconditioner and cross-rank sum are the callbacks of npb_krylov_ops, so
// the row-sharded operator and every preconditioner of the C-ABI plug in unchanged.
//
// Block Jacobi: the unknowns of block c are rows c * cs + k * ks (k < bs): cs = 1, ks = nblk
// for the reference's variable-major state arrays w[k, i, j]; cs = bs, ks = 1 for interleaved
// unknowns.  Setup gathers the bs x bs diagonal blocks from the CSR rows and inverts them
// (Gauss-Jordan with partial pivoting, one thread per block, bs <= 8); apply is one small
// dense mat-vec per block.
#include <math.h>
#include <string.h>

#include "common.cuh"

int npb_dot_multi_launch(int64_t n, int k, const double* V, int64_t ld, const double* w,
                         double* out, void* scratch, cudaStream_t st);
int npb_axpby_launch(int64_t n, double a, const double* x, double b, double* y, cudaStream_t st);
extern "C" int64_t npb_reduce_scratch_bytes(void);

namespace {

constexpr int TPB = 256;
constexpr int BJ_MAX = 8;

// y = a x + b y + c z
__global__ void __launch_bounds__(TPB)
k_axpbypcz(int64_t n, double a, const double* __restrict__ x, double b, double* y, double c,
           const double* __restrict__ z) {
    for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < n; i += (int64_t)gridDim.x * TPB)
        y[i] = fma(a, x[i], fma(c, z[i], b * y[i]));
}

__global__ void __launch_bounds__(TPB)
k_bj_setup(int64_t nblk, int bs, int64_t cs, int64_t ks, const int32_t* __restrict__ ptr,
           const int32_t* __restrict__ idx, const double* __restrict__ dat,
           double* __restrict__ dinv, int32_t* __restrict__ info) {
    const int64_t c = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (c >= nblk) return;
    double W[BJ_MAX][2 * BJ_MAX];
    for (int k = 0; k < bs; ++k)
        for (int j = 0; j < 2 * bs; ++j) W[k][j] = (j == bs + k) ? 1.0 : 0.0;
    for (int k = 0; k < bs; ++k) {
        const int64_t r = c * cs + (int64_t)k * ks;
        for (int32_t e = ptr[r], e1 = ptr[r + 1]; e < e1; ++e) {
            const int64_t d = (int64_t)idx[e] - c * cs;          // = j * ks for a block member
            if (d < 0) continue;
            const int64_t j = d / ks;
            if (j < bs && j * ks == d) W[k][j] += dat[e];
        }
    }
    bool singular = false;
    for (int k = 0; k < bs; ++k) {
        int p = k;
        double best = fabs(W[k][k]);
        for (int r = k + 1; r < bs; ++r)
            if (fabs(W[r][k]) > best) { best = fabs(W[r][k]); p = r; }
        if (best == 0.0) { singular = true; break; }
        if (p != k)
            for (int j = 0; j < 2 * bs; ++j) { const double t = W[k][j]; W[k][j] = W[p][j]; W[p][j] = t; }
        const double rinv = 1.0 / W[k][k];
        for (int j = 0; j < 2 * bs; ++j) W[k][j] *= rinv;
        for (int r = 0; r < bs; ++r) {
            if (r == k) continue;
            const double f = W[r][k];
            if (f == 0.0) continue;
            for (int j = 0; j < 2 * bs; ++j) W[r][j] = fma(-f, W[k][j], W[r][j]);
        }
    }
    if (singular) atomicAdd(info, 1);
    // a singular block falls back to the identity (the caller sees the count)
    for (int k = 0; k < bs; ++k)
        for (int j = 0; j < bs; ++j)
            dinv[((int64_t)(k * bs + j)) * nblk + c] = singular ? (k == j ? 1.0 : 0.0) : W[k][bs + j];
}

// y[row(c, k)] = sum_j Dinv[c][k][j] x[row(c, j)]; dinv stored entry-major (coalesced over c)
__global__ void __launch_bounds__(TPB)
k_bj_apply(int64_t nblk, int bs, int64_t cs, int64_t ks, const double* __restrict__ dinv,
           const double* __restrict__ x, double* __restrict__ y) {
    const int64_t c = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (c >= nblk) return;
    double xv[BJ_MAX];
    for (int j = 0; j < bs; ++j) xv[j] = x[c * cs + (int64_t)j * ks];
    for (int k = 0; k < bs; ++k) {
        double s = 0.0;
        for (int j = 0; j < bs; ++j) s = fma(dinv[((int64_t)(k * bs + j)) * nblk + c], xv[j], s);
        y[c * cs + (int64_t)k * ks] = s;
    }
}

int grid_of(int64_t n) {
    int64_t b = (n + TPB - 1) / TPB;
    const int64_t cap = (int64_t)NPB_NUM_SMS * 8;
    if (b > cap) b = cap;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace

extern "C" {

typedef int (*npb_apply_fn)(void* ctx, const double* x, double* y, void* stream);
typedef int (*npb_allreduce_fn)(void* ctx, double* buf, int count, void* stream);
struct npb_krylov_ops {
    npb_apply_fn matvec; void* matvec_ctx;
    npb_apply_fn precond; void* precond_ctx;
    npb_allreduce_fn allreduce; void* allreduce_ctx;
};
struct npb_krylov_info { int32_t iters, restarts, status, pad; double bnorm, rnorm0, rnorm; };

struct npb_block_jacobi {
    int64_t nblk;
    int32_t bs, pad;
    int64_t cs, ks;
    const double* dinv;         // bs * bs * nblk doubles, entry (k, j) of block c at (k * bs + j) * nblk + c
};
static_assert(sizeof(npb_block_jacobi) == 40, "npb_block_jacobi layout");

// *info (device) = number of singular diagonal blocks (those are replaced by the identity)
int npb_block_jacobi_setup(int64_t nblk, int bs, int64_t cs, int64_t ks, const int32_t* indptr,
                           const int32_t* indices, const double* data, double* dinv,
                           int32_t* info, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (bs < 1 || bs > BJ_MAX || cs < 1 || ks < 1 || !info) { npb_set_error("block_jacobi: bad block shape"); return NPB_ERR_ARG; }
    NPB_CUDA(cudaMemsetAsync(info, 0, sizeof(int32_t), st));
    if (nblk <= 0) return NPB_OK;
    k_bj_setup<<<(int)((nblk + TPB - 1) / TPB), TPB, 0, st>>>(nblk, bs, cs, ks, indptr, indices, data, dinv, info);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("block_jacobi_setup");
    return NPB_OK;
}

// npb_apply_fn: ctx = npb_block_jacobi
int npb_block_jacobi_apply_cb(void* ctx, const double* x, double* y, void* stream) {
    const npb_block_jacobi* b = (const npb_block_jacobi*)ctx;
    if (b->nblk <= 0) return NPB_OK;
    k_bj_apply<<<(int)((b->nblk + TPB - 1) / TPB), TPB, 0, (cudaStream_t)stream>>>(b->nblk, b->bs, b->cs, b->ks, b->dinv, x, y);
    NPB_COUNT_LAUNCH();
    NPB_COUNT_BYTES(8.0 * b->nblk * b->bs * (b->bs + 2));
    NPB_CHECK_LAUNCH("block_jacobi_apply");
    return NPB_OK;
}

int64_t npb_bicgstab_workspace_bytes(int64_t n) {
    return 8 * ((n * 8 + 255) / 256 * 256) + 64 * (int64_t)sizeof(double) + npb_reduce_scratch_bytes() + 1024;
}

// Solve A x = b (x: initial guess on entry).  Stops when the TRUE residual ||b - A x|| <=
// max(rtol ||b||, atol) (the recurrence residual is re-checked against the operator before
// success is reported), after maxiter iterations (2 operator + 2 preconditioner applications
// each) -> NPB_ERR_NOCONV, or on a breakdown of the recurrences (rho or omega vanish) ->
// NPB_ERR_BREAKDOWN.
int npb_bicgstab(int64_t n, const npb_krylov_ops* ops, const double* b, double* x, int maxiter,
                 double rtol, double atol, void* work, int64_t work_bytes,
                 npb_krylov_info* info, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!ops || !ops->matvec || n <= 0) return NPB_ERR_ARG;
    if (work_bytes < npb_bicgstab_workspace_bytes(n)) {
        npb_set_error("bicgstab: workspace too small");
        return NPB_ERR_WORKSPACE;
    }
    const int64_t ld = (n * 8 + 255) / 256 * 256 / 8;
    double* base = (double*)work;
    double *r = base, *rh = base + ld, *p = base + 2 * ld, *v = base + 3 * ld, *s = base + 4 * ld,
           *t = base + 5 * ld, *ph = base + 6 * ld, *sh = base + 7 * ld;
    double* dsmall = base + 8 * ld;
    void* scratch = (void*)(dsmall + 64);
    NPB_CUDA(cudaMemsetAsync(scratch, 0, (size_t)npb_reduce_scratch_bytes(), st));
    int rc = 0, status = NPB_ERR_NOCONV, iters = 0, checks = 0;
    double bnorm = 0, rnorm = 0, rnorm0 = 0;
#define BI_TRY(expr) do { rc = (expr); if (rc) goto done; } while (0)
    auto dots = [&](const double* a0, const double* b0, const double* a1, const double* b1, double* out) -> int {
        // out[0] = a0 . b0, out[1] = a1 . b1 (a1 == NULL: one product); cross-rank sum; host copy
        int q = npb_dot_multi_launch(n, 1, a0, ld, b0, dsmall, scratch, st);
        if (!q && a1) q = npb_dot_multi_launch(n, 1, a1, ld, b1, dsmall + 1, scratch, st);
        if (!q && ops->allreduce) q = ops->allreduce(ops->allreduce_ctx, dsmall, a1 ? 2 : 1, stream);
        if (q) return q;
        cudaError_t e = cudaMemcpyAsync(out, dsmall, sizeof(double) * (a1 ? 2 : 1), cudaMemcpyDeviceToHost, st);
        if (e == cudaSuccess) e = cudaStreamSynchronize(st);
        if (e != cudaSuccess) { npb_set_error("bicgstab fetch: %s", cudaGetErrorString(e)); return NPB_ERR_CUDA; }
        return 0;
    };
    auto apply_m = [&](const double* in, double* out) -> int {
        if (ops->precond) return ops->precond(ops->precond_ctx, in, out, stream);
        return npb_axpby_launch(n, 1.0, in, 0.0, out, st);
    };
    double h[2];
    {
        BI_TRY(dots(b, b, nullptr, nullptr, h));
        bnorm = sqrt(h[0]);
        const double tol = fmax(rtol * bnorm, atol);
        BI_TRY(ops->matvec(ops->matvec_ctx, x, r, stream));
        BI_TRY(npb_axpby_launch(n, 1.0, b, -1.0, r, st));                     // r = b - A x
        BI_TRY(dots(r, r, nullptr, nullptr, h));
        rnorm0 = rnorm = sqrt(h[0]);
        BI_TRY(npb_axpby_launch(n, 1.0, r, 0.0, rh, st));                     // shadow residual
        double rho = 1.0, alpha = 1.0, omega = 1.0;
        NPB_CUDA(cudaMemsetAsync(p, 0, (size_t)n * 8, st));
        NPB_CUDA(cudaMemsetAsync(v, 0, (size_t)n * 8, st));
        while (true) {
            if (!(rnorm == rnorm)) { status = NPB_ERR_BREAKDOWN; break; }
            if (rnorm <= tol) {
                // the recurrence says converged: confirm with the operator (at most a few times)
                BI_TRY(ops->matvec(ops->matvec_ctx, x, t, stream));
                BI_TRY(npb_axpby_launch(n, 1.0, b, -1.0, t, st));
                BI_TRY(dots(t, t, nullptr, nullptr, h));
                const double tr = sqrt(h[0]);
                if (tr <= tol || ++checks > 3) { rnorm = tr; status = tr <= tol ? 0 : NPB_ERR_NOCONV; break; }
                BI_TRY(npb_axpby_launch(n, 1.0, t, 0.0, r, st));              // restart from the true residual
                BI_TRY(npb_axpby_launch(n, 1.0, r, 0.0, rh, st));
                rho = alpha = omega = 1.0;
                NPB_CUDA(cudaMemsetAsync(p, 0, (size_t)n * 8, st));
                NPB_CUDA(cudaMemsetAsync(v, 0, (size_t)n * 8, st));
                rnorm = tr;
            }
            if (iters >= maxiter) { status = NPB_ERR_NOCONV; break; }
            BI_TRY(dots(rh, r, nullptr, nullptr, h));
            const double rho_new = h[0];
            if (rho_new == 0.0 || omega == 0.0 || !(rho_new == rho_new)) { status = NPB_ERR_BREAKDOWN; break; }
            const double beta = (rho_new / rho) * (alpha / omega);
            rho = rho_new;
            // p = r + beta (p - omega v)
            k_axpbypcz<<<grid_of(n), TPB, 0, st>>>(n, 1.0, r, beta, p, -beta * omega, v);
            NPB_COUNT_LAUNCH();
            BI_TRY(apply_m(p, ph));
            BI_TRY(ops->matvec(ops->matvec_ctx, ph, v, stream));
            BI_TRY(dots(rh, v, nullptr, nullptr, h));
            if (h[0] == 0.0 || !(h[0] == h[0])) { status = NPB_ERR_BREAKDOWN; break; }
            alpha = rho / h[0];
            // s = r - alpha v
            BI_TRY(npb_axpby_launch(n, 1.0, r, 0.0, s, st));
            BI_TRY(npb_axpby_launch(n, -alpha, v, 1.0, s, st));
            BI_TRY(apply_m(s, sh));
            BI_TRY(ops->matvec(ops->matvec_ctx, sh, t, stream));
            BI_TRY(dots(t, s, t, t, h));
            omega = h[1] != 0.0 ? h[0] / h[1] : 0.0;
            // x += alpha ph + omega sh ;  r = s - omega t
            k_axpbypcz<<<grid_of(n), TPB, 0, st>>>(n, alpha, ph, 1.0, x, omega, sh);
            NPB_COUNT_LAUNCH();
            BI_TRY(npb_axpby_launch(n, 1.0, s, 0.0, r, st));
            BI_TRY(npb_axpby_launch(n, -omega, t, 1.0, r, st));
            BI_TRY(dots(r, r, nullptr, nullptr, h));
            rnorm = sqrt(h[0]);
            ++iters;
        }
    }
done:
    if (info) {
        info->iters = iters; info->restarts = checks;
        info->status = rc ? rc : status;
        info->pad = 0;
        info->bnorm = bnorm; info->rnorm0 = rnorm0; info->rnorm = rnorm;
    }
    return rc ? rc : status;
#undef BI_TRY
}

}  // extern "C"
