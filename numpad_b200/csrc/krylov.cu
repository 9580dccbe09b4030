// SpMV and the vector kernels of the Newton linear solve that replaces
// scipy.sparse.linalg.spsolve (reference adsolve.py:193-195) and the
// factorized() solves of adsolve.py:78,141: restarted (flexible) GMRES with
// right preconditioning, CGS2 orthogonalisation done with two fused
// multi-vector kernels, all reductions deterministic (fixed-order last-block
// sums, no floating-point atomics).
//
// Every kernel is HBM-bound; algorithmic bytes per call are listed in
// DESIGN.md:  SpMV 12*nnz + 4*(n+1) + 16*n;  multi-dot 8*n*(k+1);
// multi-axpy 8*n*(k+2).
#include <math.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace {

constexpr int TPB = 256;
constexpr int RED_BLOCKS = NPB_NUM_SMS * 4;   // persistent grid of the reductions
constexpr int KC = 16;                        // basis vectors per multi-dot pass

// ---------------------------------------------------------------------- SpMV
// Epilogues shared by both SpMV kernels (s = (A x)[r]):
//   0: y = s     1: y = b - s     2: y = b + s
//   3: y = x + omega * dinv .* (b - s)      (damped-Jacobi sweep; y != x)
struct spmv_epi {
    int mode;
    const double* b;
    const double* dinv;
    double omega;
};

__device__ __forceinline__ double spmv_finish(const spmv_epi& e, int64_t r, double s,
                                              const double* x) {
    switch (e.mode) {
        case 0: return s;
        case 1: return e.b[r] - s;
        case 2: return e.b[r] + s;
        default: return x[r] + e.omega * e.dinv[r] * (e.b[r] - s);
    }
}

// (a) "stream" kernel for short rows (the Jacobians here have 4..36 nnz/row):
// a CTA owns 256 consecutive rows; it streams their CONTIGUOUS range of
// (index, value) pairs with fully coalesced, mutually independent loads -- no
// load depends on a row pointer, so memory-level parallelism is limited only
// by occupancy -- stages the products in shared memory, then thread t sums row
// t from shared memory in CSR order (deterministic, same order as a sequential
// CPU SpMV).  Rows longer than the tile are handled by looping over tiles.
constexpr int SP_ROWS = 256;
constexpr int SP_TILE = 3072;       // 24 KB of products -> 8 CTAs / SM

__global__ void __launch_bounds__(SP_ROWS)
k_spmv_stream(int64_t nrows, const int32_t* __restrict__ ptr,
              const int32_t* __restrict__ idx, const double* __restrict__ dat,
              const double* x, spmv_epi epi, double* y) {
    __shared__ int32_t sptr[SP_ROWS + 1];
    __shared__ double sprod[SP_TILE];
    const int tid = threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.x * SP_ROWS;
    const int nr = (int)min((int64_t)SP_ROWS, nrows - r0);
    if (tid <= nr) sptr[tid] = ptr[r0 + tid];
    if (tid == 0 && nr == SP_ROWS) sptr[SP_ROWS] = ptr[r0 + SP_ROWS];
    __syncthreads();
    const int32_t k0 = sptr[0], k1 = sptr[nr];
    const int32_t rs = (tid < nr) ? sptr[tid] : k1, re = (tid < nr) ? sptr[tid + 1] : k1;
    double acc = 0.0;
    for (int32_t base = k0; base < k1; base += SP_TILE) {
        const int32_t lim = min(base + SP_TILE, k1);
#pragma unroll 4
        for (int32_t k = base + tid; k < lim; k += SP_ROWS)
            sprod[k - base] = dat[k] * __ldg(x + idx[k]);
        __syncthreads();
        const int32_t s = max(rs, base), e = min(re, lim);
        for (int32_t k = s; k < e; ++k) acc += sprod[k - base];
        __syncthreads();
    }
    if (tid < nr) y[r0 + tid] = spmv_finish(epi, r0 + tid, acc, x);
}

// (b) "vector" kernel for long rows: G lanes cooperate on one row.
template <int G>
__global__ void __launch_bounds__(TPB)
k_spmv(int64_t nrows, const int32_t* __restrict__ ptr,
       const int32_t* __restrict__ idx, const double* __restrict__ dat,
       const double* x, spmv_epi epi, double* y) {
    const int lane = threadIdx.x & (G - 1);
    const int64_t r = ((int64_t)blockIdx.x * TPB + threadIdx.x) / G;
    double s = 0.0;
    if (r < nrows) {
        const int32_t e = ptr[r + 1];
        for (int32_t k = ptr[r] + lane; k < e; k += G)
            s = fma(dat[k], __ldg(x + idx[k]), s);
    }
#pragma unroll
    for (int d = G >> 1; d; d >>= 1) s += __shfl_xor_sync(0xffffffffu, s, d);
    if (r < nrows && lane == 0) y[r] = spmv_finish(epi, r, s, x);
}

// ------------------------------------------------ deterministic block reduce
// sums `nv` per-thread values over the block into smem[0..nv)
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* smem /*[8*NV]*/) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        double t = v[j];
#pragma unroll
        for (int d = 16; d; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
        if (lane == 0) smem[warp * NV + j] = t;
    }
    __syncthreads();
    if (threadIdx.x < NV) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < TPB / 32; ++w) t += smem[w * NV + threadIdx.x];
        smem[threadIdx.x] = t;   // safe: thread j only overwrites slot (0,j) it just read
    }
    __syncthreads();
}

// last block to finish adds the per-block partials in block order
__device__ __forceinline__ bool last_block(unsigned int* ticket) {
    __shared__ bool is_last;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(ticket, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    return is_last;
}

// out[j] = V_j . w for j < kc (kc <= KC), V_j = V + j*ld
__global__ void __launch_bounds__(TPB)
k_dot_multi(int64_t n, int kc, const double* __restrict__ V, int64_t ld,
            const double* __restrict__ w, double* __restrict__ partial,
            unsigned int* __restrict__ ticket, double* __restrict__ out) {
    __shared__ double smem[(TPB / 32) * KC];
    double acc[KC];
#pragma unroll
    for (int j = 0; j < KC; ++j) acc[j] = 0.0;
    for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * TPB) {
        const double wi = w[i];
#pragma unroll
        for (int j = 0; j < KC; ++j)
            if (j < kc) acc[j] = fma(V[j * ld + i], wi, acc[j]);
    }
    block_sum<KC>(acc, smem);
    if (threadIdx.x < kc) partial[(int64_t)blockIdx.x * KC + threadIdx.x] = smem[threadIdx.x];
    if (last_block(ticket)) {
        __threadfence();
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int j = warp; j < kc; j += TPB / 32) {
            double t = 0.0;
            for (int b = lane; b < (int)gridDim.x; b += 32) t += partial[(int64_t)b * KC + j];
#pragma unroll
            for (int d = 16; d; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
            if (lane == 0) out[j] = t;
        }
        if (threadIdx.x == 0) *ticket = 0;
    }
}

// w -= sum_j h[j] V_j (j < k), and out_norm2 = ||w_new||^2
__global__ void __launch_bounds__(TPB)
k_axpy_multi(int64_t n, int k, const double* __restrict__ V, int64_t ld,
             const double* __restrict__ h, double sign, double* __restrict__ w,
             double* __restrict__ partial, unsigned int* __restrict__ ticket,
             double* __restrict__ out_norm2) {
    extern __shared__ double hs[];          // k coefficients + reduce scratch
    double* smem = hs + k;
    for (int j = threadIdx.x; j < k; j += TPB) hs[j] = sign * h[j];
    __syncthreads();
    double nrm[1] = {0.0};
    for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * TPB) {
        double s0 = 0.0, s1 = 0.0;
        int j = 0;
        for (; j + 1 < k; j += 2) {
            s0 = fma(hs[j], V[j * ld + i], s0);
            s1 = fma(hs[j + 1], V[(j + 1) * ld + i], s1);
        }
        if (j < k) s0 = fma(hs[j], V[j * ld + i], s0);
        const double wi = w[i] + (s0 + s1);
        w[i] = wi;
        nrm[0] = fma(wi, wi, nrm[0]);
    }
    block_sum<1>(nrm, smem);
    if (threadIdx.x == 0) partial[blockIdx.x] = smem[0];
    if (last_block(ticket)) {
        __threadfence();
        if (threadIdx.x < 32) {
            double t = 0.0;
            for (int b = threadIdx.x; b < (int)gridDim.x; b += 32) t += partial[b];
#pragma unroll
            for (int d = 16; d; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
            if (threadIdx.x == 0) { *out_norm2 = t; *ticket = 0; }
        }
    }
}

// y = a*x + b*y   (b == 0 never reads y)
__global__ void __launch_bounds__(TPB)
k_axpby(int64_t n, double a, const double* __restrict__ x, double b, double* __restrict__ y) {
    for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * TPB)
        y[i] = (b == 0.0) ? a * x[i] : fma(a, x[i], b * y[i]);
}

int grid_for(int64_t n) {
    int64_t b = (n + TPB - 1) / TPB;
    if (b > RED_BLOCKS) b = RED_BLOCKS;
    return (int)(b < 1 ? 1 : b);
}

}  // namespace

#define ST(s) ((cudaStream_t)(s))

// ------------------------------------------------------------- internal API
// y = epilogue(A x); see spmv_epi.  `variant`: 0 auto, 1 stream, 2 vector
int npb_spmv_ex(int64_t nrows, int64_t nnz, const int32_t* ptr, const int32_t* idx,
                const double* dat, const double* x, int mode, const double* b,
                const double* dinv, double omega, double* y, int variant, cudaStream_t st) {
    if (nrows <= 0) return NPB_OK;
    spmv_epi e{mode, b, dinv, omega};
    const double mean = (double)nnz / (double)nrows;
    if (variant == 1 || (variant == 0 && mean <= 64.0)) {
        k_spmv_stream<<<npb_blocks(nrows, SP_ROWS), SP_ROWS, 0, st>>>(nrows, ptr, idx, dat, x, e, y);
    } else if (mean <= 48.0) {
        k_spmv<16><<<npb_blocks(nrows * 16, TPB), TPB, 0, st>>>(nrows, ptr, idx, dat, x, e, y);
    } else {
        k_spmv<32><<<npb_blocks(nrows * 32, TPB), TPB, 0, st>>>(nrows, ptr, idx, dat, x, e, y);
    }
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("spmv");
    return NPB_OK;
}

int npb_spmv_launch(int64_t nrows, int64_t nnz, const int32_t* ptr, const int32_t* idx,
                    const double* dat, const double* x, const double* b, double* y,
                    int mode, cudaStream_t st) {
    return npb_spmv_ex(nrows, nnz, ptr, idx, dat, x, mode, b, nullptr, 0.0, y, 0, st);
}

// scratch: RED_BLOCKS*KC doubles of partials followed by one uint ticket
static inline double* red_partial(void* scratch) { return (double*)scratch; }
static inline unsigned int* red_ticket(void* scratch) {
    return (unsigned int*)((double*)scratch + (size_t)RED_BLOCKS * KC);
}

int npb_spmv_ex(int64_t nrows, int64_t nnz, const int32_t* ptr, const int32_t* idx,
                const double* dat, const double* x, int mode, const double* b,
                const double* dinv, double omega, double* y, int variant, cudaStream_t st);

int npb_dot_multi_launch(int64_t n, int k, const double* V, int64_t ld, const double* w,
                         double* out, void* scratch, cudaStream_t st) {
    for (int j0 = 0; j0 < k; j0 += KC) {
        int kc = k - j0 < KC ? k - j0 : KC;
        k_dot_multi<<<grid_for(n), TPB, 0, st>>>(n, kc, V + (int64_t)j0 * ld, ld, w,
                                                 red_partial(scratch), red_ticket(scratch),
                                                 out + j0);
        NPB_COUNT_LAUNCH();
    }
    NPB_CHECK_LAUNCH("dot_multi");
    return NPB_OK;
}

int npb_axpy_multi_launch(int64_t n, int k, const double* V, int64_t ld, const double* h,
                          double sign, double* w, double* out_norm2, void* scratch,
                          cudaStream_t st) {
    size_t sh = sizeof(double) * ((size_t)k + TPB / 32);
    k_axpy_multi<<<grid_for(n), TPB, sh, st>>>(n, k, V, ld, h, sign, w, red_partial(scratch),
                                               red_ticket(scratch), out_norm2);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("axpy_multi");
    return NPB_OK;
}

int npb_axpby_launch(int64_t n, double a, const double* x, double b, double* y, cudaStream_t st) {
    if (n <= 0) return NPB_OK;
    k_axpby<<<grid_for(n), TPB, 0, st>>>(n, a, x, b, y);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("axpby");
    return NPB_OK;
}

extern "C" {

int64_t npb_reduce_scratch_bytes(void) {
    return (int64_t)(sizeof(double) * (size_t)RED_BLOCKS * KC + 64);
}

int npb_spmv(int64_t nrows, int64_t nnz, const int32_t* indptr, const int32_t* indices,
             const double* data, const double* x, double* y, void* stream) {
    return npb_spmv_launch(nrows, nnz, indptr, indices, data, x, nullptr, y, 0, ST(stream));
}

// same, forcing a kernel variant (1 = stream, 2 = vector); for tests / tuning
int npb_spmv_variant(int64_t nrows, int64_t nnz, const int32_t* indptr, const int32_t* indices,
                     const double* data, const double* x, double* y, int variant, void* stream) {
    return npb_spmv_ex(nrows, nnz, indptr, indices, data, x, 0, nullptr, nullptr, 0.0, y, variant,
                       ST(stream));
}

// y = b - A x
int npb_spmv_residual(int64_t nrows, int64_t nnz, const int32_t* indptr,
                      const int32_t* indices, const double* data, const double* x,
                      const double* b, double* y, void* stream) {
    return npb_spmv_launch(nrows, nnz, indptr, indices, data, x, b, y, 1, ST(stream));
}

// out[j] = V[j*ld : j*ld+n] . w ; scratch = npb_reduce_scratch_bytes() bytes,
// zero-initialised once by the caller
int npb_dot_multi(int64_t n, int k, const double* V, int64_t ld, const double* w,
                  double* out, void* scratch, void* stream) {
    return npb_dot_multi_launch(n, k, V, ld, w, out, scratch, ST(stream));
}

// w += sign * sum_j h[j] V_j ; *out_norm2 = ||w||^2 afterwards
int npb_axpy_multi(int64_t n, int k, const double* V, int64_t ld, const double* h,
                   double sign, double* w, double* out_norm2, void* scratch, void* stream) {
    return npb_axpy_multi_launch(n, k, V, ld, h, sign, w, out_norm2, scratch, ST(stream));
}

int npb_axpby(int64_t n, double a, const double* x, double b, double* y, void* stream) {
    return npb_axpby_launch(n, a, x, b, y, ST(stream));
}

}  // extern "C"
