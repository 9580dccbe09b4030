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
#include "dist.cuh"

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
    // row-sharded operators (dist.cuh): local columns >= n_own are boundary values of
    // other ranks, gathered into xh; a single-GPU operator has n_own = INT32_MAX
    const double* xh;
    int32_t n_own;
    npb_wait wait;      // HALO kernels: arrival counters to wait for before the first gather
};

template <bool HALO>
__device__ __forceinline__ double ldx(const double* x, const spmv_epi& e, int32_t c) {
    if (!HALO) return __ldg(x + c);
    // halo slots are written by PEER GPUs while this kernel may already run (the arrival
    // wait sits in the prologue): they must not go through the non-coherent path
    return c < e.n_own ? __ldg(x + c) : __ldcg(e.xh + (c - e.n_own));
}

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
constexpr int SP_THREADS = 256;
constexpr int SP_TILE = 3072;       // 24 KB of products -> 8 CTAs / SM

// ROWS rows per CTA, TPR = 256 / ROWS threads cooperate on the sum of one row.
// ROWS = 256 for the big operators (least overhead per row); 64 / 32 for the
// small multigrid levels so that their grids still cover the 148 SMs.
template <int ROWS, bool HALO>
__global__ void __launch_bounds__(SP_THREADS)
k_spmv_stream(int64_t nrows, const int32_t* __restrict__ ptr,
              const int32_t* __restrict__ idx, const double* __restrict__ dat,
              const double* x, spmv_epi epi, double* y) {
    constexpr int TPR = SP_THREADS / ROWS;
    __shared__ int32_t sptr[ROWS + 1];
    __shared__ double sprod[SP_TILE];
    if (HALO) npb_halo_begin(epi.wait, epi.xh);
    const int tid = threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.x * ROWS;
    const int nr = (int)min((int64_t)ROWS, nrows - r0);
    if (tid <= nr) sptr[tid] = ptr[r0 + tid];
    if (ROWS == SP_THREADS && tid == 0 && nr == ROWS) sptr[ROWS] = ptr[r0 + ROWS];
    __syncthreads();
    const int32_t k0 = sptr[0], k1 = sptr[nr];
    const int row = tid / TPR, sub = tid % TPR;
    const int32_t rs = (row < nr) ? sptr[row] : k1, re = (row < nr) ? sptr[row + 1] : k1;
    double acc = 0.0;
    for (int32_t base = k0; base < k1; base += SP_TILE) {
        const int32_t lim = min(base + SP_TILE, k1);
#pragma unroll 4
        for (int32_t k = base + tid; k < lim; k += SP_THREADS)
            sprod[k - base] = dat[k] * ldx<HALO>(x, epi, idx[k]);
        __syncthreads();
        const int32_t s = max(rs, base), e = min(re, lim);
        for (int32_t k = s + sub; k < e; k += TPR) acc += sprod[k - base];
        __syncthreads();
    }
    if (TPR > 1) {
#pragma unroll
        for (int d = TPR >> 1; d; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    }
    if (row < nr && sub == 0) y[r0 + row] = spmv_finish(epi, r0 + row, acc, x);
}

// (a') same decomposition, but the (index, value) range is brought into shared
// memory with 16-byte cp.async (LDGSTS): no registers are tied up by loads in
// flight, so every CTA keeps its whole tile (36 KB) outstanding at once.  Needs
// 16-byte aligned idx / dat base pointers (checked by the launcher).
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gmem));
}

template <int ROWS, bool HALO>
__global__ void __launch_bounds__(SP_THREADS)
k_spmv_stream_async(int64_t nrows, int64_t nnz, const int32_t* __restrict__ ptr,
                    const int32_t* __restrict__ idx, const double* __restrict__ dat,
                    const double* x, spmv_epi epi, double* y) {
    constexpr int TPR = SP_THREADS / ROWS;
    __shared__ int32_t sptr[ROWS + 1];
    __shared__ __align__(16) int32_t sidx[SP_TILE];
    __shared__ __align__(16) double sdat[SP_TILE];
    if (HALO) npb_halo_begin(epi.wait, epi.xh);
    const int tid = threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.x * ROWS;
    const int nr = (int)min((int64_t)ROWS, nrows - r0);
    if (tid <= nr) sptr[tid] = ptr[r0 + tid];
    if (ROWS == SP_THREADS && tid == 0 && nr == ROWS) sptr[ROWS] = ptr[r0 + ROWS];
    __syncthreads();
    const int32_t k0 = sptr[0], k1 = sptr[nr];
    const int row = tid / TPR, sub = tid % TPR;
    const int32_t rs = (row < nr) ? sptr[row] : k1, re = (row < nr) ? sptr[row + 1] : k1;
    double acc = 0.0;
    for (int32_t base = k0 & ~3; base < k1; base += SP_TILE) {
        const int32_t lim = min(base + SP_TILE, k1);            // exclusive, unaligned
        const int32_t full = (int32_t)min((int64_t)((lim - base + 3) & ~3), (nnz - base) & ~(int64_t)3);
        // whole 16-byte chunks inside the arrays
        for (int32_t c = tid * 4; c < full; c += SP_THREADS * 4) cp_async16(sidx + c, idx + base + c);
        for (int32_t c = tid * 2; c < full; c += SP_THREADS * 2) cp_async16(sdat + c, dat + base + c);
        // ragged tail at the very end of the arrays
        for (int32_t k = base + full + tid; k < lim; k += SP_THREADS) {
            sidx[k - base] = idx[k];
            sdat[k - base] = dat[k];
        }
        asm volatile("cp.async.commit_group;\n" ::);
        asm volatile("cp.async.wait_group 0;\n" ::);
        __syncthreads();
        const int32_t lo = max(base, k0);
#pragma unroll 4
        for (int32_t k = lo + tid; k < lim; k += SP_THREADS)
            sdat[k - base] *= ldx<HALO>(x, epi, sidx[k - base]);
        __syncthreads();
        const int32_t s = max(rs, base), e = min(re, lim);
        for (int32_t k = s + sub; k < e; k += TPR) acc += sdat[k - base];
        __syncthreads();
    }
    if (TPR > 1) {
#pragma unroll
        for (int d = TPR >> 1; d; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    }
    if (row < nr && sub == 0) y[r0 + row] = spmv_finish(epi, r0 + row, acc, x);
}

// (a'') same decomposition with 128-bit loads: the tile starts at the 4-entry
// boundary below the CTA's first entry, every thread fetches 4 indices (int4)
// and 4 values (2 x double2) per round and three rounds are in flight at once
// (144 B per thread outstanding instead of 48).  Needs 16-byte aligned idx /
// dat base pointers (checked by the launcher).
template <int ROWS, bool HALO>
__global__ void __launch_bounds__(SP_THREADS)
k_spmv_stream_v4(int64_t nrows, int64_t nnz, const int32_t* __restrict__ ptr,
                 const int32_t* __restrict__ idx, const double* __restrict__ dat,
                 const double* x, spmv_epi epi, double* y) {
    constexpr int TPR = SP_THREADS / ROWS;
    constexpr int ROUNDS = SP_TILE / (SP_THREADS * 4);
    __shared__ int32_t sptr[ROWS + 1];
    __shared__ __align__(16) double sprod[SP_TILE];
    if (HALO) npb_halo_begin(epi.wait, epi.xh);
    const int tid = threadIdx.x;
    const int64_t r0 = (int64_t)blockIdx.x * ROWS;
    const int nr = (int)min((int64_t)ROWS, nrows - r0);
    if (tid <= nr) sptr[tid] = ptr[r0 + tid];
    if (ROWS == SP_THREADS && tid == 0 && nr == ROWS) sptr[ROWS] = ptr[r0 + ROWS];
    __syncthreads();
    const int32_t k0 = sptr[0], k1 = sptr[nr];
    const int row = tid / TPR, sub = tid % TPR;
    const int32_t rs = (row < nr) ? sptr[row] : k1, re = (row < nr) ? sptr[row + 1] : k1;
    double acc = 0.0;
    for (int32_t base = k0 & ~3; base < k1; base += SP_TILE) {
        const int32_t lim = min(base + SP_TILE, k1);
        int4 iv[ROUNDS];
        double2 da[ROUNDS], db[ROUNDS];
#pragma unroll
        for (int j = 0; j < ROUNDS; ++j) {
            const int32_t k = base + (j * SP_THREADS + tid) * 4;
            iv[j] = make_int4(0, 0, 0, 0);
            da[j] = make_double2(0.0, 0.0);
            db[j] = make_double2(0.0, 0.0);
            if (k < lim) {
                if ((int64_t)k + 3 < nnz) {
                    iv[j] = *reinterpret_cast<const int4*>(idx + k);
                    da[j] = *reinterpret_cast<const double2*>(dat + k);
                    db[j] = *reinterpret_cast<const double2*>(dat + k + 2);
                } else {                      // last 1..3 entries of the arrays
                    iv[j].x = idx[k]; da[j].x = dat[k];
                    if ((int64_t)k + 1 < nnz) { iv[j].y = idx[k + 1]; da[j].y = dat[k + 1]; }
                    if ((int64_t)k + 2 < nnz) { iv[j].z = idx[k + 2]; db[j].x = dat[k + 2]; }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < ROUNDS; ++j) {
            const int32_t k = base + (j * SP_THREADS + tid) * 4;
            if (k < lim) {
                double2 p0, p1;
                p0.x = da[j].x * ldx<HALO>(x, epi, iv[j].x);
                p0.y = da[j].y * ldx<HALO>(x, epi, iv[j].y);
                p1.x = db[j].x * ldx<HALO>(x, epi, iv[j].z);
                p1.y = db[j].y * ldx<HALO>(x, epi, iv[j].w);
                *reinterpret_cast<double2*>(sprod + (k - base)) = p0;
                *reinterpret_cast<double2*>(sprod + (k - base) + 2) = p1;
            }
        }
        __syncthreads();
        const int32_t s = max(rs, base), e = min(re, lim);
        for (int32_t k = s + sub; k < e; k += TPR) acc += sprod[k - base];
        __syncthreads();
    }
    if (TPR > 1) {
#pragma unroll
        for (int d = TPR >> 1; d; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
    }
    if (row < nr && sub == 0) y[r0 + row] = spmv_finish(epi, r0 + row, acc, x);
}

#include "spmv_tma.cuh"

// (b) "vector" kernel for long rows: G lanes cooperate on one row.
template <int G, bool HALO>
__global__ void __launch_bounds__(TPB)
k_spmv(int64_t nrows, const int32_t* __restrict__ ptr,
       const int32_t* __restrict__ idx, const double* __restrict__ dat,
       const double* x, spmv_epi epi, double* y) {
    if (HALO) npb_halo_begin(epi.wait, epi.xh);
    const int lane = threadIdx.x & (G - 1);
    const int64_t r = ((int64_t)blockIdx.x * TPB + threadIdx.x) / G;
    double s = 0.0;
    if (r < nrows) {
        const int32_t e = ptr[r + 1];
        for (int32_t k = ptr[r] + lane; k < e; k += G)
            s = fma(dat[k], ldx<HALO>(x, epi, idx[k]), s);
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

// same for few vectors (kc <= KS: the CG inner products and norms, the first Gram-Schmidt
// steps): the general kernel keeps only kc + 1 eight-byte loads in flight per thread, which
// leaves a one- or two-vector product latency-bound; here every thread moves two elements
// with 128-bit loads and two such pairs per round.  V / w / ld 16-byte aligned; any n.
template <int KS>
__global__ void __launch_bounds__(TPB)
k_dot_few(int64_t n, int kc, const double* __restrict__ V, int64_t ld,
          const double* __restrict__ w, double* __restrict__ partial,
          unsigned int* __restrict__ ticket, double* __restrict__ out) {
    __shared__ double smem[(TPB / 32) * KS];
    double acc[KS];
#pragma unroll
    for (int j = 0; j < KS; ++j) acc[j] = 0.0;
    const int64_t n2 = n >> 1;              // element pairs; an odd last element goes to one thread
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
#pragma unroll
        for (int j = 0; j < KS; ++j)
            if (j < kc) acc[j] = V[(int64_t)j * ld + n - 1] * w[n - 1];
    }
    const int64_t stride = (int64_t)gridDim.x * TPB;
    const double2* w2 = reinterpret_cast<const double2*>(w);
    for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < n2; i += 2 * stride) {
        const int64_t i1 = i + stride;
        const bool two = i1 < n2;
        const double2 wa = w2[i];
        const double2 wb = two ? w2[i1] : make_double2(0.0, 0.0);
        double2 va[KS], vb[KS];
#pragma unroll
        for (int j = 0; j < KS; ++j) {
            va[j] = vb[j] = make_double2(0.0, 0.0);
            if (j < kc) {
                const double2* v2 = reinterpret_cast<const double2*>(V + (int64_t)j * ld);
                va[j] = v2[i];
                if (two) vb[j] = v2[i1];
            }
        }
#pragma unroll
        for (int j = 0; j < KS; ++j) {
            acc[j] = fma(va[j].x, wa.x, acc[j]);
            acc[j] = fma(va[j].y, wa.y, acc[j]);
            acc[j] = fma(vb[j].x, wb.x, acc[j]);
            acc[j] = fma(vb[j].y, wb.y, acc[j]);
        }
    }
    block_sum<KS>(acc, smem);
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

// out_r[j] = V_j . w_r for j < kc (kc <= KS) and r < NR right-hand vectors, all read ONCE:
// the Gram-Schmidt products of the Krylov loop (NR = 2: the delayed second pass over the
// previous basis vector rides on the first pass over the new one, gmres.cu).  Every thread
// moves one element pair per vector and round with 128-bit loads (KS + NR of them in flight;
// the scalar kernel above, 8 bytes per load, stays at ~4.3 of ~6.7 TB/s).  V / w_r / ld
// 16-byte aligned; any n.
template <int KS, int NR>
__global__ void __launch_bounds__(TPB, 2)
k_dot_block(int64_t n, int kc, const double* __restrict__ V, int64_t ld,
            const double* __restrict__ w0, const double* __restrict__ w1,
            double* __restrict__ partial, unsigned int* __restrict__ ticket,
            double* __restrict__ out0, double* __restrict__ out1) {
    constexpr int NV = KS * NR;
    __shared__ double smem[(TPB / 32) * NV];
    double acc[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) acc[j] = 0.0;
    const int64_t n2 = n >> 1;
    if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
        const double a0 = w0[n - 1], a1 = (NR > 1) ? w1[n - 1] : 0.0;
#pragma unroll
        for (int j = 0; j < KS; ++j)
            if (j < kc) {
                const double v = V[(int64_t)j * ld + n - 1];
                acc[j] = v * a0;
                if (NR > 1) acc[KS + j] = v * a1;
            }
    }
    const double2* p0 = reinterpret_cast<const double2*>(w0);
    const double2* p1 = reinterpret_cast<const double2*>(NR > 1 ? w1 : w0);
    for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < n2;
         i += (int64_t)gridDim.x * TPB) {
        const double2 a0 = p0[i];
        double2 a1 = make_double2(0.0, 0.0);
        if (NR > 1) a1 = p1[i];
        double2 v[KS];
#pragma unroll
        for (int j = 0; j < KS; ++j) {
            v[j] = make_double2(0.0, 0.0);
            if (j < kc) v[j] = reinterpret_cast<const double2*>(V + (int64_t)j * ld)[i];
        }
#pragma unroll
        for (int j = 0; j < KS; ++j) {
            acc[j] = fma(v[j].x, a0.x, fma(v[j].y, a0.y, acc[j]));
            if (NR > 1) acc[KS + j] = fma(v[j].x, a1.x, fma(v[j].y, a1.y, acc[KS + j]));
        }
    }
    block_sum<NV>(acc, smem);
    if (threadIdx.x < NV) partial[(int64_t)blockIdx.x * NV + threadIdx.x] = smem[threadIdx.x];
    if (last_block(ticket)) {
        __threadfence();
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        for (int q = warp; q < NR * kc; q += TPB / 32) {
            const int r = q / kc, j = q - r * kc;
            double t = 0.0;
            for (int b = lane; b < (int)gridDim.x; b += 32) t += partial[(int64_t)b * NV + r * KS + j];
#pragma unroll
            for (int d = 16; d; d >>= 1) t += __shfl_xor_sync(0xffffffffu, t, d);
            if (lane == 0) (r == 0 ? out0 : out1)[j] = t;
        }
        if (threadIdx.x == 0) *ticket = 0;
    }
}

// w += sum_j hs[j] V_j (j < k, hs = sign * h), and out_norm2 = ||w_new||^2.
// Two elements per thread (128-bit loads) and four basis vectors per round: 8 x 16 B
// in flight per thread -- the kernel streams (k + 2) vectors and has nothing to hide
// the latency with but its own loads.  VEC = false: any alignment / odd n.
template <bool VEC>
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
    if (VEC) {
        const int64_t n2 = n >> 1;
        for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < n2;
             i += (int64_t)gridDim.x * TPB) {
            double2 s0 = make_double2(0.0, 0.0), s1 = s0;
            int j = 0;
            for (; j + 3 < k; j += 4) {
                const double2 a0 = reinterpret_cast<const double2*>(V + (int64_t)j * ld)[i];
                const double2 a1 = reinterpret_cast<const double2*>(V + (int64_t)(j + 1) * ld)[i];
                const double2 a2 = reinterpret_cast<const double2*>(V + (int64_t)(j + 2) * ld)[i];
                const double2 a3 = reinterpret_cast<const double2*>(V + (int64_t)(j + 3) * ld)[i];
                s0.x = fma(hs[j], a0.x, s0.x);         s0.y = fma(hs[j], a0.y, s0.y);
                s1.x = fma(hs[j + 1], a1.x, s1.x);     s1.y = fma(hs[j + 1], a1.y, s1.y);
                s0.x = fma(hs[j + 2], a2.x, s0.x);     s0.y = fma(hs[j + 2], a2.y, s0.y);
                s1.x = fma(hs[j + 3], a3.x, s1.x);     s1.y = fma(hs[j + 3], a3.y, s1.y);
            }
            for (; j < k; ++j) {
                const double2 a0 = reinterpret_cast<const double2*>(V + (int64_t)j * ld)[i];
                s0.x = fma(hs[j], a0.x, s0.x);         s0.y = fma(hs[j], a0.y, s0.y);
            }
            double2 wi = reinterpret_cast<double2*>(w)[i];
            wi.x += s0.x + s1.x;
            wi.y += s0.y + s1.y;
            reinterpret_cast<double2*>(w)[i] = wi;
            nrm[0] = fma(wi.x, wi.x, fma(wi.y, wi.y, nrm[0]));
        }
    }
    // scalar path: everything (VEC = false) or the odd last element
    for (int64_t i = (VEC ? (n & ~(int64_t)1) : 0) + (int64_t)blockIdx.x * TPB + threadIdx.x; i < n;
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

// The update pass of the Gram-Schmidt step with delayed re-orthogonalisation (gmres.cu): ONE
// sweep over the k final basis vectors does both
//     u <- (u - sum_j c_j V_j) * inv_rho            (second pass over the previous vector)
//     w <- w - sum_j a_j V_j - b * u_old            (first pass over the new one)
// and returns ||w_new||^2.  coef = [c_0..c_{k-1}, a_0..a_{k-1}, inv_rho, b] (device).
template <bool VEC>
__global__ void __launch_bounds__(TPB)
k_axpy_multi2(int64_t n, int k, const double* __restrict__ V, int64_t ld,
              const double* __restrict__ coef, double* __restrict__ u, double* __restrict__ w,
              double* __restrict__ partial, unsigned int* __restrict__ ticket,
              double* __restrict__ out_norm2) {
    extern __shared__ double hs[];          // 2k + 2 coefficients + reduce scratch
    double* smem = hs + 2 * k + 2;
    for (int j = threadIdx.x; j < 2 * k + 2; j += TPB) hs[j] = coef[j];
    __syncthreads();
    const double* hc = hs;
    const double* ha = hs + k;
    const double inv_rho = hs[2 * k], bu = hs[2 * k + 1];
    double nrm[1] = {0.0};
    if (VEC) {
        const int64_t n2 = n >> 1;
        for (int64_t i = (int64_t)blockIdx.x * TPB + threadIdx.x; i < n2;
             i += (int64_t)gridDim.x * TPB) {
            double2 su = make_double2(0.0, 0.0), sw = su;
            int j = 0;
            for (; j + 3 < k; j += 4) {
                const double2 a0 = reinterpret_cast<const double2*>(V + (int64_t)j * ld)[i];
                const double2 a1 = reinterpret_cast<const double2*>(V + (int64_t)(j + 1) * ld)[i];
                const double2 a2 = reinterpret_cast<const double2*>(V + (int64_t)(j + 2) * ld)[i];
                const double2 a3 = reinterpret_cast<const double2*>(V + (int64_t)(j + 3) * ld)[i];
                su.x = fma(hc[j], a0.x, su.x);         su.y = fma(hc[j], a0.y, su.y);
                sw.x = fma(ha[j], a0.x, sw.x);         sw.y = fma(ha[j], a0.y, sw.y);
                su.x = fma(hc[j + 1], a1.x, su.x);     su.y = fma(hc[j + 1], a1.y, su.y);
                sw.x = fma(ha[j + 1], a1.x, sw.x);     sw.y = fma(ha[j + 1], a1.y, sw.y);
                su.x = fma(hc[j + 2], a2.x, su.x);     su.y = fma(hc[j + 2], a2.y, su.y);
                sw.x = fma(ha[j + 2], a2.x, sw.x);     sw.y = fma(ha[j + 2], a2.y, sw.y);
                su.x = fma(hc[j + 3], a3.x, su.x);     su.y = fma(hc[j + 3], a3.y, su.y);
                sw.x = fma(ha[j + 3], a3.x, sw.x);     sw.y = fma(ha[j + 3], a3.y, sw.y);
            }
            for (; j < k; ++j) {
                const double2 a0 = reinterpret_cast<const double2*>(V + (int64_t)j * ld)[i];
                su.x = fma(hc[j], a0.x, su.x);         su.y = fma(hc[j], a0.y, su.y);
                sw.x = fma(ha[j], a0.x, sw.x);         sw.y = fma(ha[j], a0.y, sw.y);
            }
            const double2 uo = reinterpret_cast<double2*>(u)[i];
            double2 wi = reinterpret_cast<double2*>(w)[i];
            wi.x = wi.x - sw.x - bu * uo.x;
            wi.y = wi.y - sw.y - bu * uo.y;
            reinterpret_cast<double2*>(u)[i] = make_double2((uo.x - su.x) * inv_rho, (uo.y - su.y) * inv_rho);
            reinterpret_cast<double2*>(w)[i] = wi;
            nrm[0] = fma(wi.x, wi.x, fma(wi.y, wi.y, nrm[0]));
        }
    }
    // scalar path: everything (VEC = false) or the odd last element
    for (int64_t i = (VEC ? (n & ~(int64_t)1) : 0) + (int64_t)blockIdx.x * TPB + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * TPB) {
        double su = 0.0, sw = 0.0;
        for (int j = 0; j < k; ++j) {
            const double v = V[(int64_t)j * ld + i];
            su = fma(hc[j], v, su);
            sw = fma(ha[j], v, sw);
        }
        const double uo = u[i];
        const double wi = w[i] - sw - bu * uo;
        u[i] = (uo - su) * inv_rho;
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
k_axpby(int64_t n, double a, const double* x, double b, double* y) {     // x may alias y
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

// ---------------------------------------------------------------------- DIA SpMV
// Stencil operators whose entries sit on a handful of constant offsets col - row (the pressure
// Laplacian of the cavity: 5) are ALSO kept in diagonal storage, fp32: ndiag values per row, no
// column indices, and the reads of x are coalesced runs instead of gathers.  20 B per row for
// the 5-point operator against 44 in fp32 CSR; a row per thread.  Same epilogues as the CSR
// kernels (sums, x, y fp64).
constexpr int DIA_MAX = 8;
struct dia_offsets { int32_t off[DIA_MAX]; };

// ND: compile-time bound on ndiag (register arrays); R rows per thread, grid-stride over rounds.
// Every load of a round has to be in flight before the first is used (the kernel has nothing
// else to hide the latency with): rounds whose rows all have their ND neighbours in range --
// all but the first and last few blocks -- take a path without bounds predicates, which ptxas
// schedules as one batch of loads (with the predicated address clamps it interleaved loads and
// their conversions, ~3 loads in flight, 3.2 TB/s); EXACT: ndiag == ND, no runtime trip count.
template <int ND, bool EXACT>
__device__ __forceinline__ void dia_load_row(bool fast, int ndiag, const dia_offsets& offs,
                                             const float* __restrict__ vals, int64_t ld,
                                             const double* __restrict__ x, int64_t r, int64_t ncols,
                                             bool live, float (&v)[ND], double (&xv)[ND]) {
    if (fast) {
#pragma unroll
        for (int k = 0; k < ND; ++k) {
            v[k] = 0.0f; xv[k] = 0.0;
            if (EXACT || k < ndiag) {
                v[k] = __ldcs(vals + (int64_t)k * ld + r);       // streamed once
                xv[k] = __ldg(x + (r + offs.off[k]));
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < ND; ++k) {
            v[k] = 0.0f; xv[k] = 0.0;
            if ((EXACT || k < ndiag) && live) {
                const int64_t c = r + offs.off[k];
                const bool ok = c >= 0 && c < ncols;
                v[k] = ok ? __ldcs(vals + (int64_t)k * ld + r) : 0.0f;
                xv[k] = ok ? __ldg(x + c) : 0.0;
            }
        }
    }
}

// rows [lo, hi) all have every neighbour in [0, ncols): block-uniform
__device__ __forceinline__ bool dia_interior(const dia_offsets& offs, int ndiag, int64_t lo, int64_t hi,
                                             int64_t nrows, int64_t ncols) {
    int32_t mn = 0, mx = 0;
#pragma unroll
    for (int k = 0; k < DIA_MAX; ++k)
        if (k < ndiag) { mn = min(mn, offs.off[k]); mx = max(mx, offs.off[k]); }
    return hi <= nrows && lo + mn >= 0 && hi - 1 + mx < ncols;
}

// sum of one value per thread over the whole grid, deterministic: per-block partials, the last
// block to finish adds them in a fixed order (scratch of the reductions: partial / ticket)
__device__ __forceinline__ void grid_sum_to(double v, double* __restrict__ partial,
                                            unsigned int* __restrict__ ticket, double* __restrict__ out) {
    __shared__ double gs_smem[TPB / 32];
    __shared__ bool gs_last;
    double a[1] = {v};
    block_sum<1>(a, gs_smem);
    // only thread 0 publishes: a fence by every thread would wait for the block's own result
    // stores to drain before the block can retire (measured: +40 us on a 40 us kernel of 8192
    // short blocks)
    if (threadIdx.x == 0) {
        partial[blockIdx.x] = gs_smem[0];
        __threadfence();
        gs_last = atomicAdd(ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (gs_last) {
        __threadfence();
        double t[1] = {0.0};
        for (int b = threadIdx.x; b < (int)gridDim.x; b += TPB) t[0] += __ldcg(partial + b);
        __syncthreads();
        block_sum<1>(t, gs_smem);
        if (threadIdx.x == 0) { *out = gs_smem[0]; *ticket = 0; }
    }
}

// DOT: the Jacobi epilogue (mode 3) also returns b . y (the r . z of a preconditioned CG step
// whose preconditioner ends with this sweep)
// NW > 0: hybrid image -- the entries that are on none of the diagonals sit in an ELL part of
// width ell_w <= NW (column-major like the diagonals: ell_idx / ell_val [k * ld + r], unused slots
// hold a zero value and a valid column): the momentum block of the cavity has 5 of its 9 entries
// per row on 7 diagonals and 4 cross-coupling entries whose offsets drift with the grid row --
// 4 gathers per row instead of 9, and no row pointers.
template <int ND, int R, bool DOT, bool EXACT, int NW>
__global__ void __launch_bounds__(TPB)
k_spmv_dia(int64_t nrows, int64_t ncols, int ndiag, int64_t ld, dia_offsets offs,
           const float* __restrict__ vals, const double* __restrict__ x, spmv_epi e,
           double* __restrict__ y, double* __restrict__ partial, unsigned int* __restrict__ ticket,
           double* __restrict__ dot_out, int ell_w, const int32_t* __restrict__ ell_idx,
           const float* __restrict__ ell_val) {
    double dacc = 0.0;
    for (int64_t base = (int64_t)blockIdx.x * (TPB * R); base < nrows;
         base += (int64_t)gridDim.x * (TPB * R)) {
        const bool fast = dia_interior(offs, ndiag, base, base + TPB * R, nrows, ncols);
        float v[R][ND];
        double xv[R][ND];
        double bv[R], dv[R], xr[R];
        int32_t ei[R][NW > 0 ? NW : 1];
        float ev[R][NW > 0 ? NW : 1];
        if (NW > 0) {
            // the column indices of the ELL part first: their gathers depend on them
#pragma unroll
            for (int j = 0; j < R; ++j) {
                const int64_t r = base + (int64_t)j * TPB + threadIdx.x;
#pragma unroll
                for (int k = 0; k < NW; ++k) {
                    ei[j][k] = 0; ev[j][k] = 0.0f;
                    if (k < ell_w && r < nrows) {
                        ei[j][k] = __ldcs(ell_idx + (int64_t)k * ld + r);
                        ev[j][k] = __ldcs(ell_val + (int64_t)k * ld + r);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const int64_t r = base + (int64_t)j * TPB + threadIdx.x;
            const bool live = r < nrows;
            bv[j] = dv[j] = xr[j] = 0.0;
            dia_load_row<ND, EXACT>(fast, ndiag, offs, vals, ld, x, r, ncols, live, v[j], xv[j]);
            if (live && e.mode >= 1) bv[j] = __ldg(e.b + r);
            if (live && e.mode == 3) { dv[j] = __ldg(e.dinv + r); xr[j] = __ldg(x + r); }
        }
        double ex[R][NW > 0 ? NW : 1];
        if (NW > 0) {
#pragma unroll
            for (int j = 0; j < R; ++j)
#pragma unroll
                for (int k = 0; k < NW; ++k) ex[j][k] = (k < ell_w) ? __ldg(x + ei[j][k]) : 0.0;
        }
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const int64_t r = base + (int64_t)j * TPB + threadIdx.x;
            if (r >= nrows) continue;
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < ND; ++k) s = fma((double)v[j][k], xv[j][k], s);
            if (NW > 0) {
#pragma unroll
                for (int k = 0; k < NW; ++k) s = fma((double)ev[j][k], ex[j][k], s);
            }
            double out;
            switch (e.mode) {
                case 0: out = s; break;
                case 1: out = bv[j] - s; break;
                case 2: out = bv[j] + s; break;
                default: out = xr[j] + e.omega * dv[j] * (bv[j] - s); break;
            }
            y[r] = out;
            if (DOT) dacc = fma(bv[j], out, dacc);
        }
    }
    if (DOT) grid_sum_to(dacc, partial, ticket, dot_out);
}

// One kernel for the direction update of a CG step on a DIA operator L, scalars on the device:
//   beta = rz_new / rz_old (first: 0) ;  p <- z + beta p ;  q <- L z + beta q  (= L p) ;  *pq = p . q
// instead of p-update, SpMV and inner product: z is read once, p and q once each.
template <int ND, int R, bool EXACT>
__global__ void __launch_bounds__(TPB)
k_dia_cg_dir(int64_t nrows, int ndiag, int64_t ld, dia_offsets offs, const float* __restrict__ vals,
             const double* __restrict__ z, const double* __restrict__ rz_new,
             const double* __restrict__ rz_old, int first, double* __restrict__ p,
             double* __restrict__ q, double* __restrict__ partial, unsigned int* __restrict__ ticket,
             double* __restrict__ pq_out) {
    double beta = 0.0;
    if (!first) {
        const double den = *rz_old;
        beta = den != 0.0 ? *rz_new / den : 0.0;
    }
    double dacc = 0.0;
    for (int64_t base = (int64_t)blockIdx.x * (TPB * R); base < nrows;
         base += (int64_t)gridDim.x * (TPB * R)) {
        const bool fast = dia_interior(offs, ndiag, base, base + TPB * R, nrows, nrows);
        float v[R][ND];
        double zv[R][ND];
        double zr[R], po[R], qo[R];
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const int64_t r = base + (int64_t)j * TPB + threadIdx.x;
            const bool live = r < nrows;
            dia_load_row<ND, EXACT>(fast, ndiag, offs, vals, ld, z, r, nrows, live, v[j], zv[j]);
            zr[j] = po[j] = qo[j] = 0.0;
            if (live) {
                zr[j] = __ldg(z + r);
                if (!first) { po[j] = p[r]; qo[j] = q[r]; }
            }
        }
#pragma unroll
        for (int j = 0; j < R; ++j) {
            const int64_t r = base + (int64_t)j * TPB + threadIdx.x;
            if (r >= nrows) continue;
            double s = 0.0;
#pragma unroll
            for (int k = 0; k < ND; ++k) s = fma((double)v[j][k], zv[j][k], s);
            const double pn = fma(beta, po[j], zr[j]), qn = fma(beta, qo[j], s);
            p[r] = pn;
            q[r] = qn;
            dacc = fma(pn, qn, dacc);
        }
    }
    grid_sum_to(dacc, partial, ticket, pq_out);
}

// rows per thread (tuning switch npb_spmv_dia_set_rows).  Two: with every load of a round issued
// up front, 2 x (2 ND + 3) loads in flight per thread reach 5.0 / 6.4 TB/s on the residual / Jacobi
// epilogue of the 2048^2 pressure operator (one row: 4.5; tools/dia_bench.py)
static int npb_dia_rows = 2;

// CSR -> DIA image: vals[k * ld + r] = entry of row r on offset offs[k] (zero-filled by the
// caller).  Entries on other offsets: ell_w == 0 -> counted in *bad; ell_w > 0 -> they go to the ELL
// part (slot order = column order; more than ell_w of them in a row count as bad).  Unused ELL
// slots: value 0, column min(r, ncols - 1).
__global__ void __launch_bounds__(TPB)
k_csr_to_dia(int64_t nrows, int64_t ncols, const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
             const double* __restrict__ dat, int ndiag, int64_t ld, dia_offsets offs,
             float* __restrict__ vals, int64_t* __restrict__ bad, int ell_w,
             int32_t* __restrict__ ell_idx, float* __restrict__ ell_val) {
    const int64_t r = (int64_t)blockIdx.x * TPB + threadIdx.x;
    if (r >= nrows) return;
    int w = 0;
    for (int32_t k = ptr[r], e = ptr[r + 1]; k < e; ++k) {
        const int32_t o = idx[k] - (int32_t)r;
        int hit = -1;
        for (int d = 0; d < ndiag; ++d)
            if (offs.off[d] == o) hit = d;
        if (hit >= 0) vals[(int64_t)hit * ld + r] = (float)dat[k];
        else if (w < ell_w) {
            ell_idx[(int64_t)w * ld + r] = idx[k];
            ell_val[(int64_t)w * ld + r] = (float)dat[k];
            ++w;
        } else atomicAdd((unsigned long long*)bad, 1ull);
    }
    for (; w < ell_w; ++w) {
        ell_idx[(int64_t)w * ld + r] = (int32_t)(r < ncols ? r : ncols - 1);
        ell_val[(int64_t)w * ld + r] = 0.0f;
    }
}

// *maxw = largest number of entries of a row that are on none of the offsets
__global__ void __launch_bounds__(TPB)
k_dia_remainder(int64_t nrows, const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx,
                int ndiag, dia_offsets offs, int32_t* __restrict__ maxw) {
    const int64_t r = (int64_t)blockIdx.x * TPB + threadIdx.x;
    int w = 0;
    if (r < nrows) {
        for (int32_t k = ptr[r], e = ptr[r + 1]; k < e; ++k) {
            const int32_t o = idx[k] - (int32_t)r;
            bool hit = false;
            for (int d = 0; d < ndiag; ++d) hit |= offs.off[d] == o;
            w += hit ? 0 : 1;
        }
    }
#pragma unroll
    for (int d = 16; d; d >>= 1) w = max(w, __shfl_xor_sync(0xffffffffu, w, d));
    if ((threadIdx.x & 31) == 0 && w > 0) atomicMax(maxw, w);
}

#define ST(s) ((cudaStream_t)(s))

// scratch: RED_BLOCKS*KC*2 doubles of partials (two right-hand sides) followed by one uint ticket
static inline double* red_partial(void* scratch) { return (double*)scratch; }
static inline unsigned int* red_ticket(void* scratch) {
    return (unsigned int*)((double*)scratch + (size_t)RED_BLOCKS * KC * 2);
}
constexpr int64_t DIA_GRID_CAP = (int64_t)NPB_NUM_SMS * 64;
// persistent grid of the kernels with a fused inner product: 0 = as many CTAs as are resident
// (2 or 3 per SM by registers; measured best, tools/dia_bench.py), else CTAs per SM (tuning)
static int npb_dia_dot_ctas = 0;
template <class K>
static int64_t dia_dot_grid(K kernel) {
    if (npb_dia_dot_ctas > 0) return (int64_t)NPB_NUM_SMS * npb_dia_dot_ctas;
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, TPB, 0) != cudaSuccess || nb < 1) nb = 2;
    return (int64_t)NPB_NUM_SMS * nb;
}
static_assert(DIA_GRID_CAP <= (int64_t)RED_BLOCKS * KC * 2, "DIA dot partials fit the reduction scratch");

// dot_out != NULL (mode 3 only): also *dot_out = b . y, through the reduction scratch `red`
int npb_spmv_dia_launch(int64_t nrows, int64_t ncols, int ndiag, int64_t ld,
                        const int32_t* offsets_host, const float* vals, const double* x, int mode,
                        const double* b, const double* dinv, double omega, double* y,
                        cudaStream_t st, double* dot_out = nullptr, void* red = nullptr,
                        int ell_w = 0, const int32_t* ell_idx = nullptr, const float* ell_val = nullptr) {
    if (nrows <= 0) return NPB_OK;
    if (ndiag < 1 || ndiag > DIA_MAX || ld < nrows || (dot_out && (mode != 3 || !red)) || ell_w < 0 ||
        ell_w > 8 || (ell_w > 0 && (!ell_idx || !ell_val || dot_out)))
        return NPB_ERR_ARG;
    dia_offsets o;
    for (int k = 0; k < DIA_MAX; ++k) o.off[k] = k < ndiag ? offsets_host[k] : 0;
    spmv_epi e{mode, b, dinv, omega, nullptr, 0x7fffffff, {nullptr, 0, 0, 0}};
    // x is only read and never aliases y here (the Jacobi epilogue requires y != x): the kernel
    // reads it through the non-coherent path
    const int R = npb_dia_rows;
    int64_t grid = (nrows + (int64_t)TPB * R - 1) / ((int64_t)TPB * R);
    if (grid > DIA_GRID_CAP) grid = DIA_GRID_CAP;
    double* pa = red ? red_partial(red) : nullptr;
    unsigned int* ti = red ? red_ticket(red) : nullptr;
    // exact instantiations for the 3- and 5-point stencils, a bounded generic one otherwise; with
    // the inner product: a persistent grid (the per-block reduction + ticket at the end of
    // thousands of one-round blocks doubled the kernel's time)
#define NPB_DIA_LAUNCH_W(ND_, R_, DOT_, EX_, NW_) do { \
        int64_t g_ = grid; \
        if (DOT_) { const int64_t c_ = dia_dot_grid(k_spmv_dia<ND_, R_, DOT_, EX_, NW_>); if (g_ > c_) g_ = c_; } \
        k_spmv_dia<ND_, R_, DOT_, EX_, NW_><<<(int)g_, TPB, 0, st>>>(nrows, ncols, ndiag, ld, o, vals, x, e, y, pa, ti, dot_out, ell_w, ell_idx, ell_val); \
    } while (0)
#define NPB_DIA_LAUNCH(ND_, R_, DOT_, EX_) NPB_DIA_LAUNCH_W(ND_, R_, DOT_, EX_, 0)
    if (ell_w > 0) {
        // hybrid image: two rows per thread, ELL width bounded by 4 or 8
        grid = (nrows + (int64_t)TPB * 2 - 1) / ((int64_t)TPB * 2);
        if (grid > DIA_GRID_CAP) grid = DIA_GRID_CAP;
        if (ell_w <= 4) {
            if (ndiag == 5) NPB_DIA_LAUNCH_W(5, 2, false, true, 4); else if (ndiag == 7) NPB_DIA_LAUNCH_W(7, 2, false, true, 4); else NPB_DIA_LAUNCH_W(8, 2, false, false, 4);
        } else {
            if (ndiag == 5) NPB_DIA_LAUNCH_W(5, 2, false, true, 8); else if (ndiag == 7) NPB_DIA_LAUNCH_W(7, 2, false, true, 8); else NPB_DIA_LAUNCH_W(8, 2, false, false, 8);
        }
    } else if (dot_out && R == 2) {
        if (ndiag == 3) NPB_DIA_LAUNCH(3, 2, true, true); else if (ndiag == 5) NPB_DIA_LAUNCH(5, 2, true, true); else NPB_DIA_LAUNCH(8, 2, true, false);
    } else if (dot_out) {
        if (ndiag == 3) NPB_DIA_LAUNCH(3, 1, true, true); else if (ndiag == 5) NPB_DIA_LAUNCH(5, 1, true, true); else NPB_DIA_LAUNCH(8, 1, true, false);
    } else if (R == 2) {
        if (ndiag == 3) NPB_DIA_LAUNCH(3, 2, false, true); else if (ndiag == 5) NPB_DIA_LAUNCH(5, 2, false, true); else NPB_DIA_LAUNCH(8, 2, false, false);
    } else {
        if (ndiag == 3) NPB_DIA_LAUNCH(3, 1, false, true); else if (ndiag == 5) NPB_DIA_LAUNCH(5, 1, false, true); else NPB_DIA_LAUNCH(8, 1, false, false);
    }
#undef NPB_DIA_LAUNCH
#undef NPB_DIA_LAUNCH_W
    NPB_COUNT_LAUNCH();
    NPB_COUNT_BYTES(8.0 * ell_w * nrows);
    NPB_COUNT_BYTES(4.0 * ndiag * nrows + 8.0 * nrows * (2 + (mode >= 1) + 2 * (mode == 3)));
    NPB_CHECK_LAUNCH("spmv_dia");
    return NPB_OK;
}

// p <- z + beta p, q <- L z + beta q, *pq = p . q with beta = *rz_new / *rz_old (first: 0); L square
int npb_dia_cg_dir_launch(int64_t nrows, int ndiag, int64_t ld, const int32_t* offsets_host,
                          const float* vals, const double* z, const double* rz_new,
                          const double* rz_old, int first, double* p, double* q, double* pq_out,
                          void* red, cudaStream_t st) {
    if (nrows <= 0) return NPB_OK;
    if (ndiag < 1 || ndiag > DIA_MAX || ld < nrows || !red) return NPB_ERR_ARG;
    dia_offsets o;
    for (int k = 0; k < DIA_MAX; ++k) o.off[k] = k < ndiag ? offsets_host[k] : 0;
    const int R = npb_dia_rows;
    int64_t grid = (nrows + (int64_t)TPB * R - 1) / ((int64_t)TPB * R);
#define NPB_DIA_LAUNCH(ND_, R_, EX_) do { \
        int64_t g_ = grid; \
        const int64_t c_ = dia_dot_grid(k_dia_cg_dir<ND_, R_, EX_>); if (g_ > c_) g_ = c_; \
        k_dia_cg_dir<ND_, R_, EX_><<<(int)g_, TPB, 0, st>>>(nrows, ndiag, ld, o, vals, z, rz_new, rz_old, first, p, q, red_partial(red), red_ticket(red), pq_out); \
    } while (0)
    if (R == 2) {
        if (ndiag == 3) NPB_DIA_LAUNCH(3, 2, true); else if (ndiag == 5) NPB_DIA_LAUNCH(5, 2, true); else NPB_DIA_LAUNCH(8, 2, false);
    } else {
        if (ndiag == 3) NPB_DIA_LAUNCH(3, 1, true); else if (ndiag == 5) NPB_DIA_LAUNCH(5, 1, true); else NPB_DIA_LAUNCH(8, 1, false);
    }
#undef NPB_DIA_LAUNCH
    NPB_COUNT_LAUNCH();
    NPB_COUNT_BYTES(4.0 * ndiag * nrows + 8.0 * nrows * (first ? 3 : 5));
    NPB_CHECK_LAUNCH("dia_cg_dir");
    return NPB_OK;
}

// 1: the auto dispatch prefers the TMA pipeline kernel (npb_spmv_set_default_tma)
static int npb_spmv_default_tma = 1;

// the non-persistent kernels (operators too small for the TMA pipeline, forced variants)
template <bool HALO>
static void spmv_small_dispatch(int64_t nrows, int64_t nnz, const int32_t* ptr, const int32_t* idx,
                                const double* dat, const double* x, const spmv_epi& e, double* y,
                                int variant, double mean, bool aligned, cudaStream_t st) {
    if ((variant == 4 || (variant == 0 && nrows >= 256 * NPB_NUM_SMS * 4 && mean <= 24.0)) && aligned) {
        k_spmv_stream_v4<256, HALO><<<npb_blocks(nrows, 256), SP_THREADS, 0, st>>>(nrows, nnz, ptr, idx, dat, x, e, y);
    } else if (variant == 3 && aligned) {
        k_spmv_stream_async<256, HALO><<<npb_blocks(nrows, 256), SP_THREADS, 0, st>>>(nrows, nnz, ptr, idx, dat, x, e, y);
    } else if (variant == 1 || variant == 3 || variant == 4 || (variant == 0 && mean <= 96.0)) {
        // rows per CTA: as many as keep >= ~4 CTAs per SM in flight and the
        // row block inside a few shared-memory tiles
        if (nrows >= 256 * NPB_NUM_SMS * 4 && mean <= 24.0)
            k_spmv_stream<256, HALO><<<npb_blocks(nrows, 256), SP_THREADS, 0, st>>>(nrows, ptr, idx, dat, x, e, y);
        else if (nrows >= 64 * NPB_NUM_SMS * 4 && mean <= 64.0)
            k_spmv_stream<64, HALO><<<npb_blocks(nrows, 64), SP_THREADS, 0, st>>>(nrows, ptr, idx, dat, x, e, y);
        else
            k_spmv_stream<32, HALO><<<npb_blocks(nrows, 32), SP_THREADS, 0, st>>>(nrows, ptr, idx, dat, x, e, y);
    } else if (mean <= 48.0) {
        k_spmv<16, HALO><<<npb_blocks(nrows * 16, TPB), TPB, 0, st>>>(nrows, ptr, idx, dat, x, e, y);
    } else {
        k_spmv<32, HALO><<<npb_blocks(nrows * 32, TPB), TPB, 0, st>>>(nrows, ptr, idx, dat, x, e, y);
    }
}

// ------------------------------------------------------------- internal API
// y = epilogue(A x); see spmv_epi.  `variant`: 0 auto, 1 stream, 2 vector, 3 stream + cp.async
int npb_spmv_ex(int64_t nrows, int64_t nnz, const int32_t* ptr, const int32_t* idx,
                const double* dat, const double* x, int mode, const double* b,
                const double* dinv, double omega, double* y, int variant, cudaStream_t st,
                const npb_halo* halo, const npb_wait* wait, const float* dat32) {
    if (nrows <= 0) return NPB_OK;
    if (dat32) {
        // operator stored with fp32 values (multigrid preconditioner, single GPU): always the
        // TMA pipeline kernel; 8 instead of 12 bytes per entry
        spmv_epi e32{mode, b, dinv, omega, nullptr, 0x7fffffff, {nullptr, 0, 0, 0}};
        if (halo && halo->dist) { e32.xh = halo->recv; e32.n_own = (int32_t)halo->n_own; }
        if (wait && e32.xh) e32.wait = *wait;
        if ((((uintptr_t)idx | (uintptr_t)dat32) & 15) != 0 || nnz >= (int64_t)0x7fffffff - 2 * TM_CAP_MAX) {
            npb_set_error("spmv: fp32 operator not usable here (alignment / size)");
            return NPB_ERR_ARG;
        }
        int rc = spmv_tma_dispatch(nrows, nnz, ptr, idx, dat32, x, e32, y, st);
        if (rc) return rc;
        NPB_COUNT_LAUNCH();
        NPB_COUNT_BYTES(8.0 * nnz + 4.0 * (nrows + 1) + 8.0 * nrows * (2 + (mode >= 1) + 2 * (mode == 3)));
        NPB_CHECK_LAUNCH("spmv_f32");
        return NPB_OK;
    }
    spmv_epi e{mode, b, dinv, omega, nullptr, 0x7fffffff, {nullptr, 0, 0, 0}};
    if (halo && halo->dist) { e.xh = halo->recv; e.n_own = (int32_t)halo->n_own; }
    if (wait && e.xh) e.wait = *wait;
    const double mean = (double)nnz / (double)nrows;
    const bool aligned = (((uintptr_t)idx | (uintptr_t)dat) & 15) == 0;
    const bool tma_ok = aligned && nnz < (int64_t)0x7fffffff - 2 * TM_CAP_MAX;
    if (tma_ok && (variant == 5 || (variant == 0 && npb_spmv_default_tma && mean <= 96.0 &&
                                    nrows >= tma_min_rows(mean)))) {
        int rc = spmv_tma_dispatch(nrows, nnz, ptr, idx, dat, x, e, y, st);
        if (rc) return rc;
    } else if (e.xh) {
        spmv_small_dispatch<true>(nrows, nnz, ptr, idx, dat, x, e, y, variant, mean, aligned, st);
    } else {
        spmv_small_dispatch<false>(nrows, nnz, ptr, idx, dat, x, e, y, variant, mean, aligned, st);
    }
    NPB_COUNT_LAUNCH();
    // CSR once + x + y (+ b, + dinv and x[row] for the Jacobi epilogue); x counted with the row count
    NPB_COUNT_BYTES(12.0 * nnz + 4.0 * (nrows + 1) + 8.0 * nrows * (2 + (mode >= 1) + 2 * (mode == 3)));
    NPB_CHECK_LAUNCH("spmv");
    return NPB_OK;
}

int npb_spmv_launch(int64_t nrows, int64_t nnz, const int32_t* ptr, const int32_t* idx,
                    const double* dat, const double* x, const double* b, double* y,
                    int mode, cudaStream_t st) {
    return npb_spmv_ex(nrows, nnz, ptr, idx, dat, x, mode, b, nullptr, 0.0, y, 0, st, nullptr, nullptr, nullptr);
}


int npb_dot_multi_launch(int64_t n, int k, const double* V, int64_t ld, const double* w,
                         double* out, void* scratch, cudaStream_t st) {
    // (the pressure space of the cavity has N^2 - 1 unknowns: odd lengths must take the fast path too)
    const bool vec = ((((uintptr_t)V | (uintptr_t)w) & 15) == 0) && (ld % 2 == 0 || k <= 1);
    for (int j0 = 0; j0 < k; j0 += KC) {
        int kc = k - j0 < KC ? k - j0 : KC;
        if (vec && kc <= 2 && n >= 4096)
            k_dot_few<2><<<grid_for(n / 4), TPB, 0, st>>>(n, kc, V + (int64_t)j0 * ld, ld, w,
                                                        red_partial(scratch), red_ticket(scratch), out + j0);
        else if (vec && kc <= 6 && n >= 4096)
            k_dot_few<6><<<grid_for(n / 4), TPB, 0, st>>>(n, kc, V + (int64_t)j0 * ld, ld, w,
                                                        red_partial(scratch), red_ticket(scratch), out + j0);
        else if (vec && n >= 4096)
            k_dot_block<KC, 1><<<grid_for(n / 2), TPB, 0, st>>>(n, kc, V + (int64_t)j0 * ld, ld, w, w,
                                                               red_partial(scratch), red_ticket(scratch),
                                                               out + j0, out + j0);
        else
            k_dot_multi<<<grid_for(n), TPB, 0, st>>>(n, kc, V + (int64_t)j0 * ld, ld, w,
                                                     red_partial(scratch), red_ticket(scratch),
                                                     out + j0);
        NPB_COUNT_LAUNCH();
    }
    NPB_COUNT_BYTES(8.0 * n * (k + 1));
    NPB_CHECK_LAUNCH("dot_multi");
    return NPB_OK;
}

// out0[j] = V_j . w0, out1[j] = V_j . w1 (j < k): V is read once for both
int npb_dot_multi2_launch(int64_t n, int k, const double* V, int64_t ld, const double* w0,
                          const double* w1, double* out0, double* out1, void* scratch,
                          cudaStream_t st) {
    const bool vec = ((((uintptr_t)V | (uintptr_t)w0 | (uintptr_t)w1) & 15) == 0) && (ld % 2 == 0 || k <= 1);
    if (!vec) {
        int rc = npb_dot_multi_launch(n, k, V, ld, w0, out0, scratch, st);
        return rc ? rc : npb_dot_multi_launch(n, k, V, ld, w1, out1, scratch, st);
    }
    for (int j0 = 0; j0 < k; j0 += KC) {
        int kc = k - j0 < KC ? k - j0 : KC;
        if (kc <= 8)
            k_dot_block<8, 2><<<grid_for(n / 2), TPB, 0, st>>>(n, kc, V + (int64_t)j0 * ld, ld, w0, w1,
                                                              red_partial(scratch), red_ticket(scratch),
                                                              out0 + j0, out1 + j0);
        else
            k_dot_block<KC, 2><<<grid_for(n / 2), TPB, 0, st>>>(n, kc, V + (int64_t)j0 * ld, ld, w0, w1,
                                                               red_partial(scratch), red_ticket(scratch),
                                                               out0 + j0, out1 + j0);
        NPB_COUNT_LAUNCH();
    }
    NPB_COUNT_BYTES(8.0 * n * (k + 2));
    NPB_CHECK_LAUNCH("dot_multi2");
    return NPB_OK;
}

int npb_axpy_multi_launch(int64_t n, int k, const double* V, int64_t ld, const double* h,
                          double sign, double* w, double* out_norm2, void* scratch,
                          cudaStream_t st) {
    size_t sh = sizeof(double) * ((size_t)k + TPB / 32);
    const bool vec = ((((uintptr_t)V | (uintptr_t)w) & 15) == 0) && (ld % 2 == 0 || k <= 1);
    if (vec)
        k_axpy_multi<true><<<grid_for((n + 1) / 2), TPB, sh, st>>>(n, k, V, ld, h, sign, w, red_partial(scratch),
                                                                  red_ticket(scratch), out_norm2);
    else
        k_axpy_multi<false><<<grid_for(n), TPB, sh, st>>>(n, k, V, ld, h, sign, w, red_partial(scratch),
                                                         red_ticket(scratch), out_norm2);
    NPB_COUNT_LAUNCH();
    NPB_COUNT_BYTES(8.0 * n * (k + 2));
    NPB_CHECK_LAUNCH("axpy_multi");
    return NPB_OK;
}

int npb_axpy_multi2_launch(int64_t n, int k, const double* V, int64_t ld, const double* coef,
                           double* u, double* w, double* out_norm2, void* scratch,
                           cudaStream_t st) {
    size_t sh = sizeof(double) * ((size_t)2 * k + 2 + TPB / 32);
    const bool vec = ((((uintptr_t)V | (uintptr_t)u | (uintptr_t)w) & 15) == 0) && (ld % 2 == 0 || k <= 1);
    if (vec)
        k_axpy_multi2<true><<<grid_for((n + 1) / 2), TPB, sh, st>>>(n, k, V, ld, coef, u, w, red_partial(scratch),
                                                                   red_ticket(scratch), out_norm2);
    else
        k_axpy_multi2<false><<<grid_for(n), TPB, sh, st>>>(n, k, V, ld, coef, u, w, red_partial(scratch),
                                                          red_ticket(scratch), out_norm2);
    NPB_COUNT_LAUNCH();
    NPB_COUNT_BYTES(8.0 * n * (k + 4));
    NPB_CHECK_LAUNCH("axpy_multi2");
    return NPB_OK;
}

int npb_axpby_launch(int64_t n, double a, const double* x, double b, double* y, cudaStream_t st) {
    if (n <= 0) return NPB_OK;
    k_axpby<<<grid_for(n), TPB, 0, st>>>(n, a, x, b, y);
    NPB_COUNT_LAUNCH();
    NPB_COUNT_BYTES(8.0 * n * (b == 0.0 ? 2 : 3));
    NPB_CHECK_LAUNCH("axpby");
    return NPB_OK;
}

extern "C" {

int64_t npb_reduce_scratch_bytes(void) {
    return (int64_t)(sizeof(double) * (size_t)RED_BLOCKS * KC * 2 + 64);
}

int npb_spmv(int64_t nrows, int64_t nnz, const int32_t* indptr, const int32_t* indices,
             const double* data, const double* x, double* y, void* stream) {
    return npb_spmv_launch(nrows, nnz, indptr, indices, data, x, nullptr, y, 0, ST(stream));
}

// tuning switch: 0 = the auto dispatch never picks the TMA pipeline kernel
int npb_spmv_set_default_tma(int on) {
    const int old = npb_spmv_default_tma;
    npb_spmv_default_tma = on ? 1 : 0;
    return old;
}

// tuning switch: ring geometry of the TMA kernel (-1: by row length (default); 0: 3072 x 3
// stages, 2 CTAs/SM; 1: 2304 x 3, 2; 2: 2304 x 2, 3; 3: 1536 x 3, 4); returns the old one
int npb_spmv_set_tma_geometry(int g) {
    const int old = npb_tma_geometry;
    if (g >= -1 && g <= 4) npb_tma_geometry = g;
    return old;
}

// tuning switch: fp32-stored operators with 5..20-entry rows on 3 resident CTAs per SM (default 1)
int npb_spmv_set_tma_f32_three(int on) {
    const int old = npb_tma_f32_three;
    npb_tma_f32_three = on ? 1 : 0;
    return old;
}

// same, forcing a kernel variant (1 = stream, 2 = vector, 3 = stream + cp.async,
// 4 = stream + 128-bit loads, 5 = persistent TMA pipeline); for tests / tuning
int npb_spmv_variant(int64_t nrows, int64_t nnz, const int32_t* indptr, const int32_t* indices,
                     const double* data, const double* x, double* y, int variant, void* stream) {
    return npb_spmv_ex(nrows, nnz, indptr, indices, data, x, 0, nullptr, nullptr, 0.0, y, variant,
                       ST(stream), nullptr, nullptr, nullptr);
}

// y = epilogue(A x) with a forced kernel variant (0 = auto):
//   mode 0: y = Ax   1: y = b - Ax   2: y = b + Ax   3: y = x + omega dinv.(b - Ax)  (y != x)
int npb_spmv_epilogue(int64_t nrows, int64_t nnz, const int32_t* indptr, const int32_t* indices,
                      const double* data, const double* x, int mode, const double* b,
                      const double* dinv, double omega, double* y, int variant, void* stream) {
    if (mode < 0 || mode > 3 || (mode >= 1 && !b) || (mode == 3 && (!dinv || x == y))) {
        npb_set_error("spmv_epilogue: bad mode / operands");
        return NPB_ERR_ARG;
    }
    return npb_spmv_ex(nrows, nnz, indptr, indices, data, x, mode, b, dinv, omega, y, variant,
                       ST(stream), nullptr, nullptr, nullptr);
}

// y = epilogue(A x) for an operator whose values are stored in fp32 (x, y, b, dinv and the
// row sums are fp64): the preconditioner's SpMV, 8 bytes per stored entry
int npb_spmv_f32(int64_t nrows, int64_t nnz, const int32_t* indptr, const int32_t* indices,
                 const float* data32, const double* x, int mode, const double* b,
                 const double* dinv, double omega, double* y, void* stream) {
    if (mode < 0 || mode > 3 || (mode >= 1 && !b) || (mode == 3 && (!dinv || x == y)) || !data32) {
        npb_set_error("spmv_f32: bad mode / operands");
        return NPB_ERR_ARG;
    }
    return npb_spmv_ex(nrows, nnz, indptr, indices, nullptr, x, mode, b, dinv, omega, y, 0,
                       ST(stream), nullptr, nullptr, data32);
}

// diagonal-storage image of a CSR operator whose entries sit on `ndiag` (<= 8) constant offsets
// col - row (offsets: [host]); vals: ndiag * ld floats (zeroed here), diagonal k at vals + k * ld
// (ld >= nrows; a multiple of 32 keeps every diagonal 128-byte aligned); *bad (device, reset
// here) = entries found on other offsets (the image is then unusable)
int npb_csr_to_dia(int64_t nrows, const int32_t* indptr, const int32_t* indices, const double* data,
                   int ndiag, const int32_t* offsets, float* vals, int64_t ld, int64_t* bad,
                   void* stream) {
    if (ndiag < 1 || ndiag > DIA_MAX || !bad || ld < nrows) return NPB_ERR_ARG;
    NPB_CUDA(cudaMemsetAsync(bad, 0, sizeof(int64_t), ST(stream)));
    if (nrows <= 0) return NPB_OK;
    NPB_CUDA(cudaMemsetAsync(vals, 0, sizeof(float) * (size_t)ndiag * (size_t)ld, ST(stream)));
    dia_offsets o;
    for (int k = 0; k < DIA_MAX; ++k) o.off[k] = k < ndiag ? offsets[k] : 0;
    k_csr_to_dia<<<npb_blocks(nrows, TPB), TPB, 0, ST(stream)>>>(nrows, nrows, indptr, indices, data, ndiag, ld, o,
                                                                vals, bad, 0, nullptr, nullptr);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("csr_to_dia");
    return NPB_OK;
}

// hybrid image (diagonals + ELL remainder) of an operator most of whose entries sit on `ndiag`
// constant offsets.  npb_csr_dia_remainder: *maxw (device, reset here) = the largest number of
// entries of a row on none of the offsets = the ELL width the image needs.  npb_csr_to_hyb: fills
// the diagonals (as npb_csr_to_dia) and the ELL part, ell_idx / ell_val: ell_w * ld entries, slot k
// of row r at [k * ld + r]; *bad = entries that fit neither.
int npb_csr_dia_remainder(int64_t nrows, const int32_t* indptr, const int32_t* indices, int ndiag,
                          const int32_t* offsets, int32_t* maxw, void* stream) {
    if (ndiag < 1 || ndiag > DIA_MAX || !maxw) return NPB_ERR_ARG;
    NPB_CUDA(cudaMemsetAsync(maxw, 0, sizeof(int32_t), ST(stream)));
    if (nrows <= 0) return NPB_OK;
    dia_offsets o;
    for (int k = 0; k < DIA_MAX; ++k) o.off[k] = k < ndiag ? offsets[k] : 0;
    k_dia_remainder<<<npb_blocks(nrows, TPB), TPB, 0, ST(stream)>>>(nrows, indptr, indices, ndiag, o, maxw);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("csr_dia_remainder");
    return NPB_OK;
}

int npb_csr_to_hyb(int64_t nrows, int64_t ncols, const int32_t* indptr, const int32_t* indices,
                   const double* data, int ndiag, const int32_t* offsets, float* vals, int64_t ld,
                   int ell_w, int32_t* ell_idx, float* ell_val, int64_t* bad, void* stream) {
    if (ndiag < 1 || ndiag > DIA_MAX || !bad || ld < nrows || ell_w < 1 || ell_w > 8 || !ell_idx || !ell_val ||
        ncols < 1)
        return NPB_ERR_ARG;
    NPB_CUDA(cudaMemsetAsync(bad, 0, sizeof(int64_t), ST(stream)));
    if (nrows <= 0) return NPB_OK;
    NPB_CUDA(cudaMemsetAsync(vals, 0, sizeof(float) * (size_t)ndiag * (size_t)ld, ST(stream)));
    dia_offsets o;
    for (int k = 0; k < DIA_MAX; ++k) o.off[k] = k < ndiag ? offsets[k] : 0;
    k_csr_to_dia<<<npb_blocks(nrows, TPB), TPB, 0, ST(stream)>>>(nrows, ncols, indptr, indices, data, ndiag, ld, o,
                                                                vals, bad, ell_w, ell_idx, ell_val);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("csr_to_hyb");
    return NPB_OK;
}

// y = epilogue(A x) from a hybrid image (same modes as npb_spmv_epilogue)
int npb_spmv_hyb(int64_t nrows, int64_t ncols, int ndiag, const int32_t* offsets, const float* vals,
                 int64_t ld, int ell_w, const int32_t* ell_idx, const float* ell_val, const double* x,
                 int mode, const double* b, const double* dinv, double omega, double* y, void* stream) {
    if (mode < 0 || mode > 3 || (mode >= 1 && !b) || (mode == 3 && (!dinv || x == y))) {
        npb_set_error("spmv_hyb: bad mode / operands");
        return NPB_ERR_ARG;
    }
    return npb_spmv_dia_launch(nrows, ncols, ndiag, ld, offsets, vals, x, mode, b, dinv, omega, y, ST(stream),
                               nullptr, nullptr, ell_w, ell_idx, ell_val);
}

// tuning switch: CTAs per SM of the DIA kernels with a fused inner product (0: resident count); returns the old value
int npb_spmv_dia_set_dot_ctas(int c) {
    const int old = npb_dia_dot_ctas;
    if (c >= 0 && c <= 64) npb_dia_dot_ctas = c;
    return old;
}

// tuning switch: rows per thread of the DIA kernel (1 or 2); returns the old value
int npb_spmv_dia_set_rows(int r) {
    const int old = npb_dia_rows;
    if (r == 1 || r == 2) npb_dia_rows = r;
    return old;
}

// y = epilogue(A x) from the diagonal-storage image (same modes as npb_spmv_epilogue)
int npb_spmv_dia(int64_t nrows, int64_t ncols, int ndiag, const int32_t* offsets, const float* vals,
                 int64_t ld, const double* x, int mode, const double* b, const double* dinv,
                 double omega, double* y, void* stream) {
    if (mode < 0 || mode > 3 || (mode >= 1 && !b) || (mode == 3 && (!dinv || x == y))) {
        npb_set_error("spmv_dia: bad mode / operands");
        return NPB_ERR_ARG;
    }
    return npb_spmv_dia_launch(nrows, ncols, ndiag, ld, offsets, vals, x, mode, b, dinv, omega, y, ST(stream), nullptr, nullptr);
}

// Jacobi sweep y = x + omega dinv (b - A x) from the diagonal-storage image that also returns
// *dot_out = b . y (device); red: npb_reduce_scratch_bytes() bytes, zeroed once
int npb_spmv_dia_jacobi_dot(int64_t nrows, int64_t ncols, int ndiag, const int32_t* offsets,
                            const float* vals, int64_t ld, const double* x, const double* b,
                            const double* dinv, double omega, double* y, double* dot_out,
                            void* red, void* stream) {
    if (!b || !dinv || x == y || !dot_out || !red) {
        npb_set_error("spmv_dia_jacobi_dot: bad operands");
        return NPB_ERR_ARG;
    }
    return npb_spmv_dia_launch(nrows, ncols, ndiag, ld, offsets, vals, x, 3, b, dinv, omega, y, ST(stream),
                               dot_out, red);
}

// direction update of a CG step on a square diagonal-storage operator L in one kernel:
// beta = *rz_new / *rz_old (first != 0: beta = 0), p <- z + beta p, q <- L z + beta q, *pq = p . q
int npb_dia_cg_dir(int64_t nrows, int ndiag, const int32_t* offsets, const float* vals, int64_t ld,
                   const double* z, const double* rz_new, const double* rz_old, int first,
                   double* p, double* q, double* pq, void* red, void* stream) {
    return npb_dia_cg_dir_launch(nrows, ndiag, ld, offsets, vals, z, rz_new, rz_old, first, p, q, pq, red,
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

// out0[j] = V_j . w0 and out1[j] = V_j . w1 for j < k in one sweep over V
int npb_dot_multi2(int64_t n, int k, const double* V, int64_t ld, const double* w0,
                   const double* w1, double* out0, double* out1, void* scratch, void* stream) {
    return npb_dot_multi2_launch(n, k, V, ld, w0, w1, out0, out1, scratch, ST(stream));
}

// u <- (u - sum_j c_j V_j) * inv_rho ; w <- w - sum_j a_j V_j - b * u_old ; *out_norm2 = ||w||^2
// coef = [c (k), a (k), inv_rho, b] on the device
int npb_axpy_multi2(int64_t n, int k, const double* V, int64_t ld, const double* coef,
                    double* u, double* w, double* out_norm2, void* scratch, void* stream) {
    return npb_axpy_multi2_launch(n, k, V, ld, coef, u, w, out_norm2, scratch, ST(stream));
}

int npb_axpby(int64_t n, double a, const double* x, double b, double* y, void* stream) {
    return npb_axpby_launch(n, a, x, b, y, ST(stream));
}

}  // extern "C"
