// Persistent TMA-pipelined CSR SpMV (included by krylov.cu after spmv_epi).
//
// Same row decomposition as the "stream" kernels (a unit = ROWS consecutive rows
// = one contiguous range of (index, value) pairs), but the streaming half of the
// traffic never touches a register: one producer thread moves each unit's range
// into a 3-stage shared-memory ring with 1-D bulk copies (cp.async.bulk,
// completion counted on an mbarrier, L2 evict-first so that the streamed matrix
// does not push x out of L2), while 256 consumer threads gather x and sum their
// rows straight from the ring.  CTAs are persistent (2 per SM, unit u -> CTA
// u mod grid): up to ~72 KB of matrix per CTA are always in flight and no load of
// the stream waits on a row pointer -- the pointers of the NEXT unit are fetched
// into registers while the current unit is summed.  Ranges longer than a stage
// are cut into chunks; producer and consumers derive the same chunk sequence
// from (k0, k1).  Consumer warps never synchronise with each other: a stage is
// handed back when its 8 warps have arrived on the `empty` barrier.
//
// Algorithmic bytes per call: 12*nnz + 4*(nrows+1) + 16*nrows.
#pragma once

constexpr int TM_CONSUMERS = 256;
constexpr int TM_THREADS = TM_CONSUMERS + 32;
constexpr int TM_CAP_MAX = 3072;

// ring geometry: entries per stage, stages, resident CTAs per SM.  Shared memory the
// ring does not take stays L1 for the gathers of x, so the ring is no larger than the
// bandwidth-delay product needs.
template <int CAP_, int STAGES_, int CTAS_>
struct tm_geom {
    static constexpr int CAP = CAP_, STAGES = STAGES_, CTAS = CTAS_;
};

template <class G, class T>
struct tm_smem {
    alignas(128) T val[G::STAGES][G::CAP];
    alignas(128) int32_t idx[G::STAGES][G::CAP];
    alignas(8) unsigned long long full[G::STAGES];
    unsigned long long empty[G::STAGES];
};

__device__ __forceinline__ unsigned smem_u32(const void* p) {
    return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(b)),
                 "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    unsigned done;
    unsigned spins = 0;
    do {
        asm volatile(
            "{\n .reg .pred p;\n"
            " mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            " selp.u32 %0, 1, 0, p;\n}\n"
            : "=r"(done) : "r"(smem_u32(b)), "r"(parity) : "memory");
        if (!done && ++spins > (1u << 26)) __trap();      // a lost arrival must not hang the GPU
    } while (!done);
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes,
                                         unsigned long long* bar, unsigned long long policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
        " [%0], [%1], %2, [%3], %4;\n"
        ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}

// T = value type of the stored matrix: double, or float for the operators of the multigrid
// preconditioner (8 instead of 12 bytes per entry; x, y and the sums stay fp64)
template <int ROWS, int RPT, class G, bool HALO, class T>
__global__ void __launch_bounds__(TM_THREADS, G::CTAS)
k_spmv_tma(int64_t nrows, int64_t nnz, int64_t nunits, const int32_t* __restrict__ ptr,
           const int32_t* __restrict__ idx, const T* __restrict__ dat,
           const double* x, spmv_epi epi, double* y) {
    constexpr int TPR = TM_CONSUMERS / ROWS;
    constexpr int UROWS = ROWS * RPT;            // rows per unit
    constexpr int GB = 12 / RPT;                 // gathers in flight per row and batch
    static_assert(RPT == 1 || TPR == 1, "several rows per thread only with one thread per row");
    constexpr int TM_CAP = G::CAP, TM_STAGES = G::STAGES;
    extern __shared__ __align__(128) unsigned char tm_raw[];
    tm_smem<G, T>& S = *reinterpret_cast<tm_smem<G, T>*>(
        (reinterpret_cast<uintptr_t>(tm_raw) + 127) & ~(uintptr_t)127);
    const int tid = threadIdx.x;
    if (HALO) npb_halo_begin(epi.wait, epi.xh);
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < TM_STAGES; ++s) {
            mbar_init(&S.full[s], 1);
            mbar_init(&S.empty[s], TM_CONSUMERS / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
    }
    __syncthreads();
    int stage = 0;
    unsigned phase = 0;
    if (tid >= TM_CONSUMERS) {
        // ---------------------------------------------------------- producer
        if (tid != TM_CONSUMERS) return;
        unsigned long long policy;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(policy));
        const int32_t nnz4 = (int32_t)(nnz & ~(int64_t)3);
        int64_t u = blockIdx.x;
        int32_t k0n = 0, k1n = 0;
        if (u < nunits) {
            const int64_t r0 = u * UROWS;
            k0n = ptr[r0];
            k1n = ptr[min(r0 + UROWS, nrows)];
        }
        for (; u < nunits; u += gridDim.x) {
            const int32_t k0 = k0n, k1 = k1n;
            const int64_t un = u + gridDim.x;
            if (un < nunits) {
                const int64_t r0 = un * UROWS;
                k0n = ptr[r0];
                k1n = ptr[min(r0 + UROWS, nrows)];
            }
            int32_t lo = k0 & ~3;
            do {
                const int32_t hi = min(lo + TM_CAP, k1);
                mbar_wait(&S.empty[stage], phase ^ 1u);
                // whole 16-byte groups that lie inside the arrays go by bulk copy,
                // the last 1..3 entries of the arrays (if any) by plain stores
                int32_t ne = (hi - lo + 3) & ~3;
                if (ne < 0) ne = 0;
                if (lo + ne > nnz4) ne = max(nnz4 - lo, 0);
                for (int32_t k = lo + ne; k < hi; ++k) {
                    S.idx[stage][k - lo] = idx[k];
                    S.val[stage][k - lo] = dat[k];
                }
                if (ne > 0) {
                    mbar_arrive_tx(&S.full[stage], (unsigned)ne * (4u + (unsigned)sizeof(T)));
                    bulk_g2s(&S.idx[stage][0], idx + lo, (unsigned)ne * 4u, &S.full[stage], policy);
                    bulk_g2s(&S.val[stage][0], dat + lo, (unsigned)ne * (unsigned)sizeof(T), &S.full[stage], policy);
                } else {
                    mbar_arrive(&S.full[stage]);
                }
                if (++stage == TM_STAGES) { stage = 0; phase ^= 1u; }
                lo += TM_CAP;
            } while (lo < k1);
        }
        return;
    }
    // ------------------------------------------------------------- consumers
    // thread (row, sub) sums the entries sub, sub + TPR, ... of the rows
    // row, row + ROWS, ... (RPT of them) of each unit
    const int row = tid / TPR, sub = tid % TPR;
    int64_t u = blockIdx.x;
    int32_t rs_n[RPT], re_n[RPT], k0_n = 0, k1_n = 0;
#pragma unroll
    for (int j = 0; j < RPT; ++j) rs_n[j] = re_n[j] = 0;
    if (u < nunits) {
        const int64_t r0 = u * UROWS;
        const int nr = (int)min((int64_t)UROWS, nrows - r0);
        k0_n = ptr[r0];
        k1_n = ptr[r0 + nr];
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            const int rr = row + j * ROWS;
            rs_n[j] = (rr < nr) ? ptr[r0 + rr] : k1_n;
            re_n[j] = (rr < nr) ? ptr[r0 + rr + 1] : k1_n;
        }
    }
    for (; u < nunits; u += gridDim.x) {
        int32_t rs[RPT], re[RPT];
#pragma unroll
        for (int j = 0; j < RPT; ++j) { rs[j] = rs_n[j]; re[j] = re_n[j]; }
        const int32_t k0 = k0_n, k1 = k1_n;
        const int64_t r0 = u * UROWS;
        const int nr = (int)min((int64_t)UROWS, nrows - r0);
        const int64_t un = u + gridDim.x;
        if (un < nunits) {
            const int64_t q0 = un * UROWS;
            const int qn = (int)min((int64_t)UROWS, nrows - q0);
            k0_n = ptr[q0];
            k1_n = ptr[q0 + qn];
#pragma unroll
            for (int j = 0; j < RPT; ++j) {
                const int rr = row + j * ROWS;
                rs_n[j] = (rr < qn) ? ptr[q0 + rr] : k1_n;
                re_n[j] = (rr < qn) ? ptr[q0 + rr + 1] : k1_n;
            }
        }
        // operands of the epilogue, requested before the stream is waited for
        // (one row per thread only: with RPT > 1 they are read at the end)
        const bool writer0 = row < nr && sub == 0;
        double bv = 0.0, dv = 0.0, xv = 0.0;
        if (RPT == 1) {
            if (writer0 && epi.mode >= 1) bv = epi.b[r0 + row];
            if (writer0 && epi.mode == 3) { dv = epi.dinv[r0 + row]; xv = x[r0 + row]; }
        }
        double acc[RPT];
#pragma unroll
        for (int j = 0; j < RPT; ++j) acc[j] = 0.0;
        int32_t lo = k0 & ~3;
        do {
            const int32_t hi = min(lo + TM_CAP, k1);
            mbar_wait(&S.full[stage], phase);
            const int32_t* si = &S.idx[stage][0];
            const T* sv = &S.val[stage][0];
            // gathers of all RPT rows are issued GB at a time before any of them is
            // used: one memory latency per batch instead of one per row / per 4 entries
            int32_t kk[RPT], ee[RPT];
            bool more = false;
#pragma unroll
            for (int j = 0; j < RPT; ++j) {
                ee[j] = min(re[j], hi) - lo;
                kk[j] = max(rs[j], lo) - lo + sub;
                more |= kk[j] < ee[j];
            }
            while (more) {
                double xv[RPT][GB];
#pragma unroll
                for (int j = 0; j < RPT; ++j)
#pragma unroll
                    for (int g = 0; g < GB; ++g) {
                        const int32_t k = kk[j] + g * TPR;
                        xv[j][g] = (k < ee[j]) ? ldx<HALO>(x, epi, si[k]) : 0.0;
                    }
                more = false;
#pragma unroll
                for (int j = 0; j < RPT; ++j) {
                    double a_ = acc[j];
#pragma unroll
                    for (int g = 0; g < GB; ++g) {
                        const int32_t k = kk[j] + g * TPR;
                        if (k < ee[j]) a_ = fma((double)sv[k], xv[j][g], a_);
                    }
                    acc[j] = a_;
                    kk[j] += GB * TPR;
                    more |= kk[j] < ee[j];
                }
            }
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(&S.empty[stage]);
            if (++stage == TM_STAGES) { stage = 0; phase ^= 1u; }
            lo += TM_CAP;
        } while (lo < k1);
#pragma unroll
        for (int j = 0; j < RPT; ++j) {
            double a_ = acc[j];
            if (TPR > 1) {
#pragma unroll
                for (int d = TPR >> 1; d; d >>= 1) a_ += __shfl_xor_sync(0xffffffffu, a_, d);
            }
            const int rr = row + j * ROWS;
            if (rr < nr && sub == 0) {
                const int64_t r = r0 + rr;
                if (RPT > 1) {
                    if (epi.mode >= 1) bv = epi.b[r];
                    if (epi.mode == 3) { dv = epi.dinv[r]; xv = x[r]; }
                }
                double out;
                switch (epi.mode) {
                    case 0: out = a_; break;
                    case 1: out = bv - a_; break;
                    case 2: out = bv + a_; break;
                    default: out = xv + epi.omega * dv * (bv - a_); break;
                }
                y[r] = out;
            }
        }
    }
}

// unit shape for a mean row length: ROWS rows are summed side by side (TPR = 256 / ROWS
// threads each), RPT such groups make a unit; picked so that a mean unit fills most of a stage
struct tma_shape { int rows, rpt; };
static inline tma_shape tma_shape_for(double mean, int cap) {
    const double f = cap / 3072.0;
    if (mean <= 2.9 * f) return {256, 4};
    if (mean <= 5.8 * f) return {256, 2};
    if (mean <= 11.5 * f) return {256, 1};
    if (mean <= 23.0 * f) return {128, 1};
    if (mean <= 46.0 * f) return {64, 1};
    return {32, 1};
}

// ring geometry: -1 = by mean row length (measured on the operators of a 2048^2 cavity
// step, tools/spmv_bench.py: very short and long rows run best on 3 CTAs x 2 stages x
// 2304 entries -- more consumer warps and more L1 for the gathers -- the 5..20 entry
// stencil rows on 2 CTAs x 3 stages x 3072), 0..3 = forced (tuning)
static int npb_tma_geometry = -1;
static inline int tma_geometry_for(double mean) {
    if (npb_tma_geometry >= 0) return npb_tma_geometry;
    return (mean <= 4.5 || mean >= 20.0) ? 2 : 0;
}
static inline int tma_cap(int g) { return (g == 0 || g == 4) ? 3072 : (g == 3 ? 1536 : 2304); }
static inline int tma_ctas(int g) { return (g == 2 || g == 4) ? 3 : (g == 3 ? 4 : 2); }
// rows below which the persistent grid cannot be filled twice over
static inline int64_t tma_min_rows(double mean) {
    const int g = tma_geometry_for(mean);
    const tma_shape s = tma_shape_for(mean, tma_cap(g));
    return (int64_t)s.rows * s.rpt * NPB_NUM_SMS * tma_ctas(g) * 2;
}

template <int ROWS, int RPT, class G, bool HALO, class T>
static int launch_spmv_tma_h(int64_t nrows, int64_t nnz, const int32_t* ptr, const int32_t* idx,
                           const T* dat, const double* x, const spmv_epi& e, double* y,
                           cudaStream_t st) {
    static bool configured = false;
    const size_t sh = sizeof(tm_smem<G, T>) + 128;
    if (!configured) {
        NPB_CUDA(cudaFuncSetAttribute(k_spmv_tma<ROWS, RPT, G, HALO, T>,
                                      cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
        configured = true;
    }
    const int64_t nunits = (nrows + ROWS * RPT - 1) / (ROWS * RPT);
    const int64_t cap = (int64_t)NPB_NUM_SMS * G::CTAS;
    const int grid = (int)(nunits < cap ? nunits : cap);
    k_spmv_tma<ROWS, RPT, G, HALO, T><<<grid, TM_THREADS, sh, st>>>(nrows, nnz, nunits, ptr, idx, dat, x, e, y);
    return NPB_OK;
}

// the halo-aware instantiation only for row-sharded operators: the extra compare / select per
// gather costs the single-GPU kernel 6 % (measured)
template <int ROWS, int RPT, class G>
static int launch_spmv_tma(int64_t nrows, int64_t nnz, const int32_t* ptr, const int32_t* idx,
                           const double* dat, const double* x, const spmv_epi& e, double* y,
                           cudaStream_t st) {
    return e.xh ? launch_spmv_tma_h<ROWS, RPT, G, true, double>(nrows, nnz, ptr, idx, dat, x, e, y, st)
                : launch_spmv_tma_h<ROWS, RPT, G, false, double>(nrows, nnz, ptr, idx, dat, x, e, y, st);
}
// fp32-stored operators (multigrid preconditioner)
template <int ROWS, int RPT, class G>
static int launch_spmv_tma(int64_t nrows, int64_t nnz, const int32_t* ptr, const int32_t* idx,
                           const float* dat, const double* x, const spmv_epi& e, double* y,
                           cudaStream_t st) {
    return e.xh ? launch_spmv_tma_h<ROWS, RPT, G, true, float>(nrows, nnz, ptr, idx, dat, x, e, y, st)
                : launch_spmv_tma_h<ROWS, RPT, G, false, float>(nrows, nnz, ptr, idx, dat, x, e, y, st);
}

template <class G, class T>
static int spmv_tma_dispatch_g(int64_t nrows, int64_t nnz, const int32_t* ptr, const int32_t* idx,
                               const T* dat, const double* x, const spmv_epi& e, double* y,
                               cudaStream_t st) {
    const tma_shape s = tma_shape_for((double)nnz / (double)nrows, G::CAP);
    if (s.rpt == 4) return launch_spmv_tma<256, 4, G>(nrows, nnz, ptr, idx, dat, x, e, y, st);
    if (s.rpt == 2) return launch_spmv_tma<256, 2, G>(nrows, nnz, ptr, idx, dat, x, e, y, st);
    switch (s.rows) {
        case 256: return launch_spmv_tma<256, 1, G>(nrows, nnz, ptr, idx, dat, x, e, y, st);
        case 128: return launch_spmv_tma<128, 1, G>(nrows, nnz, ptr, idx, dat, x, e, y, st);
        case 64: return launch_spmv_tma<64, 1, G>(nrows, nnz, ptr, idx, dat, x, e, y, st);
        default: return launch_spmv_tma<32, 1, G>(nrows, nnz, ptr, idx, dat, x, e, y, st);
    }
}

// fp32-stored operators (74 KB instead of 111 KB per 3072 x 3 ring) leave room for a third
// resident CTA, i.e. more consumer warps with gathers in flight -- the kernel is bound by the
// gathers of x, not by the stream (J in fp64 and the momentum operator in fp32 both run at
// ~345 G gathers/s on 2 CTAs/SM).  Measured on the 2048^2 cavity step (tools/krylov_time.py):
// Krylov phase 615 -> 572 ms; 4 CTAs on smaller rings (2304 x 2 / x 3): 630 ms.
// 0 = as the fp64 operators (tuning switch npb_spmv_set_tma_f32_three)
static int npb_tma_f32_three = 1;

template <class T>
static int spmv_tma_dispatch(int64_t nrows, int64_t nnz, const int32_t* ptr, const int32_t* idx,
                             const T* dat, const double* x, const spmv_epi& e, double* y,
                             cudaStream_t st) {
    const int g = tma_geometry_for((double)nnz / (double)nrows);
    if constexpr (sizeof(T) == 4) {
        if (g == 0 && npb_tma_f32_three)
            return spmv_tma_dispatch_g<tm_geom<3072, 3, 3>, T>(nrows, nnz, ptr, idx, dat, x, e, y, st);
    }
    switch (g) {
        case 1: return spmv_tma_dispatch_g<tm_geom<2304, 3, 2>, T>(nrows, nnz, ptr, idx, dat, x, e, y, st);
        case 2: return spmv_tma_dispatch_g<tm_geom<2304, 2, 3>, T>(nrows, nnz, ptr, idx, dat, x, e, y, st);
        case 3: return spmv_tma_dispatch_g<tm_geom<1536, 3, 4>, T>(nrows, nnz, ptr, idx, dat, x, e, y, st);
        case 4: return spmv_tma_dispatch_g<tm_geom<3072, 2, 3>, T>(nrows, nnz, ptr, idx, dat, x, e, y, st);
        default: return spmv_tma_dispatch_g<tm_geom<3072, 3, 2>, T>(nrows, nnz, ptr, idx, dat, x, e, y, st);
    }
}

