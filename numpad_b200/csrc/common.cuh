// Shared definitions for the numpad_b200 CUDA kernels (sm_100a only).
//
// Conventions of the C-ABI (include/numpad_b200.h):
//   * every entry point returns int: 0 = ok, <0 = npb_status error code;
//   * all pointers are DEVICE pointers owned by the caller unless stated;
//   * `stream` is a cudaStream_t passed as void*; ordering is by stream only;
//   * CSR = (indptr int32[n_rows+1], indices int32[nnz], data f64[nnz]) with
//     column indices strictly increasing inside each row ("canonical").
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#define NPB_OK 0
#define NPB_ERR_CUDA (-1)
#define NPB_ERR_ARG (-2)
#define NPB_ERR_WORKSPACE (-3)
#define NPB_ERR_BREAKDOWN (-4)
#define NPB_ERR_NOCONV (-5)

#define NPB_MAX_SCALES 6

// A chain of pending scalar multipliers.  The reference multiplies a CSR
// matrix by Python numbers one at a time (adstate.py:290-297 -> scipy
// _mul_scalar, one rounding per factor); the chain is applied in the same
// order, with explicit non-fused multiplies, when a kernel consumes the data,
// so values round exactly as the reference's and a scalar product costs no
// memory traffic at all.
struct npb_scales {
    int n;
    double s[NPB_MAX_SCALES];
};

__device__ __forceinline__ double npb_apply(const npb_scales& sc, double v) {
#pragma unroll
    for (int i = 0; i < NPB_MAX_SCALES; ++i)
        if (i < sc.n) v = __dmul_rn(v, sc.s[i]);
    return v;
}

void npb_set_error(const char* fmt, ...);

#define NPB_CHECK_LAUNCH(name)                                                 \
    do {                                                                       \
        cudaError_t e_ = cudaGetLastError();                                   \
        if (e_ != cudaSuccess) {                                               \
            npb_set_error("%s: %s", name, cudaGetErrorString(e_));             \
            return NPB_ERR_CUDA;                                               \
        }                                                                      \
    } while (0)

#define NPB_CUDA(call)                                                         \
    do {                                                                       \
        cudaError_t e_ = (call);                                               \
        if (e_ != cudaSuccess) {                                               \
            npb_set_error("%s:%d %s", __FILE__, __LINE__,                      \
                          cudaGetErrorString(e_));                             \
            return NPB_ERR_CUDA;                                               \
        }                                                                      \
    } while (0)

static inline int npb_blocks(int64_t n, int per_block) {
    int64_t b = (n + per_block - 1) / per_block;
    return (int)(b < 1 ? 1 : b);
}

// number of SMs of the B200 this library targets; persistent grids are sized
// as multiples of it
#define NPB_NUM_SMS 148

extern "C" {
// bumped by every kernel launch issued through the C-ABI (bench.py reports it
// as gpu_launches)
extern unsigned long long npb_launch_counter;
// algorithmic bytes (every operand once) of the SpMV-shaped and vector
// kernels of the solve phase launched since the last reset: bench.py divides it by the
// Krylov phase time for the whole-phase roofline
extern double npb_bytes_counter;
}
#define NPB_COUNT_LAUNCH() (++npb_launch_counter)
#define NPB_COUNT_BYTES(b) (npb_bytes_counter += (double)(b))
