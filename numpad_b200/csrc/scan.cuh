// Exclusive prefix sum of int32 counts -> int32 offsets (+ int64 total), the
// symbolic half of every two-pass CSR kernel.  Reduce-then-scan in three
// launches; traffic is 2 reads + 1 write of 4 B per row, noise next to the
// >= 100 B per row of the numeric pass it serves.
#pragma once
#include "common.cuh"

#define NPB_SCAN_THREADS 256
#define NPB_SCAN_ROUNDS 16
#define NPB_SCAN_TILE (NPB_SCAN_THREADS * NPB_SCAN_ROUNDS)

// workspace (int64 elements) needed by npb_scan_launch for n inputs
static inline int64_t npb_scan_ws_elems(int64_t n) {
    return (n + NPB_SCAN_TILE - 1) / NPB_SCAN_TILE + 2;
}

// out[0..n] = exclusive scan of in[0..n-1]; out[n] = total (clamped into
// int32), *total64 = exact total (the host checks it against INT32_MAX).
// `in` and `out` must not overlap.
int npb_scan_launch(const int32_t* in, int32_t* out, int64_t n, int64_t* ws,
                    int64_t* total64, cudaStream_t st);
