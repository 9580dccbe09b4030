#include "scan.cuh"

namespace {

__device__ __forceinline__ int64_t warp_incl_scan(int64_t v, int lane) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        int64_t o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane >= d) v += o;
    }
    return v;
}

// inclusive scan of one value per thread over a 256-thread block; returns the
// inclusive prefix, *block_total gets the sum over the block
__device__ __forceinline__ int64_t block_incl_scan(int64_t v, int64_t* smem,
                                                   int64_t* block_total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int64_t incl = warp_incl_scan(v, lane);
    if (lane == 31) smem[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int64_t w = (lane < NPB_SCAN_THREADS / 32) ? smem[lane] : 0;
        w = warp_incl_scan(w, lane);
        if (lane < NPB_SCAN_THREADS / 32) smem[lane] = w;
    }
    __syncthreads();
    int64_t off = warp ? smem[warp - 1] : 0;
    *block_total = smem[NPB_SCAN_THREADS / 32 - 1];
    __syncthreads();
    return incl + off;
}

__global__ void __launch_bounds__(NPB_SCAN_THREADS)
scan_tile_sums(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ tile_sum) {
    __shared__ int64_t smem[NPB_SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * NPB_SCAN_TILE;
    int64_t acc = 0;
#pragma unroll 4
    for (int r = 0; r < NPB_SCAN_ROUNDS; ++r) {
        int64_t i = base + (int64_t)r * NPB_SCAN_THREADS + threadIdx.x;
        if (i < n) acc += in[i];
    }
    int64_t tot;
    block_incl_scan(acc, smem, &tot);
    if (threadIdx.x == 0) tile_sum[blockIdx.x] = tot;
}

// one block: exclusive scan of the tile sums in place, total -> tile_sum[nt]
// and *total64
__global__ void __launch_bounds__(NPB_SCAN_THREADS)
scan_tile_offsets(int64_t* __restrict__ tile_sum, int64_t nt, int64_t* __restrict__ total64) {
    __shared__ int64_t smem[NPB_SCAN_THREADS / 32];
    int64_t carry = 0;
    for (int64_t b = 0; b < nt; b += NPB_SCAN_THREADS) {
        int64_t i = b + threadIdx.x;
        int64_t v = (i < nt) ? tile_sum[i] : 0;
        int64_t tot;
        int64_t incl = block_incl_scan(v, smem, &tot);
        if (i < nt) tile_sum[i] = carry + incl - v;
        carry += tot;
    }
    if (threadIdx.x == 0) {
        tile_sum[nt] = carry;
        if (total64) *total64 = carry;
    }
}

__global__ void __launch_bounds__(NPB_SCAN_THREADS)
scan_tiles(const int32_t* __restrict__ in, int32_t* __restrict__ out, int64_t n,
           const int64_t* __restrict__ tile_off, int64_t nt) {
    __shared__ int64_t smem[NPB_SCAN_THREADS / 32];
    const int64_t base = (int64_t)blockIdx.x * NPB_SCAN_TILE;
    int64_t carry = tile_off[blockIdx.x];
    for (int r = 0; r < NPB_SCAN_ROUNDS; ++r) {
        int64_t i = base + (int64_t)r * NPB_SCAN_THREADS + threadIdx.x;
        int64_t v = (i < n) ? in[i] : 0;
        int64_t tot;
        int64_t incl = block_incl_scan(v, smem, &tot);
        if (i < n) out[i] = (int32_t)(carry + incl - v);
        carry += tot;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        int64_t t = tile_off[nt];
        out[n] = (int32_t)(t > 2147483647LL ? 2147483647LL : t);
    }
}

}  // namespace

int npb_scan_launch(const int32_t* in, int32_t* out, int64_t n, int64_t* ws,
                    int64_t* total64, cudaStream_t st) {
    if (n < 0) return NPB_ERR_ARG;
    int64_t nt = (n + NPB_SCAN_TILE - 1) / NPB_SCAN_TILE;
    if (nt == 0) nt = 1;
    scan_tile_sums<<<(unsigned)nt, NPB_SCAN_THREADS, 0, st>>>(in, n, ws);
    NPB_COUNT_LAUNCH();
    scan_tile_offsets<<<1, NPB_SCAN_THREADS, 0, st>>>(ws, nt, total64);
    NPB_COUNT_LAUNCH();
    scan_tiles<<<(unsigned)nt, NPB_SCAN_THREADS, 0, st>>>(in, out, n, ws, nt);
    NPB_COUNT_LAUNCH();
    NPB_CHECK_LAUNCH("scan");
    return NPB_OK;
}
