// device_utils.cuh -- block/grid-level building blocks: deterministic reductions, exclusive scan.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace arap {

constexpr int kBlock = 256;           // threads per CTA for the vertex/row-parallel kernels
constexpr int kWarpsPerBlock = kBlock / 32;

// Programmatic dependent launch (sm_90+): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may
// become resident while its predecessor is still running. pdl_wait() blocks until the predecessor grid has completed and
// its writes are visible -- every kernel of the iteration calls it before touching any global memory; pdl_trigger() lets
// the NEXT kernel's CTAs be scheduled as soon as all CTAs of this one have started (they then sit in their own
// pdl_wait()), which hides the launch latency of the ~20 small dependent kernels of a CG iteration. Both are no-ops
// for kernels launched without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_enter() { pdl_wait(); pdl_trigger(); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Sum N doubles per thread over the block. Result valid in thread 0. Fixed tree -> deterministic.
template <int N>
__device__ __forceinline__ void block_sum(double (&v)[N]) {
    __shared__ double smem[N][kWarpsPerBlock];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < N; ++k) {
        double s = warp_sum(v[k]);
        if (lane == 0) smem[k][warp] = s;
    }
    __syncthreads();
    if (warp == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) {
            double s = (lane < kWarpsPerBlock) ? smem[k][lane] : 0.0;
            s = warp_sum(s);
            v[k] = s;
        }
    }
    __syncthreads();
}

// Grid-wide deterministic sum of N doubles per thread using the "last block finishes" pattern.
// partials: gridDim.x * N doubles; counter: one unsigned, zero on entry, left zero on exit.
// Returns true in thread 0 of the LAST block to arrive, with total[] holding the grid sums
// (accumulated in a fixed order, independent of block scheduling).
template <int N>
__device__ __forceinline__ bool grid_sum_last_block(double (&v)[N], double *partials, unsigned *counter, double (&total)[N]) {
    __shared__ bool is_last;
    block_sum<N>(v);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) partials[(size_t)blockIdx.x * N + k] = v[k];
        __threadfence();
        const unsigned ticket = atomicAdd(counter, 1u);
        is_last = (ticket == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return false;
    __threadfence();
    double acc[N];
#pragma unroll
    for (int k = 0; k < N; ++k) acc[k] = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
        for (int k = 0; k < N; ++k) acc[k] += __ldcg(&partials[(size_t)b * N + k]);
    }
    block_sum<N>(acc);
    if (threadIdx.x == 0) {
#pragma unroll
        for (int k = 0; k < N; ++k) total[k] = acc[k];
        *counter = 0u;
        return true;
    }
    return false;
}

// ---- exclusive scan of int32 (three small kernels: tile sums, spine, apply) ---------------------
constexpr int kScanItems = 8;
constexpr int kScanTile = kBlock * kScanItems;

__device__ __forceinline__ int block_exclusive_scan_int(int val, int *block_total) {
    __shared__ int warp_sums[kWarpsPerBlock];
    __shared__ int total_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = val;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int ws = (lane < kWarpsPerBlock) ? warp_sums[lane] : 0;
        int wi = ws;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        if (lane < kWarpsPerBlock) warp_sums[lane] = wi - ws;
        if (lane == kWarpsPerBlock - 1) total_s = wi;
    }
    __syncthreads();
    const int result = warp_sums[warp] + incl - val;
    if (block_total) *block_total = total_s;
    __syncthreads();
    return result;
}

__global__ void __launch_bounds__(kBlock) scan_tile_sums(const int *__restrict__ in, int n, int *__restrict__ tile_sums) {
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    int s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) if (base + k < n) s += in[base + k];
    double v[1] = {(double)s};
    block_sum<1>(v);                      // exact: tile sums < 2^31 fit a double
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = (int)v[0];
}

// single CTA: exclusive scan of the tile sums in place; tile_sums[n_tiles] = grand total
__global__ void __launch_bounds__(kBlock) scan_spine(int *tile_sums, int n_tiles) {
    int carry = 0;
    for (int base = 0; base < n_tiles; base += kBlock) {
        const int i = base + threadIdx.x;
        const int v = (i < n_tiles) ? tile_sums[i] : 0;
        int total;
        const int ex = block_exclusive_scan_int(v, &total);
        if (i < n_tiles) tile_sums[i] = carry + ex;
        carry += total;
    }
    if (threadIdx.x == 0) tile_sums[n_tiles] = carry;
}

// out[i] = exclusive prefix of in[0..i); out[n] = total. in and out may alias only if out == in is NOT used (out has n+1 entries).
__global__ void __launch_bounds__(kBlock) scan_apply(const int *__restrict__ in, int n, const int *__restrict__ tile_offsets,
                                                     int n_tiles, int *__restrict__ out) {
    const int base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
    int item[kScanItems];
    int s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) { item[k] = (base + k < n) ? in[base + k] : 0; s += item[k]; }
    int ex = block_exclusive_scan_int(s, nullptr) + tile_offsets[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) {
        if (base + k < n) out[base + k] = ex;
        ex += item[k];
    }
    if (blockIdx.x == gridDim.x - 1 && threadIdx.x == 0) out[n] = tile_offsets[n_tiles];
}

}  // namespace arap
