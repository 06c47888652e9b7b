// tile_order.cu — the order in which the megakernel's pixel queue hands out its 8x4 tiles.
//
// One pixel is one sequential chain of work (a single RNG stream runs through all its samples, raytrace.wgsl:89,161-167),
// and pixels differ by an order of magnitude in cost (RTIOW at 100 spp: 100 rays for a sky pixel, ~1100 for the heaviest).
// A frame therefore ends with a tail in which most lanes have run out of pixels while a few finish heavy ones.  Handing
// out the heaviest tiles FIRST (longest-processing-time-first list scheduling) keeps that tail short.  "Heaviest" is
// judged by what each tile cost in the previous frame of the same size: the render kernel adds every pixel's ray count
// to its tile's counter, and the three small kernels below turn the counters into the next frame's order — a counting
// sort over 1024 logarithmic cost classes (5 bits of exponent, 5 bits of mantissa: 3 % resolution at any magnitude).
// Ordering cannot change an image — pixels are independent — only when each one is computed; inside one cost class the
// order is whatever the atomics make it.

#include "kernels.cuh"

namespace bvr {

namespace {

constexpr uint32_t N_CLASSES = 1024;

__device__ __forceinline__ uint32_t cost_class(uint32_t cost) {
    const uint32_t e = 31u - (uint32_t)__clz((int)(cost | 1u));        // position of the leading one, 0..31
    const uint32_t m = e >= 5u ? (cost >> (e - 5u)) & 31u : (cost << (5u - e)) & 31u;
    return (e << 5) | m;
}

__global__ void iota_kernel(uint32_t* __restrict__ v, uint32_t n, bool reversed) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = reversed ? n - 1u - i : i;
}

__global__ void class_count_kernel(const uint32_t* __restrict__ cost, uint32_t n, uint32_t* __restrict__ classes) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&classes[cost_class(cost[i])], 1u);
}

// classes[c] := number of tiles in the classes handed out BEFORE class c (heavier ones, or lighter ones when ascending)
__global__ void __launch_bounds__(N_CLASSES) class_scan_kernel(uint32_t* __restrict__ classes, bool ascending) {
    __shared__ uint32_t warp_sums[32];
    const uint32_t t = threadIdx.x, lane = t & 31u, warp = t >> 5;
    const uint32_t c = ascending ? t : N_CLASSES - 1u - t;             // thread t owns the t-th class in hand-out order
    const uint32_t v = classes[c];
    uint32_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, o);
        if (lane >= (uint32_t)o) x += y;
    }
    if (lane == 31u) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint32_t w = warp_sums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= (uint32_t)o) w += y;
        }
        warp_sums[lane] = w;
    }
    __syncthreads();
    classes[c] = (warp ? warp_sums[warp - 1] : 0u) + x - v;
}

__global__ void class_scatter_kernel(uint32_t* __restrict__ cost, uint32_t n, uint32_t* __restrict__ classes,
                                     uint32_t* __restrict__ order) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    order[atomicAdd(&classes[cost_class(cost[i])], 1u)] = i;
    cost[i] = 0u;                                                       // the counters start the next frame at zero
}

}  // namespace

size_t tile_order_scratch_bytes(uint32_t) { return N_CLASSES * sizeof(uint32_t); }

int launch_tile_order_update(uint32_t* tile_cost, uint32_t* tile_order, void* scratch, uint32_t n_tiles, int mode, bool first,
                             cudaStream_t stream) {
    if (n_tiles == 0) return 0;
    const uint32_t blocks = (n_tiles + 255u) / 256u;
    if (mode == 1) {   // experiment: reversed row-major
        if (first) iota_kernel<<<blocks, 256, 0, stream>>>(tile_order, n_tiles, true);
        cudaMemsetAsync(tile_cost, 0, (size_t)n_tiles * sizeof(uint32_t), stream);
        return first ? 1 : 0;
    }
    uint32_t* classes = static_cast<uint32_t*>(scratch);
    cudaMemsetAsync(classes, 0, N_CLASSES * sizeof(uint32_t), stream);
    class_count_kernel<<<blocks, 256, 0, stream>>>(tile_cost, n_tiles, classes);
    class_scan_kernel<<<1, N_CLASSES, 0, stream>>>(classes, mode == 3);   // 3 = experiment: lightest first
    class_scatter_kernel<<<blocks, 256, 0, stream>>>(tile_cost, n_tiles, classes, tile_order);
    return 3;
}

}  // namespace bvr
