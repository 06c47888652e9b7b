// tile_order.cu — the order in which the megakernel's pixel queue hands out its 8x4 tiles.
//
// One pixel is one sequential chain of work (a single RNG stream runs through all its samples, raytrace.wgsl:89,161-167),
// and pixels differ by an order of magnitude in cost (RTIOW at 100 spp: 100 rays for a sky pixel, ~1100 for the heaviest).
// A frame therefore ends with a tail in which most lanes have run out of pixels while a few finish heavy ones.  Handing
// out the heaviest tiles FIRST (longest-processing-time-first list scheduling) keeps that tail short.  "Heaviest" is
// judged by what each tile cost in the previous frame of the same size: the render kernel adds every pixel's ray count
// to its tile's counter, and this file turns the counters into the next frame's order with one radix sort.  Ordering
// cannot change an image — pixels are independent — only when each one is computed.

#include <cub/device/device_radix_sort.cuh>

#include "kernels.cuh"

namespace bvr {

namespace {

__global__ void iota_kernel(uint32_t* __restrict__ v, uint32_t n, bool reversed) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = reversed ? n - 1u - i : i;
}

size_t align256(size_t x) { return (x + 255u) & ~(size_t)255u; }

size_t cub_temp_bytes(uint32_t n) {
    size_t bytes = 0;
    cub::DeviceRadixSort::SortPairsDescending(nullptr, bytes, (const uint32_t*)nullptr, (uint32_t*)nullptr,
                                              (const uint32_t*)nullptr, (uint32_t*)nullptr, (int)n);
    return bytes;
}

}  // namespace

// scratch = [iota n][sorted keys n][cub temp]
size_t tile_order_scratch_bytes(uint32_t n_tiles) {
    return 2u * align256((size_t)n_tiles * sizeof(uint32_t)) + align256(cub_temp_bytes(n_tiles));
}

int launch_tile_order_update(uint32_t* tile_cost, uint32_t* tile_order, void* scratch, uint32_t n_tiles, int mode, bool first,
                             cudaStream_t stream) {
    if (n_tiles == 0) return 0;
    int launches = 0;
    char* base = static_cast<char*>(scratch);
    uint32_t* iota = reinterpret_cast<uint32_t*>(base);
    uint32_t* keys_out = reinterpret_cast<uint32_t*>(base + align256((size_t)n_tiles * sizeof(uint32_t)));
    void* temp = base + 2u * align256((size_t)n_tiles * sizeof(uint32_t));
    if (mode == 1) {
        if (first) { iota_kernel<<<(n_tiles + 255) / 256, 256, 0, stream>>>(tile_order, n_tiles, true); launches++; }
    } else {
        if (first) { iota_kernel<<<(n_tiles + 255) / 256, 256, 0, stream>>>(iota, n_tiles, false); launches++; }
        size_t temp_bytes = cub_temp_bytes(n_tiles);
        // stable sort: tiles of equal cost keep their row-major order, so the order is a pure function of the counters.
        // Only bits 4..19 of a counter take part (two 8-bit passes instead of four): 16 rays are noise, and a tile of more
        // than 2^20 rays merely sorts as if it had fewer.
        const int lo_bit = 4, hi_bit = 20;
        if (mode == 3)
            cub::DeviceRadixSort::SortPairs(temp, temp_bytes, (const uint32_t*)tile_cost, keys_out, (const uint32_t*)iota,
                                            tile_order, (int)n_tiles, lo_bit, hi_bit, stream);
        else
            cub::DeviceRadixSort::SortPairsDescending(temp, temp_bytes, (const uint32_t*)tile_cost, keys_out,
                                                      (const uint32_t*)iota, tile_order, (int)n_tiles, lo_bit, hi_bit, stream);
        // (cub's kernels are library kernels: not counted among the launches the library reports as its own)
    }
    cudaMemsetAsync(tile_cost, 0, (size_t)n_tiles * sizeof(uint32_t), stream);
    return launches;
}

}  // namespace bvr
