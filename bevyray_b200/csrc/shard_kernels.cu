// shard_kernels.cu — the two device helpers the multi-GPU paths need around NCCL:
// de-interleaving gathered tile shards, and weighting per-seed partial averages.

#include "kernels.cuh"

namespace bvr {

namespace {

__global__ void axpby_kernel(float* __restrict__ dst, float dw, const float* __restrict__ src, float sw, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        dst[i] = __fadd_rn(__fmul_rn(dst[i], dw), __fmul_rn(src[i], sw));
}

// in-place scale (dst and src of axpby must not alias: both are __restrict__)
__global__ void scale_kernel(float* __restrict__ dst, float w, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] = __fmul_rn(dst[i], w);
}

// Sample sharding over peer memory: dst = ((slot_0 + slot_1) + slot_2) + ... in slot order, over the slots named in `mask`
// (a rank that got no samples leaves its slot untouched).  The first term is copied, not added to zero, and the order is
// fixed, so the sum is a pure function of the slots.
__global__ void sum_slots_kernel(const float4* __restrict__ slots, size_t slot_stride4, uint32_t n_slots, unsigned long long mask,
                                 float4* __restrict__ dst, size_t n4) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        bool first = true;
        for (uint32_t s = 0; s < n_slots; s++) {
            if (!((mask >> s) & 1ull)) continue;
            const float4 v = slots[s * slot_stride4 + i];
            if (first) { acc = v; first = false; }
            else acc = make_float4(__fadd_rn(acc.x, v.x), __fadd_rn(acc.y, v.y), __fadd_rn(acc.z, v.z), __fadd_rn(acc.w, v.w));
        }
        dst[i] = acc;
    }
}

// fragment's depth composite (raytrace.wgsl:104-120) as a pass of its own: used when the ray-traced colour and depth
// of one frame are the sum of several ranks' partial frames and can only be compared with the raster depth afterwards
__global__ void composite_kernel(float4* __restrict__ rgba, const float* __restrict__ rt_depth,
                                 const float4* __restrict__ raster_rgba, const float* __restrict__ raster_depth,
                                 CameraParams cam, size_t n) {
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        if (raster_wins(cam, raster_depth[i], rt_depth[i])) rgba[i] = raster_rgba[i];
}

// gathered: shard_count planes, each `shard_stride_words` words, shard-local rows of width*channels words
__global__ void unshard_kernel(const uint32_t* __restrict__ gathered, size_t shard_stride_words,
                               uint32_t* __restrict__ full, uint32_t width, uint32_t height, uint32_t channels,
                               uint32_t shard_count, uint32_t strip_rows) {
    const size_t row_words = (size_t)width * channels;
    const size_t total = row_words * height;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const uint32_t gy = (uint32_t)(i / row_words);
        const size_t within = i - (size_t)gy * row_words;
        const uint32_t strip = gy / strip_rows;
        const uint32_t shard = strip % shard_count;
        const uint32_t ly = (strip / shard_count) * strip_rows + (gy % strip_rows);
        full[i] = gathered[(size_t)shard * shard_stride_words + (size_t)ly * row_words + within];
    }
}

}  // namespace

int launch_scale(float* dst, float w, size_t n, cudaStream_t stream) {
    if (n == 0) return 0;
    scale_kernel<<<592, 256, 0, stream>>>(dst, w, n);
    return 1;
}

int launch_sum_slots(const float* slots, size_t slot_stride, uint32_t n_slots, unsigned long long mask, float* dst, size_t n,
                     cudaStream_t stream) {
    if (n == 0) return 0;
    sum_slots_kernel<<<1184, 256, 0, stream>>>(reinterpret_cast<const float4*>(slots), slot_stride / 4u, n_slots, mask,
                                               reinterpret_cast<float4*>(dst), n / 4u);
    return 1;
}

int launch_axpby(float* dst, float dw, const float* src, float sw, size_t n, cudaStream_t stream) {
    if (n == 0) return 0;
    const int block = 256;
    const int grid = (int)((n + block - 1) / block < 148 * 8 ? (n + block - 1) / block : 148 * 8);
    // a zero source weight (or src == dst) is a plain scale: src is not read, so an Inf there cannot turn into NaN
    if (sw == 0.0f || src == dst || src == nullptr) scale_kernel<<<grid, block, 0, stream>>>(dst, src == dst ? dw + sw : dw, n);
    else axpby_kernel<<<grid, block, 0, stream>>>(dst, dw, src, sw, n);
    return 1;
}

int launch_composite(float4* rgba, const float* rt_depth, const float4* raster_rgba, const float* raster_depth,
                     const CameraParams& cam, size_t n, cudaStream_t stream) {
    if (n == 0) return 0;
    const int block = 256;
    const int grid = (int)((n + block - 1) / block < 148 * 8 ? (n + block - 1) / block : 148 * 8);
    composite_kernel<<<grid, block, 0, stream>>>(rgba, rt_depth, raster_rgba, raster_depth, cam, n);
    return 1;
}

int launch_unshard(const uint32_t* gathered, size_t shard_stride_words, uint32_t* full, uint32_t width,
                   uint32_t height, uint32_t channels, uint32_t shard_count, uint32_t strip_rows,
                   cudaStream_t stream) {
    const size_t total = (size_t)width * height * channels;
    if (total == 0) return 0;
    const int block = 256;
    const int grid = (int)((total + block - 1) / block < 148 * 8 ? (total + block - 1) / block : 148 * 8);
    unshard_kernel<<<grid, block, 0, stream>>>(gathered, shard_stride_words, full, width, height, channels,
                                               shard_count, strip_rows);
    return 1;
}

}  // namespace bvr
