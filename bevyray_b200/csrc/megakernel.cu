// megakernel.cu — one kernel for the whole `fragment` entry point (assets/shaders/raytrace.wgsl:93-123):
// per-pixel seed, sample loop, bounce loop, traversal, scatter, accumulation and the fused depth composite.

#include "kernels.cuh"

namespace bvr {

namespace {

constexpr int TILE_W = 8;    // a warp covers an 8x4 pixel tile
constexpr int TILE_H = 16;   // 4 warps per block

template <bool REFERENCE_ORDER>
__device__ __forceinline__ Hit raycast(const SceneView& s, const Ray& ray) {
    if (REFERENCE_ORDER) {
        return raycast_reference_order(s, ray);
    } else {
        uint32_t sref[BVR_FAST_STACK];
        float sdst[BVR_FAST_STACK];
        return raycast_near_first(
            s, ray, [&](int i) -> uint32_t& { return sref[i]; }, [&](int i) -> float& { return sdst[i]; });
    }
}

template <bool REFERENCE_ORDER>
__global__ void __launch_bounds__(TILE_W* TILE_H) megakernel_v1(const RenderParams p) {
    const CameraParams& cam = p.cam;
    const uint32_t lx = blockIdx.x * TILE_W + threadIdx.x;
    const uint32_t ly = blockIdx.y * TILE_H + threadIdx.y;
    const uint32_t gy = shard_global_row(p.shard, ly);
    const bool active = lx < cam.width && ly < p.shard.rows && gy < cam.height;
    unsigned long long rays = 0;

    if (active) {
        const float u = pixel_u(cam, lx), v = pixel_v(cam, gy);
        uint32_t rng = pixel_seed(cam, u, v);

        // trace_multisampled, raytrace.wgsl:159-172
        V3 total = v3(0.0f, 0.0f, 0.0f);
        float total_depth = 0.0f;
        uint32_t primary_id = 0xffffffffu;
        float primary_t = BVR_INF;
        for (uint32_t sidx = 0; sidx < cam.sample_count; sidx++) {
            Ray ray = random_ray_from_uv(cam, u, v, rng);
            // raytrace, raytrace.wgsl:174-224
            float first_depth = BVR_INF;
            V3 ray_color = v3(1.0f, 1.0f, 1.0f);
            V3 light = v3(0.0f, 0.0f, 0.0f);
            uint32_t bounce = 0;
            for (; bounce <= cam.bounce_count; bounce++) {
                const Hit hit = raycast<REFERENCE_ORDER>(p.scene, ray);
                rays++;
                if (bounce == 0) {
                    first_depth = hit.t;
                    if (sidx == 0) { primary_id = hit.t == BVR_INF ? 0xffffffffu : hit.model; primary_t = hit.t; }
                }
                if (hit.t == BVR_INF) { light = background_gradient(ray); break; }
                V3 attenuation;
                const bool absorbed = scatter(p.scene, ray, hit, rng, attenuation);
                if (absorbed) break;
                ray_color = vmul(ray_color, attenuation);
            }
            if (bounce == cam.bounce_count + 1u) ray_color = v3(0.0f, 0.0f, 0.0f);
            if (first_depth == BVR_INF) first_depth = cam.fallback_far;
            const V3 lin = vmul(ray_color, light);
            total = vadd(total, v3(fsqrt(lin.x), fsqrt(lin.y), fsqrt(lin.z)));   // linear_to_gamma per sample
            total_depth = fadd(total_depth, first_depth);
        }
        const float n = (float)cam.sample_count;
        float4 out = make_float4(fdiv(total.x, n), fdiv(total.y, n), fdiv(total.z, n), 1.0f);
        const float depth_avg = fdiv(total_depth, n);

        // fused composite, raytrace.wgsl:104-120
        const size_t gpix = (size_t)gy * cam.width + lx;
        if (cam.level == 1u || cam.level == 2u) {
            if (raster_wins(cam, p.raster_depth[gpix], depth_avg)) out = p.raster_rgba[gpix];
        }
        const size_t lpix = (size_t)ly * cam.width + lx;
        if (p.out_rgba) p.out_rgba[lpix] = out;
        if (p.out_rt_depth) p.out_rt_depth[lpix] = depth_avg;
        if (p.out_primary_id) p.out_primary_id[lpix] = primary_id;
        if (p.out_primary_depth) p.out_primary_depth[lpix] = primary_t;
        if (p.out_srgb8) p.out_srgb8[lpix] = store_srgb8(out);
    }

    // one atomic per warp for the ray counter
    const unsigned mask = 0xffffffffu;
    unsigned lo = (unsigned)(rays & 0xffffffffull);
    // per-thread counts are far below 2^32 / 32 for any sane spp*bounces; reduce in 64 bits anyway
    unsigned long long sum = rays;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(mask, sum, o);
    (void)lo;
    if (((threadIdx.y * TILE_W + threadIdx.x) & 31) == 0 && p.ray_counter && sum) atomicAdd(p.ray_counter, sum);
}

// level 0 (Skip): the fragment returns the raster texel, raytrace.wgsl:97-99
__global__ void copy_raster_kernel(const RenderParams p) {
    const uint32_t lx = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t ly = blockIdx.y * blockDim.y + threadIdx.y;
    const uint32_t gy = shard_global_row(p.shard, ly);
    if (lx >= p.cam.width || ly >= p.shard.rows || gy >= p.cam.height) return;
    const size_t gpix = (size_t)gy * p.cam.width + lx, lpix = (size_t)ly * p.cam.width + lx;
    const float4 c = p.raster_rgba[gpix];
    if (p.out_rgba) p.out_rgba[lpix] = c;
    if (p.out_rt_depth) p.out_rt_depth[lpix] = 0.0f;
    if (p.out_primary_id) p.out_primary_id[lpix] = 0xffffffffu;
    if (p.out_primary_depth) p.out_primary_depth[lpix] = BVR_INF;
    if (p.out_srgb8) p.out_srgb8[lpix] = store_srgb8(c);
}

}  // namespace

int launch_megakernel(const RenderParams& p, cudaStream_t stream) {
    dim3 block(TILE_W, TILE_H);
    dim3 grid((p.cam.width + TILE_W - 1) / TILE_W, (p.shard.rows + TILE_H - 1) / TILE_H);
    if (grid.x == 0 || grid.y == 0) return 0;
    if (p.reference_order) megakernel_v1<true><<<grid, block, 0, stream>>>(p);
    else megakernel_v1<false><<<grid, block, 0, stream>>>(p);
    return 1;
}

int launch_copy_raster(const RenderParams& p, cudaStream_t stream) {
    dim3 block(32, 8);
    dim3 grid((p.cam.width + 31) / 32, (p.shard.rows + 7) / 8);
    if (grid.x == 0 || grid.y == 0) return 0;
    copy_raster_kernel<<<grid, block, 0, stream>>>(p);
    return 1;
}

}  // namespace bvr
