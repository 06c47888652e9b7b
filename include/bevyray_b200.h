/*
 * bevyray_b200.h — C ABI of the B200-native path tracer that replaces bevyray's
 * WGSL fragment shader (the hot path) behind the reference's own data contract.
 *
 * Everything here is `extern "C"`, plain pointers and sizes.  File:line citations
 * are relative to the reference tree (GrandmasterB42/bevyray).
 *
 * What each entry point replaces in the reference:
 *
 *   bvr_create / bvr_destroy   RaytracingPipeline::from_world  (src/raytracing/pipeline.rs:233-331)
 *                              — pipeline + GPU buffer ownership, one per device.
 *   bvr_upload_scene           the three queue.write_buffer calls in RayTracingNode::run
 *                              (src/raytracing/pipeline.rs:132-139) fed by prepare_buffers
 *                              (src/raytracing/extract.rs:334-336).  Accepts the encase bytes of
 *                              Vec<Model>, Vec<RaytraceMaterial>, Vec<BVHNode> unchanged.
 *   bvr_render                 bind groups + render pass + draw(0..3, 0..1)
 *                              (src/raytracing/pipeline.rs:153-217) == one invocation of the
 *                              `fragment` entry point per pixel (assets/shaders/raytrace.wgsl:93-123).
 *   bvr_render_device          same, with caller-owned DEVICE buffers (for NCCL and for timing with
 *                              inputs resident in HBM).
 *
 * Threading: a context is Send but not Sync — wrap it in a Mutex exactly like the reference wraps
 * its StorageBuffers (src/raytracing/extract.rs:252-262).  No function throws or unwinds.
 */
#ifndef BEVYRAY_B200_H
#define BEVYRAY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BVR_ABI_VERSION 1u

/* ------------------------------------------------------------------------------------------- */
/* Status codes                                                                                 */
/* ------------------------------------------------------------------------------------------- */
typedef enum BvrStatus {
    BVR_OK = 0,
    BVR_ERR_INVALID_ARGUMENT = 1,   /* null pointer, zero size, inconsistent counts            */
    BVR_ERR_CUDA = 2,               /* a CUDA runtime call failed; see bvr_last_error          */
    BVR_ERR_UNSUPPORTED_PROJECTION = 3, /* camera.projection != 0: the reference does not even
                                       extract orthographic cameras (extract.rs:148)            */
    BVR_ERR_NO_SCENE = 4,           /* bvr_render before any bvr_upload_scene                   */
    BVR_ERR_BAD_SCENE = 5,          /* node/model indices out of range, leaf too large          */
    BVR_ERR_OUT_OF_MEMORY = 6,
    BVR_ERR_NO_DEVICE = 7           /* no CUDA device: there is NO CPU fallback                 */
} BvrStatus;

/* ------------------------------------------------------------------------------------------- */
/* The flattened storage-buffer / uniform layouts (encase std430 / std140 of the Rust structs)   */
/* ------------------------------------------------------------------------------------------- */

/* struct Model — src/raytracing/extract.rs:213-218, assets/shaders/raytrace.wgsl:57-61. 32 B. */
typedef struct BvrModel {
    float    position[3];   /* @0  */
    float    radius;        /* @12 */
    uint32_t material_id;   /* @16 */
    uint32_t _pad[3];       /* @20 */
} BvrModel;

/* struct RaytraceMaterial — extract.rs:181-189, raytrace.wgsl:64-77. 32 B. */
typedef struct BvrMaterial {
    float base_color[3];          /* @0  linear RGB (extract.rs:201) */
    float metallic;               /* @12 */
    float roughness;              /* @16 = StandardMaterial::perceptual_roughness (extract.rs:203) */
    float reflectance;            /* @20 unused by the shader */
    float ior;                    /* @24 */
    float specular_transmission;  /* @28 */
} BvrMaterial;

/* struct BVHNode — extract.rs:229-237, raytrace.wgsl:80-87. 48 B.
 * Leaf iff model_count > 0 (index = first model); otherwise index / index+1 are the children.
 * Node 0 is the root (raytrace.wgsl:316-322). */
typedef struct BvrBvhNode {
    float    bounds_min[3];  /* @0  */
    uint32_t _pad0;          /* @12 */
    float    bounds_max[3];  /* @16 */
    uint32_t index;          /* @28 */
    uint32_t model_count;    /* @32 */
    uint32_t _pad1[3];       /* @36 */
} BvrBvhNode;

/* struct CameraExtract — extract.rs:83-97, raytrace.wgsl:35-47. 80 B. */
typedef struct BvrCamera {
    uint32_t sample_count;  /* @0  */
    uint32_t bounce_count;  /* @4  */
    uint32_t projection;    /* @8  0 = perspective, anything else is an error */
    float    near_plane;    /* @12 */
    float    far_plane;     /* @16 */
    float    fov;           /* @20 vertical, radians */
    float    aspect;        /* @24 width / height */
    uint32_t _pad0;         /* @28 */
    float    position[3];   /* @32 */
    uint32_t _pad1;
    float    direction[3];  /* @48 unit forward */
    uint32_t _pad2;
    float    up[3];         /* @64 unit up */
    uint32_t _pad3;
} BvrCamera;

/* enum Raytracing — src/raytracing/mod.rs:94-101 (#[repr(u32)]). */
typedef enum BvrRaytracing {
    BVR_RAYTRACING_SKIP = 0,
    BVR_RAYTRACING_FALLBACK_RASTER = 1,
    BVR_RAYTRACING_FALLBACK_RAYTRACED = 2,
    BVR_RAYTRACING_PURE = 3
} BvrRaytracing;

/* struct RaytraceLevelExtract — extract.rs:100-104, raytrace.wgsl:30-33. 32 B. */
typedef struct BvrRaytraceLevel {
    uint32_t level;         /* @0  */
    uint32_t _pad0[3];
    float    _padding[3];   /* @16 */
    uint32_t _pad1;
} BvrRaytraceLevel;

/* struct WindowExtract — extract.rs:56-61, raytrace.wgsl:50-54. 16 B. */
typedef struct BvrWindow {
    float    random_seed;   /* @0 in [0,1); the reference draws a fresh one per frame (extract.rs:72-73) */
    uint32_t height;        /* @4 physical pixels */
    float    _padding[2];
} BvrWindow;

/* Half-open element range of one of the three scene arrays that changed since the last upload. */
typedef enum BvrSceneArray {
    BVR_ARRAY_MODELS = 0,
    BVR_ARRAY_MATERIALS = 1,
    BVR_ARRAY_BVH_NODES = 2
} BvrSceneArray;

typedef struct BvrDirtyRange {
    uint32_t array;   /* BvrSceneArray */
    uint32_t first;   /* first dirty element */
    uint32_t count;   /* number of dirty elements */
} BvrDirtyRange;

/* ------------------------------------------------------------------------------------------- */
/* Render options (things the reference has no knob for; zero-initialise for reference behaviour) */
/* ------------------------------------------------------------------------------------------- */
typedef enum BvrKernel {
    BVR_KERNEL_AUTO = 0,        /* library picks (megakernel) */
    BVR_KERNEL_MEGAKERNEL = 1,  /* persistent-thread megakernel */
    BVR_KERNEL_WAVEFRONT = 2,   /* raygen / extend / shade / compact pipeline, one launch per stage */
    BVR_KERNEL_CTA_WAVEFRONT = 3 /* the same stages inside one persistent kernel, one pool of paths per CTA */
} BvrKernel;

typedef enum BvrTraversal {
    BVR_TRAVERSAL_AUTO = 0,             /* near-first ordered traversal, same closest hit */
    BVR_TRAVERSAL_REFERENCE_ORDER = 1   /* raytrace.wgsl:313-346 verbatim, incl. the 32-entry
                                           stack truncation (raytrace.wgsl:320) */
} BvrTraversal;

typedef struct BvrRenderOptions {
    uint32_t width;        /* image width in pixels (the reference derives it as height*aspect only
                              for jitter maths, raytrace.wgsl:142; that derivation is kept) */
    uint32_t kernel;       /* BvrKernel */
    uint32_t traversal;    /* BvrTraversal */
    /* Tile sharding across GPUs: rows are grouped in strips of `strip_rows` rows; this context
     * renders strips s with s % shard_count == shard_index.  shard_count 0 or 1 = whole image. */
    uint32_t shard_index;
    uint32_t shard_count;
    uint32_t strip_rows;   /* 0 = default (8) */
    uint32_t flags;        /* BvrRenderFlags */
    float    output_weight; /* 0 (or 1) = none.  Otherwise every rgba texel (all four channels) and every rt_depth texel is
                              multiplied by it as it is stored: a rank's partial frame of a sample-sharded render leaves
                              the kernel already weighted by its share of the samples */
} BvrRenderOptions;

typedef enum BvrRenderFlags {
    /* Levels 1-2: skip the depth composite (raytrace.wgsl:104-120) but keep the level's fallback depth for misses
     * (raytrace.wgsl:177-182).  For sample sharding: the partial frames of all ranks are summed first (colour AND
     * rt_depth), then bvr_composite_device compares the summed depth with the raster depth once. */
    BVR_RENDER_DEFER_COMPOSITE = 1u,
    /* Uneven sample counts, for sharing S samples out over N ranks when N does not divide S: the pixels of the 8x4 tiles
     * (tx, ty) with ((tx + ty + PHASE) % MODULUS) < COUNT take ONE sample more than camera.sample_count, with
     * MODULUS = flags bits 8-15 (1..255), PHASE = bits 16-23, COUNT = bits 24-31.  Rank g of N renders S / N samples with
     * MODULUS = N, PHASE = g, COUNT = S % N: every pixel then gets exactly S samples over the ranks, and every rank the same
     * amount of work.  With this flag output_weight is the weight of ONE sample (1 / S): a pixel with n samples is stored
     * multiplied by n x output_weight.  Megakernel only (BVR_ERR_INVALID_ARGUMENT with the other kernels / traversals);
     * camera.sample_count must be at least 1. */
    BVR_RENDER_EXTRA_SAMPLE = 2u
} BvrRenderFlags;
#define BVR_RENDER_EXTRA_SAMPLE_BITS(modulus, phase, count) \
    (BVR_RENDER_EXTRA_SAMPLE | ((uint32_t)(modulus) << 8) | ((uint32_t)(phase) << 16) | ((uint32_t)(count) << 24))

/* Output planes.  Any pointer may be NULL (plane not produced / not copied).
 * For a sharded render each plane holds only this shard's rows, strips concatenated in order:
 * bvr_shard_rows() rows of `width` pixels. */
typedef struct BvrOutputs {
    float*    rgba;           /* 4 floats / pixel: the value `fragment` returns (raytrace.wgsl:93-123)
                                 before the Rgba8UnormSrgb store (pipeline.rs:311-315) */
    float*    rt_depth;       /* 1 float / pixel: RaytraceResult.depth, averaged over samples
                                 (raytrace.wgsl:170) */
    uint32_t* primary_id;     /* 1 u32 / pixel: model index hit by sample 0's camera ray, 0xFFFFFFFF = miss */
    float*    primary_depth;  /* 1 float / pixel: hit.distance of sample 0's camera ray (INF = 3.40282347e38 = miss) */
    uint8_t*  srgb8;          /* 4 bytes / pixel: rgba after the Rgba8UnormSrgb store conversion */
} BvrOutputs;

typedef struct BvrStats {
    uint64_t rays;            /* raycast() invocations (raytrace.wgsl:190) in the last render */
    uint64_t paths;           /* camera rays (pixels x samples) in the last render */
    uint64_t kernel_launches; /* CUDA kernels launched by the library since bvr_create */
    uint64_t h2d_bytes;       /* bytes copied host->device since bvr_create */
    uint64_t d2h_bytes;       /* bytes copied device->host since bvr_create */
    float    last_render_ms;  /* device time of the last render (CUDA events on the context stream) */
    float    last_upload_ms;
    /* BVR_SELFCHECK=1 (environment, read at bvr_create): about one ray in 1024 of the last render was traced again in the
     * kernel with the reference's own traversal order and boxes (raytrace.wgsl:313-346); a different closest hit counts
     * as a mismatch.  Both stay 0 when the check is off. */
    uint64_t selfcheck_rays;
    uint64_t selfcheck_mismatches;
} BvrStats;

typedef struct BvrContext BvrContext;

/* ------------------------------------------------------------------------------------------- */
/* Entry points                                                                                  */
/* ------------------------------------------------------------------------------------------- */

uint32_t    bvr_abi_version(void);
const char* bvr_status_string(int status);

/* Create a context on CUDA device `device`.  Fails with BVR_ERR_NO_DEVICE when there is no GPU. */
int  bvr_create(int device, BvrContext** out_ctx);
void bvr_destroy(BvrContext* ctx);
/* Message of the last failing call on this context ("" if none).  Valid until the next call. */
const char* bvr_last_error(const BvrContext* ctx);

/* Run all work of this context on `cuda_stream` (a cudaStream_t; NULL = the context's own stream). */
int bvr_set_stream(BvrContext* ctx, void* cuda_stream);
int bvr_sync(BvrContext* ctx);

/* Experiment knobs (BVR_NO_TIGHT, BVR_NO_BVH4, BVR_MK_THREADS, ...) are read from the environment once,
 * in bvr_create; this re-reads them for an existing context (A/B tests).  Production never calls it. */
int bvr_reload_tuning(BvrContext* ctx);

/* Upload the scene in the reference's layout.  `ranges == NULL` uploads everything; otherwise only
 * the listed element ranges are copied (pinned staging -> HBM) and the device-side traversal layout
 * is re-derived.  Counts must match the previous upload when ranges are given.
 * n_models == 0 is legal: every ray misses.
 * The caller owns its arrays again when the call returns: pageable arrays are staged, pinned arrays are copied from
 * directly and the call waits for those copies (the device-side re-layout still runs asynchronously). */
int bvr_upload_scene(BvrContext* ctx,
                     const BvrModel* models, size_t n_models,
                     const BvrMaterial* materials, size_t n_materials,
                     const BvrBvhNode* nodes, size_t n_nodes,
                     const BvrDirtyRange* ranges, size_t n_ranges);

/* Same as bvr_upload_scene, but the BVH is BUILT ON THE GPU from the models (PLOC over the Morton order, the
 * reference's own algorithm, emitting the BVHNode contract) instead of being supplied: replaces
 * obvhs::ploc::build_ploc + the node mapping at src/raytracing/extract.rs:316-332 for callers that opt in.  `ranges` may name models / materials only.
 * `out_nodes` (nullable) receives the 2*n_models-1 nodes in the reference layout. */
int bvr_upload_scene_gpu_bvh(BvrContext* ctx,
                             const BvrModel* models, size_t n_models,
                             const BvrMaterial* materials, size_t n_materials,
                             const BvrDirtyRange* ranges, size_t n_ranges,
                             BvrBvhNode* out_nodes);

/* Same arguments; keeps the TOPOLOGY of the tree the last bvr_upload_scene_gpu_bvh built for these counts and only refits
 * its boxes to the moved / resized spheres (bottom-up, one kernel) — the cheap path for small motion; the caller decides
 * when the tree has degraded enough to rebuild.  The image does not depend on the choice (any valid tree gives the
 * reference's closest hits). */
int bvr_refit_scene_gpu_bvh(BvrContext* ctx,
                            const BvrModel* models, size_t n_models,
                            const BvrMaterial* materials, size_t n_materials,
                            const BvrDirtyRange* ranges, size_t n_ranges,
                            BvrBvhNode* out_nodes);

/* Rows this shard renders for an image of `height` rows (== height when unsharded). */
uint32_t bvr_shard_rows(uint32_t height, const BvrRenderOptions* opts);

/* Host-only helper (no context, no GPU): validates a node array against the reference's contract the way
 * bvr_upload_scene does and reports what the upload derives from it — the number of tree levels and, per model, its
 * position in the reference's traversal order (raytrace.wgsl:329-341 pops `index + 1` before `index`, so the order
 * over the leaves is fixed by the tree).  raytrace.wgsl:354 keeps the FIRST sphere that reaches a given t; the
 * kernels, which visit nodes in another order, resolve bit-exact ties in t towards the lower rank.
 * out_ranks: n_models words (0xFFFFFFFF for a model no reachable leaf holds), may be NULL; out_depth may be NULL.
 * Returns BVR_OK or BVR_ERR_BAD_SCENE / BVR_ERR_INVALID_ARGUMENT. */
int bvr_scene_traversal_ranks(const BvrBvhNode* nodes, size_t n_nodes, size_t n_models,
                              uint32_t* out_ranks, uint32_t* out_depth);

/* Render one frame for one view with HOST buffers: inputs are copied host->device, outputs
 * device->host, and the call returns when the outputs are complete.
 * raster_rgba: 4 floats / pixel (the post-tonemap main texture, pipeline.rs:166), full image;
 * raster_depth: 1 float / pixel (prepass reverse-Z depth, pipeline.rs:113,169), full image.
 * Both may be NULL for level 3 (Pure); required for levels 0-2. */
int bvr_render(BvrContext* ctx,
               const BvrCamera* camera, const BvrRaytraceLevel* level, const BvrWindow* window,
               const BvrRenderOptions* opts,
               const float* raster_rgba, const float* raster_depth,
               const BvrOutputs* host_out);

/* bvr_render without the wait: every host buffer (raster inputs, output planes) must be page-locked; the call enqueues
 * the input copies, the kernels and the output copies on the context stream and returns.  The frame is complete, and the
 * buffers are the caller's again, after bvr_sync.  Two contexts on one device (each has its own stream) pipeline an
 * animated scene: frame f renders in one while frame f+1's scene is uploaded / its BVH built in the other. */
int bvr_render_async(BvrContext* ctx,
                     const BvrCamera* camera, const BvrRaytraceLevel* level, const BvrWindow* window,
                     const BvrRenderOptions* opts,
                     const float* raster_rgba, const float* raster_depth,
                     const BvrOutputs* host_out);

/* Same, but every pointer (raster inputs and outputs) is a DEVICE pointer owned by the caller and
 * the call only enqueues work on the context stream (use bvr_sync or stream semantics). */
int bvr_render_device(BvrContext* ctx,
                      const BvrCamera* camera, const BvrRaytraceLevel* level, const BvrWindow* window,
                      const BvrRenderOptions* opts,
                      const float* d_raster_rgba, const float* d_raster_depth,
                      const BvrOutputs* device_out);

/* Sample sharding: dst[i] = (dst[i]*dst_weight + src[i]*src_weight) on device;
 * used to weight per-seed partial averages around an NCCL reduce.  n = number of floats.  src_weight == 0 (or
 * d_src == d_dst) is an in-place scale: d_src is not read. */
int bvr_axpby_device(BvrContext* ctx, float* d_dst, float dst_weight,
                     const float* d_src, float src_weight, size_t n);

/* The depth composite of `fragment` (raytrace.wgsl:104-120) as a pass of its own, for frames rendered with
 * BVR_RENDER_DEFER_COMPOSITE: rgba[i] = raster_rgba[i] where the raster depth wins against rt_depth[i].  Levels 0 and 3
 * have no depth test: no-op.  All pointers are device pointers; n_pixels texels. */
int bvr_composite_device(BvrContext* ctx, const BvrCamera* camera, const BvrRaytraceLevel* level, float* d_rgba,
                         const float* d_rt_depth, const float* d_raster_rgba, const float* d_raster_depth, size_t n_pixels);

/* Tile sharding: scatter `shard_count` gathered shard planes (each bvr_shard_rows()-max rows of
 * `width` pixels x `channels` 32-bit words, laid out back to back with `shard_stride_words`)
 * into the full image.  All pointers are device pointers. */
int bvr_unshard_device(BvrContext* ctx, const void* d_gathered, size_t shard_stride_words,
                       void* d_full, uint32_t width, uint32_t height, uint32_t channels,
                       uint32_t shard_count, uint32_t strip_rows);

/* Sample sharding over peer memory (one process per GPU of ONE node).  Instead of scaling its partial frame and
 * handing it to an NCCL reduce, every rank renders straight INTO its own slot of a buffer that lives on rank 0 — the
 * render kernel's pixel stores go over NVLink as the pixels finish (output_weight above applies the share) — and rank 0
 * adds the slots up in rank order once all ranks are done.  The exchange overlaps the rendering, the sum has a fixed
 * order (bit-reproducible, unlike a reduction tree), and what is left on the critical path is one small barrier + one pass
 * over the slots.
 *   bvr_peer_alloc   cudaMalloc on this context's device + the 64-byte handle another process opens
 *   bvr_peer_open    maps a buffer another rank allocated (peer access is enabled on first use)
 *   bvr_peer_close / bvr_peer_free
 *   bvr_sum_slots_device   d_dst[i] = ((slot_0[i] + slot_1[i]) + slot_2[i]) + ... over the slots whose bit is set in `mask` */
#define BVR_PEER_HANDLE_BYTES 64
int bvr_peer_alloc(BvrContext* ctx, size_t bytes, void** d_ptr, uint8_t* handle_out);
int bvr_peer_open(BvrContext* ctx, const uint8_t* handle, void** d_ptr);
int bvr_peer_close(BvrContext* ctx, void* d_ptr);
int bvr_peer_free(BvrContext* ctx, void* d_ptr);
int bvr_sum_slots_device(BvrContext* ctx, const float* d_slots, size_t slot_stride_floats, uint32_t n_slots, uint64_t mask,
                         float* d_dst, size_t n);

int bvr_get_stats(BvrContext* ctx, BvrStats* out);

/* Measurement helper (not on the render path): FP32 FMA throughput of `device` in TFLOP/s, the
 * denominator of the FP32 roofline (MEASURED_PEAKS.json has no FP32 figure). */
int bvr_bench_fp32_peak(int device, float* tflops_out);
/* Measurement helper: L2 -> SM read bandwidth of `device` in GB/s (a 32 MiB L2-resident buffer streamed by every
 * CTA): the roofline denominator for scenes whose nodes are fetched from L2 (BASELINE configs[3]). */
int bvr_bench_l2_bandwidth(int device, float* gbs_out);

#ifdef __cplusplus
} /* extern "C" */

static_assert(sizeof(BvrModel) == 32, "Model stride (encase std430)");
static_assert(sizeof(BvrMaterial) == 32, "Material stride");
static_assert(sizeof(BvrBvhNode) == 48, "BVHNode stride");
static_assert(sizeof(BvrCamera) == 80, "Camera uniform size");
static_assert(sizeof(BvrRaytraceLevel) == 32, "RaytraceLevel uniform size");
static_assert(sizeof(BvrWindow) == 16, "Window uniform size");
static_assert(offsetof(BvrModel, radius) == 12 && offsetof(BvrModel, material_id) == 16, "Model offsets");
static_assert(offsetof(BvrMaterial, metallic) == 12 && offsetof(BvrMaterial, specular_transmission) == 28, "Material offsets");
static_assert(offsetof(BvrBvhNode, bounds_max) == 16 && offsetof(BvrBvhNode, index) == 28 &&
              offsetof(BvrBvhNode, model_count) == 32, "BVHNode offsets");
static_assert(offsetof(BvrCamera, near_plane) == 12 && offsetof(BvrCamera, aspect) == 24 &&
              offsetof(BvrCamera, position) == 32 && offsetof(BvrCamera, direction) == 48 &&
              offsetof(BvrCamera, up) == 64, "Camera offsets");
static_assert(offsetof(BvrRaytraceLevel, _padding) == 16, "RaytraceLevel offsets");
static_assert(offsetof(BvrWindow, height) == 4, "Window offsets");
#endif

#endif /* BEVYRAY_B200_H */
