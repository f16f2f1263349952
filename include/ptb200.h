/* ptb200.h — C ABI of libptb200.so: the B200 (sm_100a) CUDA wavefront path tracer that replaces the render loop
 * GLSL-PathTracer drives through src/shaders/tile.glsl.
 *
 * Every entry point cites the reference interface it replaces (paths relative to the reference tree).
 * Conventions: plain pointers and sizes only, no C++/torch types; every function returns a PtbStatus (0 = ok) and
 * never throws; one context per GPU; the calling thread is the launching thread (the reference's Renderer is
 * single-threaded on the GL-context thread, Main.cpp:613,650-653).  All work is enqueued on the context's CUDA
 * stream; functions that return host data synchronise that stream.
 *
 * There is NO CPU fallback: if no CUDA device is usable ptb_create fails with PTB_ERR_NO_DEVICE.
 */
#ifndef PTB200_H
#define PTB200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum PtbStatus {
    PTB_OK = 0,
    PTB_ERR_INVALID_ARGUMENT = 1,   /* null pointer / bad size (reference: printf("No Scene Found"), Renderer.cpp:72-76) */
    PTB_ERR_NO_DEVICE = 2,          /* no usable CUDA device: the product has no CPU path */
    PTB_ERR_CUDA = 3,               /* a CUDA runtime call failed; see ptb_last_error() */
    PTB_ERR_UNSUPPORTED = 4,        /* scene exceeds an encoding limit (leaf size, stack depth, node count) */
    PTB_ERR_OUT_OF_MEMORY = 5
} PtbStatus;

/* Feature bits = the `#define OPT_*` set Renderer::InitShaders derives (Renderer.cpp:401-459). */
enum {
    PTB_OPT_ENVMAP = 1 << 0, PTB_OPT_LIGHTS = 1 << 1, PTB_OPT_RR = 1 << 2, PTB_OPT_UNIFORM_LIGHT = 1 << 3,
    PTB_OPT_OPENGL_NORMALMAP = 1 << 4, PTB_OPT_HIDE_EMITTERS = 1 << 5, PTB_OPT_BACKGROUND = 1 << 6,
    PTB_OPT_TRANSPARENT_BACKGROUND = 1 << 7, PTB_OPT_ALPHA_TEST = 1 << 8, PTB_OPT_ROUGHNESS_MOLLIFICATION = 1 << 9,
    PTB_OPT_MEDIUM = 1 << 10, PTB_OPT_VOL_MIS = 1 << 11
};

/* The arrays Renderer::InitGPUDataBuffers uploads (Renderer.cpp:135-249), exactly as Scene holds them
 * (Scene.h:87-125).  All pointers are HOST pointers; ptb_create copies them to device buffers. */
typedef struct PtbSceneDesc {
    const float*   nodes;        int32_t numNodes;      /* scene->bvhTranslator.nodes: 36 B = 9 floats each (bvh_translator.h:45-50) */
    int32_t        topLevelIndex;                       /* scene->bvhTranslator.topLevelIndex (Renderer.cpp:504) */
    const int32_t* vertIndices;  int32_t numIndices;    /* scene->vertIndices, 3 ints each (Renderer.cpp:150) */
    const float*   verticesUVX;  int32_t numVertices;   /* scene->verticesUVX, xyz+u (Renderer.cpp:158) */
    const float*   normalsUVY;                          /* scene->normalsUVY,  xyz+v (Renderer.cpp:166), numVertices entries */
    const float*   materials;    int32_t numMaterials;  /* scene->materials, 128 B = 32 floats each (Material.h:87-119) */
    const float*   transforms;   int32_t numInstances;  /* scene->transforms, Mat4 = 16 floats each (Renderer.cpp:182) */
    const float*   lights;       int32_t numLights;     /* scene->lights, 60 B = 15 floats each (Scene.h:50-59) */
    const uint8_t* textures;     int32_t numTextures, texW, texH;  /* scene->textureMapsArray RGBA8 (Renderer.cpp:200-208) */
    const float*   envImg;       const float* envCdf;   /* scene->envMap->img (RGB32F) / ->cdf (R32F) (Renderer.cpp:211-226) */
    int32_t        envW, envH;   float envTotalSum;     /* envMapRes / envMapTotalSum uniforms (Renderer.cpp:498-502) */
} PtbSceneDesc;

/* RenderOptions fields the path reads (Renderer.h:37-100) + the derived feature mask. */
typedef struct PtbOptions {
    int32_t  renderW, renderH;         /* renderOptions.renderResolution */
    int32_t  tileW, tileH;             /* renderOptions.tileWidth/Height (Renderer.cpp:290-297) */
    int32_t  maxDepth;                 /* uniform maxDepth (Renderer.cpp:779) */
    int32_t  rrDepth;                  /* OPT_RR_DEPTH (Renderer.cpp:413) */
    uint32_t features;                 /* PTB_OPT_* mask; ptb_derive_features() reproduces Renderer.cpp:401-459 */
    float    envMapIntensity;          /* Renderer.cpp:777 */
    float    envMapRot;                /* renderOptions.envMapRot in degrees; divided by 360 as at Renderer.cpp:778 */
    float    roughnessMollificationAmt;/* Renderer.cpp:782 */
    float    uniformLightCol[3];       /* Renderer.cpp:781 */
    float    backgroundCol[3];         /* tonemap uniform (Renderer.cpp:810) */
    int32_t  enableTonemap, enableAces, simpleAcesFit;  /* Renderer.cpp:807-809 */
    int32_t  samplesPerWave;           /* B200 tuning: samples kept in flight per wavefront (0 = auto) */
} PtbOptions;

/* camera.* uniforms (Renderer.cpp:769-775); fov in radians as Camera::fov. */
typedef struct PtbCamera {
    float position[3], right[3], up[3], forward[3];
    float fov, focalDist, aperture;
} PtbCamera;

/* Per-ray closest-hit record for the G1 parity gate. */
typedef struct PtbHit {
    float   t;          /* 1e6 (INF, globals.glsl:31) on miss */
    int32_t kind;       /* 0 miss, 1 triangle, 2 analytic light */
    int32_t instance;   /* -leaf-1 of the TLAS leaf for triangle hits, else -1 */
    int32_t matID;      /* TLAS-leaf LRLeaf.y for triangle hits, else -1 */
    int32_t primSlot;   /* leftIndex+i of the winning triangle (closest_hit.glsl:120), else -1 */
    int32_t triIDx;     /* vertIndices[primSlot].x, else -1 */
    float   bary[3];    /* uvt.wxy (closest_hit.glsl:148) */
    int32_t lightIdx;   /* winning light for kind 2, else -1 */
} PtbHit;

typedef struct PtbBsdfQuery {            /* G2 parity gate: DisneyEval / DisneySample on a fixed sample set */
    float mat[32];                        /* one Material row as uploaded */
    float V[3], N[3], L[3];               /* view dir, face-forward normal, light dir (world) */
    float eta;                            /* state.eta (pathtrace.glsl:114) */
    float r1, r2, r3;                     /* the three rand() draws of DisneySample (disney.glsl:146-147,192) */
} PtbBsdfQuery;
typedef struct PtbBsdfResult { float f[3]; float pdf; float L[3]; } PtbBsdfResult;

typedef struct PtbStats {
    uint64_t pathSegments;     /* closest-hit rays traced (camera, bounce, alpha-skip re-traces, transmittance steps) */
    uint64_t shadowRays;       /* any-hit rays traced */
    uint64_t samplesRendered;  /* full-frame sample passes completed since the last reset */
    uint64_t kernelLaunches;   /* CUDA kernels launched by this context since creation */
    float    lastRenderMs;     /* device time of the last ptb_render_* call (CUDA events on the context stream) */
    float    lastTraceMs;      /* device time of the closest-hit launches within it (0 unless profiling is enabled) */
    uint64_t lastTraceRays;    /* rays those launches traced */
    float    lastCameraMs, lastSortMs, lastShadeMs, lastShadowMs, lastAccumMs;   /* the other kernel classes of that call (profiling only) */
    float    reserved_;
} PtbStats;

typedef struct PtbCtx PtbCtx;

/* Renderer::Renderer (Renderer.cpp:41-87): InitGPUDataBuffers -> device buffers (+ derived traversal layout),
 * InitFBOs -> accumulation/output buffers, InitShaders -> feature selection.  device = CUDA ordinal. */
int  ptb_create(const PtbSceneDesc* scene, const PtbOptions* opts, int device, PtbCtx** out);
/* Renderer::~Renderer (Renderer.cpp:89-133). */
int  ptb_destroy(PtbCtx* ctx);
const char* ptb_last_error(void);

/* Reproduces the define derivation of Renderer::InitShaders (Renderer.cpp:401-459) from scene + option booleans.
 * optionBools bit i set = {enableEnvMap, enableRR, enableUniformLight, openglNormalMap, hideEmitters, enableBackground,
 * transparentBackground, enableRoughnessMollification, enableVolumeMIS}[i]. */
uint32_t ptb_derive_features(const PtbSceneDesc* scene, uint32_t optionBools);

/* Renderer::ReloadShaders / InitShaders (Renderer.cpp:381-392): change options/features without re-uploading. */
int  ptb_set_options(PtbCtx* ctx, const PtbOptions* opts);
/* Renderer::ResizeRenderer (Renderer.cpp:251-279): new render/tile size; clears accumulation, counters restart. */
int  ptb_resize(PtbCtx* ctx, int32_t w, int32_t h, int32_t tileW, int32_t tileH);
/* camera uniforms pushed by Renderer::Update (Renderer.cpp:769-775). */
int  ptb_set_camera(PtbCtx* ctx, const PtbCamera* cam);
/* Renderer::Update instancesModified branch (Renderer.cpp:649-665): re-upload transforms, materials and the TLAS slice
 * nodes[topLevelIndex..numNodes) (9 floats each, numTlasNodes = numNodes - topLevelIndex). */
int  ptb_update_instances(PtbCtx* ctx, const float* transforms, int32_t numInstances, const float* materials, int32_t numMaterials,
                          const float* tlasNodes, int32_t numTlasNodes);
/* Scene::RebuildInstances (Scene.cpp:200-214) + the upload above, from the TRANSFORMS alone: the TLAS is rebuilt by the library — on the device
 * (k_tlas_build: instance world boxes of Scene::createTLAS, Scene.cpp:148-187; Bvh(10, 64, false)::Build, RadeonRays/bvh.cpp:68-243;
 * BvhTranslator::ProcessTLASNodes, bvh_translator.cpp:58-86) — into the canonical node array, byte-identical to what the reference's host code builds
 * (ptb_read_nodes shows it).  instanceMaterialIDs: per-instance material id (MeshInstance::materialID), NULL = unchanged.  onHost != 0, and inputs where the
 * reference's in-place partition order matters (coincident centroids, -0.0 / non-finite boxes), use the exact sequential builder on the host instead;
 * ptb_last_rebuild_info: where = 0 device, 1 host as asked, 2 host as fallback; buildMs = the TLAS build alone (kernel time / host builder time),
 * totalMs = the whole call incl. the derivation and upload of the packed layouts. */
int  ptb_rebuild_instances(PtbCtx* ctx, const float* transforms, int32_t numInstances, const float* materials, int32_t numMaterials,
                           const int32_t* instanceMaterialIDs, int32_t onHost);
int  ptb_last_rebuild_info(PtbCtx* ctx, int32_t* where, float* buildMs, float* totalMs);
/* Renderer::Update envMapModified branch (Renderer.cpp:668-692). */
int  ptb_update_envmap(PtbCtx* ctx, const float* img, const float* cdf, int32_t w, int32_t h, float totalSum);

/* Renderer::Update dirty branch (Renderer.cpp:733-744): clear the accumulation buffer. */
int  ptb_reset_accum(PtbCtx* ctx);
/* One Renderer::Render() tile pass (Renderer.cpp:566-580): 1 spp for tile (tx,ty) with uniform frameNum (Renderer.cpp:783),
 * added into the accumulation buffer.  Off-image pixels of over-hanging tiles are skipped. */
int  ptb_render_tile(PtbCtx* ctx, int32_t tx, int32_t ty, int32_t frameNum);
/* Whole-frame fast path: nSamples sample passes starting at 1-based pass firstSample, every tile of a pass in one
 * wavefront, with the frameNum/tile-local RNG seeding the reference's tile schedule would have used
 * (Renderer.cpp:745-783, tile.glsl:45), so the sum equals nSamples x numTiles Render() calls.
 * sampleStride > 1 renders passes firstSample, firstSample+stride, ... (multi-GPU sample sharding). */
int  ptb_render_samples(PtbCtx* ctx, int32_t firstSample, int32_t nSamples, int32_t sampleStride);
/* One sample pass of the WHOLE frame for a host that keeps the reference's one-tile-per-Render() loop (Main.cpp:175-212,
 * Renderer.cpp:566-589, 745-762): called at the first tile of pass `sample`, it replaces the numTiles Render() draws of that pass —
 * only the completed buffer is observable (Renderer.cpp:628-633), so the other tiles' calls become no-ops.  The library traces a
 * wave of up to maxLookahead passes sample, sample+sampleStride, ... at once (<= 0: its own wave size) and adds only pass `sample`
 * to the running sum; the following calls (sample+sampleStride, ...) add their pass from the resident wave, unless camera, options
 * or scene arrays changed in between — then the wave is discarded and traced again.  Sum, seeds and frameNum schedule equal
 * ptb_render_samples(sample, 1, 1).  sampleStride > 1: the other passes belong to other GPUs (ptb_mgpu_render_pass). */
int  ptb_render_pass(PtbCtx* ctx, int32_t sample, int32_t sampleStride, int32_t maxLookahead);
/* preview.glsl:41-71 + Renderer.cpp:555-565,798: quarter-resolution 1-spp depth-2 render, no accumulation.
 * out = (w*h*4) floats, w = windowW*0.25, h = windowH*0.25. */
int  ptb_render_preview(PtbCtx* ctx, int32_t w, int32_t h, float* outRgba);

/* Linear running sum (the accumTexture, Renderer.cpp:335-341): w*h*4 floats, row 0 = bottom. Host copy. */
int  ptb_read_accum_f32(PtbCtx* ctx, float* outRgba);
/* Overwrite the running sum from the host (checkpoint restore / multi-GPU merge). */
int  ptb_write_accum_f32(PtbCtx* ctx, const float* rgba);
/* Device pointer of the running sum (float4[w*h]) so a host framework can reduce it with NCCL over NVLink. */
int  ptb_accum_device_ptr(PtbCtx* ctx, void** devPtr, uint64_t* nbytes);
/* tonemap.glsl:97-133 with invSampleCounter (Renderer.cpp:806) + glGetTexImage(GL_RGBA, GL_UNSIGNED_BYTE)
 * (Renderer.cpp:619-634): tonemapped gamma-2.2 RGBA8, bottom row first, w*h*4 bytes. */
int  ptb_read_output_rgba8(PtbCtx* ctx, float invSampleCounter, uint8_t* outRgba8);
/* Same, tonemapping a caller-provided device buffer (float4[w*h], e.g. the NCCL-reduced sum of several contexts' running sums)
 * instead of the context's own; devAccum = NULL is ptb_read_output_rgba8. */
int  ptb_read_output_rgba8_from(PtbCtx* ctx, const void* devAccum, float invSampleCounter, uint8_t* outRgba8);
/* The tonemap pass Render() runs after a tile (Renderer.cpp:584-588) into tileOutputTexture[currentBuffer], kept on the DEVICE:
 * ptb_snapshot_output freezes the tonemapped image of the pass just completed (asynchronous, no host copy);
 * ptb_read_snapshot_rgba8 is the glGetTexImage of GetOutputBuffer (Renderer.cpp:619-634) on that frozen image (zeros before
 * the first snapshot). */
int  ptb_snapshot_output(PtbCtx* ctx, float invSampleCounter);
int  ptb_snapshot_output_from(PtbCtx* ctx, const void* devAccum, float invSampleCounter);   /* ... of a caller-provided device sum */
int  ptb_read_snapshot_rgba8(PtbCtx* ctx, uint8_t* outRgba8);
/* Denoiser hook support (Renderer.cpp:695-728 reads tileOutputTexture[1-currentBuffer] as GL_RGB / GL_FLOAT): when enabled, a
 * snapshot also keeps the RGBA32F form of the tonemapped image; ptb_read_snapshot_rgb32f copies its rgb to the host (w*h*3 floats). */
int  ptb_set_snapshot_float(PtbCtx* ctx, int32_t enable);
int  ptb_read_snapshot_rgb32f(PtbCtx* ctx, float* outRgb);
/* Page-locked host memory for readback targets (a pageable target costs ~3 % of the end-to-end rate at 1080p). */
int  ptb_host_alloc(uint64_t nbytes, void** out);
int  ptb_host_free(void* p);

int  ptb_get_stats(PtbCtx* ctx, PtbStats* out);
int  ptb_reset_stats(PtbCtx* ctx);
/* enable per-kernel CUDA-event timing of the closest-hit launches (adds events, no syncs in the loop). */
int  ptb_set_profiling(PtbCtx* ctx, int32_t enable);
/* Use an externally owned cudaStream_t (e.g. torch's current stream) for all subsequent work; NULL = own stream. */
int  ptb_set_stream(PtbCtx* ctx, void* cudaStream);
int  ptb_synchronize(PtbCtx* ctx);
/* Traversal variant.  enable = 1 (default): child boxes whose entry distance exceeds the current hit distance x 1.00001 are skipped
 * (SURVEY H3); the visiting order of the remaining nodes is the reference's.  enable = 0: the reference's unculled traversal
 * (closest_hit.glsl:173-205 never compares box distances with t).  Each variant is bit-identical (IDs and t) to the oracle's host
 * traversal of the same variant.  The two variants give identical hits except when two triangles share an edge and their computed t
 * differ by an ulp: the skipped box may then hold the one the unculled order keeps (observed once in 1.3e5 primary rays of the
 * instancing scene, |dt| = 2 ulp; never on the in-repo scenes).  Culling removes ~30 % of the closest-hit kernel's time. */
int  ptb_set_cull(PtbCtx* ctx, int32_t enable);

/* Parity entry points: run the production traversal / BSDF device code on caller-provided inputs.
 * rays: n x 6 floats (origin, direction), HOST pointers; depth = state.depth seen by OPT_HIDE_EMITTERS. */
int  ptb_trace_closest(PtbCtx* ctx, const float* rays, int64_t n, int32_t depth, PtbHit* out);
int  ptb_trace_any(PtbCtx* ctx, const float* rays, const float* maxDist, int64_t n, int32_t* outOccluded);
int  ptb_bsdf_eval(PtbCtx* ctx, const PtbBsdfQuery* q, int64_t n, PtbBsdfResult* out);
int  ptb_bsdf_sample(PtbCtx* ctx, const PtbBsdfQuery* q, int64_t n, PtbBsdfResult* out);
/* lambert.glsl:25-46 (LambertEval / LambertSample with r1, r2 of the query).  The reference includes these in tile.glsl but never
 * calls them from PathTrace, so they are not on the render path here either; exposed for parity only. */
int  ptb_lambert_eval(PtbCtx* ctx, const PtbBsdfQuery* q, int64_t n, PtbBsdfResult* out);
int  ptb_lambert_sample(PtbCtx* ctx, const PtbBsdfQuery* q, int64_t n, PtbBsdfResult* out);
/* Camera rays of 1-based sample pass `sample` for every pixel (tile.glsl:41-68): w*h*6 floats to the host. */
int  ptb_camera_rays(PtbCtx* ctx, int32_t sample, float* outRays);
/* Device-resident variants for benchmarking the traversal kernel alone (inputs/outputs already in HBM). */
int  ptb_trace_closest_device(PtbCtx* ctx, const void* devRays, int64_t n, int32_t depth, void* devHits);
/* Read back a device copy of the canonical node array (byte-exact G1 check after upload/update). */
int  ptb_read_nodes(PtbCtx* ctx, float* outNodes, int32_t numNodes);
/* Required traversal stack depth computed from the uploaded hierarchy (TLAS height + marker + max BLAS height). */
int  ptb_stack_depth(PtbCtx* ctx, int32_t* out);

/* ---- several GPUs of one box behind one handle -------------------------------------------------------
 * For a host that is one process and one thread, as the reference's Renderer is (Renderer.h:169-179 knows no device): N contexts
 * (one scene replica per GPU) + one NCCL communicator.  Sample passes are sharded round-robin (context r renders passes
 * first+r, first+r+N, ... with the seeds a 1-GPU run would use); there is no data-path collective.  A readback combines the N
 * running sums with ONE ncclReduce over NVLink into a scratch buffer on devices[0] — the per-GPU sums stay untouched, so
 * progressive readbacks are correct — and tonemaps there.  NCCL is bound at run time (dlopen) and only when numDevices > 1.
 * devices = NULL means ordinals 0..numDevices-1.  One process per GPU with torch.distributed (bench.py) is the other supported
 * arrangement; it uses ptb_accum_device_ptr + ptb_read_output_rgba8_from. */
typedef struct PtbMgpu PtbMgpu;
int  ptb_mgpu_create(const PtbSceneDesc* scene, const PtbOptions* opts, const int32_t* devices, int32_t numDevices, PtbMgpu** out);
int  ptb_mgpu_destroy(PtbMgpu* m);
int  ptb_mgpu_num_devices(PtbMgpu* m);
PtbCtx* ptb_mgpu_context(PtbMgpu* m, int32_t i);                                    /* the i-th per-GPU context (stats, parity entry points) */
int  ptb_mgpu_set_options(PtbMgpu* m, const PtbOptions* opts);                      /* ptb_set_options on every context */
int  ptb_mgpu_set_camera(PtbMgpu* m, const PtbCamera* cam);
int  ptb_mgpu_set_cull(PtbMgpu* m, int32_t enable);
int  ptb_mgpu_update_instances(PtbMgpu* m, const float* transforms, int32_t numInstances, const float* materials, int32_t numMaterials,
                               const float* tlasNodes, int32_t numTlasNodes);
int  ptb_mgpu_rebuild_instances(PtbMgpu* m, const float* transforms, int32_t numInstances, const float* materials, int32_t numMaterials,
                                const int32_t* instanceMaterialIDs, int32_t onHost);                /* ptb_rebuild_instances on every GPU */
int  ptb_mgpu_update_envmap(PtbMgpu* m, const float* img, const float* cdf, int32_t w, int32_t h, float totalSum);
int  ptb_mgpu_reset_accum(PtbMgpu* m);
int  ptb_mgpu_render_samples(PtbMgpu* m, int32_t firstSample, int32_t nSamples);    /* passes [first, first+n) over all GPUs, asynchronous */
int  ptb_mgpu_render_pass(PtbMgpu* m, int32_t sample, int32_t maxLookahead);         /* ptb_render_pass on GPU (sample-1) mod N, stride N */
int  ptb_mgpu_read_output_rgba8(PtbMgpu* m, float invSampleCounter, uint8_t* outRgba8);   /* reduce -> tonemap -> host (GetOutputBuffer) */
/* ptb_snapshot_output for N GPUs: every GPU freezes its own running sum (asynchronous device copy, no cross-GPU synchronisation while the GPUs
 * render their waves); the ONE ncclReduce + tonemap of the frozen sums happens when the image is read. */
int  ptb_mgpu_snapshot_output(PtbMgpu* m, float invSampleCounter);
int  ptb_mgpu_read_snapshot_rgba8(PtbMgpu* m, uint8_t* outRgba8);
int  ptb_mgpu_read_snapshot_rgb32f(PtbMgpu* m, float* outRgb);                      /* for the denoiser hook (ptb_set_snapshot_float on context 0) */
int  ptb_mgpu_read_accum_f32(PtbMgpu* m, float* outRgba);                           /* reduced linear sum (parity) */
int  ptb_mgpu_get_stats(PtbMgpu* m, PtbStats* out);                                 /* counters summed, lastRenderMs = max over GPUs */
int  ptb_mgpu_synchronize(PtbMgpu* m);

#ifdef __cplusplus
}
#endif
#endif
