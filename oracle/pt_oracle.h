/* pt_oracle.h — C API of the CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A line-by-line CPU restatement of the reference's fragment-shader path tracer
 * (src/shaders/tile.glsl + every file under src/shaders/common) operating on the arrays the
 * unmodified reference host code produces (oracle/ref_host/scene_dump.cpp).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product (libptb200.so) never links or calls it.
 *
 * Parity pinning: PINNED against the executing reference.  The reference ships no tests or golden vectors
 * and no GL/Mesa exists in this image, but its shader text is C-like: oracle/glsl_ref compiles the reference's own
 * tile.glsl / preview.glsl / tonemap.glsl (+ the common/ includes), read in place from /root/reference, with g++ (lexical rewrites
 * only) and runs them on the host.  This restatement is BIT-IDENTICAL to that executing reference on 8 scenes + 15 feature
 * variants (accumulation buffers, previews, RGBA8 readbacks), on ClosestHit / AnyHit per ray and on DisneyEval per query
 * (tests/test_glsl_ref.py; golden vectors tests/golden/glslref_golden.npz).  The flattened BVH / mesh arrays consumed here
 * are produced by the reference's own host code and checked against SURVEY §8(c) FNV hashes.  What stays implementation-
 * defined in GL (normalize, inverse, bilinear weights) is pinned to the spec formula in both.
 */
#ifndef PT_ORACLE_H
#define PT_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* Raw views of the arrays Renderer::InitGPUDataBuffers uploads (Renderer.cpp:135-249). */
typedef struct OrcSceneDesc {
    const float*   nodes;        int32_t numNodes;     /* 9 floats per node (bvh_translator.h:45-50) */
    int32_t        topLevelIndex;
    const int32_t* vertIndices;  int32_t numIndices;   /* 3 ints  (Scene.h:61-64)  */
    const float*   verticesUVX;  int32_t numVertices;  /* 4 floats */
    const float*   normalsUVY;                          /* 4 floats */
    const float*   materials;    int32_t numMaterials; /* 32 floats (Material.h:87-119) */
    const float*   transforms;   int32_t numInstances; /* 16 floats (Mat4.h) */
    const float*   lights;       int32_t numLights;    /* 15 floats (Scene.h:50-59) */
    const uint8_t* textures;     int32_t numTextures, texW, texH; /* RGBA8 array */
    const float*   envImg;       const float* envCdf;  int32_t envW, envH; float envTotalSum;
} OrcSceneDesc;

/* Feature defines of Renderer.cpp:401-459 + uniforms of Renderer.cpp:766-811. */
typedef struct OrcOptions {
    int32_t optEnvMap, optLights, optRR, rrDepth, optUniformLight, optOpenglNormalMap, optHideEmitters,
            optBackground, optTransparentBackground, optAlphaTest, optRoughnessMollification, optMedium, optVolMis;
    int32_t maxDepth;
    float   envMapIntensity, envMapRot /* already /360 */, roughnessMollificationAmt;
    float   uniformLightCol[3];
    int32_t renderW, renderH, tileW, tileH;
    /* camera uniforms (Renderer.cpp:769-775) */
    float   camPosition[3], camRight[3], camUp[3], camForward[3], camFov, camFocalDist, camAperture;
    /* traversal variant: 0 = reference-faithful (no t-culling), 1 = cull child boxes with entry t > current t */
    int32_t cullBoxes;
} OrcOptions;

typedef struct OrcHit {
    float   t;          /* INF (1e6) on miss */
    int32_t kind;       /* 0 miss, 1 triangle, 2 analytic light */
    int32_t instance;   /* -leaf-1 of the TLAS leaf (triangle hits), else -1 */
    int32_t matID;      /* TLAS leaf LRLeaf.y (triangle hits), else -1 */
    int32_t primSlot;   /* leftIndex + i of the winning triangle, else -1 */
    int32_t triIDx;     /* vertIndices[primSlot].x, else -1 */
    float   bary[3];    /* uvt.wxy */
    int32_t lightIdx;   /* winning light for kind 2, else -1 */
} OrcHit;

typedef struct OrcStats {
    uint64_t closestRays, anyRays;            /* path segments / shadow rays */
    uint64_t nodeVisits, internalSteps, triTests, tlasLeaves;   /* closest-hit traversals */
    uint64_t surfaceHits;
    uint64_t anyNodeVisits, anyInternalSteps, anyTriTests, anyTlasLeaves;   /* any-hit traversals */
} OrcStats;

/* BSDF probe input: material row (32 floats as uploaded) + geometry. */
typedef struct OrcBsdfQuery {
    float mat[32];
    float V[3], N[3] /* ffnormal */, L[3];
    float eta;           /* state.eta */
    float r1, r2, r3;    /* for sampling */
} OrcBsdfQuery;
typedef struct OrcBsdfResult { float f[3]; float pdf; float L[3]; } OrcBsdfResult;

typedef struct OrcCtx OrcCtx;

OrcCtx* orc_create(const OrcSceneDesc* scene, const OrcOptions* opts);
void    orc_destroy(OrcCtx*);
void    orc_set_options(OrcCtx*, const OrcOptions* opts);

/* rays: n x 6 floats (origin, direction); depth: state.depth seen by OPT_HIDE_EMITTERS. */
void orc_trace_closest(OrcCtx*, const float* rays, int64_t n, int32_t depth, OrcHit* out);
/* any-hit without alpha test (alpha needs the path RNG); out[i] = 1 if occluded. */
void orc_trace_any(OrcCtx*, const float* rays, const float* maxDist, int64_t n, int32_t* out);
/* brute force over every instance x triangle + lights, no BVH: used to validate the traversal itself. */
void orc_trace_closest_brute(OrcCtx*, const float* rays, int64_t n, int32_t depth, OrcHit* out);

void orc_bsdf_eval(OrcCtx*, const OrcBsdfQuery* q, int64_t n, OrcBsdfResult* out);
void orc_bsdf_sample(OrcCtx*, const OrcBsdfQuery* q, int64_t n, OrcBsdfResult* out);
/* lambert.glsl:25-46 (dead code in the reference's PathTrace): sample = 0 LambertEval, 1 LambertSample with the query's r1, r2. */
void orc_lambert(OrcCtx*, const OrcBsdfQuery* q, int64_t n, int32_t sample, OrcBsdfResult* out);

/* Camera rays of sample pass `sample` (1-based) for every pixel: n = w*h, 6 floats each (tile.glsl:41-68). */
void orc_camera_rays(OrcCtx*, int32_t sample, float* rays);

/* Analysis aid: the closest-hit ray each pixel's path traces at loop depth `depth` of pass `sample`; valid[i] = 0 if the path ended before. */
void orc_capture_rays(OrcCtx*, int32_t sample, int32_t depth, float* rays, uint8_t* valid);
/* Analysis aid: the light-NEE shadow ray (origin, direction, maxDist, light index = 8 floats) traced while shading at loop depth `depth`. */
void orc_capture_shadow_rays(OrcCtx*, int32_t sample, int32_t depth, float* rays8, uint8_t* valid);

/* Adds `nSamples` full-frame passes (all tiles, reference frameNum schedule, Renderer.cpp:745-783)
 * starting at 1-based pass `firstSample` to accum (w*h*4 floats, row 0 = bottom). */
void orc_render_samples(OrcCtx*, int32_t firstSample, int32_t nSamples, float* accum);
/* Same for a pixel sub-rectangle (bounded CPU-baseline samples). */
void orc_render_samples_rect(OrcCtx*, int32_t firstSample, int32_t nSamples, int32_t x0, int32_t y0, int32_t x1, int32_t y1, float* accum);
/* One Renderer::Render() tile draw (Renderer.cpp:566-580). */
void orc_render_tile(OrcCtx*, int32_t tx, int32_t ty, int32_t frameNum, float* accum);

/* preview.glsl:41-71: w x h preview (1 spp, depth 2, frame 1), out = w*h*4 floats, no accumulation. */
void orc_render_preview(OrcCtx*, int32_t w, int32_t h, float* out);

/* tonemap.glsl:97-133 + GL float->unorm8 conversion; out RGBA8, row 0 = bottom. */
void orc_tonemap(const float* accum, int32_t w, int32_t h, float invSampleCounter, int32_t enableTonemap, int32_t enableAces,
                 int32_t simpleAcesFit, const float* backgroundCol, int32_t optBackground, int32_t optTransparentBackground, uint8_t* out);

void orc_get_stats(OrcCtx*, OrcStats* out);
void orc_reset_stats(OrcCtx*);
int  orc_num_threads(void);
void orc_set_num_threads(int n);   /* OpenMP threads of the CPU arms (process-wide) */

#ifdef __cplusplus
}
#endif
#endif
