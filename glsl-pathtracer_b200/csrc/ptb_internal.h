// ptb_internal.h — POD structures shared by the host side (ptb_api.cpp, g++) and the kernels (ptb_kernels.cu, nvcc).
#pragma once
// 1: inner nodes are laid out in (x,y)/(z,z) pairs and the slab test runs on Blackwell's packed fp32 pipe (add/mul.rn.f32x2 ->
// FADD2/FMUL2: two IEEE-rounded operations per issue slot, bit-identical results).  0: scalar __f*_rn operations.
#ifndef PTB_PACKED_SLAB
#define PTB_PACKED_SLAB 0
#endif
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#include <cuda_runtime.h>
#else
#include <vector_types.h>
#include <vector_functions.h>
#endif

// ---- child / node "meta" word of the derived traversal layout -------------------------------------------------
// bits 31..30 kind; INNER: canonical node index (30 bits); LEAF: count (4 bits, 29..26) | first leaf-ref slot (26 bits);
// INST: bit 29 = translation-only transform, instance index (29 bits); NONE: stack sentinel / BLAS marker (the reference's -1,
// closest_hit.glsl:91,166).
enum : uint32_t { PTB_K_INNER = 0u, PTB_K_LEAF = 1u, PTB_K_INST = 2u, PTB_K_NONE = 3u };
#define PTB_META_NONE 0xFFFFFFFFu
#define PTB_LIGHT_GRID 16
#define PTB_INST_TRANSLATION_ONLY (1u << 29)
#define PTB_INST_INDEX_MASK ((1u << 29) - 1)
#define PTB_MAX_LEAF_TRIS 15
#define PTB_MAX_LEAF_SLOT ((1u << 26) - 1)

// Device-resident scene.  Canonical arrays are byte-identical copies of what Renderer::InitGPUDataBuffers uploads
// (reference Renderer.cpp:135-249); "derived" arrays are re-packings with identical content/indices for 128-bit fetches.
struct DevScene
{
    // canonical (layout unchanged)
    const float*  nodes;        // 9 floats / node
    const int*    vertIndices;  // 3 ints / leaf-ref slot
    const float4* verticesUVX;
    const float4* normalsUVY;
    const float4* materials;    // 8 float4 / material
    const float4* transforms;   // 4 float4 / instance
    const float*  lights;       // 15 floats / light
    const uchar4* textures;     // RGBA8 [layer][y][x]
    const float*  envImg;       // RGB32F
    const float*  envCdf;       // R32F
    // derived
    const float4* inner;        // 4 float4 / canonical node index: child boxes + child metas (valid for internal nodes)
    const float4* tris;         // 3 float4 / leaf-ref slot: v0, e0, e1, vertIndices.x
    const float4* triShade;     // 4 float4 / leaf-ref slot: n0 n1 n2 (9 floats), (u,v) x 3 (ptbd_build_tri_shade)
    const float4* instTrav;     // 4 float4 / instance: rows of inverse(transform) (xyz) + {0, matID, wide rootMeta, rootMeta} in .w
    const float4* instShade;    // 8 float4 / instance: transform rows (4) + inverse(mat3) rows (3) + pad
    const float4* lightsPre;    // 8 float4 / light (see buildLightsPre in ptb_api.cpp)
    const float4* lightGroups;  // 3 float4 / group of consecutive lights (shared plane + padded bounds)
    const uint32_t* lightGrid;  // PTB_LIGHT_GRID^2 member masks per gridded group (ptbd_build_lights)
    const float4* wide;         // 8 float4 / node of the 4-wide any-hit hierarchy (ptbd_build_wide); null = not available
    uint32_t rootMetaWide;      // TLAS root in it
    int numLightGroups;
    uint32_t rootMeta;          // meta of the TLAS root
    int numNodes, topLevelIndex, numIndices, numVertices, numMaterials, numInstances, numLights;
    int numTextures, texW, texH, envW, envH;
    float envTotalSum;
    const uint32_t* envGuide;   // guide table of the CDF search (ptbd_build_env_guide): envGuideN + 1 entries; null = the reference's two binary searches
    float envGuideScale; int envGuideN;
    int stackDepth;             // traversal stack entries needed (bottom sentinel + TLAS path + marker + BLAS path) in the binary hierarchy
    int stackDepthAny;          // ... by the any-hit kernels: max of the binary and the 4-wide hierarchy's bound
};

struct FrameParams
{
    // options
    uint32_t features;
    int maxDepth, rrDepth;
    float envMapIntensity, envMapRot /* already /360 */, roughnessMollificationAmt;
    float uniformLightCol[3];
    int renderW, renderH, tileW, tileH, numTilesX, numTilesY;
    float invNumTilesX, invNumTilesY;
    // camera
    float camPos[3], camRight[3], camUp[3], camFwd[3];
    float camScale /* tan(fov/2), computed on the host */, camFocalDist, camAperture;
    float aspect;        // float(renderH) / float(renderW)
    // derived switches
    int cullBoxes;       // cull child boxes whose entry distance exceeds the current hit distance
    int inlineShadow;    // shadow rays consume path RNG draws (BLEND alpha in AnyHit / EvalTransmittance) -> traced inside shade
    int general;         // shade kernel specialisation: 0 lights only, 1 + env/textures/emission, 2 + media/alpha (k_shade<2> when inlineShadow, else k_shade<3>)
    int deferTransmit;   // OPT_MEDIUM + OPT_VOL_MIS without BLEND materials: EvalTransmittance draws no random number -> NEE rays are queued and k_transmit evaluates them
    // per-column / per-row tables of the pixel -> (frame texture coordinate, tile-local coordinate, tile) mapping (ptbd_build_pixel_tables); null = evaluate per pixel
    const float2* pixTabX; const float2* pixTabY;
};

// One wavefront: S sample passes of a pixel rectangle.
struct WaveParams
{
    int x0, y0, rw, rh;        // pixel rectangle
    int vw, vh;                // rectangle padded to 8x4 pixel blocks
    int nSamples;              // sample passes in this wave
    int firstSample, sampleStride;
    int fixedFrame;            // >= 0: use this frameNum for every pixel (ptb_render_tile); < 0: reference schedule
    int previewMode;           // preview.glsl: InitRNG(gl_FragCoord, 1), TexCoords over the whole image, depth 2
    uint32_t nSlots;           // vw*vh*nSamples
    int accFirst, accCount;    // k_accumulate adds the wave's passes [accFirst, accFirst+accCount) to the running sum (accCount 0 = all of them)
    int lps, lpw;              // block-major only: log2(sample passes per 32-slot group), log2(width of the group's pixel sub-block); see groupToPixel in ptb_kernels.cu
    int blockMajor;            // order of the 32-slot groups (one 8x4 pixel block of one sample pass each): 0 = sample-major (all blocks of pass 0, then pass 1, ...),
                               // 1 = block-major (all passes of block 0, then block 1, ...): a 2048-slot tile then holds the paths of a few neighbouring pixels,
                               // so the tile-local grouping by direction yields warps whose rays share origin AND direction
};

// Path state fields are addressed as base + slot * stride.  stride = sizeof(T) gives SoA arrays; a common 128-byte (lights-only)
// or 192-byte (general) stride interleaves all fields of one path in one or two cache lines, so that the material-sorted (scattered)
// shade passes touch whole sectors instead of 16 bytes out of every 32-byte sector.
#if defined(__CUDACC__)
#define PTB_HD __host__ __device__ __forceinline__
#else
#define PTB_HD inline
#endif
template <class T> struct StateField
{
    char* base; uint32_t stride;
    PTB_HD T& operator[](size_t i) const { return *reinterpret_cast<T*>(base + i * stride); }
};
struct PathState
{
    StateField<float4> rayO;     // origin.xyz, prev scatter pdf
    StateField<float4> rayD;     // direction.xyz, bits: depth (low 16, signed) | flags (high 16)
    StateField<float4> thr;      // throughput.xyz, previous roughness (mollification)
    StateField<float4> rad;      // radiance.xyz, alpha
    StateField<uint4>  rng;      // pcg4d state
    StateField<float4> hit;      // t, bary u, bary v, bits(primSlot)
    StateField<int>    hitInst;  // >=0 instance (triangle hit); -1 miss; <= -2: light -(idx+2)
    StateField<float4> med;      // general: medium density, anisotropy, bits(type), bits(prevMatID)
    StateField<float4> medCol;   // general: medium color.xyz, -
    StateField<float2> prevUV;   // general: stale texCoord (Q2)
    // shadow queues A (env NEE) and B (light NEE), dense by queue slot
    float4* shO[2];   // origin.xyz, maxDist
    float4* shD[2];   // direction.xyz, bits(path slot)
    float4* shC[2];   // contribution.rgb (already multiplied by throughput)
    uint32_t* queue[2];
};

#define PTB_HIT_DEAD ((int)0x80000000)      // hitInst of a path slot that holds no path (off-image pixel of a padded 8x4 block)
#define PTB_FLAG_INMEDIUM   (1u << 16)
#define PTB_FLAG_SURFSCAT   (1u << 17)

#define PTB_MAX_ITERS 256
#define PTB_CTR_STRIDE 8
// counter slots per iteration
enum { CTR_NPATHS = 0, CTR_NSHA = 1, CTR_NSHB = 2, CTR_FETCH_TRACE = 3, CTR_FETCH_SHADE = 4, CTR_FETCH_SHA = 5, CTR_FETCH_SHB = 6 };

struct DevStats { unsigned long long pathSegments, shadowRays; };

// ---- launchers implemented in ptb_kernels.cu ------------------------------------------------------------------
struct LaunchCfg
{
    int numSMs; void* stream;
    int traceBlocks;            // resident k_trace blocks per SM on this context's device (ptbk_configure_device)
    int shadowBlocks;           // same for k_shadow (its stack may be deeper: 4-wide hierarchy)
    int shadeBlocks[4];         // same for the k_shade specialisations (0, 1, 2 = inline shadow rays, 3 = media / alpha with deferred shadow rays)
    int transmitBlocks;         // same for k_transmit
    unsigned long long* launches;   // per-context count of kernel launches (may be null)
};

void ptbk_camera(const LaunchCfg&, const DevScene&, const FrameParams&, const WaveParams&, const PathState&, uint32_t* ctr0);
void ptbk_trace(const LaunchCfg&, const DevScene&, const FrameParams&, const PathState&, const uint32_t* queue,
                const uint32_t* countPtr, uint32_t* fetchCtr, int depthForLights, DevStats* stats, uint32_t* keys, uint32_t* hist, uint32_t nOverride = 0, uint32_t holeKey = 0, uint32_t tflags = 0);
void ptbk_trace_primary(const LaunchCfg&, const DevScene&, const FrameParams&, const WaveParams&, const PathState&, uint32_t* ctr0, int depthForLights, DevStats* stats,
                        uint32_t* keys, uint32_t* hist, uint32_t liveCount, uint32_t holeKey, uint32_t tflags = 0);
void ptbk_sort(const LaunchCfg&, const uint32_t* queue, const uint32_t* keys, const uint32_t* countPtr, uint32_t* hist, uint32_t* cursor, int numKeys,
               uint32_t* sorted);
void ptbk_sort_tile_local(const LaunchCfg&, const uint32_t* queue, const uint32_t* keys, const uint32_t* countPtr, int numKeys, uint32_t* sorted, int holeKey = -1, uint32_t nOverride = 0);
void ptbk_shade(const LaunchCfg&, const DevScene&, const FrameParams&, const PathState&, const uint32_t* queue,
                uint32_t* ctrThis, uint32_t* ctrNext, uint32_t* nextQueue, DevStats* stats, int firstIter, uint32_t* slotKeys = nullptr, uint32_t nOverride = 0,
                uint32_t flags = 0,      // flags: SHADE_* of ptb_kernels.cu (1 identity queue, 2 static chunks, 4 count the continuing paths only, 8 octahedral direction classes)
                uint32_t* shadowKeysA = nullptr, uint32_t* shadowKeysB = nullptr, int lightKeys = 0);   // slot-ordered shadow queues (SlotShadow in ptb_kernels.cu)
void ptbk_shadow(const LaunchCfg&, const DevScene&, const FrameParams&, const PathState&, int which, const uint32_t* countPtr,
                 uint32_t* fetchCtr, DevStats* stats, const uint32_t* idxQueue = nullptr, uint32_t nOverride = 0);
void ptbk_transmit(const LaunchCfg&, const DevScene&, const FrameParams&, const PathState&, int which, const uint32_t* countPtr,
                   uint32_t* fetchCtr, DevStats* stats);
void ptbk_accumulate(const LaunchCfg&, const FrameParams&, const WaveParams&, const PathState&, float4* accum, float4* previewOut);
void ptbk_tonemap(const LaunchCfg&, const float4* accum, int w, int h, float invSampleCounter, int enableTonemap, int enableAces,
                  int simpleAcesFit, const float* backgroundCol3, uint32_t features, uchar4* out, float4* outF = nullptr);
void ptbk_trace_closest_batch(const LaunchCfg&, const DevScene&, const FrameParams&, const float* rays, long long n, int depth, void* hitsOut);
void ptbk_trace_any_batch(const LaunchCfg&, const DevScene&, const FrameParams&, const float* rays, const float* maxDist, long long n, int* out);
void ptbk_bsdf_batch(const LaunchCfg&, const void* queries, long long n, void* results, int sample);
void ptbk_camera_rays(const LaunchCfg&, const FrameParams&, const WaveParams&, float* outRays);
int  ptbk_tlas_build(const LaunchCfg&, float* nodes, int top, const float4* transforms, int n, const int* blasRoot, const int* materialID, float* instBounds, float* cent,
                     int* nodeOf, void* recA, void* recB, void* recC, int* remap, int* result);     // rec buffers: (n + 2) records of ptbk_tlas_rec_size() bytes; remap: n + 2 ints
int  ptbk_tlas_rec_size();
int  ptbk_configure_device(const DevScene&, int* traceBlocks, int* shadowBlocks, int shadeBlocks[4], int* transmitBlocks);   // per-device attributes; returns a cudaError_t value
