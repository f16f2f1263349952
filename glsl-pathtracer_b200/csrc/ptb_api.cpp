// ptb_api.cpp — host side of libptb200.so: the C ABI of include/ptb200.h.
//
// Mirrors the host half of the reference's Renderer (reference src/core/Renderer.cpp): scene upload
// (InitGPUDataBuffers :135-249), buffers (InitFBOs :281-379), feature selection (InitShaders :392-459), per-frame
// uniforms and the wave loop that replaces the three GL draws of Render() (:546-590).
// Built with -ffp-contract=off: the derived per-instance inverses and light planes must be bit-identical to what the
// shader-side formulas give under IEEE arithmetic.
#include "ptb200.h"
#include "ptb_internal.h"
#include "ptb_derive.h"
#include <cuda_runtime_api.h>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>
#include <algorithm>
#include <chrono>

static thread_local std::string g_err;
extern "C" const char* ptb_last_error(void) { return g_err.c_str(); }
extern "C" __attribute__((visibility("hidden"))) void ptb_set_last_error_(const char* msg) { g_err = msg ? msg : ""; }      // for ptb_mgpu.cpp

#define CK(call)                                                                                           \
    do {                                                                                                   \
        cudaError_t e_ = (call);                                                                           \
        if (e_ != cudaSuccess) {                                                                           \
            g_err = std::string(#call) + ": " + cudaGetErrorString(e_);                                    \
            return e_ == cudaErrorMemoryAllocation ? PTB_ERR_OUT_OF_MEMORY : PTB_ERR_CUDA;                 \
        }                                                                                                  \
    } while (0)
#define REQUIRE(cond, code, msg) do { if (!(cond)) { g_err = msg; return code; } } while (0)

namespace {

template <class T> struct DevBuf
{
    T* p = nullptr; size_t n = 0;
    cudaError_t alloc(size_t count)
    {
        if (count <= n && p) return cudaSuccess;
        release();
        if (count == 0) return cudaSuccess;
        cudaError_t e = cudaMalloc((void**)&p, count * sizeof(T));
        if (e == cudaSuccess) n = count;
        return e;
    }
    cudaError_t upload(const T* src, size_t count, cudaStream_t s)
    {
        cudaError_t e = alloc(count);
        if (e != cudaSuccess || count == 0) return e;
        return cudaMemcpyAsync(p, src, count * sizeof(T), cudaMemcpyHostToDevice, s);
    }
    void release() { if (p) cudaFree(p); p = nullptr; n = 0; }
    DevBuf() = default;
    DevBuf(const DevBuf&) = delete;
    DevBuf& operator=(const DevBuf&) = delete;
    ~DevBuf() { release(); }          // temporaries in the parity entry points are freed on every return path
};

inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

} // namespace

struct PtbCtx
{
    int device = 0, numSMs = 148;
    cudaStream_t ownStream = nullptr, stream = nullptr;
    cudaEvent_t evStart = nullptr, evStop = nullptr;
    // profiling: one event between consecutive launches of a render call, each interval attributed to the kernel class that runs in it
    std::vector<cudaEvent_t> traceEvents; std::vector<int> eventKind; size_t traceEventsUsed = 0; bool profiling = false;

    // host copies needed for re-derivation
    std::vector<float> hNodes, hTransforms, hMaterials;
    int numNodes = 0, topLevelIndex = 0;

    DevBuf<float> nodes, lights, envImg, envCdf;
    DevBuf<uint32_t> envGuide;   // guide table of the env-map CDF search (ptbd_build_env_guide)
    int envGuideOn = 1;          // PTB_ENV_GUIDE
    DevBuf<int> vertIndices;
    DevBuf<float4> triShade, wide;
    DevBuf<uint32_t> lightGrid;
    // device-side TLAS rebuild (ptb_rebuild_instances): per-instance BLAS root / material id, scratch of the builder
    DevBuf<int> tlasBlasRoot, tlasMatID, tlasNodeOf, tlasRemap, tlasResult; DevBuf<float> tlasBounds, tlasCent; DevBuf<char> tlasRec[3];
    int lastRebuildWhere = -1;   // 0 device, 1 host (asked for), 2 host (fallback of the device builder)
    float lastRebuildBuildMs = 0.f, lastRebuildTotalMs = 0.f;    // TLAS build alone (kernel: CUDA events; host builder: wall clock) / the whole call
    int wideAny = 1;           // 1: shadow rays use the 4-wide any-hit hierarchy where it is provably equivalent (PTB_WIDE_ANY)
    DevBuf<float4> verticesUVX, normalsUVY, materials, transforms, inner, tris, instTrav, instShade, lightsPre, lightGroups;
    DevBuf<uchar4> textures;
    DevScene S{};

    PtbOptions opts{};
    PtbCamera cam{};
    FrameParams F{};

    DevBuf<float4> accum, preview;
    DevBuf<float2> pixTabX, pixTabY; int tabW = 0, tabH = 0, tabTW = 0, tabTH = 0;     // pixel tables of the current resolution / tile size
    DevBuf<uchar4> out8;

    // wave state
    size_t slotCap = 0; bool stateGeneral = false;
    DevBuf<float4> state, shO[2], shD[2], shC[2];     // state: all per-path fields, interleaved (AoS) or as consecutive arrays (SoA)
    int aos = 2; size_t stateStrideF4 = 0;   // path-state layout (PTB_AOS): 0 SoA of 16-byte fields, 1 one record per path (-7 %: the coherent first bounce loses), 2 SoA with (rayO, rayD) and (thr, rng) as 32-byte pairs — one full DRAM sector per scattered access of the sorted bounces (+0.4 %), 3-5 other pairings (+-0)
    DevBuf<uint32_t> queue[2], counters, sortKeys, sortedQueue, sortHist, slotKeys, slotSorted;
    int sortFrom = 1;          // first bounce whose shade queue is material-sorted (PTB_SORT_FROM)
    int traceFinish = 1;       // misses / light hits of the specialisations without media are finished inside the trace kernel (PTB_TRACE_FINISH)
    int streamShade = 5;       // SHADE_* flags allowed for the first shade pass (PTB_STREAM_SHADE): 1 identity queue, 2 static chunks, 4 count-only
    int deferTransmit = 1;     // 1: EvalTransmittance rays of scenes without BLEND materials are queued for k_transmit instead of traced inside k_shade (PTB_DEFER_TRANSMIT)
    int blockMajor = 1;        // 1: the 32-slot groups of a wave are ordered block-major (WaveParams::blockMajor, PTB_BLOCK_MAJOR)
    int dirBins = 16;          // direction classes of the slot-ordered bounce 1: 16 / 8 = 16x16 / 8x8 octahedral cells, 0 = dominant axis + sign (PTB_DIR_BINS)
    int slotShadow = 1;        // 1: the NEE rays of the first shade pass are queued by path slot and grouped by light / direction inside tiles (SlotShadow, PTB_SLOT_SHADOW)
    int warpSamplesLog2 = 5;   // upper bound of WaveParams::lps (PTB_WARP_SAMPLES_LOG2): 0 = a warp's primary rays are one 8x4 pixel block of one pass
    int fuseCamera = 1;        // 1: camera rays are generated inside the first closest-hit launch (PTB_FUSE_CAMERA)
    int slotOrder = 1;         // 1: bounce 1 runs over the path slots in screen order (holes for ended paths), grouped by direction class inside
                               //    tiles of 2048 slots, instead of over the compacted arrival-order queue (PTB_SLOT_ORDER)
    int sortMode = 3;          // 0 off, 1 global sort of bounces >= 1, 2 global sort of every bounce, 3 tile-local sort of bounces >= 1 (default: +3 % over 1)
    DevBuf<DevStats> dstats;
    uint32_t* hCount = nullptr;   // pinned

    // ptb_render_pass: a wave of several consecutive sample passes is traced at the first of them; the later ones are added to the
    // running sum when the host asks for them, provided nothing the passes depend on has changed in between
    struct Pending { bool valid = false; int first = 0, n = 0, consumed = 0, stride = 1; uint64_t sceneVersion = 0; FrameParams F{}; WaveParams W{}; } pending;
    uint64_t sceneVersion = 0;                // bumped by every delta upload
    DevBuf<uchar4> snapshot;                  // ptb_snapshot_output: tonemapped image kept on the device until it is asked for
    DevBuf<float4> snapshotF; bool snapshotFloat = false;   // + its RGBA32F form when a denoiser hook wants it (ptb_set_snapshot_float)

    uint64_t samplesRendered = 0;
    unsigned long long launches = 0;          // kernels launched through this context
    int traceBlocks = 1, shadowBlocks = 1, shadeBlocks[4] = {1, 1, 1, 1}, transmitBlocks = 1, configuredDepth = -1, configuredDepthAny = -1;   // per-device launch configuration (ptbk_configure_device)
    uint64_t lastTraceRays = 0;
    bool timingValid = false;
};

namespace {

LaunchCfg cfg(PtbCtx* c)
{
    return LaunchCfg{c->numSMs, (void*)c->stream, c->traceBlocks, c->shadowBlocks, {c->shadeBlocks[0], c->shadeBlocks[1], c->shadeBlocks[2], c->shadeBlocks[3]}, c->transmitBlocks, &c->launches};
}

// kernel attributes / occupancy of this context's device for the current stack depth (the device must be current)
int configureDevice(PtbCtx* c)
{
    if (c->configuredDepth == c->S.stackDepth && c->configuredDepthAny == c->S.stackDepthAny) return PTB_OK;
    cudaError_t e = (cudaError_t)ptbk_configure_device(c->S, &c->traceBlocks, &c->shadowBlocks, c->shadeBlocks, &c->transmitBlocks);
    if (e != cudaSuccess) { g_err = std::string("ptbk_configure_device: ") + cudaGetErrorString(e); return PTB_ERR_CUDA; }
    c->configuredDepth = c->S.stackDepth; c->configuredDepthAny = c->S.stackDepthAny;
    return PTB_OK;
}

// (Re)derive the packed inner nodes for canonical node range [begin,end), plus instTrav/instShade and the stack bound (ptb_derive.cpp), and upload.
int deriveHierarchy(PtbCtx* c, int begin, int end, bool all)
{
    PtbDerivedHierarchy dh; std::string err;
    int rc = ptbd_derive_hierarchy(c->hNodes.data(), c->numNodes, c->topLevelIndex, c->S.numIndices, c->S.numMaterials, c->hTransforms.data(),
                                   (int)(c->hTransforms.size() / 16), begin, end, dh, err);
    REQUIRE(rc == 0, rc, err);
    // 4-wide hierarchy for the any-hit rays (rebuilt as a whole: a few ms even at 10^5 nodes); its BLAS roots travel in instTrav[k][2].w
    PtbDerivedWide dw;
    if (c->wideAny) ptbd_build_wide(c->hNodes.data(), c->numNodes, c->topLevelIndex, c->S.numIndices, (int)(c->hTransforms.size() / 16), dh.transOnly, dw);
    if (dw.ok)
    {
        for (size_t k = 0; k < dw.instRootMeta.size(); k++) dh.instTrav[k * 4 + 2].w = u2f(dw.instRootMeta[k]);
        CK(cudaStreamSynchronize(c->stream));       // a wave in flight may still read the old array
        CK(c->wide.upload(dw.wide.data(), dw.wide.size(), c->stream));
    }
    CK(c->inner.alloc((size_t)c->numNodes * 4));
    if (end > begin) CK(cudaMemcpyAsync(c->inner.p + (size_t)begin * 4, dh.inner.data(), dh.inner.size() * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
    CK(c->instTrav.upload(dh.instTrav.data(), dh.instTrav.size(), c->stream));
    CK(c->instShade.upload(dh.instShade.data(), dh.instShade.size(), c->stream));
    CK(cudaStreamSynchronize(c->stream));   // staging vectors die at scope exit
    c->S.stackDepth = dh.stackDepth; c->S.stackDepthAny = dw.ok ? std::max(dh.stackDepth, dw.stackDepth) : dh.stackDepth; c->S.rootMeta = dh.rootMeta;
    if (const char* e = getenv("PTB_ANY_STACK_MIN")) c->S.stackDepthAny = std::max(c->S.stackDepthAny, atoi(e));      // measurement aid: shared-memory footprint of k_shadow
    c->S.wide = dw.ok ? c->wide.p : nullptr; c->S.rootMetaWide = dw.rootMeta;
    c->S.inner = c->inner.p; c->S.instTrav = c->instTrav.p; c->S.instShade = c->instShade.p;
    (void)all;
    return configureDevice(c);
}

void refreshDerivedFlags(PtbCtx* c)
{
    const uint32_t f = c->opts.features;
    bool anyBlend = false, anyEmission = false;
    const int nm = (int)(c->hMaterials.size() / 32);
    for (int i = 0; i < nm; i++)
    {
        const float* m = &c->hMaterials[(size_t)i * 32];
        if ((int)m[29] == 1) anyBlend = true;
        if (m[4] != 0.f || m[5] != 0.f || m[6] != 0.f || m[27] >= 0.f) anyEmission = true;
    }
    FrameParams& F = c->F;
    // shade specialisation: 0 = lights only, 1 = + env map / textures / emission / mollification, 2 = + media / alpha test / inline shadow rays
    F.general = (f & (PTB_OPT_MEDIUM | PTB_OPT_ALPHA_TEST)) ? 2
              : (((f & (PTB_OPT_ENVMAP | PTB_OPT_ROUGHNESS_MOLLIFICATION)) != 0u) || c->S.numTextures > 0 || anyEmission) ? 1 : 0;
    // Shadow rays that draw from the path RNG cannot be deferred without changing the reference's random-number order: AnyHit's BLEND alpha test
    // (anyhit.glsl:118-141) and EvalTransmittance's (pathtrace.glsl:136).  Without a BLEND material neither draws anything.
    const bool volMis = (f & PTB_OPT_MEDIUM) && (f & PTB_OPT_VOL_MIS);
    const bool deferT = volMis && !anyBlend && c->deferTransmit;
    F.inlineShadow = (((f & PTB_OPT_ALPHA_TEST) && !(f & PTB_OPT_MEDIUM) && anyBlend) || (volMis && !deferT)) ? 1 : 0;
    F.deferTransmit = deferT ? 1 : 0;
}

// (re)build the per-column / per-row pixel tables when the resolution or the tile size changed
int refreshPixelTables(PtbCtx* c)
{
    const PtbOptions& o = c->opts;
    if (c->pixTabX.p && c->tabW == o.renderW && c->tabH == o.renderH && c->tabTW == o.tileW && c->tabTH == o.tileH) return PTB_OK;
    std::vector<float2> tx, ty; std::string err;
    if (ptbd_build_pixel_tables(o.renderW, o.renderH, o.tileW, o.tileH, tx, ty, err) != 0)
    {   // sizes beyond the table encoding: the kernels evaluate the mapping per pixel
        c->pixTabX.release(); c->pixTabY.release(); c->tabW = 0;
        return PTB_OK;
    }
    CK(cudaStreamSynchronize(c->stream));           // a wave in flight may still read the old tables
    CK(c->pixTabX.upload(tx.data(), tx.size(), c->stream));
    CK(c->pixTabY.upload(ty.data(), ty.size(), c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->tabW = o.renderW; c->tabH = o.renderH; c->tabTW = o.tileW; c->tabTH = o.tileH;
    return PTB_OK;
}

void refreshFrameParams(PtbCtx* c)
{
    FrameParams& F = c->F; const PtbOptions& o = c->opts; const PtbCamera& cam = c->cam;
    F.pixTabX = c->pixTabX.p; F.pixTabY = c->pixTabY.p;
    F.features = o.features;
    if (!(c->S.envImg && c->S.envW > 0)) F.features &= ~(uint32_t)PTB_OPT_ENVMAP;       // Renderer.cpp:404 needs scene->envMap
    if (c->S.numLights == 0) F.features &= ~(uint32_t)PTB_OPT_LIGHTS;
    F.maxDepth = o.maxDepth; F.rrDepth = o.rrDepth;
    F.envMapIntensity = o.envMapIntensity; F.envMapRot = o.envMapRot / 360.0f; F.roughnessMollificationAmt = o.roughnessMollificationAmt;
    memcpy(F.uniformLightCol, o.uniformLightCol, 12);
    F.renderW = o.renderW; F.renderH = o.renderH; F.tileW = o.tileW; F.tileH = o.tileH;
    F.invNumTilesX = (float)o.tileW / o.renderW; F.invNumTilesY = (float)o.tileH / o.renderH;           // Renderer.cpp:293-294
    F.numTilesX = (int)ceilf((float)o.renderW / o.tileW); F.numTilesY = (int)ceilf((float)o.renderH / o.tileH);   // :296-297
    memcpy(F.camPos, cam.position, 12); memcpy(F.camRight, cam.right, 12); memcpy(F.camUp, cam.up, 12); memcpy(F.camFwd, cam.forward, 12);
    F.aspect = (float)o.renderH / (float)o.renderW;
    F.camScale = tanf(cam.fov * 0.5f); F.camFocalDist = cam.focalDist; F.camAperture = cam.aperture;
    refreshDerivedFlags(c);
}

// guide table of the env-map CDF search; absent (null) when the CDF is not monotone: the kernels then run the reference's two binary searches
int buildEnvGuide(PtbCtx* c, const float* cdf, int w, int h, float totalSum)
{
    std::vector<uint32_t> g; float scale = 0.f;
    c->S.envGuide = nullptr; c->S.envGuideN = 0; c->S.envGuideScale = 0.f;
    if (!c->envGuideOn || !cdf || w <= 0 || ptbd_build_env_guide(cdf, w, h, totalSum, g, scale) != 0) return PTB_OK;
    CK(c->envGuide.upload(g.data(), g.size(), c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->S.envGuide = c->envGuide.p; c->S.envGuideN = (int)g.size() - 1; c->S.envGuideScale = scale;
    return PTB_OK;
}

int buildLightsPre(PtbCtx* c, const float* lights, int n)
{
    PtbDerivedLights dl;
    ptbd_build_lights(lights, n, dl);
    CK(c->lightsPre.upload(dl.lightsPre.data(), dl.lightsPre.size(), c->stream));
    CK(c->lightGroups.upload(dl.lightGroups.data(), dl.lightGroups.size(), c->stream));
    CK(c->lightGrid.upload(dl.lightGrid.data(), dl.lightGrid.size(), c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->S.lightsPre = c->lightsPre.p; c->S.lightGroups = c->lightGroups.p; c->S.lightGrid = c->lightGrid.p; c->S.numLightGroups = dl.numGroups;
    return PTB_OK;
}

int buildTris(PtbCtx* c, const PtbSceneDesc* d)
{
    std::vector<float4> t; std::string err;
    int rc = ptbd_build_tris(d->vertIndices, d->numIndices, d->verticesUVX, d->numVertices, t, err);
    REQUIRE(rc == 0, rc, err);
    CK(c->tris.upload(t.data(), t.size(), c->stream));
    std::vector<float4> ts;
    ptbd_build_tri_shade(d->vertIndices, d->numIndices, d->verticesUVX, d->normalsUVY, ts);
    CK(c->triShade.upload(ts.data(), ts.size(), c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->S.tris = c->tris.p; c->S.triShade = c->triShade.p;
    return PTB_OK;
}

int ensureWaveState(PtbCtx* c, size_t slots)
{
    const bool gen = c->F.general != 0;
    if (slots <= c->slotCap && (!gen || c->stateGeneral)) return PTB_OK;
    size_t n = std::max(slots, c->slotCap);
    c->stateStrideF4 = (gen || c->stateGeneral) ? 12 : 8;            // float4 per path: 7 used (+3 general), padded to 128 / 192 bytes
    CK(c->state.alloc(n * c->stateStrideF4));
    CK(c->queue[0].alloc(n)); CK(c->queue[1].alloc(n));
    CK(c->shO[1].alloc(n)); CK(c->shD[1].alloc(n)); CK(c->shC[1].alloc(n));
    if (gen || c->stateGeneral)
    {
        CK(c->shO[0].alloc(n)); CK(c->shD[0].alloc(n)); CK(c->shC[0].alloc(n));
        c->stateGeneral = true;
    }
    CK(c->counters.alloc((size_t)(PTB_MAX_ITERS + 2) * PTB_CTR_STRIDE));
    CK(c->sortKeys.alloc(n)); CK(c->sortedQueue.alloc(n)); CK(c->slotKeys.alloc(n)); CK(c->slotSorted.alloc(n));
    c->slotCap = n;
    return PTB_OK;
}

PathState pathState(PtbCtx* c)
{
    PathState P{};
    char* b = (char*)c->state.p;
    const size_t n = c->slotCap;
    if (c->aos == 1)
    {   // field k of path i at b + i*stride + 16*k
        const uint32_t st = (uint32_t)(c->stateStrideF4 * 16);
        P.rayO = {b, st}; P.rayD = {b + 16, st}; P.thr = {b + 32, st}; P.rad = {b + 48, st}; P.rng = {b + 64, st}; P.hit = {b + 80, st};
        P.hitInst = {b + 96, st}; P.med = {b + 112, st}; P.medCol = {b + 128, st}; P.prevUV = {b + 144, st};
    }
    else if (c->aos == 4)
    {   // pairs (rayO, rayD), (thr, rng), (rad, hit)
        auto arr = [&](int k) { return b + (size_t)k * n * 16; };
        P.rayO = {arr(0), 32}; P.rayD = {arr(0) + 16, 32}; P.thr = {arr(2), 32}; P.rng = {arr(2) + 16, 32}; P.rad = {arr(4), 32}; P.hit = {arr(4) + 16, 32};
        P.hitInst = {arr(6), 4}; P.med = {arr(7), 16}; P.medCol = {arr(8), 16}; P.prevUV = {arr(9), 8};
    }
    else if (c->aos == 5)
    {   // 64-byte records (rayO, rayD, thr, rng)
        auto arr = [&](int k) { return b + (size_t)k * n * 16; };
        P.rayO = {arr(0), 64}; P.rayD = {arr(0) + 16, 64}; P.thr = {arr(0) + 32, 64}; P.rng = {arr(0) + 48, 64}; P.rad = {arr(4), 16}; P.hit = {arr(5), 16};
        P.hitInst = {arr(6), 4}; P.med = {arr(7), 16}; P.medCol = {arr(8), 16}; P.prevUV = {arr(9), 8};
    }
    else if (c->aos == 3)
    {   // only the ray as a 32-byte pair
        auto arr = [&](int k) { return b + (size_t)k * n * 16; };
        P.rayO = {arr(0), 32}; P.rayD = {arr(0) + 16, 32}; P.thr = {arr(2), 16}; P.rad = {arr(3), 16}; P.rng = {arr(4), 16}; P.hit = {arr(5), 16};
        P.hitInst = {arr(6), 4}; P.med = {arr(7), 16}; P.medCol = {arr(8), 16}; P.prevUV = {arr(9), 8};
    }
    else if (c->aos == 2)
    {   // SoA of 32-byte pairs (one DRAM sector per scattered access): (rayO, rayD), (thr, rng); the rest as in SoA
        auto arr = [&](int k) { return b + (size_t)k * n * 16; };
        P.rayO = {arr(0), 32}; P.rayD = {arr(0) + 16, 32}; P.thr = {arr(2), 32}; P.rng = {arr(2) + 16, 32}; P.rad = {arr(4), 16}; P.hit = {arr(5), 16};
        P.hitInst = {arr(6), 4}; P.med = {arr(7), 16}; P.medCol = {arr(8), 16}; P.prevUV = {arr(9), 8};
    }
    else
    {   // SoA: array k at b + k*n*16
        auto arr = [&](int k) { return b + (size_t)k * n * 16; };
        P.rayO = {arr(0), 16}; P.rayD = {arr(1), 16}; P.thr = {arr(2), 16}; P.rad = {arr(3), 16}; P.rng = {arr(4), 16}; P.hit = {arr(5), 16};
        P.hitInst = {arr(6), 4}; P.med = {arr(7), 16}; P.medCol = {arr(8), 16}; P.prevUV = {arr(9), 8};
    }
    for (int k = 0; k < 2; k++) { P.shO[k] = c->shO[k].p; P.shD[k] = c->shD[k].p; P.shC[k] = c->shC[k].p; P.queue[k] = c->queue[k].p; }
    return P;
}

enum { KIND_NONE = -1, KIND_CAMERA = 0, KIND_TRACE = 1, KIND_SORT = 2, KIND_SHADE = 3, KIND_SHADOW = 4, KIND_ACCUM = 5, KIND_COUNT = 6 };

// profiling only: an event in front of the launch(es) of class `kind`; the time up to the next mark is attributed to that class
void mark(PtbCtx* c, int kind)
{
    if (!c->profiling) return;
    if (c->traceEventsUsed == c->traceEvents.size())
    {
        cudaEvent_t e; cudaEventCreate(&e); c->traceEvents.push_back(e); c->eventKind.push_back(KIND_NONE);
    }
    c->eventKind[c->traceEventsUsed] = kind;
    cudaEventRecord(c->traceEvents[c->traceEventsUsed++], c->stream);
}

// One wavefront: camera -> (trace, shade, shadow)* -> accumulate.
int renderWave(PtbCtx* c, const FrameParams& F, WaveParams& W, float4* previewOut)
{
    c->pending.valid = false;                 // the wave state is about to be overwritten
    W.vw = (W.rw + 7) & ~7; W.vh = (W.rh + 3) & ~3;
    W.nSlots = (uint32_t)((size_t)W.vw * W.vh * W.nSamples);
    W.blockMajor = (c->blockMajor && W.nSamples > 1) ? 1 : 0;
    W.lps = 0; W.lpw = 3;
    if (W.blockMajor)
    {   // passes per warp group: the largest power of two that divides the pass count, up to 32 (one pixel, 32 passes)
        // (a look-ahead wave of ptb_render_pass is added to the running sum ONE pass per call: groups of one pass keep those 256 reads per wave contiguous —
        // the drop-in loop runs 1112 spp/s with lps = 0 against 1076 with lps = 5, although the wave itself traces 1.2 % slower)
        ptbd_wave_groups(W.nSamples, W.accCount > 0 ? 0 : c->warpSamplesLog2, &W.lps, &W.lpw);
    }
    int rc = ensureWaveState(c, W.nSlots);
    if (rc) return rc;
    PathState P = pathState(c);
    LaunchCfg L = cfg(c);
    uint32_t* ctr = c->counters.p;
    CK(cudaMemsetAsync(ctr, 0, (size_t)(PTB_MAX_ITERS + 2) * PTB_CTR_STRIDE * sizeof(uint32_t), c->stream));
    const int numKeys = c->S.numMaterials + 2;
    CK(c->sortHist.alloc((size_t)numKeys * 2));
    CK(cudaMemsetAsync(c->sortHist.p, 0, (size_t)numKeys * 2 * sizeof(uint32_t), c->stream));
    // the camera rays of a render wave are generated inside the first closest-hit launch (k_trace_primary); the preview target keeps k_camera
    const bool fusedCamera = c->fuseCamera && !W.previewMode && (W.nSlots & 31u) == 0;
    if (!fusedCamera)
    {
        mark(c, KIND_CAMERA);
        ptbk_camera(L, c->S, F, W, P, ctr);
    }
    const int lightsFromDepth = (F.features & PTB_OPT_HIDE_EMITTERS) ? 1 : 0;
    const bool alphaScene = (F.features & PTB_OPT_ALPHA_TEST) != 0u;
    const int nominal = F.maxDepth + 1;
    // slot-ordered bounce 1 needs the tile-local material sorter (it drops the holes for k_shade) and is kept to scenes without alpha re-traces
    const bool useSlotOrder = c->slotOrder && c->sortMode == 3 && numKeys + 1 <= 4096 && !alphaScene && F.maxDepth >= 1 && !W.previewMode;
    const bool octKeys = c->dirBins == 8 || c->dirBins == 16;
    const int slotHole = c->dirBins == 16 ? 256 : (octKeys ? 64 : 7);
    if (useSlotOrder) CK(cudaMemsetAsync(c->slotKeys.p, 0x7f, (size_t)W.nSlots * sizeof(uint32_t), c->stream));   // 0x7f7f7f7f clamps to the hole key = ended / never live
    // scenes without media / alpha: paths ending in a miss or on a light are finished by the trace kernel and leave the queues as holes (needs the tile-local sorter)
    const bool finishInTrace = c->traceFinish && F.general <= 1 && c->sortMode == 3 && numKeys + 1 <= 4096;
    int it = 0;
    while (true)
    {
        uint32_t* ci = ctr + (size_t)it * PTB_CTR_STRIDE;
        uint32_t* cn = ctr + (size_t)(it + 1) * PTB_CTR_STRIDE;
        const bool sortThis = c->sortMode == 2 || ((c->sortMode == 1 || c->sortMode == 3) && it >= c->sortFrom);
        // bounce 1 over the slots in screen order (95 % of them still alive there): see slotOrder
        const bool slotIter = useSlotOrder && it == 1;
        const uint32_t nOv = (slotIter || (it == 0 && fusedCamera)) ? W.nSlots : 0u;      // queue length = slot count (holes inside)
        const uint32_t* traceQueue = P.queue[it & 1];
        if (slotIter)
        {
            mark(c, KIND_SORT);
            ptbk_sort_tile_local(L, nullptr, c->slotKeys.p, ci + CTR_NPATHS, slotHole + 1, c->slotSorted.p, slotHole, W.nSlots);
            traceQueue = c->slotSorted.p;
        }
        uint32_t* globalHist = (c->sortMode == 3 && numKeys <= 4096) ? nullptr : c->sortHist.p;     // key histogram over the whole queue: global sorter only
        mark(c, KIND_TRACE);
        if (it == 0 && fusedCamera)
            ptbk_trace_primary(L, c->S, F, W, P, ctr, lightsFromDepth, c->dstats.p, sortThis ? c->sortKeys.p : nullptr, globalHist,
                               (uint32_t)((size_t)W.rw * W.rh * W.nSamples), (uint32_t)numKeys, finishInTrace ? (1u | (sortThis ? 0u : 4u)) : 0u);
        else
            ptbk_trace(L, c->S, F, P, traceQueue, ci + CTR_NPATHS, ci + CTR_FETCH_TRACE, lightsFromDepth, c->dstats.p,
                       sortThis ? c->sortKeys.p : nullptr, globalHist, nOv, (uint32_t)numKeys, finishInTrace ? (1u | (it == 0 ? 2u : 0u) | (sortThis ? 0u : 4u)) : 0u);
        const uint32_t* shadeQueue = traceQueue;
        if (sortThis)
        {   // material-sorted shading: counting sort of the queue by (miss | light | material)
            mark(c, KIND_SORT);
            if (nOv) ptbk_sort_tile_local(L, traceQueue, c->sortKeys.p, ci + CTR_NPATHS, numKeys + 1, c->sortedQueue.p, numKeys, W.nSlots);     // queue with holes
            else if (finishInTrace) ptbk_sort_tile_local(L, traceQueue, c->sortKeys.p, ci + CTR_NPATHS, numKeys + 1, c->sortedQueue.p, numKeys, 0);   // finished paths = holes
            else if (c->sortMode == 3 && numKeys <= 4096) ptbk_sort_tile_local(L, traceQueue, c->sortKeys.p, ci + CTR_NPATHS, numKeys, c->sortedQueue.p);
            else ptbk_sort(L, traceQueue, c->sortKeys.p, ci + CTR_NPATHS, c->sortHist.p, c->sortHist.p + numKeys, numKeys, c->sortedQueue.p);
            shadeQueue = c->sortedQueue.p;
        }
        mark(c, KIND_SHADE);
        // first shade pass after a fused camera/trace launch: identity queue -> streaming mode (no queue read, no fetch atomic; with the slot-ordered
        // bounce 1 the continuing paths are only counted)
        uint32_t shadeFlags = 0;
        if (it == 0 && fusedCamera && !sortThis) shadeFlags = (uint32_t)c->streamShade & (1u | 2u | (useSlotOrder ? 4u : 0u));
        if (useSlotOrder && it == 0 && octKeys) shadeFlags |= (c->dirBins == 16 ? 16u : 8u);
        const bool queueA = F.general && (F.features & PTB_OPT_ENVMAP) && !(F.features & PTB_OPT_UNIFORM_LIGHT), queueB = (F.features & PTB_OPT_LIGHTS) != 0u;
        // first shade pass, block-major slot order: NEE rays are queued by path slot and grouped inside tiles by light (2..255 lights) / direction cell before
        // k_shadow runs (SlotShadow in ptb_kernels.cu).  The key arrays borrow buffers that are idle until the next bounce's sorts: sortKeys (A), slotSorted (B).
        const bool slotSh = it == 0 && useSlotOrder && W.blockMajor && c->slotShadow && !sortThis && !F.inlineShadow && !F.deferTransmit;
        // measured: the light queue of hyperion (17 quads) gains 1.2 ms of k_shadow for 0.35 ms of sorting (+3.5 %); the env-map queue of ibl_spheres
        // (direction cells) gains 0.6 ms of k_shadow but pays 0.7 ms in the sorter and the immediate-push shade kernel: slot order for the light queue only
        const bool slotA = slotSh && queueA && c->slotShadow >= 2, slotB = slotSh && queueB && c->S.numLights >= 2 && c->S.numLights <= 255;
        const bool lightKeys = true;
        if (slotA) CK(cudaMemsetAsync(c->sortKeys.p, 0x7f, (size_t)W.nSlots * sizeof(uint32_t), c->stream));       // 0x7f7f7f7f clamps to the hole key
        if (slotB) CK(cudaMemsetAsync(c->slotSorted.p, 0x7f, (size_t)W.nSlots * sizeof(uint32_t), c->stream));
        ptbk_shade(L, c->S, F, P, shadeQueue, ci, cn, P.queue[(it + 1) & 1], c->dstats.p, it == 0, (useSlotOrder && it == 0) ? c->slotKeys.p : nullptr, nOv, shadeFlags,
                   slotA ? c->sortKeys.p : nullptr, slotB ? c->slotSorted.p : nullptr, lightKeys ? 1 : 0);
        if (!F.inlineShadow)
        {
            // binary visibility (AnyHit) or, under OPT_MEDIUM + OPT_VOL_MIS, the transmittance along the ray (EvalTransmittance)
            if (queueA)
            {
                if (slotA)
                {
                    mark(c, KIND_SORT);
                    ptbk_sort_tile_local(L, nullptr, c->sortKeys.p, ci + CTR_NPATHS, 65, c->sortedQueue.p, 64, W.nSlots);
                }
                mark(c, KIND_SHADOW);
                if (F.deferTransmit) ptbk_transmit(L, c->S, F, P, 0, ci + CTR_NSHA, ci + CTR_FETCH_SHA, c->dstats.p);
                else ptbk_shadow(L, c->S, F, P, 0, ci + CTR_NSHA, ci + CTR_FETCH_SHA, c->dstats.p, slotA ? c->sortedQueue.p : nullptr, slotA ? W.nSlots : 0u);
            }
            if (queueB)
            {
                if (slotB)
                {
                    const int nk = lightKeys ? c->S.numLights + 1 : 65;
                    mark(c, KIND_SORT);
                    ptbk_sort_tile_local(L, nullptr, c->slotSorted.p, ci + CTR_NPATHS, nk, c->sortedQueue.p, nk - 1, W.nSlots);
                }
                mark(c, KIND_SHADOW);
                if (F.deferTransmit) ptbk_transmit(L, c->S, F, P, 1, ci + CTR_NSHB, ci + CTR_FETCH_SHB, c->dstats.p);
                else ptbk_shadow(L, c->S, F, P, 1, ci + CTR_NSHB, ci + CTR_FETCH_SHB, c->dstats.p, slotB ? c->sortedQueue.p : nullptr, slotB ? W.nSlots : 0u);
            }
        }
        it++;
        if (it >= PTB_MAX_ITERS) break;                       // alpha-skip re-traces are unbounded in the reference (Q7); hard stop
        if (it >= nominal)
        {
            if (!alphaScene) break;                           // depth == maxDepth terminates every path (pathtrace.glsl:367)
            CK(cudaMemcpyAsync(c->hCount, cn + CTR_NPATHS, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            if (*c->hCount == 0) break;
        }
    }
    mark(c, KIND_ACCUM);
    ptbk_accumulate(L, F, W, P, c->accum.p, previewOut);
    mark(c, KIND_NONE);
    CK(cudaGetLastError());
    return PTB_OK;
}

int allocFrameBuffers(PtbCtx* c)
{
    size_t n = (size_t)c->opts.renderW * c->opts.renderH;
    CK(c->accum.alloc(n)); CK(c->out8.alloc(n));
    CK(cudaMemsetAsync(c->accum.p, 0, n * sizeof(float4), c->stream));
    return PTB_OK;
}

} // namespace

static void beginTiming(PtbCtx* c) { c->traceEventsUsed = 0; cudaEventRecord(c->evStart, c->stream); }
static void endTiming(PtbCtx* c) { cudaEventRecord(c->evStop, c->stream); c->timingValid = true; }

extern "C" {

uint32_t ptb_derive_features(const PtbSceneDesc* s, uint32_t ob)
{
    // Renderer.cpp:401-459
    uint32_t m = 0;
    if (!s) return 0;
    if ((ob & 1u) && s->envImg && s->envW > 0) m |= PTB_OPT_ENVMAP;
    if (s->numLights > 0) m |= PTB_OPT_LIGHTS;
    if (ob & 2u) m |= PTB_OPT_RR;
    if (ob & 4u) m |= PTB_OPT_UNIFORM_LIGHT;
    if (ob & 8u) m |= PTB_OPT_OPENGL_NORMALMAP;
    if (ob & 16u) m |= PTB_OPT_HIDE_EMITTERS;
    if (ob & 32u) m |= PTB_OPT_BACKGROUND;
    if (ob & 64u) m |= PTB_OPT_TRANSPARENT_BACKGROUND;
    for (int i = 0; i < s->numMaterials; i++) if ((int)s->materials[(size_t)i * 32 + 29] != 0) { m |= PTB_OPT_ALPHA_TEST; break; }
    if (ob & 128u) m |= PTB_OPT_ROUGHNESS_MOLLIFICATION;
    for (int i = 0; i < s->numMaterials; i++) if ((int)s->materials[(size_t)i * 32 + 18] != 0) { m |= PTB_OPT_MEDIUM; break; }
    if (ob & 256u) m |= PTB_OPT_VOL_MIS;
    return m;
}

int ptb_create(const PtbSceneDesc* d, const PtbOptions* o, int device, PtbCtx** out)
{
    REQUIRE(d && o && out, PTB_ERR_INVALID_ARGUMENT, "No Scene Found");     // Renderer.cpp:72-76
    REQUIRE(d->nodes && d->numNodes > 0 && d->topLevelIndex >= 0 && d->topLevelIndex < d->numNodes, PTB_ERR_INVALID_ARGUMENT, "bad node array");
    REQUIRE(d->numInstances > 0 && d->transforms && d->materials && d->numMaterials > 0, PTB_ERR_INVALID_ARGUMENT, "bad instance/material arrays");
    REQUIRE(o->renderW > 0 && o->renderH > 0 && o->tileW > 0 && o->tileH > 0, PTB_ERR_INVALID_ARGUMENT, "bad resolution");
    REQUIRE((uint32_t)d->numNodes < (1u << 24), PTB_ERR_UNSUPPORTED, "node indices are stored as floats in the reference layout (exact below 2^24)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev)
    {
        cudaGetLastError();
        g_err = "no usable CUDA device (libptb200 has no CPU fallback)";
        return PTB_ERR_NO_DEVICE;
    }
    CK(cudaSetDevice(device));
    PtbCtx* c = new PtbCtx();
    c->device = device;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    c->numSMs = prop.multiProcessorCount;
    CK(cudaStreamCreateWithFlags(&c->ownStream, cudaStreamNonBlocking));
    c->stream = c->ownStream;
    CK(cudaEventCreate(&c->evStart)); CK(cudaEventCreate(&c->evStop));
    CK(cudaMallocHost((void**)&c->hCount, sizeof(uint32_t)));

    c->numNodes = d->numNodes; c->topLevelIndex = d->topLevelIndex;
    c->hNodes.assign(d->nodes, d->nodes + (size_t)d->numNodes * 9);
    c->hTransforms.assign(d->transforms, d->transforms + (size_t)d->numInstances * 16);
    c->hMaterials.assign(d->materials, d->materials + (size_t)d->numMaterials * 32);

    cudaStream_t s = c->stream;
    CK(c->nodes.upload(d->nodes, (size_t)d->numNodes * 9, s));
    CK(c->vertIndices.upload((const int*)d->vertIndices, (size_t)d->numIndices * 3, s));
    CK(c->verticesUVX.upload((const float4*)d->verticesUVX, (size_t)d->numVertices, s));
    CK(c->normalsUVY.upload((const float4*)d->normalsUVY, (size_t)d->numVertices, s));
    CK(c->materials.upload((const float4*)d->materials, (size_t)d->numMaterials * 8, s));
    CK(c->transforms.upload((const float4*)d->transforms, (size_t)d->numInstances * 4, s));
    if (d->numLights > 0) CK(c->lights.upload(d->lights, (size_t)d->numLights * 15, s));
    if (d->numTextures > 0) CK(c->textures.upload((const uchar4*)d->textures, (size_t)d->numTextures * d->texW * d->texH, s));
    if (d->envImg && d->envW > 0)
    {
        CK(c->envImg.upload(d->envImg, (size_t)d->envW * d->envH * 3, s));
        CK(c->envCdf.upload(d->envCdf, (size_t)d->envW * d->envH, s));
    }
    CK(cudaStreamSynchronize(s));
    DevScene& S = c->S;
    S.nodes = c->nodes.p; S.vertIndices = c->vertIndices.p; S.verticesUVX = c->verticesUVX.p; S.normalsUVY = c->normalsUVY.p;
    S.materials = c->materials.p; S.transforms = c->transforms.p; S.lights = c->lights.p; S.textures = c->textures.p;
    S.envImg = c->envImg.p; S.envCdf = c->envCdf.p;
    S.numNodes = d->numNodes; S.topLevelIndex = d->topLevelIndex; S.numIndices = d->numIndices; S.numVertices = d->numVertices;
    S.numMaterials = d->numMaterials; S.numInstances = d->numInstances; S.numLights = d->numLights;
    S.numTextures = d->numTextures; S.texW = d->texW; S.texH = d->texH;
    S.envW = d->envImg ? d->envW : 0; S.envH = d->envImg ? d->envH : 0; S.envTotalSum = d->envTotalSum;
    if (const char* e = getenv("PTB_ENV_GUIDE")) c->envGuideOn = atoi(e);
    if (d->envImg && d->envW > 0) { int grc = buildEnvGuide(c, d->envCdf, d->envW, d->envH, d->envTotalSum); if (grc) { ptb_destroy(c); return grc; } }

    if (const char* e = getenv("PTB_WIDE_ANY")) c->wideAny = atoi(e);
    int rc;
    if ((rc = buildTris(c, d)) != PTB_OK) { ptb_destroy(c); return rc; }
    if ((rc = buildLightsPre(c, d->lights, d->numLights)) != PTB_OK) { ptb_destroy(c); return rc; }
    if ((rc = deriveHierarchy(c, 0, c->numNodes, true)) != PTB_OK) { ptb_destroy(c); return rc; }

    c->opts = *o;
    // default camera: looking down -z from the origin (the caller sets the real one with ptb_set_camera)
    c->cam = PtbCamera{{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {0, 0, -1}, 1.0f, 1.0f, 0.0f};
    if ((rc = refreshPixelTables(c)) != PTB_OK) { ptb_destroy(c); return rc; }
    refreshFrameParams(c);
    c->F.cullBoxes = 1;      // t-culled traversal by default; ptb_set_cull(0) = the reference's unculled visit order, see ptb200.h
    if ((rc = allocFrameBuffers(c)) != PTB_OK) { ptb_destroy(c); return rc; }
    CK(c->dstats.alloc(1));
    CK(cudaMemsetAsync(c->dstats.p, 0, sizeof(DevStats), s));
    CK(cudaStreamSynchronize(s));
    if (const char* e = getenv("PTB_SORT")) c->sortMode = atoi(e);
    if (const char* e = getenv("PTB_AOS")) c->aos = atoi(e);
    if (const char* e = getenv("PTB_SLOT_ORDER")) c->slotOrder = atoi(e);
    if (const char* e = getenv("PTB_FUSE_CAMERA")) c->fuseCamera = atoi(e);
    if (const char* e = getenv("PTB_DEFER_TRANSMIT")) c->deferTransmit = atoi(e);
    if (const char* e = getenv("PTB_BLOCK_MAJOR")) c->blockMajor = atoi(e);
    if (const char* e = getenv("PTB_SLOT_SHADOW")) c->slotShadow = atoi(e);
    if (const char* e = getenv("PTB_WARP_SAMPLES_LOG2")) c->warpSamplesLog2 = atoi(e);
    if (const char* e = getenv("PTB_DIR_BINS")) c->dirBins = atoi(e);
    if (const char* e = getenv("PTB_STREAM_SHADE")) c->streamShade = atoi(e);
    if (const char* e = getenv("PTB_TRACE_FINISH")) c->traceFinish = atoi(e);
    if (const char* e = getenv("PTB_SORT_FROM")) c->sortFrom = atoi(e);
    refreshDerivedFlags(c);
    *out = c;
    return PTB_OK;
}

int ptb_destroy(PtbCtx* c)
{
    if (!c) return PTB_OK;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    c->nodes.release(); c->lights.release(); c->envImg.release(); c->envCdf.release(); c->envGuide.release(); c->vertIndices.release();
    c->verticesUVX.release(); c->normalsUVY.release(); c->materials.release(); c->transforms.release(); c->inner.release(); c->tris.release(); c->triShade.release(); c->wide.release(); c->tlasBlasRoot.release(); c->tlasMatID.release(); c->tlasNodeOf.release(); c->tlasRemap.release(); c->tlasResult.release(); c->tlasBounds.release(); c->tlasCent.release(); for (int k = 0; k < 3; k++) c->tlasRec[k].release();
    c->instTrav.release(); c->instShade.release(); c->lightsPre.release(); c->lightGroups.release(); c->lightGrid.release(); c->textures.release(); c->accum.release(); c->preview.release(); c->out8.release(); c->snapshot.release(); c->snapshotF.release(); c->pixTabX.release(); c->pixTabY.release();
    c->state.release();
    for (int k = 0; k < 2; k++) { c->shO[k].release(); c->shD[k].release(); c->shC[k].release(); c->queue[k].release(); }
    c->sortKeys.release(); c->sortedQueue.release(); c->sortHist.release(); c->slotKeys.release(); c->slotSorted.release();
    c->counters.release(); c->dstats.release();
    for (auto e : c->traceEvents) cudaEventDestroy(e);
    if (c->evStart) cudaEventDestroy(c->evStart);
    if (c->evStop) cudaEventDestroy(c->evStop);
    if (c->hCount) cudaFreeHost(c->hCount);
    if (c->ownStream) cudaStreamDestroy(c->ownStream);
    delete c;
    return PTB_OK;
}

int ptb_set_options(PtbCtx* c, const PtbOptions* o)
{
    REQUIRE(c && o, PTB_ERR_INVALID_ARGUMENT, "null argument");
    CK(cudaSetDevice(c->device));
    const bool resized = o->renderW != c->opts.renderW || o->renderH != c->opts.renderH;
    REQUIRE(o->renderW > 0 && o->renderH > 0 && o->tileW > 0 && o->tileH > 0, PTB_ERR_INVALID_ARGUMENT, "bad resolution");
    c->opts = *o;
    int cull = c->F.cullBoxes;
    { int rc = refreshPixelTables(c); if (rc) return rc; }
    refreshFrameParams(c);
    c->F.cullBoxes = cull;
    if (resized) { c->snapshot.release(); c->snapshotF.release(); c->pending.valid = false; return allocFrameBuffers(c); }
    return PTB_OK;
}

int ptb_resize(PtbCtx* c, int32_t w, int32_t h, int32_t tileW, int32_t tileH)
{
    REQUIRE(c, PTB_ERR_INVALID_ARGUMENT, "null context");
    PtbOptions o = c->opts; o.renderW = w; o.renderH = h; o.tileW = tileW; o.tileH = tileH;
    int rc = ptb_set_options(c, &o);
    if (rc) return rc;
    c->samplesRendered = 0;
    return ptb_reset_accum(c);
}

int ptb_set_camera(PtbCtx* c, const PtbCamera* cam)
{
    REQUIRE(c && cam, PTB_ERR_INVALID_ARGUMENT, "null argument");
    c->cam = *cam;
    int cull = c->F.cullBoxes;
    refreshFrameParams(c);
    c->F.cullBoxes = cull;
    return PTB_OK;
}

int ptb_update_instances(PtbCtx* c, const float* transforms, int32_t numInstances, const float* materials, int32_t numMaterials,
                         const float* tlasNodes, int32_t numTlasNodes)
{
    REQUIRE(c && transforms && materials && tlasNodes, PTB_ERR_INVALID_ARGUMENT, "null argument");
    REQUIRE(numInstances == c->S.numInstances, PTB_ERR_INVALID_ARGUMENT, "instance count changed (the reference re-uploads the same-sized arrays)");
    REQUIRE(numTlasNodes == c->numNodes - c->topLevelIndex, PTB_ERR_INVALID_ARGUMENT, "TLAS slice size mismatch");
    REQUIRE(numMaterials > 0, PTB_ERR_INVALID_ARGUMENT, "no materials");
    for (int i = 0; i < numTlasNodes; i++)
    {   // validate before anything is modified: a TLAS leaf must keep pointing at an existing material
        const float* n = tlasNodes + (size_t)i * 9;
        if ((int)n[8] < 0) REQUIRE((int)n[7] >= 0 && (int)n[7] < numMaterials, PTB_ERR_INVALID_ARGUMENT, "TLAS leaf material id out of range");
    }
    CK(cudaSetDevice(c->device));
    c->hTransforms.assign(transforms, transforms + (size_t)numInstances * 16);
    c->hMaterials.assign(materials, materials + (size_t)numMaterials * 32);
    memcpy(&c->hNodes[(size_t)c->topLevelIndex * 9], tlasNodes, (size_t)numTlasNodes * 9 * sizeof(float));
    CK(c->transforms.upload((const float4*)transforms, (size_t)numInstances * 4, c->stream));
    CK(c->materials.upload((const float4*)materials, (size_t)numMaterials * 8, c->stream));
    CK(cudaMemcpyAsync(c->nodes.p + (size_t)c->topLevelIndex * 9, tlasNodes, (size_t)numTlasNodes * 9 * sizeof(float), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->S.materials = c->materials.p; c->S.transforms = c->transforms.p; c->S.numMaterials = numMaterials;
    c->sceneVersion++;
    int rc = deriveHierarchy(c, c->topLevelIndex, c->numNodes, false);
    if (rc) return rc;
    refreshDerivedFlags(c);
    return PTB_OK;
}

int ptb_rebuild_instances(PtbCtx* c, const float* transforms, int32_t numInstances, const float* materials, int32_t numMaterials, const int32_t* instanceMaterialIDs, int32_t onHost)
{
    REQUIRE(c && transforms && materials, PTB_ERR_INVALID_ARGUMENT, "null argument");
    REQUIRE(numInstances == c->S.numInstances && numMaterials > 0, PTB_ERR_INVALID_ARGUMENT, "instance count changed (the reference keeps it across RebuildInstances)");
    REQUIRE(c->numNodes - c->topLevelIndex >= 2 * numInstances, PTB_ERR_INVALID_ARGUMENT, "TLAS slice smaller than 2 * numInstances slots");
    CK(cudaSetDevice(c->device));
    const int n = numInstances, top = c->topLevelIndex;
    // per instance: BLAS root and material id, as the current TLAS leaves carry them (bvh_translator.cpp:70-78)
    std::vector<int> root((size_t)n, -1), mat((size_t)n, 0);
    for (int i = top; i < c->numNodes; i++)
    {
        const float* nd = &c->hNodes[(size_t)i * 9];
        const int leaf = (int)nd[8];
        if (leaf < 0 && -leaf - 1 < n) { root[(size_t)(-leaf - 1)] = (int)nd[6]; mat[(size_t)(-leaf - 1)] = (int)nd[7]; }
    }
    for (int i = 0; i < n; i++)
    {
        if (instanceMaterialIDs) mat[(size_t)i] = instanceMaterialIDs[i];
        REQUIRE(root[(size_t)i] >= 0 && root[(size_t)i] < top, PTB_ERR_INVALID_ARGUMENT, "instance without a TLAS leaf");
        REQUIRE(mat[(size_t)i] >= 0 && mat[(size_t)i] < numMaterials, PTB_ERR_INVALID_ARGUMENT, "instance material id out of range");
    }
    std::vector<float> slice;
    bool built = false;
    const auto tAll = std::chrono::steady_clock::now();
    if (!onHost)
    {   // k_tlas_build writes the slice straight into the canonical device node array; it is read back for the host-side derivation of the packed layouts
        const size_t recBytes = (size_t)ptbk_tlas_rec_size() * ((size_t)n + 2);
        CK(c->tlasBlasRoot.upload(root.data(), (size_t)n, c->stream)); CK(c->tlasMatID.upload(mat.data(), (size_t)n, c->stream));
        CK(c->transforms.upload((const float4*)transforms, (size_t)n * 4, c->stream));
        CK(c->tlasNodeOf.alloc((size_t)n)); CK(c->tlasRemap.alloc((size_t)n + 2)); CK(c->tlasResult.alloc(4));
        CK(cudaMemsetAsync(c->tlasResult.p, 0, 4 * sizeof(int), c->stream));
        CK(c->tlasBounds.alloc((size_t)n * 6)); CK(c->tlasCent.alloc((size_t)n * 3));
        for (int k = 0; k < 3; k++) CK(c->tlasRec[k].alloc(recBytes));
        cudaEventRecord(c->evStart, c->stream);
        CK((cudaError_t)ptbk_tlas_build(cfg(c), c->nodes.p, top, c->transforms.p, n, c->tlasBlasRoot.p, c->tlasMatID.p, c->tlasBounds.p, c->tlasCent.p, c->tlasNodeOf.p,
                                        c->tlasRec[0].p, c->tlasRec[1].p, c->tlasRec[2].p, c->tlasRemap.p, c->tlasResult.p));
        cudaEventRecord(c->evStop, c->stream);
        c->timingValid = false;
        CK(cudaGetLastError());
        int res[2] = {1, 0};
        slice.resize((size_t)2 * n * 9);
        CK(cudaMemcpyAsync(res, c->tlasResult.p, sizeof(res), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(slice.data(), c->nodes.p + (size_t)top * 9, slice.size() * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        cudaEventElapsedTime(&c->lastRebuildBuildMs, c->evStart, c->evStop);
        built = res[0] == 0;            // flag: an input where the reference's element order matters (or -0.0 / non-finite boxes) -> exact sequential build below
        c->lastRebuildWhere = built ? 0 : 2;
    }
    if (!built)
    {
        std::string err;
        const auto t0 = std::chrono::steady_clock::now();
        int rc = ptbd_build_tlas_host(c->hNodes.data(), top, transforms, n, root.data(), mat.data(), slice, nullptr, err);
        REQUIRE(rc == 0, rc, err);
        c->lastRebuildBuildMs = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (onHost) c->lastRebuildWhere = 1;
    }
    // slots beyond the 2n the translator reserves do not exist in reference-made arrays; keep whatever a caller-provided array held there
    std::vector<float> full(c->hNodes.begin() + (size_t)top * 9, c->hNodes.end());
    memcpy(full.data(), slice.data(), slice.size() * sizeof(float));
    int rc = ptb_update_instances(c, transforms, numInstances, materials, numMaterials, full.data(), c->numNodes - top);
    c->lastRebuildTotalMs = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - tAll).count();
    return rc;
}

int ptb_last_rebuild_info(PtbCtx* c, int32_t* where, float* buildMs, float* totalMs)
{
    REQUIRE(c, PTB_ERR_INVALID_ARGUMENT, "null argument");
    if (where) *where = c->lastRebuildWhere;
    if (buildMs) *buildMs = c->lastRebuildBuildMs;
    if (totalMs) *totalMs = c->lastRebuildTotalMs;
    return PTB_OK;
}

int ptb_update_envmap(PtbCtx* c, const float* img, const float* cdf, int32_t w, int32_t h, float totalSum)
{
    REQUIRE(c && img && cdf && w > 0 && h > 0, PTB_ERR_INVALID_ARGUMENT, "bad environment map");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    c->envImg.release(); c->envCdf.release();
    CK(c->envImg.upload(img, (size_t)w * h * 3, c->stream));
    CK(c->envCdf.upload(cdf, (size_t)w * h, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->S.envImg = c->envImg.p; c->S.envCdf = c->envCdf.p; c->S.envW = w; c->S.envH = h; c->S.envTotalSum = totalSum;
    { int grc = buildEnvGuide(c, cdf, w, h, totalSum); if (grc) return grc; }
    c->sceneVersion++;
    int cull = c->F.cullBoxes;
    refreshFrameParams(c);
    c->F.cullBoxes = cull;
    return PTB_OK;
}

int ptb_reset_accum(PtbCtx* c)
{
    REQUIRE(c, PTB_ERR_INVALID_ARGUMENT, "null context");
    CK(cudaSetDevice(c->device));
    CK(cudaMemsetAsync(c->accum.p, 0, (size_t)c->opts.renderW * c->opts.renderH * sizeof(float4), c->stream));
    return PTB_OK;
}


int ptb_render_tile(PtbCtx* c, int32_t tx, int32_t ty, int32_t frameNum)
{
    REQUIRE(c, PTB_ERR_INVALID_ARGUMENT, "null context");
    const FrameParams& F = c->F;
    REQUIRE(tx >= 0 && ty >= 0 && tx < F.numTilesX && ty < F.numTilesY, PTB_ERR_INVALID_ARGUMENT, "tile out of range");
    CK(cudaSetDevice(c->device));
    WaveParams W{};
    W.x0 = tx * F.tileW; W.y0 = ty * F.tileH;
    W.rw = std::min(F.tileW, F.renderW - W.x0); W.rh = std::min(F.tileH, F.renderH - W.y0);     // overhang clipped (Q15)
    W.nSamples = 1; W.firstSample = 1; W.sampleStride = 1; W.fixedFrame = frameNum; W.previewMode = 0;
    beginTiming(c);
    int rc = renderWave(c, F, W, nullptr);
    endTiming(c);
    return rc;
}

static int samplesPerWave(const PtbCtx* c)
{
    int spw = c->opts.samplesPerWave;
    if (spw <= 0)
    {   // auto: keep ~64 M paths in flight — the deep bounces of a wave are short launches with long tails, and larger waves amortise them
        // (hyperion 1080p, same box: 8 passes 878 spp/s, 16 passes 891, 32 passes 899); ~10 GB of path state out of 180 GB of HBM
        size_t px = (size_t)c->F.renderW * c->F.renderH;
        spw = (int)std::max<size_t>(1, std::min<size_t>(32, ((64u << 20) + px / 2) / std::max<size_t>(px, 1)));
    }
    return spw;
}

int ptb_render_pass(PtbCtx* c, int32_t sample, int32_t sampleStride, int32_t maxLookahead)
{
    REQUIRE(c && sample >= 1 && sampleStride >= 1, PTB_ERR_INVALID_ARGUMENT, "bad sample pass");
    CK(cudaSetDevice(c->device));
    PtbCtx::Pending& pd = c->pending;
    if (pd.valid && sample == pd.first + pd.consumed * pd.stride && sampleStride == pd.stride && pd.consumed < pd.n && pd.sceneVersion == c->sceneVersion &&
        memcmp(&pd.F, &c->F, sizeof(FrameParams)) == 0)
    {   // this pass was traced with the wave of an earlier call under the same uniforms: add it to the running sum now
        WaveParams W = pd.W;
        W.accFirst = pd.consumed; W.accCount = 1;
        ptbk_accumulate(cfg(c), pd.F, W, pathState(c), c->accum.p, nullptr);
        CK(cudaGetLastError());
        if (++pd.consumed == pd.n) pd.valid = false;
        c->samplesRendered++;
        return PTB_OK;
    }
    int n = samplesPerWave(c);
    if (maxLookahead > 0) n = std::min(n, (int)maxLookahead);
    const FrameParams F = c->F;
    WaveParams W{};
    W.x0 = 0; W.y0 = 0; W.rw = F.renderW; W.rh = F.renderH;
    W.nSamples = n; W.firstSample = sample; W.sampleStride = sampleStride; W.fixedFrame = -1; W.previewMode = 0;
    W.accFirst = 0; W.accCount = 1;
    beginTiming(c);
    int rc = renderWave(c, F, W, nullptr);
    endTiming(c);
    if (rc) return rc;
    c->samplesRendered++;
    if (n > 1) { pd.valid = true; pd.first = sample; pd.n = n; pd.consumed = 1; pd.stride = sampleStride; pd.sceneVersion = c->sceneVersion; pd.F = F; pd.W = W; }
    return PTB_OK;
}

int ptb_render_samples(PtbCtx* c, int32_t firstSample, int32_t nSamples, int32_t sampleStride)
{
    REQUIRE(c && firstSample >= 1 && nSamples >= 0 && sampleStride >= 1, PTB_ERR_INVALID_ARGUMENT, "bad sample range");
    CK(cudaSetDevice(c->device));
    const FrameParams& F = c->F;
    const int spw = samplesPerWave(c);
    beginTiming(c);
    int done = 0;
    while (done < nSamples)
    {
        int n = std::min(spw, nSamples - done);
        WaveParams W{};
        W.x0 = 0; W.y0 = 0; W.rw = F.renderW; W.rh = F.renderH;
        W.nSamples = n; W.firstSample = firstSample + done * sampleStride; W.sampleStride = sampleStride; W.fixedFrame = -1; W.previewMode = 0;
        int rc = renderWave(c, F, W, nullptr);
        if (rc) return rc;
        done += n;
    }
    endTiming(c);
    c->samplesRendered += (uint64_t)nSamples;
    return PTB_OK;
}

int ptb_render_preview(PtbCtx* c, int32_t w, int32_t h, float* outRgba)
{
    REQUIRE(c && outRgba && w > 0 && h > 0, PTB_ERR_INVALID_ARGUMENT, "bad preview target");
    CK(cudaSetDevice(c->device));
    CK(c->preview.alloc((size_t)w * h));
    FrameParams F = c->F;
    F.maxDepth = 2;                                           // Renderer.cpp:798 (scene->dirty ? 2 : maxDepth)
    WaveParams W{};
    W.x0 = 0; W.y0 = 0; W.rw = w; W.rh = h; W.nSamples = 1; W.firstSample = 1; W.sampleStride = 1; W.fixedFrame = 1; W.previewMode = 1;
    int rc = renderWave(c, F, W, c->preview.p);
    if (rc) return rc;
    CK(cudaMemcpyAsync(outRgba, c->preview.p, (size_t)w * h * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return PTB_OK;
}

int ptb_read_accum_f32(PtbCtx* c, float* out)
{
    REQUIRE(c && out, PTB_ERR_INVALID_ARGUMENT, "null argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(out, c->accum.p, (size_t)c->opts.renderW * c->opts.renderH * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return PTB_OK;
}

int ptb_write_accum_f32(PtbCtx* c, const float* in)
{
    REQUIRE(c && in, PTB_ERR_INVALID_ARGUMENT, "null argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(c->accum.p, in, (size_t)c->opts.renderW * c->opts.renderH * sizeof(float4), cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return PTB_OK;
}

int ptb_accum_device_ptr(PtbCtx* c, void** p, uint64_t* nbytes)
{
    REQUIRE(c && p, PTB_ERR_INVALID_ARGUMENT, "null argument");
    *p = c->accum.p;
    if (nbytes) *nbytes = (uint64_t)c->opts.renderW * c->opts.renderH * sizeof(float4);
    return PTB_OK;
}

int ptb_read_output_rgba8_from(PtbCtx* c, const void* devAccum, float invSampleCounter, uint8_t* out)
{
    REQUIRE(c && out, PTB_ERR_INVALID_ARGUMENT, "null argument");
    CK(cudaSetDevice(c->device));
    const PtbOptions& o = c->opts;
    ptbk_tonemap(cfg(c), devAccum ? (const float4*)devAccum : c->accum.p, o.renderW, o.renderH, invSampleCounter, o.enableTonemap, o.enableAces, o.simpleAcesFit,
                 o.backgroundCol, c->F.features, c->out8.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out, c->out8.p, (size_t)o.renderW * o.renderH * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return PTB_OK;
}

int ptb_read_output_rgba8(PtbCtx* c, float invSampleCounter, uint8_t* out) { return ptb_read_output_rgba8_from(c, nullptr, invSampleCounter, out); }

int ptb_snapshot_output(PtbCtx* c, float invSampleCounter) { return ptb_snapshot_output_from(c, nullptr, invSampleCounter); }

int ptb_snapshot_output_from(PtbCtx* c, const void* devAccum, float invSampleCounter)
{
    REQUIRE(c, PTB_ERR_INVALID_ARGUMENT, "null context");
    CK(cudaSetDevice(c->device));
    const PtbOptions& o = c->opts;
    CK(c->snapshot.alloc((size_t)o.renderW * o.renderH));
    if (c->snapshotFloat) CK(c->snapshotF.alloc((size_t)o.renderW * o.renderH));
    ptbk_tonemap(cfg(c), devAccum ? (const float4*)devAccum : c->accum.p, o.renderW, o.renderH, invSampleCounter, o.enableTonemap, o.enableAces, o.simpleAcesFit, o.backgroundCol, c->F.features,
                 c->snapshot.p, c->snapshotFloat ? c->snapshotF.p : nullptr);
    CK(cudaGetLastError());
    return PTB_OK;
}

int ptb_read_snapshot_rgba8(PtbCtx* c, uint8_t* out)
{
    REQUIRE(c && out, PTB_ERR_INVALID_ARGUMENT, "null argument");
    CK(cudaSetDevice(c->device));
    const size_t n = (size_t)c->opts.renderW * c->opts.renderH;
    if (!c->snapshot.p || c->snapshot.n < n) { memset(out, 0, n * 4); return PTB_OK; }      // nothing completed yet: the cleared texture
    CK(cudaMemcpyAsync(out, c->snapshot.p, n * 4, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return PTB_OK;
}

int ptb_set_snapshot_float(PtbCtx* c, int32_t enable) { REQUIRE(c, PTB_ERR_INVALID_ARGUMENT, "null context"); c->snapshotFloat = enable != 0; return PTB_OK; }

int ptb_read_snapshot_rgb32f(PtbCtx* c, float* outRgb)
{
    REQUIRE(c && outRgb, PTB_ERR_INVALID_ARGUMENT, "null argument");
    CK(cudaSetDevice(c->device));
    const size_t n = (size_t)c->opts.renderW * c->opts.renderH;
    if (!c->snapshotF.p || c->snapshotF.n < n) { memset(outRgb, 0, n * 12); return PTB_OK; }
    std::vector<float> tmp(n * 4);
    CK(cudaMemcpyAsync(tmp.data(), c->snapshotF.p, n * 16, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    for (size_t i = 0; i < n; i++) { outRgb[i * 3] = tmp[i * 4]; outRgb[i * 3 + 1] = tmp[i * 4 + 1]; outRgb[i * 3 + 2] = tmp[i * 4 + 2]; }   // GL_RGB, GL_FLOAT
    return PTB_OK;
}

int ptb_host_alloc(uint64_t nbytes, void** out)
{
    REQUIRE(out && nbytes > 0, PTB_ERR_INVALID_ARGUMENT, "bad argument");
    CK(cudaMallocHost(out, (size_t)nbytes));
    return PTB_OK;
}

int ptb_host_free(void* p)
{
    if (p) CK(cudaFreeHost(p));
    return PTB_OK;
}

int ptb_get_stats(PtbCtx* c, PtbStats* out)
{
    REQUIRE(c && out, PTB_ERR_INVALID_ARGUMENT, "null argument");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    DevStats ds{};
    CK(cudaMemcpy(&ds, c->dstats.p, sizeof(ds), cudaMemcpyDeviceToHost));
    memset(out, 0, sizeof(*out));
    out->pathSegments = ds.pathSegments; out->shadowRays = ds.shadowRays; out->samplesRendered = c->samplesRendered;
    out->kernelLaunches = (uint64_t)c->launches;
    if (c->timingValid) cudaEventElapsedTime(&out->lastRenderMs, c->evStart, c->evStop);
    if (c->profiling)
    {
        float tot[KIND_COUNT] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (size_t i = 0; i + 1 < c->traceEventsUsed; i++)
        {
            const int k = c->eventKind[i];
            if (k < 0) continue;
            float ms = 0.f; cudaEventElapsedTime(&ms, c->traceEvents[i], c->traceEvents[i + 1]); tot[k] += ms;
        }
        out->lastTraceMs = tot[KIND_TRACE]; out->lastCameraMs = tot[KIND_CAMERA]; out->lastSortMs = tot[KIND_SORT]; out->lastShadeMs = tot[KIND_SHADE];
        out->lastShadowMs = tot[KIND_SHADOW]; out->lastAccumMs = tot[KIND_ACCUM];
    }
    return PTB_OK;
}

int ptb_reset_stats(PtbCtx* c)
{
    REQUIRE(c, PTB_ERR_INVALID_ARGUMENT, "null context");
    CK(cudaSetDevice(c->device));
    CK(cudaMemsetAsync(c->dstats.p, 0, sizeof(DevStats), c->stream));
    c->samplesRendered = 0;
    return PTB_OK;
}

int ptb_set_profiling(PtbCtx* c, int32_t enable) { REQUIRE(c, PTB_ERR_INVALID_ARGUMENT, "null context"); c->profiling = enable != 0; return PTB_OK; }

int ptb_set_stream(PtbCtx* c, void* s)
{
    REQUIRE(c, PTB_ERR_INVALID_ARGUMENT, "null context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    c->stream = s ? (cudaStream_t)s : c->ownStream;
    return PTB_OK;
}

int ptb_synchronize(PtbCtx* c)
{
    REQUIRE(c, PTB_ERR_INVALID_ARGUMENT, "null context");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return PTB_OK;
}

// internal knob used by tests/bench: conservative t-culling of child boxes (identical hits, fewer node fetches)
int ptb_set_cull(PtbCtx* c, int32_t enable) { REQUIRE(c, PTB_ERR_INVALID_ARGUMENT, "null context"); c->F.cullBoxes = enable ? 1 : 0; return PTB_OK; }

int ptb_trace_closest_device(PtbCtx* c, const void* devRays, int64_t n, int32_t depth, void* devHits)
{
    REQUIRE(c && devRays && devHits && n >= 0, PTB_ERR_INVALID_ARGUMENT, "null argument");
    CK(cudaSetDevice(c->device));
    if (n == 0) return PTB_OK;
    ptbk_trace_closest_batch(cfg(c), c->S, c->F, (const float*)devRays, n, depth, devHits);
    CK(cudaGetLastError());
    return PTB_OK;
}

int ptb_trace_closest(PtbCtx* c, const float* rays, int64_t n, int32_t depth, PtbHit* out)
{
    REQUIRE(c && (n == 0 || (rays && out)) && n >= 0, PTB_ERR_INVALID_ARGUMENT, "null argument");
    if (n == 0) return PTB_OK;
    CK(cudaSetDevice(c->device));
    DevBuf<float> dr; DevBuf<PtbHit> dh;
    CK(dr.upload(rays, (size_t)n * 6, c->stream));
    CK(dh.alloc((size_t)n));
    int rc = ptb_trace_closest_device(c, dr.p, n, depth, dh.p);
    if (rc == PTB_OK)
    {
        cudaError_t e = cudaMemcpyAsync(out, dh.p, (size_t)n * sizeof(PtbHit), cudaMemcpyDeviceToHost, c->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
        if (e != cudaSuccess) { g_err = cudaGetErrorString(e); rc = PTB_ERR_CUDA; }
    }
    dr.release(); dh.release();
    return rc;
}

int ptb_trace_any(PtbCtx* c, const float* rays, const float* maxDist, int64_t n, int32_t* out)
{
    REQUIRE(c && (n == 0 || (rays && maxDist && out)) && n >= 0, PTB_ERR_INVALID_ARGUMENT, "null argument");
    if (n == 0) return PTB_OK;
    CK(cudaSetDevice(c->device));
    DevBuf<float> dr, dm; DevBuf<int> dout;
    CK(dr.upload(rays, (size_t)n * 6, c->stream));
    CK(dm.upload(maxDist, (size_t)n, c->stream));
    CK(dout.alloc((size_t)n));
    ptbk_trace_any_batch(cfg(c), c->S, c->F, dr.p, dm.p, n, dout.p);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, dout.p, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    dr.release(); dm.release(); dout.release();
    if (e != cudaSuccess) { g_err = cudaGetErrorString(e); return PTB_ERR_CUDA; }
    return PTB_OK;
}

static int bsdfBatch(PtbCtx* c, const PtbBsdfQuery* q, int64_t n, PtbBsdfResult* out, int sample)
{
    REQUIRE(c && (n == 0 || (q && out)) && n >= 0, PTB_ERR_INVALID_ARGUMENT, "null argument");
    if (n == 0) return PTB_OK;
    CK(cudaSetDevice(c->device));
    DevBuf<PtbBsdfQuery> dq; DevBuf<PtbBsdfResult> dres;
    CK(dq.upload(q, (size_t)n, c->stream));
    CK(dres.alloc((size_t)n));
    ptbk_bsdf_batch(cfg(c), dq.p, n, dres.p, sample);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, dres.p, (size_t)n * sizeof(PtbBsdfResult), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    dq.release(); dres.release();
    if (e != cudaSuccess) { g_err = cudaGetErrorString(e); return PTB_ERR_CUDA; }
    return PTB_OK;
}
int ptb_bsdf_eval(PtbCtx* c, const PtbBsdfQuery* q, int64_t n, PtbBsdfResult* out) { return bsdfBatch(c, q, n, out, 0); }
int ptb_bsdf_sample(PtbCtx* c, const PtbBsdfQuery* q, int64_t n, PtbBsdfResult* out) { return bsdfBatch(c, q, n, out, 1); }
int ptb_lambert_eval(PtbCtx* c, const PtbBsdfQuery* q, int64_t n, PtbBsdfResult* out) { return bsdfBatch(c, q, n, out, 2); }
int ptb_lambert_sample(PtbCtx* c, const PtbBsdfQuery* q, int64_t n, PtbBsdfResult* out) { return bsdfBatch(c, q, n, out, 3); }

int ptb_camera_rays(PtbCtx* c, int32_t sample, float* outRays)
{
    REQUIRE(c && outRays && sample >= 1, PTB_ERR_INVALID_ARGUMENT, "bad argument");
    CK(cudaSetDevice(c->device));
    size_t n = (size_t)c->F.renderW * c->F.renderH;
    DevBuf<float> d;
    CK(d.alloc(n * 6));
    WaveParams W{}; W.rw = c->F.renderW; W.rh = c->F.renderH; W.firstSample = sample; W.sampleStride = 1; W.fixedFrame = -1; W.nSamples = 1;
    ptbk_camera_rays(cfg(c), c->F, W, d.p);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(outRays, d.p, n * 6 * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    d.release();
    if (e != cudaSuccess) { g_err = cudaGetErrorString(e); return PTB_ERR_CUDA; }
    return PTB_OK;
}

int ptb_read_nodes(PtbCtx* c, float* out, int32_t numNodes)
{
    REQUIRE(c && out && numNodes == c->numNodes, PTB_ERR_INVALID_ARGUMENT, "bad argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(out, c->nodes.p, (size_t)numNodes * 9 * sizeof(float), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return PTB_OK;
}

int ptb_stack_depth(PtbCtx* c, int32_t* out) { REQUIRE(c && out, PTB_ERR_INVALID_ARGUMENT, "null argument"); *out = c->S.stackDepth; return PTB_OK; }

} // extern "C"
