// trav_host.cpp — TEST INFRASTRUCTURE (see ptb_host_shim.h): the product's traversal / light-test / camera-ray source
// (csrc/ptb_device.cuh) and its host-side layout derivation (csrc/ptb_derive.cpp), compiled for the host and exposed through a small
// C interface with the same record layouts as the C ABI's parity entry points (PtbHit).  Used only by tests/test_host_traversal.py.
#define PTB_HOST_HARNESS 1
#include "ptb_device.cuh"
#include "ptb_derive.h"
#include <string>
#include <vector>

using namespace ptb;

struct HHScene
{
    DevScene S{};
    PtbDerivedHierarchy dh; PtbDerivedLights dl; PtbDerivedWide dw; std::vector<float4> tris;
    std::vector<float> nodes, transforms; std::vector<int> vertIndices; std::vector<float4> verticesUVX;
    int anyStackHigh = 0;      // deepest any-hit stack over all hh_trace_any calls
    unsigned long long anyBytes = 0;   // bytes the last hh_trace_any call fetched (nodes, triangles, instance rows, lights)
};
struct BigStack      // thread-local stack sized for the wide hierarchy's bound (96)
{
    uint32_t a[128]; int sp = 0, high = 0;      // high: deepest the stack has been (checked against the bound the kernels reserve shared memory for)
    void reset() { sp = 0; }
    void push(uint32_t v) { a[sp++] = v; if (sp > high) high = sp; }
    uint32_t pop() { return a[--sp]; }
};
struct HHHit { float t; int kind, instance, matID, primSlot, triIDx; float bary[3]; int lightIdx; };

extern "C" {

HHScene* hh_create(const float* nodes, int numNodes, int topLevelIndex, const int32_t* vertIndices, int numIndices, const float* verticesUVX, int numVertices,
                   const float* transforms, int numInstances, int numMaterials, const float* lights, int numLights, char* errOut, int errCap)
{
    HHScene* h = new HHScene();
    std::string err;
    h->nodes.assign(nodes, nodes + (size_t)numNodes * 9); h->transforms.assign(transforms, transforms + (size_t)numInstances * 16);
    h->vertIndices.assign(vertIndices, vertIndices + (size_t)numIndices * 3);
    int rc = ptbd_build_tris(vertIndices, numIndices, verticesUVX, numVertices, h->tris, err);
    if (!rc) rc = ptbd_derive_hierarchy(nodes, numNodes, topLevelIndex, numIndices, numMaterials, transforms, numInstances, 0, numNodes, h->dh, err);
    if (rc) { if (errOut) { strncpy(errOut, err.c_str(), errCap - 1); errOut[errCap - 1] = 0; } delete h; return nullptr; }
    ptbd_build_lights(lights, numLights, h->dl);
    ptbd_build_wide(nodes, numNodes, topLevelIndex, numIndices, numInstances, h->dh.transOnly, h->dw);
    if (h->dw.ok) for (size_t k = 0; k < h->dw.instRootMeta.size(); k++) h->dh.instTrav[k * 4 + 2].w = __uint_as_float(h->dw.instRootMeta[k]);
    DevScene& S = h->S;
    S.nodes = h->nodes.data(); S.vertIndices = h->vertIndices.data();
    S.inner = h->dh.inner.data(); S.tris = h->tris.data(); S.instTrav = h->dh.instTrav.data(); S.instShade = h->dh.instShade.data();
    S.lightsPre = h->dl.lightsPre.data(); S.lightGroups = h->dl.lightGroups.data(); S.lightGrid = h->dl.lightGrid.data(); S.numLightGroups = h->dl.numGroups;
    S.rootMeta = h->dh.rootMeta; S.stackDepth = h->dh.stackDepth;
    S.wide = h->dw.ok ? h->dw.wide.data() : nullptr; S.rootMetaWide = h->dw.rootMeta;
    S.stackDepthAny = (h->dw.ok && h->dw.stackDepth > S.stackDepth) ? h->dw.stackDepth : S.stackDepth;
    S.numNodes = numNodes; S.topLevelIndex = topLevelIndex; S.numIndices = numIndices; S.numVertices = numVertices; S.numMaterials = numMaterials;
    S.numInstances = numInstances; S.numLights = numLights;
    return h;
}
void hh_destroy(HHScene* h) { delete h; }
int hh_stack_depth(HHScene* h) { return h->S.stackDepthAny; }
int hh_any_stack_high(HHScene* h) { return h->anyStackHigh; }
double hh_any_bytes(HHScene* h) { return (double)h->anyBytes; }
int hh_wide_nodes(HHScene* h) { return h->dw.ok ? (int)(h->dw.wide.size() / 8) : -1; }

// k_trace_batch of ptb_kernels.cu, one ray after the other
void hh_trace_closest(HHScene* h, const float* rays, long long n, int lights, int cull, HHHit* out)
{
    const DevScene& S = h->S;
#pragma omp parallel for schedule(dynamic, 4096)
    for (long long i = 0; i < n; i++)
    {
        LocalStack stk;
        const float3 o = f3(rays[i * 6 + 0], rays[i * 6 + 1], rays[i * 6 + 2]), d = f3(rays[i * 6 + 3], rays[i * 6 + 4], rays[i * 6 + 5]);
        HitRec hr; hr.t = PTB_INF; hr.prim = -1; hr.inst = -1; hr.light = -1; hr.bu = hr.bv = 0.f;
        float t = PTB_INF;
        if (lights) closestLights(S, o, d, t, hr.light);
        if (cull) traverse<false, false, true>(S, o, d, t, stk, hr, NoAlpha());
        else traverse<false, false, false>(S, o, d, t, stk, hr, NoAlpha());
        HHHit r;
        r.t = hr.t;
        if (hr.t == PTB_INF) { r.kind = 0; r.instance = r.matID = r.primSlot = r.triIDx = r.lightIdx = -1; r.bary[0] = r.bary[1] = r.bary[2] = 0.f; }
        else if (hr.inst >= 0)
        {
            r.kind = 1; r.instance = hr.inst; r.matID = __float_as_int(S.instTrav[(size_t)hr.inst * 4 + 1].w); r.primSlot = hr.prim;
            r.triIDx = S.vertIndices[(size_t)hr.prim * 3];
            r.bary[0] = xs(xs(1.0f, hr.bu), hr.bv); r.bary[1] = hr.bu; r.bary[2] = hr.bv; r.lightIdx = -1;
        }
        else { r.kind = 2; r.instance = r.matID = r.primSlot = r.triIDx = -1; r.bary[0] = r.bary[1] = r.bary[2] = 0.f; r.lightIdx = hr.light; }
        out[i] = r;
    }
}

// wide: 1 = the production path of k_shadow (4-wide hierarchy where admissible, binary otherwise); 0 = binary only.  fallbacks counts the rays the wide
// path handed back to the binary traversal
void hh_trace_any(HHScene* h, const float* rays, const float* maxDist, long long n, int lights, int cull, int wide, int* out, long long* fallbacks)
{
    const DevScene& S = h->S;
    long long fb = 0; int high = 0; unsigned long long bytes = 0;
#pragma omp parallel for schedule(dynamic, 4096) reduction(+ : fb) reduction(max : high) reduction(+ : bytes)
    for (long long i = 0; i < n; i++)
    {
        BigStack stk;
        const unsigned long long b0 = g_hh_ldg_bytes;
        const float3 o = f3(rays[i * 6 + 0], rays[i * 6 + 1], rays[i * 6 + 2]), d = f3(rays[i * 6 + 3], rays[i * 6 + 4], rays[i * 6 + 5]);
        bool occ = lights && anyLights(S, o, d, maxDist[i]);
        if (!occ)
        {
            int r = 2;
            if (wide && S.wide)
            {
                if (wideRayOk(d)) r = cull ? traverseWideAny<false, true>(S, o, d, maxDist[i], stk, NoAlpha()) : traverseWideAny<false, false>(S, o, d, maxDist[i], stk, NoAlpha());
                if (r == 2) fb++;
            }
            if (r == 2) { HitRec hr; occ = cull ? traverse<true, false, true>(S, o, d, maxDist[i], stk, hr, NoAlpha()) : traverse<true, false, false>(S, o, d, maxDist[i], stk, hr, NoAlpha()); }
            else occ = r != 0;
        }
        out[i] = occ ? 1 : 0;
        if (stk.high > high) high = stk.high;
        bytes += g_hh_ldg_bytes - b0;
    }
    h->anyBytes = bytes;
    h->anyStackHigh = std::max(h->anyStackHigh, high);
    if (fallbacks) *fallbacks = fb;
}

// uniforms as refreshFrameParams (ptb_api.cpp) derives them from PtbOptions / PtbCamera
void hh_camera_rays(int renderW, int renderH, int tileW, int tileH, const float* pos, const float* right, const float* up, const float* fwd, float fov, float focalDist,
                    float aperture, int sample, int useTables, float* out)
{
    FrameParams F{};
    F.renderW = renderW; F.renderH = renderH; F.tileW = tileW; F.tileH = tileH;
    F.invNumTilesX = (float)tileW / renderW; F.invNumTilesY = (float)tileH / renderH;
    F.numTilesX = (int)ceilf((float)renderW / tileW); F.numTilesY = (int)ceilf((float)renderH / tileH);
    memcpy(F.camPos, pos, 12); memcpy(F.camRight, right, 12); memcpy(F.camUp, up, 12); memcpy(F.camFwd, fwd, 12);
    F.aspect = (float)renderH / (float)renderW;
    F.camScale = tanf(fov * 0.5f); F.camFocalDist = focalDist; F.camAperture = aperture;
    std::vector<float2> tx, ty; std::string err;
    if (useTables && ptbd_build_pixel_tables(renderW, renderH, tileW, tileH, tx, ty, err) == 0) { F.pixTabX = tx.data(); F.pixTabY = ty.data(); }
    WaveParams W{}; W.rw = renderW; W.rh = renderH; W.firstSample = sample; W.sampleStride = 1; W.fixedFrame = -1; W.nSamples = 1;
    for (int i = 0; i < renderW * renderH; i++)
    {
        int x = i % renderW, y = i / renderW;
        Rng rng; float3 ro, rd;
        cameraRay(F, W, x, y, sample, rng, ro, rd);
        out[i * 6 + 0] = ro.x; out[i * 6 + 1] = ro.y; out[i * 6 + 2] = ro.z; out[i * 6 + 3] = rd.x; out[i * 6 + 4] = rd.y; out[i * 6 + 5] = rd.z;
    }
}

// Environment-map CDF search (envBinarySearch): with the guide table (ptbd_build_env_guide) and as the reference's two binary searches.  Returns 1 when a guide
// exists for this CDF (0: the device would run the reference search; outFast then repeats it).
int hh_env_search(const float* cdf, int w, int h, float totalSum, const float* values, int n, float* outFast, float* outRef)
{
    DevScene S{};
    S.envCdf = cdf; S.envW = w; S.envH = h; S.envTotalSum = totalSum;
    std::vector<uint32_t> guide; float scale = 0.f;
    const int have = ptbd_build_env_guide(cdf, w, h, totalSum, guide, scale) == 0;
    for (int i = 0; i < n; i++) { const float2 r = envBinarySearch(S, values[i]); outRef[i * 2] = r.x; outRef[i * 2 + 1] = r.y; }
    if (have) { S.envGuide = guide.data(); S.envGuideN = (int)guide.size() - 1; S.envGuideScale = scale; }
    for (int i = 0; i < n; i++) { const float2 r = envBinarySearch(S, values[i]); outFast[i * 2] = r.x; outFast[i * 2 + 1] = r.y; }
    return have;
}

// Slot layout of a wave (groupToPixel / slotOfSample, ptbd_wave_groups): for every slot of a w x h rectangle with nSamples passes, out[slot] = {pass, px, py, slotOfSample(pixel index, pass)}
void hh_slot_map(int w, int h, int nSamples, int blockMajor, int maxLps, int32_t* out, int32_t* lpsOut)
{
    WaveParams W{};
    W.rw = w; W.rh = h; W.vw = (w + 7) & ~7; W.vh = (h + 3) & ~3; W.nSamples = nSamples; W.nSlots = (uint32_t)((size_t)W.vw * W.vh * nSamples);
    W.blockMajor = (blockMajor && nSamples > 1) ? 1 : 0; W.lps = 0; W.lpw = 3;
    if (W.blockMajor) ptbd_wave_groups(nSamples, maxLps, &W.lps, &W.lpw);
    *lpsOut = W.lps;
    for (uint32_t slot = 0; slot < W.nSlots; slot++)
    {
        int s, px, py;
        slotToPixel(W, slot, s, px, py);
        const uint32_t idx = (uint32_t)(((py >> 2) * (W.vw >> 3) + (px >> 3)) * 32 + (py & 3) * 8 + (px & 7));
        out[(size_t)slot * 4 + 0] = s; out[(size_t)slot * 4 + 1] = px; out[(size_t)slot * 4 + 2] = py; out[(size_t)slot * 4 + 3] = (int32_t)slotOfSample(W, idx, (uint32_t)s);
    }
}

// TLAS rebuild (ptbd_build_tlas_host) from the scene's own arrays: blasRoot / materialID per instance are read from the current TLAS leaves
int hh_build_tlas(const float* nodes, int numNodes, int topLevelIndex, const float* transforms, int numInstances, const int32_t* materialIDs, float* tlasOut, int* heightOut)
{
    std::vector<int32_t> root(numInstances, -1), mat(numInstances, 0);
    for (int i = topLevelIndex; i < numNodes; i++)
    {
        const float* n = nodes + (size_t)i * 9;
        if ((int)n[8] < 0 && -(int)n[8] - 1 < numInstances) { root[-(int)n[8] - 1] = (int)n[6]; mat[-(int)n[8] - 1] = (int)n[7]; }
    }
    if (materialIDs) mat.assign(materialIDs, materialIDs + numInstances);         // meshInstances[i].materialID as the application edited it
    std::vector<float> out; std::string err;
    int rc = ptbd_build_tlas_host(nodes, topLevelIndex, transforms, numInstances, root.data(), mat.data(), out, heightOut, err);
    if (rc) return rc;
    memcpy(tlasOut, out.data(), out.size() * sizeof(float));
    return 0;
}

} // extern "C"
