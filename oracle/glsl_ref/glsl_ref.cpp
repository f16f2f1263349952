// glsl_ref.cpp — runs the reference's OWN fragment shaders on the CPU.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// tile.glsl, preview.glsl and tonemap.glsl are taken from /root/reference/src/shaders at build time,
// made compilable by glsl2cpp.py (lexical rewrites only) and #included below, one namespace per shader
// program.  What this file adds is the part of the reference that is OpenGL state and therefore cannot be
// compiled: binding the scene arrays to the samplers (Renderer.cpp:135-249), the uniforms
// (Renderer.cpp:479-543, 766-811), the viewport / fragment coordinates of the full-screen quad
// (Quad.cpp:38-46, vertex.glsl) and the tile schedule (Renderer.cpp:566-589, 733-762).
//
// One shared object per set of `#define OPT_*` (GLSL_REF_DEFINES), like Renderer::InitShaders compiles one
// program per option set.  It exists to PIN the hand-written oracle (oracle/pt_oracle.cpp) and the CUDA
// path to the executing reference shader text: tests/test_glsl_ref.py, scripts/make_glsl_ref_golden.py.
//
// Fixed-function behaviour the harness supplies:
//   * gl_FragCoord = pixel centre of the bound viewport (x+0.5, y+0.5); TexCoords = the quad's texCoords
//     attribute interpolated at that centre = ((x+0.5)/vpW, (y+0.5)/vpH);
//   * accumTexture / pathTraceTexture are sampled at exact texel centres; with the <= 8-bit filter weights
//     GL implementations use (GL 4.6 §8.14.2 leaves the precision open) that returns the texel itself, so
//     they are bound as NEAREST here;
//   * glGetTexImage(GL_RGBA, GL_UNSIGNED_BYTE) of the float output = round(clamp(c,0,1)*255) (GL 4.6 §2.3.5).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../pt_oracle.h"   // OrcSceneDesc / OrcOptions: the plain-data scene and option records (shared with the oracle binding)

#ifndef TILE_INC
#error "build through oracle/glsl_ref/build.py"
#endif

namespace tile_prog {
#include "glsl_compat.h"
thread_local vec4 gl_FragCoord;
#include TILE_INC
}
namespace preview_prog {
#include "glsl_compat.h"
thread_local vec4 gl_FragCoord;
#include PREVIEW_INC
}
namespace tonemap_prog {
#include "glsl_compat.h"
thread_local vec4 gl_FragCoord;
#include TONEMAP_INC
}
// the shader text's own macros must not leak into the harness
#undef PI
#undef INF
#undef EPS

namespace {

struct Scene
{
    std::vector<float> nodes, verticesUVX, normalsUVY, materials, transforms, lights, envImg, envCdf;
    std::vector<int32_t> vertIndices;
    std::vector<uint8_t> textures;
    int numMaterials = 0, numInstances = 0, numLights = 0, numTextures = 0, texW = 0, texH = 0, envW = 0, envH = 0, topLevelIndex = 0;
    float envTotalSum = 0.0f;
    OrcOptions opt;
};
Scene* g_scene = nullptr;

// Which OPT_* this object was compiled with (Renderer.cpp:401-459).
struct Defs { int envMap, lights, rr, rrDepth, uniformLight, openglNormalMap, hideEmitters, background, transparentBackground, alphaTest, mollification, medium, volMis; };
constexpr Defs kDefs = {
#ifdef OPT_ENVMAP
    1,
#else
    0,
#endif
#ifdef OPT_LIGHTS
    1,
#else
    0,
#endif
#ifdef OPT_RR
    1, OPT_RR_DEPTH,
#else
    0, 0,
#endif
#ifdef OPT_UNIFORM_LIGHT
    1,
#else
    0,
#endif
#ifdef OPT_OPENGL_NORMALMAP
    1,
#else
    0,
#endif
#ifdef OPT_HIDE_EMITTERS
    1,
#else
    0,
#endif
#ifdef OPT_BACKGROUND
    1,
#else
    0,
#endif
#ifdef OPT_TRANSPARENT_BACKGROUND
    1,
#else
    0,
#endif
#ifdef OPT_ALPHA_TEST
    1,
#else
    0,
#endif
#ifdef OPT_ROUGHNESS_MOLLIFICATION
    1,
#else
    0,
#endif
#ifdef OPT_MEDIUM
    1,
#else
    0,
#endif
#ifdef OPT_VOL_MIS
    1,
#else
    0,
#endif
};

bool optsMatchDefines(const OrcOptions& o)
{
    return !!o.optEnvMap == kDefs.envMap && !!o.optLights == kDefs.lights && !!o.optRR == kDefs.rr && (!o.optRR || o.rrDepth == kDefs.rrDepth) &&
           !!o.optUniformLight == kDefs.uniformLight && !!o.optOpenglNormalMap == kDefs.openglNormalMap && !!o.optHideEmitters == kDefs.hideEmitters &&
           !!o.optBackground == kDefs.background && !!o.optTransparentBackground == kDefs.transparentBackground && !!o.optAlphaTest == kDefs.alphaTest &&
           !!o.optRoughnessMollification == kDefs.mollification && !!o.optMedium == kDefs.medium && !!o.optVolMis == kDefs.volMis;
}

// glUniform* + texture-unit bindings of one path-tracing program (Renderer.cpp:479-543 static part, :766-799 per-frame part).
#define BIND_PATHTRACE_PROGRAM(NS, S, O, MAXDEPTH)                                                                                  \
    do {                                                                                                                            \
        NS::BVH.data = (S).nodes.data();              NS::BVH.channels = 3;               /* GL_RGB32F  :146 */                     \
        NS::vertexIndicesTex.data = (S).vertIndices.data(); NS::vertexIndicesTex.channels = 3; /* GL_RGB32I  :154 */                \
        NS::verticesTex.data = (S).verticesUVX.data(); NS::verticesTex.channels = 4;      /* GL_RGBA32F :162 */                     \
        NS::normalsTex.data = (S).normalsUVY.data();  NS::normalsTex.channels = 4;        /* GL_RGBA32F :170 */                     \
        NS::materialsTex = NS::sampler2D{(S).materials.data(), 8 * (S).numMaterials, 1, 4, false};   /* :175 */                    \
        NS::transformsTex = NS::sampler2D{(S).transforms.data(), 4 * (S).numInstances, 1, 4, false}; /* :183 */                    \
        NS::lightsTex = NS::sampler2D{(S).lights.data(), 5 * (S).numLights, 1, 3, false};            /* :194 */                    \
        NS::textureMapsArrayTex = NS::sampler2DArray{(S).textures.data(), (S).texW, (S).texH, (S).numTextures}; /* :204 */          \
        NS::envMapTex = NS::sampler2D{(S).envImg.data(), (S).envW, (S).envH, 3, true};               /* :215 LINEAR */              \
        NS::envMapCDFTex = NS::sampler2D{(S).envCdf.data(), (S).envW, (S).envH, 1, false};           /* :222 NEAREST */             \
        NS::envMapRes = NS::vec2((float)(S).envW, (float)(S).envH);                                                                 \
        NS::envMapTotalSum = (S).envTotalSum;                                                                                       \
        NS::topBVHIndex = (S).topLevelIndex;                                                                                        \
        NS::resolution = NS::vec2((float)(O).renderW, (float)(O).renderH);                                                          \
        NS::numOfLights = (S).numLights;                                                                                            \
        NS::camera.position = NS::vec3((O).camPosition[0], (O).camPosition[1], (O).camPosition[2]);                                 \
        NS::camera.right = NS::vec3((O).camRight[0], (O).camRight[1], (O).camRight[2]);                                             \
        NS::camera.up = NS::vec3((O).camUp[0], (O).camUp[1], (O).camUp[2]);                                                         \
        NS::camera.forward = NS::vec3((O).camForward[0], (O).camForward[1], (O).camForward[2]);                                     \
        NS::camera.fov = (O).camFov;  NS::camera.focalDist = (O).camFocalDist;  NS::camera.aperture = (O).camAperture;              \
        NS::envMapIntensity = (O).envMapIntensity;                                                                                  \
        NS::envMapRot = (O).envMapRot;                                                                                              \
        NS::maxDepth = (MAXDEPTH);                                                                                                  \
        NS::uniformLightCol = NS::vec3((O).uniformLightCol[0], (O).uniformLightCol[1], (O).uniformLightCol[2]);                     \
        NS::roughnessMollificationAmt = (O).roughnessMollificationAmt;                                                              \
    } while (0)

void bindAll(Scene& s)
{
    const OrcOptions& o = s.opt;
    BIND_PATHTRACE_PROGRAM(tile_prog, s, o, o.maxDepth);
    BIND_PATHTRACE_PROGRAM(preview_prog, s, o, 2);                       // Renderer.cpp:798: maxDepth = dirty ? 2 : maxDepth
    tile_prog::invNumTiles = tile_prog::vec2((float)o.tileW / o.renderW, (float)o.tileH / o.renderH);   // Renderer.cpp:293-294
}

int numTilesX(const OrcOptions& o) { return (int)std::ceil((float)o.renderW / o.tileW); }   // Renderer.cpp:296-297
int numTilesY(const OrcOptions& o) { return (int)std::ceil((float)o.renderH / o.tileH); }

// One Renderer::Render() of the tile branch (Renderer.cpp:566-580), restricted to global pixels [x0,x1) x [y0,y1).
void drawTile(Scene& s, int tx, int ty, int frameNum, float* accum, int x0, int y0, int x1, int y1)
{
    const OrcOptions& o = s.opt;
    const int W = o.renderW, H = o.renderH, tw = o.tileW, th = o.tileH;
    tile_prog::tileOffset = tile_prog::vec2((float)tx * tile_prog::invNumTiles.x, (float)ty * tile_prog::invNumTiles.y);   // :777
    tile_prog::frameNum = frameNum;                                                                                       // :783
    tile_prog::accumTexture = tile_prog::sampler2D{accum, W, H, 4, false};
    const int lx0 = std::max(0, x0 - tx * tw), lx1 = std::min(tw, std::min(x1, W) - tx * tw);
    const int ly0 = std::max(0, y0 - ty * th), ly1 = std::min(th, std::min(y1, H) - ty * th);
    if (lx1 <= lx0 || ly1 <= ly0) return;
    const int nw = lx1 - lx0, nh = ly1 - ly0;
    std::vector<float> pathTraceTexture((size_t)nw * nh * 4);
#pragma omp parallel for schedule(dynamic, 4) collapse(2)
    for (int ly = ly0; ly < ly1; ly++)
        for (int lx = lx0; lx < lx1; lx++)
        {
            tile_prog::gl_FragCoord = tile_prog::vec4((float)lx + 0.5f, (float)ly + 0.5f, 0.5f, 1.0f);
            tile_prog::TexCoords = tile_prog::vec2(((float)lx + 0.5f) / (float)tw, ((float)ly + 0.5f) / (float)th);
            tile_prog::glsl_main();
            float* p = &pathTraceTexture[((size_t)(ly - ly0) * nw + (lx - lx0)) * 4];
            p[0] = tile_prog::color.x; p[1] = tile_prog::color.y; p[2] = tile_prog::color.z; p[3] = tile_prog::color.w;
        }
    // outputShader blit of pathTraceTexture into accumFBO at the tile's viewport (:574-578), clipped to the FBO
    for (int ly = ly0; ly < ly1; ly++)
        memcpy(&accum[((size_t)(ty * th + ly) * W + tx * tw + lx0) * 4], &pathTraceTexture[(size_t)(ly - ly0) * nw * 4], (size_t)nw * 16);
}

}  // namespace

extern "C" {

void gref_defines(int32_t* out13) { memcpy(out13, &kDefs, sizeof(kDefs)); }

// Returns 0 on success, -1 if the options' feature set is not the one this object was compiled for.
int gref_create(const OrcSceneDesc* d, const OrcOptions* o)
{
    if (!optsMatchDefines(*o)) return -1;
    delete g_scene;
    Scene* s = g_scene = new Scene();
    s->nodes.assign(d->nodes, d->nodes + (size_t)d->numNodes * 9);
    s->vertIndices.assign(d->vertIndices, d->vertIndices + (size_t)d->numIndices * 3);
    s->verticesUVX.assign(d->verticesUVX, d->verticesUVX + (size_t)d->numVertices * 4);
    s->normalsUVY.assign(d->normalsUVY, d->normalsUVY + (size_t)d->numVertices * 4);
    s->materials.assign(d->materials, d->materials + (size_t)d->numMaterials * 32);
    s->transforms.assign(d->transforms, d->transforms + (size_t)d->numInstances * 16);
    if (d->numLights) s->lights.assign(d->lights, d->lights + (size_t)d->numLights * 15);
    if (d->numTextures) s->textures.assign(d->textures, d->textures + (size_t)d->numTextures * d->texW * d->texH * 4);
    if (d->envImg) { s->envImg.assign(d->envImg, d->envImg + (size_t)d->envW * d->envH * 3); s->envCdf.assign(d->envCdf, d->envCdf + (size_t)d->envW * d->envH); }
    s->numMaterials = d->numMaterials; s->numInstances = d->numInstances; s->numLights = d->numLights;
    s->numTextures = d->numTextures; s->texW = d->texW; s->texH = d->texH; s->envW = d->envW; s->envH = d->envH;
    s->envTotalSum = d->envTotalSum; s->topLevelIndex = d->topLevelIndex;
    s->opt = *o;
    bindAll(*s);
    return 0;
}
int gref_set_options(const OrcOptions* o)
{
    if (!g_scene || !optsMatchDefines(*o)) return -1;
    g_scene->opt = *o;
    bindAll(*g_scene);
    return 0;
}
void gref_destroy(void) { delete g_scene; g_scene = nullptr; }

// One tile draw with the frameNum uniform the caller's Update() would have set.
void gref_render_tile(int32_t tx, int32_t ty, int32_t frameNum, float* accum)
{
    drawTile(*g_scene, tx, ty, frameNum, accum, 0, 0, g_scene->opt.renderW, g_scene->opt.renderH);
}

// Full-frame sample passes firstSample .. firstSample+n-1 (1-based) with the reference's tile walk and frame counter
// (Renderer.cpp:745-762: x fastest, y from the top row down, frameCounter++ per Update; the first Update is the dirty one,
// so tile ordinal j of pass s is drawn with frameNum = 2 + (s-1)*T + j), restricted to a pixel rectangle.
void gref_render_samples_rect(int32_t firstSample, int32_t nSamples, int32_t x0, int32_t y0, int32_t x1, int32_t y1, float* accum)
{
    Scene& s = *g_scene;
    const int ntx = numTilesX(s.opt), nty = numTilesY(s.opt), T = ntx * nty;
    for (int sp = firstSample; sp < firstSample + nSamples; sp++)
    {
        int j = 0;
        for (int ty = nty - 1; ty >= 0; ty--)
            for (int tx = 0; tx < ntx; tx++, j++)
                drawTile(s, tx, ty, 2 + (sp - 1) * T + j, accum, x0, y0, x1, y1);
    }
}
void gref_render_samples(int32_t firstSample, int32_t nSamples, float* accum)
{
    gref_render_samples_rect(firstSample, nSamples, 0, 0, g_scene->opt.renderW, g_scene->opt.renderH, accum);
}

// Renderer.cpp:555-560: the dirty-scene preview draw into a w x h target (windowSize * pixelRatio).
void gref_render_preview(int32_t w, int32_t h, float* out)
{
#pragma omp parallel for schedule(dynamic, 4) collapse(2)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
        {
            preview_prog::gl_FragCoord = preview_prog::vec4((float)x + 0.5f, (float)y + 0.5f, 0.5f, 1.0f);
            preview_prog::TexCoords = preview_prog::vec2(((float)x + 0.5f) / (float)w, ((float)y + 0.5f) / (float)h);
            preview_prog::glsl_main();
            float* p = &out[((size_t)y * w + x) * 4];
            p[0] = preview_prog::color.x; p[1] = preview_prog::color.y; p[2] = preview_prog::color.z; p[3] = preview_prog::color.w;
        }
}

// Renderer.cpp:582-588 (tonemap draw over the whole frame) + :619-634 (RGBA8 readback).  OPT_BACKGROUND /
// OPT_TRANSPARENT_BACKGROUND are compile-time, as in the reference (tonemapDefines, Renderer.cpp:425-435).
void gref_tonemap(const float* accum, int32_t w, int32_t h, float invSampleCounter, int32_t enableTonemap, int32_t enableAces,
                  int32_t simpleAcesFit, const float* backgroundCol, uint8_t* out)
{
    tonemap_prog::pathTraceTexture = tonemap_prog::sampler2D{accum, w, h, 4, false};
    tonemap_prog::invSampleCounter = invSampleCounter;
    tonemap_prog::enableTonemap = enableTonemap != 0;
    tonemap_prog::enableAces = enableAces != 0;
    tonemap_prog::simpleAcesFit = simpleAcesFit != 0;
    tonemap_prog::backgroundCol = tonemap_prog::vec3(backgroundCol[0], backgroundCol[1], backgroundCol[2]);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
        {
            tonemap_prog::gl_FragCoord = tonemap_prog::vec4((float)x + 0.5f, (float)y + 0.5f, 0.5f, 1.0f);
            tonemap_prog::TexCoords = tonemap_prog::vec2(((float)x + 0.5f) / (float)w, ((float)y + 0.5f) / (float)h);
            tonemap_prog::glsl_main();
            const float c[4] = {tonemap_prog::outCol.x, tonemap_prog::outCol.y, tonemap_prog::outCol.z, tonemap_prog::outCol.w};
            uint8_t* p = &out[((size_t)y * w + x) * 4];
            for (int k = 0; k < 4; k++)
            {
                float f = c[k];
                f = !(f > 0.0f) ? 0.0f : (f > 1.0f ? 1.0f : f);     // NaN -> 0
                p[k] = (uint8_t)std::floor(f * 255.0f + 0.5f);
            }
        }
}

// ---- probes: the shader's own ClosestHit / AnyHit / DisneyEval called directly (G1 / G2 pins) -----------------------------
// rays: n x 6 floats.  t = state.hitDist (1e6 on a miss), kind 0 miss / 1 triangle / 2 analytic light (state.isEmitter),
// matID = state.matID for triangle hits (else -1).
void gref_trace_closest(const float* rays, int64_t n, int32_t depth, float* t, int32_t* kind, int32_t* matID)
{
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t i = 0; i < n; i++)
    {
        using namespace tile_prog;
        Ray r = Ray(vec3(rays[i * 6 + 0], rays[i * 6 + 1], rays[i * 6 + 2]), vec3(rays[i * 6 + 3], rays[i * 6 + 4], rays[i * 6 + 5]));
        State state; state.depth = depth; state.matID = -1;
        LightSampleRec lightSample;
        bool hit = ClosestHit(r, state, lightSample);
        t[i] = hit ? state.hitDist : 1000000.0f;
        kind[i] = !hit ? 0 : (state.isEmitter ? 2 : 1);
        matID[i] = (hit && !state.isEmitter) ? state.matID : -1;
    }
}
// AnyHit(r, maxDist) for rays whose alpha test (if compiled in) draws nothing from the RNG: MASK / OPAQUE materials.
void gref_trace_any(const float* rays, const float* maxDist, int64_t n, int32_t* out)
{
#pragma omp parallel for schedule(dynamic, 1024)
    for (int64_t i = 0; i < n; i++)
    {
        using namespace tile_prog;
        Ray r = Ray(vec3(rays[i * 6 + 0], rays[i * 6 + 1], rays[i * 6 + 2]), vec3(rays[i * 6 + 3], rays[i * 6 + 4], rays[i * 6 + 5]));
        out[i] = AnyHit(r, maxDist[i]) ? 1 : 0;
    }
}
// DisneyEval(state, V, N, L, pdf) with state.mat produced by the shader's own GetMaterial from the query's material row
// (texture slots disabled), state.eta as given.  Serial: GetMaterial reads the materialsTex uniform.
void gref_bsdf_eval(const OrcBsdfQuery* q, int64_t n, OrcBsdfResult* out)
{
    using namespace tile_prog;
    const sampler2D saved = materialsTex;
    for (int64_t i = 0; i < n; i++)
    {
        float row[32]; memcpy(row, q[i].mat, sizeof(row));
        row[24] = row[25] = row[26] = row[27] = -1.0f;
        materialsTex = sampler2D{row, 8, 1, 4, false};
        State state; state.matID = 0; state.depth = 0;
        state.normal = vec3(q[i].N[0], q[i].N[1], q[i].N[2]); state.ffnormal = state.normal;
        Ray r = Ray(vec3(0.0f), -vec3(q[i].V[0], q[i].V[1], q[i].V[2]));
        GetMaterial(state, r);
        state.eta = q[i].eta;
        float pdf = 0.0f;
        vec3 f = DisneyEval(state, vec3(q[i].V[0], q[i].V[1], q[i].V[2]), state.ffnormal, vec3(q[i].L[0], q[i].L[1], q[i].L[2]), pdf);
        out[i].f[0] = f.x; out[i].f[1] = f.y; out[i].f[2] = f.z; out[i].pdf = pdf;
        out[i].L[0] = q[i].L[0]; out[i].L[1] = q[i].L[1]; out[i].L[2] = q[i].L[2];
    }
    materialsTex = saved;
}

// The shader's own LambertEval (lambert.glsl:41-46) on the same queries.
void gref_lambert_eval(const OrcBsdfQuery* q, int64_t n, OrcBsdfResult* out)
{
    using namespace tile_prog;
    for (int64_t i = 0; i < n; i++)
    {
        State state;
        state.mat.baseColor = vec3(q[i].mat[0], q[i].mat[1], q[i].mat[2]);
        float pdf = 0.0f;
        vec3 f = LambertEval(state, vec3(q[i].V[0], q[i].V[1], q[i].V[2]), vec3(q[i].N[0], q[i].N[1], q[i].N[2]), vec3(q[i].L[0], q[i].L[1], q[i].L[2]), pdf);
        out[i].f[0] = f.x; out[i].f[1] = f.y; out[i].f[2] = f.z; out[i].pdf = pdf;
        out[i].L[0] = q[i].L[0]; out[i].L[1] = q[i].L[1]; out[i].L[2] = q[i].L[2];
    }
}
// The shader's own LambertSample (lambert.glsl:25-39); it draws r1, r2 from the shader RNG, so the seed is given per query and the
// two draws are returned for the caller to feed the oracle / CUDA entry points.
void gref_lambert_sample(const OrcBsdfQuery* q, int64_t n, const uint32_t* seeds4, OrcBsdfResult* out, float* r12)
{
    using namespace tile_prog;
    for (int64_t i = 0; i < n; i++)
    {
        State state;
        state.mat.baseColor = vec3(q[i].mat[0], q[i].mat[1], q[i].mat[2]);
        seed = uvec4(seeds4[i * 4 + 0], seeds4[i * 4 + 1], seeds4[i * 4 + 2], seeds4[i * 4 + 3]);
        r12[i * 2 + 0] = tile_prog::rand(); r12[i * 2 + 1] = tile_prog::rand();
        seed = uvec4(seeds4[i * 4 + 0], seeds4[i * 4 + 1], seeds4[i * 4 + 2], seeds4[i * 4 + 3]);
        float pdf = 0.0f; vec3 L;
        vec3 f = LambertSample(state, vec3(q[i].V[0], q[i].V[1], q[i].V[2]), vec3(q[i].N[0], q[i].N[1], q[i].N[2]), L, pdf);
        out[i].f[0] = f.x; out[i].f[1] = f.y; out[i].f[2] = f.z; out[i].pdf = pdf; out[i].L[0] = L.x; out[i].L[1] = L.y; out[i].L[2] = L.z;
    }
}

int gref_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

}  // extern "C"
