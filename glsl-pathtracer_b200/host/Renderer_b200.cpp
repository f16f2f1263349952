// Renderer_b200.cpp — drop-in replacement for the reference's src/core/Renderer.cpp.
//
// It implements the class `GLSLPT::Renderer` exactly as DECLARED by the reference's own, unmodified src/core/Renderer.h
// (Renderer.h:104-185): same constructor, same public methods, same state machine.  Only the body changes: the OpenGL texture
// uploads, FBOs, GLSL compilation and the three draws per tile are replaced by calls into libptb200.so (include/ptb200.h).
// A maintainer swaps this file for Renderer.cpp and links -lptb200 instead of OpenGL/OIDN; Scene, the loaders, RadeonRays and
// Main.cpp's call sites (Main.cpp:160,168,177,180,212,289,308,520) compile unchanged.
//
// The GL handle members of the class stay zero.  State the reference class has no member for (the PtbCtx, the tonemap uniform,
// the completed image) lives in a side table keyed by `this`.
#include "Config.h"
#include "Renderer.h"
#include "Scene.h"
#include "Camera.h"
#include "ptb200.h"
#include "RendererB200Ext.h"
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <stdexcept>
#include <vector>

namespace GLSLPT
{
    namespace
    {
        struct Side
        {
            PtbMgpu* m = nullptr;                          // one context per configured GPU (default: one, device PTB_DEVICE or 0)
            PtbCtx* ctx0 = nullptr;                        // context of the first GPU: previews, single tiles, the frozen output image
            float invSampleCounter = 1.0f;                 // tonemap uniform (Renderer.cpp:806)
            std::vector<float> preview;                    // pathTraceTextureLowRes
            int samplesPerWave = 0;
            uint32_t features = 0;                         // OPT_* set of the last InitShaders / ReloadShaders (Renderer.cpp:401-459)
            bool coalesce = true;                          // a whole sample pass per first-tile Render() (PTB_COALESCE=0: one wave per tile)
            bool passCoalesced = false;                    // the pass being walked was rendered at its first tile
            std::vector<float> denoiseIn, denoiseOut;      // denoiserInputFramePtr / frameOutputPtr (Renderer.cpp:347-348)
            PtbCamera sentCam; PtbOptions sentOpts; bool sentValid = false;     // uniforms last pushed to the contexts: Update() runs once per TILE, most calls change nothing
        };
        std::map<const Renderer*, Side>& table() { static std::map<const Renderer*, Side> t; return t; }

        // process-wide configuration set before a Renderer is constructed (RendererB200Ext.h)
        std::vector<int>& configuredDevices() { static std::vector<int> d; return d; }
        DenoiseFnB200& denoiser() { static DenoiseFnB200 f = nullptr; return f; }
        void*& denoiserUser() { static void* u = nullptr; return u; }

        std::vector<int> devicesFromEnv()
        {   // SetDevicesB200() > PTB_DEVICES="0,1,2,3" > PTB_DEVICE="1" > device 0
            if (!configuredDevices().empty()) return configuredDevices();
            std::vector<int> d;
            if (const char* e = getenv("PTB_DEVICES"))
                for (const char* p = e; *p;) { d.push_back(atoi(p)); while (*p && *p != ',') p++; if (*p == ',') p++; }
            else if (const char* e1 = getenv("PTB_DEVICE")) d.push_back(atoi(e1));
            if (d.empty()) d.push_back(0);
            return d;
        }

        void check(int rc, const char* what)
        {
            if (rc != PTB_OK)
            {   // the reference throws std::runtime_error on shader failures (Shader.cpp:54-55); same convention for device failures
                printf("%s failed: %s\n", what, ptb_last_error());
                throw std::runtime_error(std::string(what) + ": " + ptb_last_error());
            }
        }

        PtbSceneDesc describe(Scene* s)
        {   // the arrays InitGPUDataBuffers uploads (Renderer.cpp:135-249)
            PtbSceneDesc d; memset(&d, 0, sizeof(d));
            d.nodes = (const float*)s->bvhTranslator.nodes.data(); d.numNodes = (int)s->bvhTranslator.nodes.size();
            d.topLevelIndex = s->bvhTranslator.topLevelIndex;
            d.vertIndices = (const int32_t*)s->vertIndices.data(); d.numIndices = (int)s->vertIndices.size();
            d.verticesUVX = (const float*)s->verticesUVX.data(); d.numVertices = (int)s->verticesUVX.size();
            d.normalsUVY = (const float*)s->normalsUVY.data();
            d.materials = (const float*)s->materials.data(); d.numMaterials = (int)s->materials.size();
            d.transforms = (const float*)s->transforms.data(); d.numInstances = (int)s->transforms.size();
            d.lights = s->lights.empty() ? nullptr : (const float*)s->lights.data(); d.numLights = (int)s->lights.size();
            d.textures = s->textures.empty() ? nullptr : s->textureMapsArray.data(); d.numTextures = (int)s->textures.size();
            d.texW = s->renderOptions.texArrayWidth; d.texH = s->renderOptions.texArrayHeight;
            if (s->envMap) { d.envImg = s->envMap->img; d.envCdf = s->envMap->cdf; d.envW = s->envMap->width; d.envH = s->envMap->height; d.envTotalSum = s->envMap->totalSum; }
            return d;
        }

        uint32_t deriveFeatures(Scene* s);

        PtbOptions options(Scene* s, int samplesPerWave, uint32_t features)
        {   // the uniforms of Update (Renderer.cpp:776-782, 806-810) + the feature set the shaders were last compiled with
            const RenderOptions& ro = s->renderOptions;
            PtbOptions o; memset(&o, 0, sizeof(o));
            o.renderW = ro.renderResolution.x; o.renderH = ro.renderResolution.y; o.tileW = ro.tileWidth; o.tileH = ro.tileHeight;
            o.maxDepth = ro.maxDepth; o.rrDepth = ro.RRDepth;
            o.features = features;
            o.envMapIntensity = ro.envMapIntensity; o.envMapRot = ro.envMapRot; o.roughnessMollificationAmt = ro.roughnessMollificationAmt;
            o.uniformLightCol[0] = ro.uniformLightCol.x; o.uniformLightCol[1] = ro.uniformLightCol.y; o.uniformLightCol[2] = ro.uniformLightCol.z;
            o.backgroundCol[0] = ro.backgroundCol.x; o.backgroundCol[1] = ro.backgroundCol.y; o.backgroundCol[2] = ro.backgroundCol.z;
            o.enableTonemap = ro.enableTonemap; o.enableAces = ro.enableAces; o.simpleAcesFit = ro.simpleAcesFit;
            o.samplesPerWave = samplesPerWave;
            return o;
        }

        uint32_t deriveFeatures(Scene* s)
        {   // the #defines of InitShaders (Renderer.cpp:401-459): evaluated where the reference compiles shaders, not per frame
            const RenderOptions& ro = s->renderOptions;
            uint32_t bools = (ro.enableEnvMap ? 1u : 0u) | (ro.enableRR ? 2u : 0u) | (ro.enableUniformLight ? 4u : 0u) | (ro.openglNormalMap ? 8u : 0u) |
                             (ro.hideEmitters ? 16u : 0u) | (ro.enableBackground ? 32u : 0u) | (ro.transparentBackground ? 64u : 0u) |
                             (ro.enableRoughnessMollification ? 128u : 0u) | (ro.enableVolumeMIS ? 256u : 0u);
            PtbSceneDesc d = describe(s);
            return ptb_derive_features(&d, bools);
        }

        PtbCamera camera(Scene* s)
        {   // camera.* uniforms (Renderer.cpp:769-775)
            Camera* c = s->camera;
            PtbCamera p;
            p.position[0] = c->position.x; p.position[1] = c->position.y; p.position[2] = c->position.z;
            p.right[0] = c->right.x; p.right[1] = c->right.y; p.right[2] = c->right.z;
            p.up[0] = c->up.x; p.up[1] = c->up.y; p.up[2] = c->up.z;
            p.forward[0] = c->forward.x; p.forward[1] = c->forward.y; p.forward[2] = c->forward.z;
            p.fov = c->fov; p.focalDist = c->focalDist; p.aperture = c->aperture;
            return p;
        }
    }

    Renderer::Renderer(Scene* scene, const std::string& shadersDirectory)
        : scene(scene), quad(nullptr), BVHBuffer(0), BVHTex(0), vertexIndicesBuffer(0), vertexIndicesTex(0), verticesBuffer(0), verticesTex(0),
          normalsBuffer(0), normalsTex(0), materialsTex(0), transformsTex(0), lightsTex(0), textureMapsArrayTex(0), envMapTex(0), envMapCDFTex(0),
          pathTraceFBO(0), pathTraceFBOLowRes(0), accumFBO(0), outputFBO(0), shadersDirectory(shadersDirectory), pathTraceShader(nullptr),
          pathTraceShaderLowRes(nullptr), outputShader(nullptr), tonemapShader(nullptr), pathTraceTextureLowRes(0), pathTraceTexture(0), accumTexture(0),
          tileOutputTexture(), denoisedTexture(0), denoiserInputFramePtr(nullptr), frameOutputPtr(nullptr), denoised(false), initialized(false)
    {
        if (scene == nullptr)
        {
            printf("No Scene Found\n");                      // Renderer.cpp:72-76
            return;
        }
        if (!scene->initialized)
            scene->ProcessScene();                           // Renderer.cpp:78-79

        Side& sd = table()[this];
        PtbSceneDesc d = describe(scene);                    // InitGPUDataBuffers
        sd.features = deriveFeatures(scene);                 // InitShaders: feature selection instead of GLSL compilation
        PtbOptions o = options(scene, sd.samplesPerWave, sd.features);
        std::vector<int> devs = devicesFromEnv();
        check(ptb_mgpu_create(&d, &o, devs.data(), (int)devs.size(), &sd.m), "ptb_mgpu_create");
        sd.ctx0 = ptb_mgpu_context(sd.m, 0);
        if (const char* e = getenv("PTB_COALESCE")) sd.coalesce = atoi(e) != 0;
        pixelRatio = 0.25f;
        InitFBOs();
        initialized = true;
    }

    Renderer::~Renderer()
    {
        auto it = table().find(this);
        if (it != table().end()) { ptb_mgpu_destroy(it->second.m); table().erase(it); }
    }

    void Renderer::InitGPUDataBuffers() {}                   // done by ptb_create
    void Renderer::InitShaders() { ReloadShaders(); }

    void Renderer::InitFBOs()
    {   // counters and tile grid of Renderer.cpp:281-300; the buffers themselves live in the PtbCtx
        sampleCounter = 1; currentBuffer = 0; frameCounter = 1;
        renderSize = scene->renderOptions.renderResolution; windowSize = scene->renderOptions.windowResolution;
        tileWidth = scene->renderOptions.tileWidth; tileHeight = scene->renderOptions.tileHeight;
        invNumTiles.x = (float)tileWidth / renderSize.x; invNumTiles.y = (float)tileHeight / renderSize.y;
        numTiles.x = ceil((float)renderSize.x / tileWidth); numTiles.y = ceil((float)renderSize.y / tileHeight);
        tile.x = -1; tile.y = numTiles.y - 1;
        Side& sd = table()[this];
        sd.invSampleCounter = 1.0f; sd.passCoalesced = false;
        printf("Window Resolution : %d %d\n", windowSize.x, windowSize.y);
        printf("Render Resolution : %d %d\n", renderSize.x, renderSize.y);
        printf("Preview Resolution : %d %d\n", (int)((float)windowSize.x * pixelRatio), (int)((float)windowSize.y * pixelRatio));
        printf("Tile Size : %d %d\n", tileWidth, tileHeight);
    }

    void Renderer::ResizeRenderer()
    {   // Renderer.cpp:251-279
        Side& sd = table()[this];
        PtbOptions o = options(scene, sd.samplesPerWave, sd.features);
        check(ptb_mgpu_set_options(sd.m, &o), "ptb_set_options");     // new size: buffers reallocated, frozen image dropped
        sd.sentValid = false;
        check(ptb_mgpu_reset_accum(sd.m), "ptb_reset_accum");         // the reference re-creates (clears) every FBO texture
        InitFBOs();
    }

    void Renderer::ReloadShaders()
    {   // Renderer.cpp:381-390: re-derive the OPT_* set
        Side& sd = table()[this];
        sd.features = deriveFeatures(scene);
        PtbOptions o = options(scene, sd.samplesPerWave, sd.features);
        check(ptb_mgpu_set_options(sd.m, &o), "ptb_set_options");
        sd.sentValid = false;
    }

    void Renderer::Render()
    {   // Renderer.cpp:546-590
        if (!scene->dirty && scene->renderOptions.maxSpp != -1 && sampleCounter >= scene->renderOptions.maxSpp)
            return;
        Side& sd = table()[this];
        if (scene->dirty)
        {
            int w = (int)(windowSize.x * pixelRatio), h = (int)(windowSize.y * pixelRatio);
            sd.preview.resize((size_t)w * h * 4);
            check(ptb_render_preview(sd.ctx0, w, h, sd.preview.data()), "ptb_render_preview");
            scene->instancesModified = false;
            scene->dirty = false;
            scene->envMapModified = false;
            return;
        }
        // Only the buffer of a COMPLETED pass is observable (GetOutputBuffer reads tileOutputTexture[1-currentBuffer],
        // Renderer.cpp:628-633), so the numTiles draws of a pass are coalesced: the whole pass is rendered as one wavefront when its
        // first tile is drawn — same frameNum per tile and same tile-local seeds as the tile walk would use (ptb_render_pass) — and
        // the remaining tiles of the pass draw nothing.  A pass that does not start on the regular schedule falls back to one wave per tile.
        const int T = numTiles.x * numTiles.y;
        const bool firstTile = tile.x == 0 && tile.y == numTiles.y - 1;
        if (firstTile)
        {
            sd.passCoalesced = sd.coalesce && frameCounter == 2 + (sampleCounter - 1) * T;
            if (sd.passCoalesced)
            {
                const int maxSpp = scene->renderOptions.maxSpp;
                const int N = ptb_mgpu_num_devices(sd.m);
                const int remaining = maxSpp == -1 ? 0 : (maxSpp - sampleCounter + N - 1) / N;      // passes this GPU still has to render (Q1: maxSpp-1 in total)
                check(ptb_mgpu_render_pass(sd.m, sampleCounter, remaining), "ptb_render_pass");
            }
        }
        if (!sd.passCoalesced)
            check(ptb_render_tile(sd.ctx0, tile.x, tile.y, frameCounter), "ptb_render_tile");
    }

    void Renderer::Present() {}                              // Renderer.cpp:592-611 draws to the window: nothing to do headless

    float Renderer::GetProgress()
    {
        int maxSpp = scene->renderOptions.maxSpp;
        return maxSpp <= 0 ? 0.0f : sampleCounter * 100.0f / maxSpp;
    }

    void Renderer::GetOutputBuffer(unsigned char** data, int& w, int& h)
    {   // Renderer.cpp:619-634: caller delete[]s
        w = renderSize.x; h = renderSize.y;
        *data = new unsigned char[w * h * 4];
        Side& sd = table()[this];
        check(ptb_mgpu_read_snapshot_rgba8(sd.m, *data), "ptb_read_snapshot_rgba8");     // the glGetTexImage: the only host copy of the image
    }

    int Renderer::GetSampleCount() { return sampleCounter; }

    void Renderer::Update(float secondsElapsed)
    {   // Renderer.cpp:641-812
        (void)secondsElapsed;
        if (!scene->dirty && scene->renderOptions.maxSpp != -1 && sampleCounter >= scene->renderOptions.maxSpp)
            return;
        Side& sd = table()[this];
        if (scene->instancesModified)
        {
            int top = scene->bvhTranslator.topLevelIndex;
            check(ptb_mgpu_update_instances(sd.m, (const float*)scene->transforms.data(), (int)scene->transforms.size(), (const float*)scene->materials.data(),
                                       (int)scene->materials.size(), (const float*)&scene->bvhTranslator.nodes[top],
                                       (int)scene->bvhTranslator.nodes.size() - top), "ptb_update_instances");
        }
        if (scene->envMapModified && scene->envMap != nullptr)
            check(ptb_mgpu_update_envmap(sd.m, scene->envMap->img, scene->envMap->cdf, scene->envMap->width, scene->envMap->height, scene->envMap->totalSum),
                  "ptb_update_envmap");
        // Denoiser hook (Renderer.cpp:695-730): same trigger and cadence as the reference; the filter itself (OIDN there) is whatever
        // the host registered with SetDenoiserB200 — without one the branch leaves `denoised` false, as a build without OIDN would.
        check(ptb_set_snapshot_float(sd.ctx0, (scene->renderOptions.enableDenoiser && denoiser()) ? 1 : 0), "ptb_set_snapshot_float");   // keep the RGBA32F form of frozen images for the hook
        if (scene->renderOptions.enableDenoiser && sampleCounter > 1 && denoiser())
        {
            if (!denoised || (frameCounter % (scene->renderOptions.denoiserFrameCnt * (numTiles.x * numTiles.y)) == 0))
            {
                const size_t n = (size_t)renderSize.x * renderSize.y * 3;
                sd.denoiseIn.resize(n); sd.denoiseOut.resize(n);
                check(ptb_mgpu_read_snapshot_rgb32f(sd.m, sd.denoiseIn.data()), "ptb_read_snapshot_rgb32f");    // glGetTexImage(GL_RGB, GL_FLOAT)
                denoiser()(sd.denoiseIn.data(), sd.denoiseOut.data(), renderSize.x, renderSize.y, denoiserUser());
                denoised = true;
            }
        }
        else
            denoised = false;
        if (scene->dirty)
        {
            tile.x = -1; tile.y = numTiles.y - 1;
            sampleCounter = 1; denoised = false; frameCounter = 1;
            check(ptb_mgpu_reset_accum(sd.m), "ptb_reset_accum");
        }
        else
        {
            frameCounter++;
            tile.x++;
            if (tile.x >= numTiles.x)
            {
                tile.x = 0;
                tile.y--;
                if (tile.y < 0)
                {
                    tile.x = 0;
                    tile.y = numTiles.y - 1;
                    // the buffer tonemapped with the finished pass's uniform becomes the displayed one (Renderer.cpp:584-588,758)
                    // Frozen on the device (asynchronous; with several GPUs: one NCCL reduce into a scratch sum first); copied to the host only
                    // when GetOutputBuffer asks.
                    check(ptb_mgpu_snapshot_output(sd.m, sd.invSampleCounter), "ptb_snapshot_output");
                    sampleCounter++;
                    currentBuffer = 1 - currentBuffer;
                }
            }
        }
        PtbCamera cam = camera(scene);
        PtbOptions o = options(scene, sd.samplesPerWave, sd.features);      // uniforms only: the OPT_* set changes in ReloadShaders, as in the reference
        if (!sd.sentValid || memcmp(&cam, &sd.sentCam, sizeof(cam)) != 0 || memcmp(&o, &sd.sentOpts, sizeof(o)) != 0)
        {   // the reference sets its uniforms on every Update (Renderer.cpp:766-811); here they are pushed to the N contexts when they changed
            check(ptb_mgpu_set_camera(sd.m, &cam), "ptb_set_camera");
            check(ptb_mgpu_set_options(sd.m, &o), "ptb_set_options");
            sd.sentCam = cam; sd.sentOpts = o; sd.sentValid = true;
        }
        sd.invSampleCounter = 1.0f / (sampleCounter);
    }

    // ---- extension (RendererB200Ext.h): whole-frame passes with the seeds of the tile schedule -----------------------
    void RenderSamplesB200(Renderer& r, Scene* scene, int n)
    {
        Side& sd = table()[&r];
        PtbCamera cam = camera(scene);
        check(ptb_mgpu_set_camera(sd.m, &cam), "ptb_set_camera");
        sd.sentValid = false;
        int first = r.GetSampleCount();
        check(ptb_mgpu_render_samples(sd.m, first, n), "ptb_render_samples");
        RendererB200Access::advance(r, n);
        check(ptb_mgpu_snapshot_output(sd.m, 1.0f / (float)(first + n - 1)), "ptb_snapshot_output");
    }
    void RebuildInstancesB200(Renderer& r, Scene* scene)
    {
        Side& sd = table()[&r];
        const int n = (int)scene->meshInstances.size();
        std::vector<int32_t> mats((size_t)n);
        for (int i = 0; i < n; i++) { scene->transforms[i] = scene->meshInstances[i].transform; mats[(size_t)i] = scene->meshInstances[i].materialID; }     // Scene.cpp:208-210
        check(ptb_mgpu_rebuild_instances(sd.m, (const float*)scene->transforms.data(), n, (const float*)scene->materials.data(), (int)scene->materials.size(), mats.data(), 0),
              "ptb_rebuild_instances");
        // keep the Scene's copy of the node array in step with the devices' (what bvhTranslator.UpdateTLAS would have written)
        std::vector<float> nodes(scene->bvhTranslator.nodes.size() * 9);
        check(ptb_read_nodes(sd.ctx0, nodes.data(), (int)scene->bvhTranslator.nodes.size()), "ptb_read_nodes");
        const int top = scene->bvhTranslator.topLevelIndex;
        memcpy(&scene->bvhTranslator.nodes[top], &nodes[(size_t)top * 9], (scene->bvhTranslator.nodes.size() - (size_t)top) * 9 * sizeof(float));
        scene->instancesModified = false;
        scene->dirty = true;
    }
    PtbCtx* ContextOfB200(Renderer& r) { return table()[&r].ctx0; }
    PtbMgpu* MgpuOfB200(Renderer& r) { return table()[&r].m; }
    void SetDevicesB200(const int* devices, int n) { configuredDevices().assign(devices, devices + (n > 0 ? n : 0)); }
    void SetDenoiserB200(DenoiseFnB200 fn, void* user) { denoiser() = fn; denoiserUser() = user; }
    const float* DenoisedImageB200(Renderer& r) { Side& sd = table()[&r]; return sd.denoiseOut.empty() ? nullptr : sd.denoiseOut.data(); }
}
