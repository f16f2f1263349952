// scene_dump — TEST INFRASTRUCTURE (oracle side), not product code.
//
// Links the UNMODIFIED reference host code, compiled from the sources where they
// lie under /root/reference (src/core/Scene.cpp, Mesh.cpp, Camera.cpp,
// EnvironmentMap.cpp, Texture.cpp, src/loaders/Loader.cpp, GLTFLoader.cpp,
// thirdparty/RadeonRays/*.cpp), loads a .scene file exactly as Main.cpp:121-155
// does, runs Scene::ProcessScene() (Scene.cpp:216) and writes every array that
// Renderer::InitGPUDataBuffers (Renderer.cpp:135-249) would upload into a
// ".ptscene" blob.  The blob is the data contract of include/ptb200.h and the
// bit-exact "G1" fixture (flattened BVH + mesh arrays) for the parity tests.
//
// Blob layout (little endian):
//   char[8]  "PTBSCN01"
//   u32      nsections
//   nsections x { char[16] name ; u64 nbytes ; u64 offset-from-file-start }
//   payload (each section 16-byte aligned)
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "Scene.h"
#include "Loader.h"
#include "GLTFLoader.h"
#include "Camera.h"

using namespace GLSLPT;

struct Section { char name[16]; std::vector<unsigned char> data; };
static std::vector<Section> g_sections;

static void add(const char* name, const void* p, size_t n)
{
    Section s; memset(s.name, 0, 16); strncpy(s.name, name, 15);
    s.data.assign((const unsigned char*)p, (const unsigned char*)p + n);
    g_sections.push_back(std::move(s));
}

static uint64_t fnv1a64(const void* p, size_t n)
{
    const unsigned char* b = (const unsigned char*)p; uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; }
    return h;
}

// Scalars the Renderer reads besides the arrays (Renderer.cpp:404-459,498-507,769-811).
struct Scalars
{
    int32_t topLevelIndex, numNodes, numIndices, numVertices, numMaterials, numInstances, numLights, numTextures;
    int32_t texW, texH, envW, envH;
    float   envTotalSum;
    // camera (Camera.h:46-54)
    float camPosition[3], camUp[3], camRight[3], camForward[3], camFov, camFocalDist, camAperture;
    // RenderOptions (Renderer.h:37-100)
    int32_t renderW, renderH, windowW, windowH, tileW, tileH, maxDepth, maxSpp, RRDepth, denoiserFrameCnt;
    float   uniformLightCol[3], backgroundCol[3], envMapIntensity, envMapRot, roughnessMollificationAmt;
    int32_t enableRR, enableDenoiser, enableTonemap, enableAces, simpleAcesFit, openglNormalMap, enableEnvMap,
            enableUniformLight, hideEmitters, enableBackground, transparentBackground, independentRenderSize,
            enableRoughnessMollification, enableVolumeMIS;
    float   sceneBoundsMin[3], sceneBoundsMax[3];
    int32_t tlasHeight, maxBlasHeight;
};

int main(int argc, char** argv)
{
    if (argc < 3) { fprintf(stderr, "usage: scene_dump <file.scene|.gltf|.glb> <out.ptscene>\n"); return 2; }
    static_assert(sizeof(RadeonRays::BvhTranslator::Node) == 36, "node layout");
    static_assert(sizeof(Material) == 128 && sizeof(Light) == 60 && sizeof(Mat4) == 64 && sizeof(Indices) == 12, "layouts");

    Scene* scene = new Scene();
    RenderOptions ro;                       // Main.cpp:71 (process-global in the app)
    ro.simpleAcesFit = false;               // uninitialised in the reference ctor (Renderer.h:39-70); pin it
    {   // Main.cpp:125-134: dispatch on the file extension
        std::string name = argv[1], ext = name.substr(name.find_last_of(".") + 1);
        Mat4 xform;
        bool ok = false;
        if (ext == "scene") ok = LoadSceneFromFile(name, scene, ro);
        else if (ext == "gltf") ok = LoadGLTF(name, scene, ro, xform, false);
        else if (ext == "glb") ok = LoadGLTF(name, scene, ro, xform, true);
        if (!ok) { fprintf(stderr, "load failed\n"); return 1; }
    }
    scene->renderOptions = ro;              // Main.cpp:154
    scene->ProcessScene();                  // Renderer.cpp:78-79

    Scalars s; memset(&s, 0, sizeof(s));
    auto& bt = scene->bvhTranslator;
    s.topLevelIndex = bt.topLevelIndex; s.numNodes = (int)bt.nodes.size();
    s.numIndices = (int)scene->vertIndices.size(); s.numVertices = (int)scene->verticesUVX.size();
    s.numMaterials = (int)scene->materials.size(); s.numInstances = (int)scene->meshInstances.size();
    s.numLights = (int)scene->lights.size(); s.numTextures = (int)scene->textures.size();
    s.texW = ro.texArrayWidth; s.texH = ro.texArrayHeight;
    if (scene->envMap) { s.envW = scene->envMap->width; s.envH = scene->envMap->height; s.envTotalSum = scene->envMap->totalSum; }
    Camera* c = scene->camera;
    memcpy(s.camPosition, &c->position, 12); memcpy(s.camUp, &c->up, 12); memcpy(s.camRight, &c->right, 12); memcpy(s.camForward, &c->forward, 12);
    s.camFov = c->fov; s.camFocalDist = c->focalDist; s.camAperture = c->aperture;
    s.renderW = ro.renderResolution.x; s.renderH = ro.renderResolution.y; s.windowW = ro.windowResolution.x; s.windowH = ro.windowResolution.y;
    s.tileW = ro.tileWidth; s.tileH = ro.tileHeight; s.maxDepth = ro.maxDepth; s.maxSpp = ro.maxSpp; s.RRDepth = ro.RRDepth; s.denoiserFrameCnt = ro.denoiserFrameCnt;
    memcpy(s.uniformLightCol, &ro.uniformLightCol, 12); memcpy(s.backgroundCol, &ro.backgroundCol, 12);
    s.envMapIntensity = ro.envMapIntensity; s.envMapRot = ro.envMapRot; s.roughnessMollificationAmt = ro.roughnessMollificationAmt;
    s.enableRR = ro.enableRR; s.enableDenoiser = ro.enableDenoiser; s.enableTonemap = ro.enableTonemap; s.enableAces = ro.enableAces; s.simpleAcesFit = ro.simpleAcesFit;
    s.openglNormalMap = ro.openglNormalMap; s.enableEnvMap = ro.enableEnvMap; s.enableUniformLight = ro.enableUniformLight; s.hideEmitters = ro.hideEmitters;
    s.enableBackground = ro.enableBackground; s.transparentBackground = ro.transparentBackground; s.independentRenderSize = ro.independentRenderSize;
    s.enableRoughnessMollification = ro.enableRoughnessMollification; s.enableVolumeMIS = ro.enableVolumeMIS;
    memcpy(s.sceneBoundsMin, &scene->sceneBounds.pmin, 12); memcpy(s.sceneBoundsMax, &scene->sceneBounds.pmax, 12);
    int maxh = 0; for (auto* m : scene->meshes) maxh = std::max(maxh, m->bvh->GetHeight());
    s.maxBlasHeight = maxh;
    // TLAS height: walk the flattened TLAS part.
    {
        std::vector<std::pair<int,int>> st; st.push_back({bt.topLevelIndex, 1}); int h = 0;
        while (!st.empty()) { auto [i, d] = st.back(); st.pop_back(); h = std::max(h, d);
            const auto& n = bt.nodes[i]; if (n.LRLeaf.z == 0) { st.push_back({(int)n.LRLeaf.x, d + 1}); st.push_back({(int)n.LRLeaf.y, d + 1}); } }
        s.tlasHeight = h;
    }

    add("scalars", &s, sizeof(s));
    add("nodes", bt.nodes.data(), bt.nodes.size() * 36);
    add("vertIndices", scene->vertIndices.data(), scene->vertIndices.size() * 12);
    add("verticesUVX", scene->verticesUVX.data(), scene->verticesUVX.size() * 16);
    add("normalsUVY", scene->normalsUVY.data(), scene->normalsUVY.size() * 16);
    {   // Material::padding1 / padding2 are never initialised by the reference (Material.h:56,84) and never read by a shader
        // (pathtrace.glsl:31-67 takes .rgb of texel 1 and .xyz of texel 7): zero them so that the blob is reproducible
        std::vector<Material> mats = scene->materials;
        for (auto& m : mats) { m.padding1 = 0.0f; m.padding2 = 0.0f; }
        add("materials", mats.data(), mats.size() * 128);
    }
    add("transforms", scene->transforms.data(), scene->transforms.size() * 64);
    add("lights", scene->lights.data(), scene->lights.size() * 60);
    add("textures", scene->textureMapsArray.data(), scene->textureMapsArray.size());
    if (scene->envMap) {
        add("envImg", scene->envMap->img, (size_t)s.envW * s.envH * 12);
        add("envCdf", scene->envMap->cdf, (size_t)s.envW * s.envH * 4);
    }
    std::vector<int32_t> inst; for (auto& mi : scene->meshInstances) { inst.push_back(mi.meshID); inst.push_back(mi.materialID); }
    add("instances", inst.data(), inst.size() * 4);

    FILE* f = fopen(argv[2], "wb"); if (!f) { perror("open out"); return 1; }
    uint32_t n = (uint32_t)g_sections.size();
    uint64_t off = 8 + 4 + (uint64_t)n * 32; off = (off + 15) & ~15ull;
    fwrite("PTBSCN01", 1, 8, f); fwrite(&n, 4, 1, f);
    std::vector<uint64_t> offs;
    for (auto& sec : g_sections) { uint64_t nb = sec.data.size(); fwrite(sec.name, 1, 16, f); fwrite(&nb, 8, 1, f); fwrite(&off, 8, 1, f); offs.push_back(off); off = (off + nb + 15) & ~15ull; }
    for (size_t i = 0; i < g_sections.size(); i++) { fseek(f, (long)offs[i], SEEK_SET); if (!g_sections[i].data.empty()) fwrite(g_sections[i].data.data(), 1, g_sections[i].data.size(), f); }
    // pad file end
    fseek(f, 0, SEEK_END); long endpos = ftell(f); while (endpos & 15) { fputc(0, f); endpos++; }
    fclose(f);

    printf("PTSCENE %s nodes=%d top=%d fnv1a64(nodes)=%016llx indices=%d verts=%d mats=%d inst=%d lights=%d tex=%d env=%dx%d tlasH=%d blasH=%d\n",
           argv[1], s.numNodes, s.topLevelIndex, (unsigned long long)fnv1a64(bt.nodes.data(), bt.nodes.size() * 36),
           s.numIndices, s.numVertices, s.numMaterials, s.numInstances, s.numLights, s.numTextures, s.envW, s.envH, s.tlasHeight, s.maxBlasHeight);
    return 0;
}
