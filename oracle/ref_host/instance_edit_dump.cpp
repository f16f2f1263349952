// instance_edit_dump — TEST INFRASTRUCTURE (oracle side), not product code.
//
// Links the UNMODIFIED reference host code (same objects as scene_dump), loads a scene, runs ProcessScene(), then performs an
// interactive-style instance edit exactly as the application does (Main.cpp:495-510: write meshInstances[k].transform, call
// Scene::RebuildInstances() — Scene.cpp:200-214: new RadeonRays::Bvh over the instance bounds, BvhTranslator::UpdateTLAS, transforms
// copied) and writes what Renderer::Update then re-uploads (Renderer.cpp:649-665): transforms, materials, nodes[topLevelIndex..].
//
//   instance_edit_dump <file.scene> <out.bin> <instance> <tx> <ty> <tz> <scale> [materialOfInstance]
// out.bin: int32 numInstances, numMaterials, numTlasNodes, topLevelIndex; then transforms (16 f32 each), materials (32 f32 each),
// TLAS nodes (9 f32 each).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include "Scene.h"
#include "Loader.h"

using namespace GLSLPT;

int main(int argc, char** argv)
{
    if (argc < 8) { fprintf(stderr, "usage: instance_edit_dump <file.scene> <out.bin> <instance> <tx> <ty> <tz> <scale> [material]\n"); return 2; }
    Scene* scene = new Scene();
    RenderOptions ro; ro.simpleAcesFit = false;
    if (!LoadSceneFromFile(argv[1], scene, ro)) { fprintf(stderr, "load failed\n"); return 1; }
    scene->renderOptions = ro;
    scene->ProcessScene();
    const int k = atoi(argv[3]);
    if (k < 0 || k >= (int)scene->meshInstances.size()) { fprintf(stderr, "no such instance\n"); return 1; }
    const float tx = (float)atof(argv[4]), ty = (float)atof(argv[5]), tz = (float)atof(argv[6]), sc = (float)atof(argv[7]);
    Mat4 S, T;
    S.data[0][0] = sc; S.data[1][1] = sc; S.data[2][2] = sc;
    T.data[3][0] = tx; T.data[3][1] = ty; T.data[3][2] = tz;
    scene->meshInstances[k].transform = scene->meshInstances[k].transform * S * T;     // row-vector convention: scale, then translate in world space
    if (argc > 8) scene->meshInstances[k].materialID = atoi(argv[8]);
    scene->RebuildInstances();

    auto& bt = scene->bvhTranslator;
    int32_t hdr[4] = {(int32_t)scene->transforms.size(), (int32_t)scene->materials.size(), (int32_t)(bt.nodes.size() - bt.topLevelIndex), (int32_t)bt.topLevelIndex};
    FILE* f = fopen(argv[2], "wb"); if (!f) { perror("open"); return 1; }
    fwrite(hdr, 4, 4, f);
    fwrite(scene->transforms.data(), 64, scene->transforms.size(), f);
    std::vector<Material> mats = scene->materials;
    for (auto& m : mats) { m.padding1 = 0.0f; m.padding2 = 0.0f; }                          // as scene_dump: uninitialised in the reference, never read
    fwrite(mats.data(), 128, mats.size(), f);
    fwrite(&bt.nodes[bt.topLevelIndex], 36, bt.nodes.size() - bt.topLevelIndex, f);
    fclose(f);
    printf("INSTEDIT %s inst=%d tlasNodes=%d top=%d\n", argv[1], k, hdr[2], hdr[3]);
    return 0;
}
