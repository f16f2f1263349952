// headless_main.cpp — replaces the SDL/ImGui application (reference src/Main.cpp) for batch rendering:
// LoadScene (Main.cpp:121-155) -> Renderer (Main.cpp:157-162) -> Update/Render loop (Main.cpp:175-213,263) -> SaveFrame (Main.cpp:164-173).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <chrono>
#include "Scene.h"
#include "Loader.h"
#include "GLTFLoader.h"
#include "Renderer.h"
#include "RendererB200Ext.h"
#include "ptb200.h"
#include "stb_image_write.h"

using namespace GLSLPT;

int main(int argc, char** argv)
{
    std::string sceneFile, out = "out.png", accumOut;
    int spp = 16, w = 0, h = 0, depth = -1; bool whole = false;
    for (int i = 1; i < argc; i++)
    {
        std::string a = argv[i];
        if ((a == "-s" || a == "--scene") && i + 1 < argc) sceneFile = argv[++i];      // Main.cpp:544-556
        else if (a == "-o" && i + 1 < argc) out = argv[++i];
        else if (a == "--spp" && i + 1 < argc) spp = atoi(argv[++i]);
        else if (a == "--res" && i + 2 < argc) { w = atoi(argv[++i]); h = atoi(argv[++i]); }
        else if (a == "--depth" && i + 1 < argc) depth = atoi(argv[++i]);
        else if (a == "--accum" && i + 1 < argc) accumOut = argv[++i];
        else if (a == "--whole-frame") whole = true;
        else { printf("usage: ptb_headless -s file.scene [-o out.png] [--spp N] [--res W H] [--depth D] [--whole-frame] [--accum file.f32]\n"); return 2; }
    }
    if (sceneFile.empty()) { printf("no scene\n"); return 2; }

    Scene* scene = new Scene();
    RenderOptions renderOptions;
    renderOptions.simpleAcesFit = false;
    {   // Main.cpp:125-141: dispatch on the file extension
        std::string ext = sceneFile.substr(sceneFile.find_last_of(".") + 1);
        Mat4 xform;
        bool success = false;
        if (ext == "scene") success = LoadSceneFromFile(sceneFile, scene, renderOptions);
        else if (ext == "gltf") success = LoadGLTF(sceneFile, scene, renderOptions, xform, false);
        else if (ext == "glb") success = LoadGLTF(sceneFile, scene, renderOptions, xform, true);
        if (!success) { printf("Unable to load scene\n"); return 1; }
    }
    if (w > 0) { renderOptions.renderResolution = iVec2(w, h); renderOptions.windowResolution = iVec2(w, h); }
    if (depth >= 0) renderOptions.maxDepth = depth;
    renderOptions.maxSpp = spp + 1;                      // Q1: maxSpp = M renders M-1 passes
    scene->renderOptions = renderOptions;                // Main.cpp:154

    Renderer* renderer = new Renderer(scene, "shaders/");
    auto t0 = std::chrono::steady_clock::now();
    if (whole)
    {
        renderer->Update(0.f); renderer->Render();       // the dirty / preview frame
        RenderSamplesB200(*renderer, scene, spp);
    }
    else
        while (renderer->GetSampleCount() < renderOptions.maxSpp) { renderer->Update(0.016f); renderer->Render(); renderer->Present(); }
    PtbStats st; ptb_get_stats(ContextOfB200(*renderer), &st);
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("rendered %d spp in %.3f s (%.1f spp/s), %llu path segments, %llu shadow rays, %llu kernel launches\n", spp, sec, spp / sec,
           (unsigned long long)st.pathSegments, (unsigned long long)st.shadowRays, (unsigned long long)st.kernelLaunches);

    unsigned char* data = nullptr; int ow, oh;
    renderer->GetOutputBuffer(&data, ow, oh);            // SaveFrame, Main.cpp:164-173
    stbi_flip_vertically_on_write(true);
    stbi_write_png(out.c_str(), ow, oh, 4, data, ow * 4);
    delete[] data;
    if (!accumOut.empty())
    {
        std::vector<float> acc((size_t)ow * oh * 4);
        ptb_read_accum_f32(ContextOfB200(*renderer), acc.data());
        FILE* f = fopen(accumOut.c_str(), "wb"); fwrite(acc.data(), 4, acc.size(), f); fclose(f);
    }
    printf("wrote %s (%dx%d)\n", out.c_str(), ow, oh);
    delete renderer; delete scene;
    return 0;
}
