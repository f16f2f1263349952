// headless_main.cpp — replaces the SDL/ImGui application (reference src/Main.cpp) for batch rendering:
// LoadScene (Main.cpp:121-155) -> Renderer (Main.cpp:157-162) -> Update/Render loop (Main.cpp:175-213,263) -> SaveFrame (Main.cpp:164-173).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <chrono>
#include <vector>
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
    int spp = 16, w = 0, h = 0, depth = -1, warmup = 0, denoiseEvery = 0; bool whole = false;
    static int denoiseCalls = 0; static double denoiseInSum = 0.0;
    int editInst = -1; float editD[3] = {0, 0, 0}; bool editOnDevice = false;
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
        else if (a == "--warmup" && i + 1 < argc) warmup = atoi(argv[++i]);          // passes rendered before the clock starts (same loop)
        else if (a == "--no-coalesce") setenv("PTB_COALESCE", "0", 1);               // one wavefront per Render() tile, as the reference draws
        else if (a == "--devices" && i + 1 < argc) setenv("PTB_DEVICES", argv[++i], 1);   // "0,1,2,3"
        else if (a == "--edit-instance" && i + 5 < argc)      // K DX DY DZ device|host: move instance K before rendering, TLAS rebuilt by the library on the GPU or by the reference on the host
        { editInst = atoi(argv[++i]); for (int k = 0; k < 3; k++) editD[k] = (float)atof(argv[++i]); editOnDevice = std::string(argv[++i]) == "device"; }
        else if (a == "--denoise-every" && i + 1 < argc) denoiseEvery = atoi(argv[++i]);  // exercise the denoiser hook with a stand-in filter
        else { printf("usage: ptb_headless -s file.scene [-o out.png] [--spp N] [--warmup N] [--res W H] [--depth D] [--whole-frame] [--no-coalesce] "
                      "[--devices 0,1,..] [--denoise-every N] [--accum file.f32]\n"); return 2; }
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
    renderOptions.maxSpp = warmup + spp + 1;             // Q1: maxSpp = M renders M-1 passes
    if (denoiseEvery > 0) { renderOptions.enableDenoiser = true; renderOptions.denoiserFrameCnt = denoiseEvery; }
    scene->renderOptions = renderOptions;                // Main.cpp:154
    if (denoiseEvery > 0)      // stand-in for OIDN: the hook, its trigger and its cadence are the deliverable, not the filter
        SetDenoiserB200([](const float* in, float* out, int w_, int h_, void*) { denoiseCalls++; denoiseInSum = 0.0; for (size_t i = 0; i < (size_t)w_ * h_ * 3; i++) denoiseInSum += in[i]; memcpy(out, in, (size_t)w_ * h_ * 12); }, nullptr);

    Renderer* renderer = new Renderer(scene, "shaders/");
    if (editInst >= 0 && editInst < (int)scene->meshInstances.size())
    {   // what the application does when an instance is dragged (Main.cpp:495-510): edit the transform, then rebuild the instances
        Mat4& m = scene->meshInstances[editInst].transform;
        m[3][0] += editD[0]; m[3][1] += editD[1]; m[3][2] += editD[2];
        if (editOnDevice) RebuildInstancesB200(*renderer, scene);       // TLAS rebuilt on the device from the transforms
        else scene->RebuildInstances();                                  // the reference's host rebuild; Update() uploads its slice (instancesModified)
    }
    auto t0 = std::chrono::steady_clock::now();
    unsigned char* data = nullptr; int ow, oh;
    PtbStats st0{}; long long updates = 0;
    if (whole)
    {
        renderer->Update(0.f); renderer->Render();       // the dirty / preview frame
        if (warmup > 0) { RenderSamplesB200(*renderer, scene, warmup); ptb_mgpu_synchronize(MgpuOfB200(*renderer)); ptb_mgpu_get_stats(MgpuOfB200(*renderer), &st0); t0 = std::chrono::steady_clock::now(); }
        RenderSamplesB200(*renderer, scene, spp);
    }
    else
    {   // the application loop of Main.cpp:175-213, 263: one Update + one Render (one tile) + Present per iteration
        bool clockStarted = warmup == 0;
        while (renderer->GetSampleCount() < renderOptions.maxSpp)
        {
            renderer->Update(0.016f);
            if (!clockStarted && renderer->GetSampleCount() >= warmup + 1)
            {   // pass `warmup` has just completed and nothing of the next pass has been traced yet (its wave starts in the Render() below):
                // drain the GPU and start the clock here, so that every pass counted is traced inside the timed region
                renderer->GetOutputBuffer(&data, ow, oh); delete[] data;
                ptb_mgpu_get_stats(MgpuOfB200(*renderer), &st0);
                t0 = std::chrono::steady_clock::now(); clockStarted = true;
            }
            renderer->Render(); renderer->Present(); updates++;
        }
    }
    renderer->GetOutputBuffer(&data, ow, oh);            // SaveFrame, Main.cpp:164-173: the one device->host copy (inside the timed region)
    double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    PtbStats st; ptb_mgpu_get_stats(MgpuOfB200(*renderer), &st);
    printf("rendered %d spp in %.3f s (%.1f spp/s), %llu path segments, %llu shadow rays, %llu kernel launches\n", spp, sec, spp / sec,
           (unsigned long long)(st.pathSegments - st0.pathSegments), (unsigned long long)(st.shadowRays - st0.shadowRays), (unsigned long long)(st.kernelLaunches - st0.kernelLaunches));
    printf("BENCH {\"spp\": %d, \"seconds\": %.6f, \"path_segments\": %llu, \"shadow_rays\": %llu, \"kernel_launches\": %llu, \"updates\": %lld, \"gpus\": %d, "
           "\"coalesced\": %s, \"whole_frame\": %s, \"denoiser_calls\": %d, \"denoiser_input_sum\": %.3f}\n", spp, sec, (unsigned long long)(st.pathSegments - st0.pathSegments),
           (unsigned long long)(st.shadowRays - st0.shadowRays), (unsigned long long)(st.kernelLaunches - st0.kernelLaunches), updates, ptb_mgpu_num_devices(MgpuOfB200(*renderer)),
           (getenv("PTB_COALESCE") && atoi(getenv("PTB_COALESCE")) == 0) ? "false" : "true", whole ? "true" : "false", denoiseCalls, denoiseInSum);

    stbi_flip_vertically_on_write(true);
    stbi_write_png(out.c_str(), ow, oh, 4, data, ow * 4);
    delete[] data;
    if (!accumOut.empty())
    {
        std::vector<float> acc((size_t)ow * oh * 4);
        ptb_mgpu_read_accum_f32(MgpuOfB200(*renderer), acc.data());
        FILE* f = fopen(accumOut.c_str(), "wb"); fwrite(acc.data(), 4, acc.size(), f); fclose(f);
    }
    printf("wrote %s (%dx%d)\n", out.c_str(), ow, oh);
    delete renderer; delete scene;
    return 0;
}
