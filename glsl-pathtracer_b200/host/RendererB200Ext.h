// RendererB200Ext.h — optional additions next to the reference's Renderer.h (which is used unmodified).
#pragma once
#include "Renderer.h"
struct PtbCtx;
struct PtbMgpu;
namespace GLSLPT
{
    // GPUs a Renderer constructed afterwards will use (CUDA ordinals; default: PTB_DEVICES="0,1,..", else PTB_DEVICE, else 0).  With
    // more than one, sample passes are sharded over the GPUs and GetOutputBuffer sees their NCCL-reduced sum.
    void SetDevicesB200(const int* devices, int n);
    // Denoiser hook in place of the reference's OIDN block (Renderer.cpp:695-728): called from Update() with the tonemapped image of
    // the last completed pass (w*h*3 floats, GL_RGB/GL_FLOAT as the reference reads it) whenever the reference would run its filter
    // (renderOptions.enableDenoiser, sampleCounter > 1, every denoiserFrameCnt passes).  nullptr (default) = no denoiser linked.
    typedef void (*DenoiseFnB200)(const float* rgbIn, float* rgbOut, int w, int h, void* user);
    void SetDenoiserB200(DenoiseFnB200 fn, void* user);
    const float* DenoisedImageB200(Renderer& r);          // what the reference uploads to denoisedTexture (nullptr before the first run)
    PtbMgpu* MgpuOfB200(Renderer& r);
    // Scene::RebuildInstances (Scene.cpp:200-214) without the host BVH build: copies meshInstances[i].transform into scene->transforms, lets the library rebuild the
    // TLAS ON THE DEVICE(S) from the transforms and material ids (ptb_mgpu_rebuild_instances: byte-identical to what sceneBvh->Build + bvhTranslator.UpdateTLAS produce),
    // reads the slice back into scene->bvhTranslator.nodes so that the Scene stays consistent, and marks the scene dirty.  Call it where the application calls
    // scene->RebuildInstances() (Main.cpp:505); instancesModified stays false — the contexts are already up to date.
    void RebuildInstancesB200(Renderer& r, Scene* scene);

    // Equivalent to n * numTiles.x * numTiles.y Update()+Render() pairs on a non-dirty scene, as ONE wavefront per batch of passes.
    void RenderSamplesB200(Renderer& r, Scene* scene, int n);
    // The C-ABI context behind a Renderer (stats, profiling).
    PtbCtx* ContextOfB200(Renderer& r);

    // Renderer's counters are protected (Renderer.h:150-160); this accessor advances them the way n full passes would.
    struct RendererB200Access : public Renderer
    {
        static void advance(Renderer& r, int passes)
        {
            RendererB200Access& a = static_cast<RendererB200Access&>(r);
            a.frameCounter += passes * a.numTiles.x * a.numTiles.y;
            a.sampleCounter += passes;
        }
    };
}
