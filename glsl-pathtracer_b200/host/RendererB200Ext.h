// RendererB200Ext.h — optional additions next to the reference's Renderer.h (which is used unmodified).
#pragma once
#include "Renderer.h"
struct PtbCtx;
namespace GLSLPT
{
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
