// ptb_derive.h — host-side derivation of the traversal layout from the canonical arrays (no CUDA calls): what ptb_create /
// ptb_update_instances compute before they upload.  Kept apart from ptb_api.cpp so that the derivation and the device headers can
// also be compiled for the host by the test harness (tests/host_harness), which checks them against the oracle without a GPU.
#pragma once
#include "ptb_internal.h"
#include <string>
#include <vector>

struct PtbDerivedHierarchy
{
    std::vector<float4> inner;        // rows for canonical nodes [begin, end): 4 float4 each
    std::vector<float4> instTrav;     // 4 float4 / instance
    std::vector<float4> instShade;    // 8 float4 / instance
    std::vector<char> transOnly;      // per instance: inverse(transform) is [I | -translation] bit for bit (flag of the instance metas)
    uint32_t rootMeta = PTB_META_NONE;
    int stackDepth = 4;
};
struct PtbDerivedLights { std::vector<float4> lightsPre, lightGroups; std::vector<uint32_t> lightGrid; int numGroups = 0; };

// status codes: 0 ok, 1 invalid argument, 4 unsupported (same values as PtbStatus)
int ptbd_derive_hierarchy(const float* nodes, int numNodes, int topLevelIndex, int numIndices, int numMaterials, const float* transforms, int numInstances,
                          int begin, int end, PtbDerivedHierarchy& out, std::string& err);
int ptbd_build_tris(const int32_t* vertIndices, int numIndices, const float* verticesUVX, int numVertices, std::vector<float4>& tris, std::string& err);
// per leaf-ref slot: the three vertex normals and texture coordinates the hit-attribute code interpolates (closest_hit.glsl:226-241), gathered
// through vertIndices at upload: one 64-byte record instead of an index fetch followed by six scattered 16-byte fetches
void ptbd_build_tri_shade(const int32_t* vertIndices, int numIndices, const float* verticesUVX, const float* normalsUVY, std::vector<float4>& out);
// 4-wide hierarchy for the ANY-HIT traversal (anyhit.glsl:65-213).  In any-hit the distance bound never shrinks, so which leaves a ray reaches does not
// depend on the visiting order; and because every node box of the reference's hierarchy contains its children's boxes (verified here) and the slab test
// is monotone in the box bounds, "child box passes => parent box passes": a leaf is reached exactly when ITS OWN box passes.  Any hierarchy over the same
// leaf boxes therefore returns the same boolean.  This one collapses the binary tree (largest-area inner child first) into nodes of up to four children
// whose boxes are the exact boxes of the corresponding binary nodes; empty child slots carry NaN boxes (every comparison fails) and the NONE meta.
//   wide[w*8 + 0..5] = 4 boxes x {min.xyz, max.xyz} (24 floats), wide[w*8 + 6] = 4 child metas (INNER -> wide index, LEAF / INST as in the binary layout)
// ok = false (containment violated somewhere, or an encoding limit) disables the wide path; the binary traversal is always available.
struct PtbDerivedWide
{
    std::vector<float4> wide;               // 8 float4 per wide node
    std::vector<uint32_t> instRootMeta;     // per instance: meta of its BLAS root in the wide hierarchy
    uint32_t rootMeta = PTB_META_NONE;      // TLAS root
    int stackDepth = 4;                     // sentinel + TLAS siblings + marker + BLAS siblings
    bool ok = false;
};
void ptbd_build_wide(const float* nodes, int numNodes, int topLevelIndex, int numIndices, int numInstances, const std::vector<char>& transOnly, PtbDerivedWide& out);
// TLAS rebuild after an instance edit — what Scene::RebuildInstances (Scene.cpp:200-214) does on the reference's host: instance world boxes from the BLAS
// root boxes and the transforms (Scene::createTLAS, Scene.cpp:148-187), Bvh(10, 64, usesah = false)::Build (bvh.cpp:68-243: split at the centre of the
// centroid bounds along their widest axis, one instance per leaf, children in DFS pre-order) and BvhTranslator::ProcessTLASNodes (bvh_translator.cpp:58-86).
// Output: the canonical TLAS slice nodes[topLevelIndex ..] (9 floats per node, 2 * numInstances slots of which 2n-1 are used), byte-identical to the reference's.
// blasRoot / materialID: per instance, the LRLeaf.x / LRLeaf.y its TLAS leaf carries.  This host version is the exact sequential algorithm (it is also the
// fallback of the device builder for degenerate inputs, where the reference's in-place partition order matters).
int ptbd_build_tlas_host(const float* nodes, int topLevelIndex, const float* transforms, int numInstances, const int32_t* blasRoot, const int32_t* materialID,
                         std::vector<float>& tlasOut, int* heightOut, std::string& err);
// the per-instance world boxes alone (6 floats each: min, max), the arithmetic of Scene.cpp:154-184
void ptbd_instance_bounds(const float* nodes, const float* transforms, int numInstances, const int32_t* blasRoot, std::vector<float>& boundsOut);
void ptbd_build_lights(const float* lights, int n, PtbDerivedLights& out);
// per-column / per-row pixel tables: {frame texture coordinate of the pixel centre (tile.glsl:43), bits(tile-local coordinate | tile index << 16)}
int ptbd_build_pixel_tables(int renderW, int renderH, int tileW, int tileH, std::vector<float2>& tabX, std::vector<float2>& tabY, std::string& err);
// Guide table for the environment-map CDF search (envmap.glsl:28-55, envBinarySearch in ptb_device.cuh).  The CDF is ONE running sum over the image
// (EnvironmentMap.cpp:52-58), so when it is non-decreasing the reference's two binary searches (row, then column) return the position of the first
// texel whose CDF exceeds the value, whatever the probing sequence.  guide[b] = number of texels whose CDF value falls in a bucket below b, with
// bucket(v) = (int)clamp(v * scale, 0, G - 1), G = guide.size() - 1: the first texel above a value of bucket b lies in [guide[b], guide[b + 1]].
// Returns 0 and fills guide / scale, or 1 when the table does not apply (CDF not monotone / not finite): the device then runs the reference's search.
int ptbd_build_env_guide(const float* cdf, int w, int h, float totalSum, std::vector<uint32_t>& guide, float& scale);
// Warp groups of a block-major wave (WaveParams::lps / lpw, groupToPixel in ptb_device.cuh): passes per group = the largest power of two that divides the pass
// count, at most 2^maxLps (<= 32); the group's pixel sub-block is 8x4, 4x4, 4x2, 2x2, 2x1, 1x1 for 1, 2, 4, 8, 16, 32 passes.
void ptbd_wave_groups(int nSamples, int maxLps, int* lps, int* lpw);
