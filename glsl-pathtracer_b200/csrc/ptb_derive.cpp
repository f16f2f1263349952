// ptb_derive.cpp — see ptb_derive.h.  Built with -ffp-contract=off: the per-instance inverses, light planes and triangle edges must be
// bit-identical to what the shader-side formulas give under IEEE arithmetic.
#include "ptb_derive.h"
#include <algorithm>
#include <cmath>
#include <cstring>
#include <utility>
#include <cstdlib>
#include <limits>

namespace {

// inverse(mat4) of the row-major reference Mat4 (adjugate / determinant, fp32, no contraction) — the GLSL `inverse(transMat)`
// of closest_hit.glsl:163-164 evaluated once per instance at upload instead of twice per TLAS-leaf visit per ray.
void inverse4(const float* a, float* b)
{
    float a00 = a[0], a01 = a[1], a02 = a[2], a03 = a[3], a10 = a[4], a11 = a[5], a12 = a[6], a13 = a[7];
    float a20 = a[8], a21 = a[9], a22 = a[10], a23 = a[11], a30 = a[12], a31 = a[13], a32 = a[14], a33 = a[15];
    float s0 = a00 * a11 - a10 * a01, s1 = a00 * a12 - a10 * a02, s2 = a00 * a13 - a10 * a03;
    float s3 = a01 * a12 - a11 * a02, s4 = a01 * a13 - a11 * a03, s5 = a02 * a13 - a12 * a03;
    float c5 = a22 * a33 - a32 * a23, c4 = a21 * a33 - a31 * a23, c3 = a21 * a32 - a31 * a22;
    float c2 = a20 * a33 - a30 * a23, c1 = a20 * a32 - a30 * a22, c0 = a20 * a31 - a30 * a21;
    float det = s0 * c5 - s1 * c4 + s2 * c3 + s3 * c2 - s4 * c1 + s5 * c0;
    float id = 1.0f / det;
    b[0] = (a11 * c5 - a12 * c4 + a13 * c3) * id;  b[1] = (-a01 * c5 + a02 * c4 - a03 * c3) * id;
    b[2] = (a31 * s5 - a32 * s4 + a33 * s3) * id;  b[3] = (-a21 * s5 + a22 * s4 - a23 * s3) * id;
    b[4] = (-a10 * c5 + a12 * c2 - a13 * c1) * id; b[5] = (a00 * c5 - a02 * c2 + a03 * c1) * id;
    b[6] = (-a30 * s5 + a32 * s2 - a33 * s1) * id; b[7] = (a20 * s5 - a22 * s2 + a23 * s1) * id;
    b[8] = (a10 * c4 - a11 * c2 + a13 * c0) * id;  b[9] = (-a00 * c4 + a01 * c2 - a03 * c0) * id;
    b[10] = (a30 * s4 - a31 * s2 + a33 * s0) * id; b[11] = (-a20 * s4 + a21 * s2 - a23 * s0) * id;
    b[12] = (-a10 * c3 + a11 * c1 - a12 * c0) * id; b[13] = (a00 * c3 - a01 * c1 + a02 * c0) * id;
    b[14] = (-a30 * s3 + a31 * s1 - a32 * s0) * id; b[15] = (a20 * s3 - a21 * s1 + a22 * s0) * id;
}
// inverse(mat3(transform)) (closest_hit.glsl:244), row-major 3x3
void inverse3(const float* m, float* b)
{
    float a00 = m[0], a01 = m[1], a02 = m[2], a10 = m[4], a11 = m[5], a12 = m[6], a20 = m[8], a21 = m[9], a22 = m[10];
    float c00 = a11 * a22 - a12 * a21, c01 = a12 * a20 - a10 * a22, c02 = a10 * a21 - a11 * a20;
    float det = a00 * c00 + a01 * c01 + a02 * c02;
    float id = 1.0f / det;
    b[0] = c00 * id; b[1] = (a02 * a21 - a01 * a22) * id; b[2] = (a01 * a12 - a02 * a11) * id;
    b[3] = c01 * id; b[4] = (a00 * a22 - a02 * a20) * id; b[5] = (a02 * a10 - a00 * a12) * id;
    b[6] = c02 * id; b[7] = (a01 * a20 - a00 * a21) * id; b[8] = (a00 * a11 - a01 * a10) * id;
}

inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }
inline float u2f(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }


uint32_t metaOf(const float* nodes, int idx, int numIndices, std::string& err)
{
    const float* n = nodes + (size_t)idx * 9;
    int leaf = (int)n[8];
    if (leaf == 0) return (PTB_K_INNER << 30) | (uint32_t)idx;
    if (leaf > 0)
    {
        int first = (int)n[6], cnt = (int)n[7];
        if (cnt > PTB_MAX_LEAF_TRIS || cnt < 0 || first < 0 || (uint32_t)first > PTB_MAX_LEAF_SLOT) { err = "leaf exceeds encoding limits"; return PTB_META_NONE; }
        if ((long long)first + cnt > (long long)numIndices) { err = "leaf references triangles past the end of vertIndices"; return PTB_META_NONE; }
        return (PTB_K_LEAF << 30) | ((uint32_t)cnt << 26) | (uint32_t)first;
    }
    return (PTB_K_INST << 30) | (uint32_t)(-leaf - 1);
}

// number of internal nodes on the deepest root-to-leaf path of the subtree at idx (iterative DFS)
int innerDepth(const float* nodes, int root, int numNodes)
{
    int best = 0;
    std::vector<std::pair<int, int>> st; st.push_back({root, 0});
    while (!st.empty())
    {
        auto [i, d] = st.back(); st.pop_back();
        if (i < 0 || i >= numNodes) continue;
        const float* n = nodes + (size_t)i * 9;
        if ((int)n[8] == 0) { st.push_back({(int)n[6], d + 1}); st.push_back({(int)n[7], d + 1}); }
        else best = std::max(best, d);
    }
    return best;
}


} // namespace

#define DREQ(cond, code, msg) do { if (!(cond)) { err = msg; return code; } } while (0)

int ptbd_derive_hierarchy(const float* N, int numNodes, int topLevelIndex, int numIndices, int numMaterials, const float* transforms, int numInstances,
                          int begin, int end, PtbDerivedHierarchy& out, std::string& err)
{
    std::string merr;
    auto metaCode = [&]() { return merr.find("past the end") != std::string::npos ? 1 : 4; };
    out.inner.assign((size_t)(end - begin) * 4, make_float4(0, 0, 0, 0));
    for (int i = begin; i < end; i++)
    {
        const float* n = N + (size_t)i * 9;
        float4* q = &out.inner[(size_t)(i - begin) * 4];
        if ((int)n[8] != 0) continue;
        int l = (int)n[6], r = (int)n[7];
        DREQ(l >= 0 && l < numNodes && r >= 0 && r < numNodes, 1, "child index out of range");
        const float* L = N + (size_t)l * 9; const float* R = N + (size_t)r * 9;
        uint32_t lm = metaOf(N, l, numIndices, merr), rm = metaOf(N, r, numIndices, merr);
        DREQ(merr.empty(), metaCode(), merr);
#if PTB_PACKED_SLAB
        // pairs for the packed (f32x2) slab test: {Lmin.xy | Lmax.xy}, {Lmin.z Lmax.z | Rmin.z Rmax.z}, {Rmin.xy | Rmax.xy}
        q[0] = make_float4(L[0], L[1], L[3], L[4]);
        q[1] = make_float4(L[2], L[5], R[2], R[5]);
        q[2] = make_float4(R[0], R[1], R[3], R[4]);
#else
        q[0] = make_float4(L[0], L[1], L[2], L[3]);
        q[1] = make_float4(L[4], L[5], R[0], R[1]);
        q[2] = make_float4(R[2], R[3], R[4], R[5]);
#endif
        q[3] = make_float4(u2f(lm), u2f(rm), 0.f, 0.f);
    }

    // instance tables from the TLAS leaves (bvh_translator.cpp:70-78: LRLeaf = (blasRoot, materialID, -(inst+1)))
    const int ni = numInstances;
    out.instTrav.assign((size_t)ni * 4, make_float4(0, 0, 0, 0)); out.instShade.assign((size_t)ni * 8, make_float4(0, 0, 0, 0));
    std::vector<uint32_t> rootMeta(ni, PTB_META_NONE); std::vector<int> matID(ni, 0), blasRoot(ni, -1);
    for (int i = topLevelIndex; i < numNodes; i++)
    {
        const float* n = N + (size_t)i * 9;
        int leaf = (int)n[8];
        if (leaf < 0)
        {
            int k = -leaf - 1;
            if (k >= ni) continue;
            int root = (int)n[6];
            DREQ(root >= 0 && root < numNodes, 1, "BLAS root out of range");
            rootMeta[k] = metaOf(N, root, numIndices, merr); matID[k] = (int)n[7]; blasRoot[k] = root;
            DREQ(merr.empty(), metaCode(), merr);
            DREQ(matID[k] >= 0 && matID[k] < numMaterials, 1, "TLAS leaf material id out of range");
        }
    }
    int maxBlas = 0;
    {
        std::vector<int> seen;
        for (int k = 0; k < ni; k++)
            if (blasRoot[k] >= 0 && std::find(seen.begin(), seen.end(), blasRoot[k]) == seen.end())
            { seen.push_back(blasRoot[k]); maxBlas = std::max(maxBlas, innerDepth(N, blasRoot[k], numNodes)); }
    }
    int tlasDepth = innerDepth(N, topLevelIndex, numNodes);
    out.stackDepth = std::max(4, 1 + tlasDepth + 1 + maxBlas + 1);
    DREQ(out.stackDepth <= 64, 4, "BVH deeper than the 64-entry traversal stack of the reference shader");

    std::vector<char>& transOnly = out.transOnly; transOnly.assign((size_t)ni, 0);
    for (int k = 0; k < ni; k++)
    {
        const float* D = transforms + (size_t)k * 16;
        float inv[16], inv3[9];
        inverse4(D, inv); inverse3(D, inv3);
        float4* it = &out.instTrav[(size_t)k * 4]; float4* is = &out.instShade[(size_t)k * 8];
        it[0] = make_float4(inv[0], inv[1], inv[2], 0.f);
        it[1] = make_float4(inv[4], inv[5], inv[6], u2f((uint32_t)matID[k]));
        it[2] = make_float4(inv[8], inv[9], inv[10], 0.f);
        it[3] = make_float4(inv[12], inv[13], inv[14], u2f(rootMeta[k]));
        // translation-only: the linear part of the computed inverse is the identity bit for bit (off-diagonals +-0), so the traversal may
        // take the short instance entry (Trav::round) — the flag travels in the TLAS-leaf metas patched below
        transOnly[k] = inv[0] == 1.f && inv[5] == 1.f && inv[10] == 1.f && inv[1] == 0.f && inv[2] == 0.f && inv[4] == 0.f && inv[6] == 0.f &&
                       inv[8] == 0.f && inv[9] == 0.f && std::isfinite(inv[12]) && std::isfinite(inv[13]) && std::isfinite(inv[14]);
        for (int r = 0; r < 4; r++) is[r] = make_float4(D[r * 4 + 0], D[r * 4 + 1], D[r * 4 + 2], D[r * 4 + 3]);
        for (int r = 0; r < 3; r++) is[4 + r] = make_float4(inv3[r * 3 + 0], inv3[r * 3 + 1], inv3[r * 3 + 2], 0.f);
    }
    out.rootMeta = metaOf(N, topLevelIndex, numIndices, merr);
    DREQ(merr.empty(), metaCode(), merr);
    DREQ((uint32_t)ni <= PTB_INST_INDEX_MASK, 4, "too many instances for the meta encoding");
    auto flagInst = [&](uint32_t m) { return ((m >> 30) == PTB_K_INST && (m & PTB_INST_INDEX_MASK) < (uint32_t)ni && transOnly[m & PTB_INST_INDEX_MASK]) ? (m | PTB_INST_TRANSLATION_ONLY) : m; };
    out.rootMeta = flagInst(out.rootMeta);
    for (int i = std::max(begin, topLevelIndex); i < end; i++)
    {   // child metas of the TLAS nodes in the derived range
        float4& q3 = out.inner[(size_t)(i - begin) * 4 + 3];
        if ((int)N[(size_t)i * 9 + 8] != 0) continue;
        q3.x = u2f(flagInst(f2u(q3.x))); q3.y = u2f(flagInst(f2u(q3.y)));
    }
    return 0;
}

int ptbd_build_tris(const int32_t* vertIndices, int numIndices, const float* verticesUVX, int numVertices, std::vector<float4>& tris, std::string& err)
{
    std::vector<float4>& t = tris; t.assign((size_t)numIndices * 3, make_float4(0, 0, 0, 0));
    for (int s = 0; s < numIndices; s++)
    {
        const int32_t* vi = vertIndices + (size_t)s * 3;
        DREQ(vi[0] >= 0 && vi[0] < numVertices && vi[1] >= 0 && vi[1] < numVertices && vi[2] >= 0 && vi[2] < numVertices,
                1, "vertex index out of range");
        const float* v0 = verticesUVX + (size_t)vi[0] * 4; const float* v1 = verticesUVX + (size_t)vi[1] * 4; const float* v2 = verticesUVX + (size_t)vi[2] * 4;
        // e0 = v1 - v0, e1 = v2 - v0 (closest_hit.glsl:128-129): one IEEE subtraction each, identical wherever it is evaluated
        float e0[3] = {v1[0] - v0[0], v1[1] - v0[1], v1[2] - v0[2]}, e1[3] = {v2[0] - v0[0], v2[1] - v0[1], v2[2] - v0[2]};
        t[s * 3 + 0] = make_float4(v0[0], v0[1], v0[2], e0[0]);
        t[s * 3 + 1] = make_float4(e0[1], e0[2], e1[0], e1[1]);
        t[s * 3 + 2] = make_float4(e1[2], u2f((uint32_t)vi[0]), 0.f, 0.f);
    }
    return 0;
}

void ptbd_build_lights(const float* lights, int n, PtbDerivedLights& out)
{
    std::vector<float4>& lp = out.lightsPre; lp.assign((size_t)n * 8, make_float4(0, 0, 0, 0));
    for (int i = 0; i < n; i++)
    {
        const float* p = lights + (size_t)i * 15;
        float pos[3] = {p[0], p[1], p[2]}, em[3] = {p[3], p[4], p[5]}, u[3] = {p[6], p[7], p[8]}, v[3] = {p[9], p[10], p[11]};
        float radius = p[12], area = p[13], type = p[14];
        // closest_hit.glsl:49-53: normal = normalize(cross(u,v)); plane = (normal, dot(normal,position)); u *= 1/dot(u,u); v *= 1/dot(v,v)
        float cx = u[1] * v[2] - u[2] * v[1], cy = u[2] * v[0] - u[0] * v[2], cz = u[0] * v[1] - u[1] * v[0];
        float len = sqrtf(cx * cx + cy * cy + cz * cz);
        float nx = cx / len, ny = cy / len, nz = cz / len;
        float planeW = nx * pos[0] + ny * pos[1] + nz * pos[2];
        float su = 1.0f / (u[0] * u[0] + u[1] * u[1] + u[2] * u[2]), sv = 1.0f / (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        lp[i * 8 + 0] = make_float4(pos[0], pos[1], pos[2], type);
        lp[i * 8 + 1] = make_float4(em[0], em[1], em[2], area);
        lp[i * 8 + 2] = make_float4(u[0], u[1], u[2], radius);
        // samePlaneAsPrevious: this quad's plane (normal, plane.w) is bit-identical to the previous light's plane
        float same = 0.f;
        if (i > 0 && type == 0.f && lp[(i - 1) * 8 + 0].w == 0.f)
        {
            const float4& pe = lp[(i - 1) * 8 + 4];
            if (f2u(pe.x) == f2u(nx) && f2u(pe.y) == f2u(ny) && f2u(pe.z) == f2u(nz) && f2u(pe.w) == f2u(planeW)) same = 1.f;
        }
        lp[i * 8 + 3] = make_float4(v[0], v[1], v[2], same);
        lp[i * 8 + 4] = make_float4(nx, ny, nz, planeW);
        lp[i * 8 + 5] = make_float4(u[0] * su, u[1] * su, u[2] * su, 0.f);
        lp[i * 8 + 6] = make_float4(v[0] * sv, v[1] * sv, v[2] * sv, 0.f);
        lp[i * 8 + 7] = make_float4(0, 0, 0, 0);
    }
    // groups of consecutive quads on one plane (+ singleton groups for everything else), with padded bounds of the member quads
    std::vector<float4>& groups = out.lightGroups; groups.clear(); out.lightGrid.clear();
    for (int i = 0; i < n;)
    {
        int j = i + 1;
        const bool quad = lp[i * 8 + 0].w == 0.f;
        if (quad) while (j < n && lp[j * 8 + 3].w == 1.f) j++;
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300}, mag = 1e-3;
        for (int k = i; k < j && quad; k++)
        {
            const float* p = lights + (size_t)k * 15;
            for (int cu = 0; cu < 2; cu++) for (int cv = 0; cv < 2; cv++) for (int a = 0; a < 3; a++)
            {
                double v = (double)p[a] + cu * (double)p[6 + a] + cv * (double)p[9 + a];
                lo[a] = std::min(lo[a], v); hi[a] = std::max(hi[a], v); mag = std::max(mag, fabs(v));
            }
        }
        double ext = quad ? std::max({hi[0] - lo[0], hi[1] - lo[1], hi[2] - lo[2]}) : 0.0;
        double pad = 1e-4 * (1.0 + mag + ext);
        // Cell grid over the group's padded box in the plane's two widest axes (groups of 4..32 quads): cell -> bit mask of the quads whose own padded box
        // touches the cell (+-1 cell for the rounding of the device's cell index).  A plane hit inside a cell can only pass the inside test of those
        // quads (same argument as for the group box), so the others are skipped without changing the result; bit order = index order.
        uint32_t kindWord = quad ? 0u : 1u; float gridOff = 0.f, invA = 0.f, invB = 0.f;
        const int cnt = j - i;
        if (quad && cnt >= 4 && cnt <= 32)
        {
            const float bmin[3] = {(float)(lo[0] - pad), (float)(lo[1] - pad), (float)(lo[2] - pad)}, bmax[3] = {(float)(hi[0] + pad), (float)(hi[1] + pad), (float)(hi[2] + pad)};
            int ax[3] = {0, 1, 2};
            std::sort(ax, ax + 3, [&](int a, int b) { return (hi[a] - lo[a]) > (hi[b] - lo[b]); });
            const int A = ax[0], B = ax[1];
            const float ia = (float)PTB_LIGHT_GRID / (bmax[A] - bmin[A]), ib = (float)PTB_LIGHT_GRID / (bmax[B] - bmin[B]);
            if (std::isfinite(ia) && std::isfinite(ib) && ia > 0.f && ib > 0.f)
            {
                const size_t off = out.lightGrid.size();
                out.lightGrid.resize(off + PTB_LIGHT_GRID * PTB_LIGHT_GRID, 0u);
                for (int k = i; k < j; k++)
                {
                    const float* p = lights + (size_t)k * 15;
                    double l2[3] = {1e300, 1e300, 1e300}, h2[3] = {-1e300, -1e300, -1e300};
                    for (int cu = 0; cu < 2; cu++) for (int cv = 0; cv < 2; cv++) for (int a = 0; a < 3; a++)
                    { double v = (double)p[a] + cu * (double)p[6 + a] + cv * (double)p[9 + a]; l2[a] = std::min(l2[a], v); h2[a] = std::max(h2[a], v); }
                    auto cell = [&](double v, int axis, float inv) { return (int)std::floor((v - (double)bmin[axis]) * (double)inv); };
                    const int a0 = std::max(0, cell(l2[A] - pad, A, ia) - 1), a1 = std::min(PTB_LIGHT_GRID - 1, cell(h2[A] + pad, A, ia) + 1);
                    const int b0 = std::max(0, cell(l2[B] - pad, B, ib) - 1), b1 = std::min(PTB_LIGHT_GRID - 1, cell(h2[B] + pad, B, ib) + 1);
                    for (int cb = b0; cb <= b1; cb++) for (int ca = a0; ca <= a1; ca++) out.lightGrid[off + (size_t)cb * PTB_LIGHT_GRID + ca] |= 1u << (k - i);
                }
                kindWord = 2u | ((uint32_t)A << 4) | ((uint32_t)B << 8);
                gridOff = u2f((uint32_t)off); invA = ia; invB = ib;
            }
        }
        groups.push_back(make_float4(u2f((uint32_t)i), u2f((uint32_t)cnt), u2f(kindWord), gridOff));
        if (quad)
        {
            groups.push_back(make_float4((float)(lo[0] - pad), (float)(lo[1] - pad), (float)(lo[2] - pad), invA));
            groups.push_back(make_float4((float)(hi[0] + pad), (float)(hi[1] + pad), (float)(hi[2] + pad), invB));
        }
        else { groups.push_back(make_float4(0, 0, 0, 0)); groups.push_back(make_float4(0, 0, 0, 0)); }
        i = j;
    }
    out.numGroups = (int)(groups.size() / 3);
}

int ptbd_build_pixel_tables(int renderW, int renderH, int tileW, int tileH, std::vector<float2>& tabX, std::vector<float2>& tabY, std::string& err)
{
    DREQ(renderW > 0 && renderH > 0 && tileW > 0 && tileH > 0 && tileW < 65536 && tileH < 65536 && renderW / tileW < 65536 && renderH / tileH < 65536, 4,
         "resolution / tile size beyond the pixel-table encoding");
    // the arithmetic of pixelSeed() in ptb_device.cuh (tile.glsl:43 with the uniforms of Renderer.cpp:293-294,780), one IEEE operation per step
    auto fill = [](int n, int tile, int res, std::vector<float2>& tab)
    {
        const float invNumTiles = (float)tile / res;
        tab.resize((size_t)n);
        for (int x = 0; x < n; x++)
        {
            const int t = x / tile, l = x - t * tile;
            const float tc = ((float)l + 0.5f) / (float)tile;
            const float off = (float)t * invNumTiles;
            const float a = off * (1.0f - tc), b = (off + invNumTiles) * tc;
            tab[(size_t)x] = make_float2(a + b, u2f((uint32_t)l | ((uint32_t)t << 16)));
        }
    };
    fill(renderW, tileW, renderW, tabX);
    fill(renderH, tileH, renderH, tabY);
    return 0;
}

void ptbd_build_tri_shade(const int32_t* vertIndices, int numIndices, const float* verticesUVX, const float* normalsUVY, std::vector<float4>& out)
{   // indices were validated by ptbd_build_tris
    out.assign((size_t)numIndices * 4, make_float4(0, 0, 0, 0));
    for (int s = 0; s < numIndices; s++)
    {
        const int32_t* vi = vertIndices + (size_t)s * 3;
        const float* n0 = normalsUVY + (size_t)vi[0] * 4; const float* n1 = normalsUVY + (size_t)vi[1] * 4; const float* n2 = normalsUVY + (size_t)vi[2] * 4;
        const float u0 = verticesUVX[(size_t)vi[0] * 4 + 3], u1 = verticesUVX[(size_t)vi[1] * 4 + 3], u2 = verticesUVX[(size_t)vi[2] * 4 + 3];
        float4* q = &out[(size_t)s * 4];
        q[0] = make_float4(n0[0], n0[1], n0[2], n1[0]);
        q[1] = make_float4(n1[1], n1[2], n2[0], n2[1]);
        q[2] = make_float4(n2[2], u0, n0[3], u1);
        q[3] = make_float4(n1[3], u2, n2[3], 0.f);
    }
}

namespace {

struct WideBuilder
{
    const float* N; int numNodes, numIndices;
    std::vector<float4>& wide;
    const std::vector<char>& transOnly;
    bool ok = true;
    int areaOrder = 0, areaOrderTlas = 0, topLevel = 1 << 30;      // areaOrderTlas: a different order for the TLAS nodes (study aid: PTB_WIDE_ORDER = 10 * tlas + blas)
    bool orderByNeed = false;       // permute the slots of a node by ascending stack need of the children (smaller worst-case stack, same boolean)
    std::vector<int> memo;          // binary inner node -> wide index (a BLAS shared by several instances is collapsed once)
    std::vector<int> need;          // stack entries needed below a wide node (siblings pushed while descending)

    WideBuilder(const float* n, int nn, int ni, std::vector<float4>& w, const std::vector<char>& to) : N(n), numNodes(nn), numIndices(ni), wide(w), transOnly(to), memo((size_t)nn, -1) {}
    uint32_t leafMeta(int node)
    {
        std::string err;
        uint32_t m = metaOf(N, node, numIndices, err);
        if (!err.empty()) ok = false;
        if ((m >> 30) == PTB_K_INST)
        {
            const uint32_t k = m & PTB_INST_INDEX_MASK;
            if (k >= transOnly.size()) ok = false;
            else if (transOnly[k]) m |= PTB_INST_TRANSLATION_ONLY;
        }
        return m;
    }

    static float area(const float* n) { float dx = n[3] - n[0], dy = n[4] - n[1], dz = n[5] - n[2]; return dx * dy + dy * dz + dz * dx; }
    bool inner(int i) const { return (int)N[(size_t)i * 9 + 8] == 0; }
    bool contains(int p, int c) const
    {
        const float* a = N + (size_t)p * 9; const float* b = N + (size_t)c * 9;
        return a[0] <= b[0] && a[1] <= b[1] && a[2] <= b[2] && a[3] >= b[3] && a[4] >= b[4] && a[5] >= b[5];
    }

    // collapse the binary subtree rooted at inner node r; returns the wide index
    int build(int r, int depth)
    {
        if (memo[(size_t)r] >= 0) return memo[(size_t)r];
        if (depth > 200) { ok = false; return 0; }
        int kids[4]; int nk = 0;
        auto addChildren = [&](int p, int at)
        {   // replace slot `at` (or append when at == nk) by the two children of p, left first
            int l = (int)N[(size_t)p * 9 + 6], rr = (int)N[(size_t)p * 9 + 7];
            if (l < 0 || l >= numNodes || rr < 0 || rr >= numNodes) { ok = false; return; }
            if (p != r && !(contains(p, l) && contains(p, rr))) ok = false;     // the dropped box must contain what replaces it
            if (at == nk) { kids[nk++] = l; kids[nk++] = rr; }
            else { for (int k = nk; k > at + 1; k--) kids[k] = kids[k - 1]; kids[at] = l; kids[at + 1] = rr; nk++; }
        };
        addChildren(r, 0);
        while (ok && nk < 4)
        {
            int best = -1; float bestA = -1.f;
            for (int k = 0; k < nk; k++) if (inner(kids[k])) { float a = area(N + (size_t)kids[k] * 9); if (a > bestA || best < 0) { bestA = a; best = k; } }
            if (best < 0) break;
            addChildren(kids[best], best);
        }
        const int w = (int)(wide.size() / 8);
        memo[(size_t)r] = w;
        wide.resize(wide.size() + 8, make_float4(0, 0, 0, 0));
        need.push_back(0);
        float box[24]; uint32_t meta[4];
        const float qnan = std::nanf("");
        int childNeed[4] = {0, 0, 0, 0};
        const int areaOrder = r >= topLevel ? (areaOrderTlas ? areaOrderTlas : this->areaOrder) : this->areaOrder;
        if (ok && areaOrder)
        {   // slots by a score, best first (stable): 2 / 3 box area descending / ascending; study modes of scripts/anyhit_order.py: 4 / 9 leaves first, then box area
            // descending / ascending; 8 leaves first, binary order otherwise
            float score[4];
            for (int k = 0; k < nk; k++)
            {
                const float a = area(N + (size_t)kids[k] * 9);
                const bool leaf = !inner(kids[k]);
                switch (areaOrder)
                {
                case 2: score[k] = a; break;
                case 3: score[k] = -a; break;
                case 4: score[k] = leaf ? 1e30f + a : a; break;
                case 8: score[k] = leaf ? 1.0f : 0.0f; break;
                default: score[k] = leaf ? 1e30f - a : -a; break;
                }
            }
            for (int a = 1; a < nk; a++) for (int b = a; b > 0 && score[b - 1] < score[b]; b--) { std::swap(score[b - 1], score[b]); std::swap(kids[b - 1], kids[b]); }
        }
        if (ok && orderByNeed)
        {   // Any-hit visits the hit children in slot order and keeps the later ones on the stack meanwhile: with the deepest subtree in the LAST slot nothing
            // waits on the stack while it is traversed.  The visiting order does not change the boolean (ptb_device.cuh: traverseWideAny), so the slots may be
            // permuted freely; a stable sort by the children's own stack need keeps the binary order among equals.
            int nd[4];
            for (int k = 0; k < nk; k++) nd[k] = inner(kids[k]) ? need[(size_t)build(kids[k], depth + 1)] + 1 : 0;
            for (int a = 1; a < nk; a++) for (int b = a; b > 0 && nd[b - 1] > nd[b]; b--) { std::swap(nd[b - 1], nd[b]); std::swap(kids[b - 1], kids[b]); }
        }
        for (int k = 0; k < 4; k++)
        {
            if (k >= nk || !ok) { for (int j = 0; j < 6; j++) box[k * 6 + j] = qnan; meta[k] = PTB_META_NONE; continue; }
            const float* n = N + (size_t)kids[k] * 9;
            // every binary node's box must contain its children's boxes for the equivalence argument (checked for the nodes kept as children too)
            if (inner(kids[k])) { int l = (int)n[6], rr = (int)n[7]; if (l < 0 || l >= numNodes || rr < 0 || rr >= numNodes || !contains(kids[k], l) || !contains(kids[k], rr)) ok = false; }
            for (int j = 0; j < 6; j++) box[k * 6 + j] = n[j];
            if (inner(kids[k])) { int cw = build(kids[k], depth + 1); meta[k] = (PTB_K_INNER << 30) | (uint32_t)cw; childNeed[k] = need[(size_t)cw]; }
            else meta[k] = leafMeta(kids[k]);
        }
        float4* q = &wide[(size_t)w * 8];
        for (int j = 0; j < 6; j++) q[j] = make_float4(box[j * 4 + 0], box[j * 4 + 1], box[j * 4 + 2], box[j * 4 + 3]);
        q[6] = make_float4(u2f(meta[0]), u2f(meta[1]), u2f(meta[2]), u2f(meta[3]));
        // stack entries needed below this node: while child k is traversed, the hit children in the slots after it wait on the stack
        int nd = 0;
        for (int k = 0; k < nk; k++) nd = std::max(nd, (nk - 1 - k) + childNeed[k]);
        need[(size_t)w] = nd;
        return w;
    }
};

} // namespace

// mode: slot order inside the wide nodes — 0 binary order, 1 ascending stack need, 2 / 3 descending / ascending box area
static void buildWideOnce(const float* N, int numNodes, int topLevelIndex, int numIndices, int numInstances, const std::vector<char>& transOnly, int mode, PtbDerivedWide& out)
{
    out = PtbDerivedWide();
    out.instRootMeta.assign((size_t)numInstances, PTB_META_NONE);
    WideBuilder B(N, numNodes, numIndices, out.wide, transOnly);
    B.orderByNeed = mode == 1; B.areaOrder = mode >= 2 ? mode % 10 : 0; B.areaOrderTlas = mode >= 10 ? mode / 10 : 0; B.topLevel = topLevelIndex;
    auto rootOf = [&](int node, int& needOut) -> uint32_t
    {
        if (node < 0 || node >= numNodes) { B.ok = false; return PTB_META_NONE; }
        if (B.inner(node)) { int w = B.build(node, 0); needOut = std::max(needOut, B.need[(size_t)w]); return (PTB_K_INNER << 30) | (uint32_t)w; }
        return B.leafMeta(node);
    };
    int needBlas = 0, needTlas = 0;
    for (int i = topLevelIndex; i < numNodes; i++)
    {
        const float* n = N + (size_t)i * 9;
        const int leaf = (int)n[8];
        if (leaf < 0 && -leaf - 1 < numInstances) out.instRootMeta[(size_t)(-leaf - 1)] = rootOf((int)n[6], needBlas);
    }
    out.rootMeta = rootOf(topLevelIndex, needTlas);
    out.stackDepth = std::max(4, 1 + needTlas + 1 + needBlas + 1);
    out.ok = B.ok && out.stackDepth <= 96 && out.wide.size() / 8 < (1u << 30);
}

// Slot order inside the wide nodes.  Any-hit may visit the children in any order (same boolean), and the order decides how soon an occluded ray finds its occluder:
// measured on hyperion's k_shadow — largest box first 5.84 ms, binary order 6.30, ascending stack need 6.62, smallest box first 6.55.  Largest box first ships.  Nearly all of it
// is decided at the TLAS (largest instance first + binary order inside the BLASes: 5.79; + leaves first inside the BLASes: 5.76 — PTB_WIDE_ORDER = 10 * tlas + blas).
// The order also sets the stack bound, i.e. how much of the SM's 256 KB is left as L1 next to 5 resident blocks of 256 stacks: up to 31 entries fit the 164 KB
// shared-memory configuration (92 KB of L1), more takes the 196 / 228 KB ones (60 / 28 KB of L1), which k_shadow feels.  When the shipped order
// needs more than 31 entries and sorting by stack need (deepest subtree last: nothing waits on the stack while it is traversed) needs fewer, that order is used
// instead — the 10 001-instance scene: 32 entries, k_shadow 123.6 -> 112.9 ms against binary order.  PTB_WIDE_ORDER = 0..3 forces one order (measurement aid).
void ptbd_build_wide(const float* N, int numNodes, int topLevelIndex, int numIndices, int numInstances, const std::vector<char>& transOnly, PtbDerivedWide& out)
{
    const char* e = getenv("PTB_WIDE_ORDER");
    if (e) { buildWideOnce(N, numNodes, topLevelIndex, numIndices, numInstances, transOnly, atoi(e), out); return; }
    buildWideOnce(N, numNodes, topLevelIndex, numIndices, numInstances, transOnly, 2, out);
    if (out.ok && out.stackDepth > 31)
    {
        PtbDerivedWide alt;
        buildWideOnce(N, numNodes, topLevelIndex, numIndices, numInstances, transOnly, 1, alt);
        if (alt.ok && alt.stackDepth < out.stackDepth) out = std::move(alt);
    }
}

// ------------------------------------------------------------------ TLAS rebuild (host, exact) ------------------------------------------------------------------
void ptbd_instance_bounds(const float* N, const float* transforms, int numInstances, const int32_t* blasRoot, std::vector<float>& out)
{
    out.resize((size_t)numInstances * 6);
    for (int i = 0; i < numInstances; i++)
    {   // Scene.cpp:154-184: rows of the transform scaled by the box corners, component-wise min / max, sums left to right, then the translation
        const float* b = N + (size_t)blasRoot[i] * 9;          // meshes[meshID]->bvh->Bounds() is the BLAS root's box
        const float* M = transforms + (size_t)i * 16;
        float lo[3], hi[3];
        for (int c = 0; c < 3; c++)
        {
            const float xa = M[0 + c] * b[0], xb = M[0 + c] * b[3];
            const float ya = M[4 + c] * b[1], yb = M[4 + c] * b[4];
            const float za = M[8 + c] * b[2], zb = M[8 + c] * b[5];
            lo[c] = ((std::min(xa, xb) + std::min(ya, yb)) + std::min(za, zb)) + M[12 + c];
            hi[c] = ((std::max(xa, xb) + std::max(ya, yb)) + std::max(za, zb)) + M[12 + c];
        }
        float* o = &out[(size_t)i * 6];
        o[0] = lo[0]; o[1] = lo[1]; o[2] = lo[2]; o[3] = hi[0]; o[4] = hi[1]; o[5] = hi[2];
    }
}

namespace {

struct Box3
{
    float lo[3], hi[3];
    Box3() { for (int a = 0; a < 3; a++) { lo[a] = std::numeric_limits<float>::max(); hi[a] = -std::numeric_limits<float>::max(); } }     // bbox() of bbox.h
    void grow(const float* p) { for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], p[a]); hi[a] = std::max(hi[a], p[a]); } }
    void growBox(const float* b) { for (int a = 0; a < 3; a++) { lo[a] = std::min(lo[a], b[a]); hi[a] = std::max(hi[a], b[3 + a]); } }
    int maxdim() const
    {   // bbox::maxdim (bbox.h)
        const float ex = hi[0] - lo[0], ey = hi[1] - lo[1], ez = hi[2] - lo[2];
        if (ex >= ey && ex >= ez) return 0;
        if (ey >= ex && ey >= ez) return 1;
        if (ez >= ex && ez >= ey) return 2;
        return 0;
    }
};

struct TlasBuilder
{
    const float* bounds; const float* cent; int* prim; float* out; int top; const int32_t* blasRoot; const int32_t* materialID;
    int cur = 0, height = 0;

    // Bvh::BuildNode (bvh.cpp:68-243, m_usesah == false) fused with ProcessTLASNodes (bvh_translator.cpp:58-86): both visit the nodes in the same pre-order
    int build(int startidx, int numprims, const Box3& nb, const Box3& cb, int level)
    {
        height = std::max(height, level);
        const int index = cur;
        float* n = out + (size_t)index * 9;
        n[0] = nb.lo[0]; n[1] = nb.lo[1]; n[2] = nb.lo[2]; n[3] = nb.hi[0]; n[4] = nb.hi[1]; n[5] = nb.hi[2];
        if (numprims < 2)
        {
            const int inst = prim[startidx];
            n[6] = (float)blasRoot[inst]; n[7] = (float)materialID[inst]; n[8] = (float)(-inst - 1);
            return index;
        }
        const int axis = cb.maxdim();
        const float border = (cb.hi[axis] + cb.lo[axis]) * 0.5f;                  // centroid_bounds.center()[axis]
        Box3 lb, rb, lcb, rcb;
        int splitidx = startidx;
        const bool near2far = ((numprims + startidx) & 1) != 0;
        if (cb.hi[axis] - cb.lo[axis] > 0.f)
        {
            int first = startidx, last = startidx + numprims;
            auto C = [&](int i) { return cent[(size_t)prim[i] * 3 + axis]; };
            auto growL = [&](int i) { lb.growBox(bounds + (size_t)prim[i] * 6); lcb.grow(cent + (size_t)prim[i] * 3); };
            auto growR = [&](int i) { rb.growBox(bounds + (size_t)prim[i] * 6); rcb.grow(cent + (size_t)prim[i] * 3); };
            while (true)
            {   // the reference's in-place partition, both orientations (`near2far` puts the smaller centroids left)
                while (first != last && (near2far ? C(first) < border : C(first) >= border)) { growL(first); ++first; }
                if (first == last--) break;
                growR(first);
                while (first != last && (near2far ? C(last) >= border : C(last) < border)) { growR(last); --last; }
                if (first == last) break;
                growL(last);
                std::swap(prim[first++], prim[last]);
            }
            splitidx = first;
        }
        if (splitidx == startidx || splitidx == startidx + numprims)
        {
            splitidx = startidx + (numprims >> 1);
            // (the reference grows the boxes accumulated above further here, without resetting them)
            for (int i = startidx; i < splitidx; ++i) { lb.growBox(bounds + (size_t)prim[i] * 6); lcb.grow(cent + (size_t)prim[i] * 3); }
            for (int i = splitidx; i < startidx + numprims; ++i) { rb.growBox(bounds + (size_t)prim[i] * 6); rcb.grow(cent + (size_t)prim[i] * 3); }
        }
        n[8] = 0.f;
        cur++;
        const int l = build(startidx, splitidx - startidx, lb, lcb, level + 1);
        cur++;
        const int r = build(splitidx, numprims - (splitidx - startidx), rb, rcb, level + 1);
        n = out + (size_t)index * 9;
        n[6] = (float)(top + l); n[7] = (float)(top + r);
        return index;
    }
};

} // namespace

int ptbd_build_tlas_host(const float* N, int topLevelIndex, const float* transforms, int numInstances, const int32_t* blasRoot, const int32_t* materialID,
                         std::vector<float>& tlasOut, int* heightOut, std::string& err)
{
    DREQ(numInstances > 0, 1, "no instances");
    std::vector<float> bounds; ptbd_instance_bounds(N, transforms, numInstances, blasRoot, bounds);
    std::vector<float> cent((size_t)numInstances * 3);
    std::vector<int> prim((size_t)numInstances);
    Box3 all, call;
    for (int i = 0; i < numInstances; i++)
    {
        const float* b = &bounds[(size_t)i * 6];
        all.growBox(b);                                                             // Bvh::Build: m_bounds
        for (int a = 0; a < 3; a++) cent[(size_t)i * 3 + a] = (b[3 + a] + b[a]) * 0.5f;       // bbox::center
        call.grow(&cent[(size_t)i * 3]);
        prim[(size_t)i] = i;
    }
    tlasOut.assign((size_t)numInstances * 2 * 9, 0.f);                              // BvhTranslator reserves 2 * numInstances slots (bvh_translator.cpp:97), 2n - 1 used
    TlasBuilder B{bounds.data(), cent.data(), prim.data(), tlasOut.data(), topLevelIndex, blasRoot, materialID};
    B.build(0, numInstances, all, call, 0);
    if (heightOut) *heightOut = B.height;
    return 0;
}

int ptbd_build_env_guide(const float* cdf, int w, int h, float totalSum, std::vector<uint32_t>& guide, float& scale)
{
    guide.clear(); scale = 0.f;
    const size_t n = (size_t)w * h;
    if (!cdf || n < 2 || n >= (1u << 30) || !(totalSum > 0.f) || !std::isfinite(totalSum)) return 1;
    for (size_t i = 0; i < n; i++)
        if (!std::isfinite(cdf[i]) || (i && cdf[i] < cdf[i - 1])) return 1;
    uint32_t G = 1024;
    while (G < (1u << 20) && (size_t)G * 2 <= n) G *= 2;          // about two texels per bucket, 4 KB .. 4 MB
    scale = (float)G / totalSum;
    if (!std::isfinite(scale) || !(scale > 0.f)) return 1;
    guide.assign((size_t)G + 1, 0u);
    const float top = (float)(G - 1);
    for (size_t i = 0; i < n; i++)
    {
        const float x = std::fmin(std::fmax(cdf[i] * scale, 0.0f), top);     // the same single-precision product and clamp as the device
        guide[(size_t)(int)x + 1]++;
    }
    for (uint32_t b = 0; b < G; b++) guide[b + 1] += guide[b];
    return 0;
}

void ptbd_wave_groups(int nSamples, int maxLps, int* lpsOut, int* lpwOut)
{
    int lps = 0;
    while (lps < 5 && lps < maxLps && (nSamples & ((2 << lps) - 1)) == 0) lps++;
    static const int lpwOf[6] = {3, 2, 2, 1, 1, 0};
    *lpsOut = lps; *lpwOut = lpwOf[lps];
}
