// pt_oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE (see pt_oracle.h).
//
// Restates, function by function, the reference's GLSL path tracer.  Every function cites the
// reference file:line it follows (paths relative to /root/reference/src/shaders unless noted).
// Built with -O2 -ffp-contract=off (no FMA contraction) so + - * / sqrt are IEEE and agree
// bit-for-bit with the CUDA traversal compiled with -fmad=false.
//
// GLSL built-ins whose precision/behaviour GL leaves open are pinned here:
//   normalize(v) = v / sqrt(dot(v,v));  inverse(mat4/mat3) = adjugate / det (fp32);
//   mat*vec sums left to right;  texture() = manual fp32 bilinear, REPEAT wrap, texel centres at +0.5;
//   min/max = fminf/fmaxf (return the non-NaN operand);  float->int conversion truncates.
#include "pt_oracle.h"
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <atomic>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

// ---------------------------------------------------------------- math -------------------------------------------
struct vec2 { float x, y; };
struct vec3 { float x, y, z; };
struct vec4 { float x, y, z, w; };

inline vec3 V3(float a) { return {a, a, a}; }
inline vec3 V3(float x, float y, float z) { return {x, y, z}; }
inline vec3 operator+(vec3 a, vec3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline vec3 operator-(vec3 a, vec3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline vec3 operator*(vec3 a, vec3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }
inline vec3 operator/(vec3 a, vec3 b) { return {a.x / b.x, a.y / b.y, a.z / b.z}; }
inline vec3 operator*(vec3 a, float s) { return {a.x * s, a.y * s, a.z * s}; }
inline vec3 operator*(float s, vec3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline vec3 operator/(vec3 a, float s) { return {a.x / s, a.y / s, a.z / s}; }
inline vec3 operator+(vec3 a, float s) { return {a.x + s, a.y + s, a.z + s}; }
inline vec3 operator-(vec3 a) { return {-a.x, -a.y, -a.z}; }
inline vec3& operator+=(vec3& a, vec3 b) { a = a + b; return a; }
inline vec3& operator*=(vec3& a, vec3 b) { a = a * b; return a; }
inline vec3& operator*=(vec3& a, float s) { a = a * s; return a; }
inline vec3& operator/=(vec3& a, float s) { a = a / s; return a; }
inline vec2 operator+(vec2 a, vec2 b) { return {a.x + b.x, a.y + b.y}; }
inline vec2 operator-(vec2 a, vec2 b) { return {a.x - b.x, a.y - b.y}; }
inline vec2 operator*(vec2 a, float s) { return {a.x * s, a.y * s}; }
inline float dot(vec3 a, vec3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline vec3 cross(vec3 a, vec3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline float length(vec3 a) { return sqrtf(dot(a, a)); }
inline vec3 normalize(vec3 a) { return a / sqrtf(dot(a, a)); }
inline float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
inline vec3 mix(vec3 a, vec3 b, float t) { return a * (1.0f - t) + b * t; }
inline float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
inline vec3 vmin(vec3 a, vec3 b) { return {fminf(a.x, b.x), fminf(a.y, b.y), fminf(a.z, b.z)}; }
inline vec3 vmax(vec3 a, vec3 b) { return {fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z)}; }
inline vec3 vpow(vec3 a, float e) { return {powf(a.x, e), powf(a.y, e), powf(a.z, e)}; }
inline vec3 vexp(vec3 a) { return {expf(a.x), expf(a.y), expf(a.z)}; }
inline vec3 reflect(vec3 I, vec3 N) { return I - N * (2.0f * dot(N, I)); }
inline vec3 refract(vec3 I, vec3 N, float eta)
{
    float d = dot(N, I);
    float k = 1.0f - eta * eta * (1.0f - d * d);
    if (k < 0.0f) return V3(0.0f);
    return I * eta - N * (eta * d + sqrtf(k));
}

// globals.glsl:25-31
const float PI = 3.14159265358979323f;
const float INV_PI = 0.31830988618379067f;
const float TWO_PI = 6.28318530717958648f;
const float INV_TWO_PI = 0.15915494309189533f;
const float INV_4_PI = 0.07957747154594766f;
const float EPS = 0.0003f;
const float INF = 1000000.0f;
enum { QUAD_LIGHT = 0, SPHERE_LIGHT = 1, DISTANT_LIGHT = 2 };
enum { ALPHA_MODE_OPAQUE = 0, ALPHA_MODE_BLEND = 1, ALPHA_MODE_MASK = 2 };
enum { MEDIUM_NONE = 0, MEDIUM_ABSORB = 1, MEDIUM_SCATTER = 2, MEDIUM_EMISSIVE = 3 };

// globals.glsl:46-139
struct Ray { vec3 origin, direction; };
struct Medium { int type; float density; vec3 color; float anisotropy; };
struct Material
{
    vec3 baseColor; float opacity; int alphaMode; float alphaCutoff; vec3 emission; float anisotropic, metallic, roughness,
        subsurface, specularTint, sheen, sheenTint, clearcoat, clearcoatRoughness, specTrans, ior, ax, ay;
    Medium medium;
};
struct Light { vec3 position, emission, u, v; float radius, area, type; };
struct State
{
    int depth; float eta, hitDist; vec3 fhp, normal, ffnormal, tangent, bitangent; bool isEmitter; vec2 texCoord; int matID;
    Material mat; Medium medium;
};
struct ScatterSampleRec { vec3 L, f; float pdf; };
struct LightSampleRec { vec3 normal, emission, direction; float dist, pdf; };

// globals.glsl:144-166 — pcg4d RNG; rand() in [0,1] inclusive.
struct Rng
{
    uint32_t x, y, z, w;
    void init(float px, float py, int frame)
    {   // InitRNG: seed = uvec4(p, uint(frame), uint(p.x) + uint(p.y))
        x = (uint32_t)px; y = (uint32_t)py; z = (uint32_t)frame; w = (uint32_t)px + (uint32_t)py;
    }
    void pcg4d()
    {
        x = x * 1664525u + 1013904223u; y = y * 1664525u + 1013904223u; z = z * 1664525u + 1013904223u; w = w * 1664525u + 1013904223u;
        x += y * w; y += z * x; z += x * y; w += y * z;
        x ^= x >> 16; y ^= y >> 16; z ^= z >> 16; w ^= w >> 16;
        x += y * w; y += z * x; z += x * y; w += y * z;
    }
    float rand() { pcg4d(); return (float)x / (float)0xffffffffu; }
};

inline float Luminance(vec3 c) { return 0.212671f * c.x + 0.715160f * c.y + 0.072169f * c.z; }   // globals.glsl:173

// ---------------------------------------------------------------- context ----------------------------------------
struct Counters { uint64_t closestRays = 0, anyRays = 0, nodeVisits = 0, internalSteps = 0, triTests = 0, tlasLeaves = 0, surfaceHits = 0,
                  anyNodeVisits = 0, anyInternalSteps = 0, anyTriTests = 0, anyTlasLeaves = 0; };

} // namespace

struct OrcCtx
{
    std::vector<float> nodes; int topLevelIndex = 0;
    std::vector<int32_t> vertIndices; std::vector<float> verticesUVX, normalsUVY, materials, transforms, lights;
    std::vector<uint8_t> textures; int numTextures = 0, texW = 0, texH = 0;
    std::vector<float> envImg, envCdf; int envW = 0, envH = 0; float envTotalSum = 0;
    int numLights = 0, numMaterials = 0, numInstances = 0;
    OrcOptions o;
    Counters total;
};

namespace {

struct Ctx   // per-evaluation view: scene + options + RNG of the current path ("global" shader state)
{
    const OrcCtx* s; const OrcOptions* o; Rng rng; Counters* cnt;
    float* capRay = nullptr; int capDepth = -1; bool capDone = false;     // orc_capture_rays: the ray traced at loop depth capDepth
    float* capShadow = nullptr; bool capShadowDone = false;              // orc_capture_shadow_rays: the light-NEE shadow ray of that depth
    float rand() { return rng.rand(); }
};

inline vec3 texel3(const std::vector<float>& a, int i) { return {a[i * 3 + 0], a[i * 3 + 1], a[i * 3 + 2]}; }

// inverse(mat4) of a row-major 4x4 (adjugate / determinant, fp32).
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

// inverse of the upper-left 3x3 of a row-major 4x4 (rows 0..2, cols 0..2), row-major 3x3 out.
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

// GLSL `mat4(r1..r4) * vec4(p, w)` where column k = row k of the reference's row-vector Mat4 (closest_hit.glsl:156-164):
// result[j] = p.x*D[0][j] + p.y*D[1][j] + p.z*D[2][j] + w*D[3][j]
inline vec3 xformPoint(const float* D, vec3 p, float w)
{
    return {p.x * D[0] + p.y * D[4] + p.z * D[8] + w * D[12], p.x * D[1] + p.y * D[5] + p.z * D[9] + w * D[13],
            p.x * D[2] + p.y * D[6] + p.z * D[10] + w * D[14]};
}

// ---------------------------------------------------------------- intersection.glsl ------------------------------
float SphereIntersect(float rad, vec3 pos, const Ray& r)   // intersection.glsl:25-45
{
    vec3 op = pos - r.origin;
    float eps = 0.001f;
    float b = dot(op, r.direction);
    float det = b * b - dot(op, op) + rad * rad;
    if (det < 0.0f) return INF;
    det = sqrtf(det);
    float t1 = b - det;
    if (t1 > eps) return t1;
    float t2 = b + det;
    if (t2 > eps) return t2;
    return INF;
}

float RectIntersect(vec3 pos, vec3 u, vec3 v, vec4 plane, const Ray& r)   // intersection.glsl:47-66
{
    vec3 n = V3(plane.x, plane.y, plane.z);
    float dt = dot(r.direction, n);
    float t = (plane.w - dot(n, r.origin)) / dt;
    if (t > EPS)
    {
        vec3 p = r.origin + r.direction * t;
        vec3 vi = p - pos;
        float a1 = dot(u, vi);
        if (a1 >= 0.0f && a1 <= 1.0f)
        {
            float a2 = dot(v, vi);
            if (a2 >= 0.0f && a2 <= 1.0f) return t;
        }
    }
    return INF;
}

// intersection.glsl:68-82.  Also returns the entry distance t0 (used only by the optional culled variant).
inline float AABBIntersect(vec3 minCorner, vec3 maxCorner, const Ray& r, float* entry)
{
    vec3 invDir = V3(1.0f) / r.direction;
    vec3 f = (maxCorner - r.origin) * invDir;
    vec3 n = (minCorner - r.origin) * invDir;
    vec3 tmax = vmax(f, n);
    vec3 tmin = vmin(f, n);
    float t1 = fminf(tmax.x, fminf(tmax.y, tmax.z));
    float t0 = fmaxf(tmin.x, fmaxf(tmin.y, tmin.z));
    *entry = t0;
    return (t1 >= t0) ? (t0 > 0.f ? t0 : t1) : -1.0f;
}

Light fetchLight(const OrcCtx* s, int i)
{
    Light l; const float* p = &s->lights[(size_t)i * 15];
    l.position = {p[0], p[1], p[2]}; l.emission = {p[3], p[4], p[5]}; l.u = {p[6], p[7], p[8]}; l.v = {p[9], p[10], p[11]};
    l.radius = p[12]; l.area = p[13]; l.type = p[14];
    return l;
}

struct HitInfo { int kind = 0, instance = -1, matID = -1, primSlot = -1, lightIdx = -1; int triID[3] = {-1, -1, -1}; vec3 bary{0, 0, 0}; };

// closest_hit.glsl:25-266
bool ClosestHit(Ctx& c, Ray r, State& state, LightSampleRec& lightSample, HitInfo* info = nullptr)
{
    const OrcCtx* s = c.s;
    if (c.cnt) c.cnt->closestRays++;
    float t = INF;
    float d;
    int lightIdx = -1;

    if (c.o->optLights)                                              // :28 OPT_LIGHTS
        if (!c.o->optHideEmitters || state.depth > 0)                // :31-34
            for (int i = 0; i < s->numLights; i++)
            {
                Light L = fetchLight(s, i);                          // :37-45
                vec3 u = L.u, v = L.v;
                if (L.type == (float)QUAD_LIGHT)
                {
                    vec3 normal = normalize(cross(u, v));
                    if (dot(normal, r.direction) > 0.f) continue;    // :50 hide backfacing quad light
                    vec4 plane = {normal.x, normal.y, normal.z, dot(normal, L.position)};
                    u *= 1.0f / dot(u, u);
                    v *= 1.0f / dot(v, v);
                    d = RectIntersect(L.position, u, v, plane, r);
                    if (d < 0.f) d = INF;
                    if (d < t)
                    {
                        t = d;
                        float cosTheta = dot(-r.direction, normal);
                        lightSample.pdf = (t * t) / (L.area * cosTheta);
                        lightSample.emission = L.emission;
                        state.isEmitter = true;
                        lightIdx = i;
                    }
                }
                if (L.type == (float)SPHERE_LIGHT)
                {
                    d = SphereIntersect(L.radius, L.position, r);
                    if (d < 0.f) d = INF;
                    if (d < t)
                    {
                        t = d;
                        vec3 hitPt = r.origin + t * r.direction;
                        float cosTheta = dot(-r.direction, normalize(hitPt - L.position));
                        lightSample.pdf = (t * t) / (L.area * cosTheta * 0.5f);
                        lightSample.emission = L.emission;
                        state.isEmitter = true;
                        lightIdx = i;
                    }
                }
            }

    // :88-218 BVH traversal
    int stack[64];
    int ptr = 0;
    stack[ptr++] = -1;
    int index = s->topLevelIndex;
    float leftHit = 0.0f, rightHit = 0.0f;
    int currMatID = 0, currInst = -1;
    bool BLAS = false;
    int triID[3] = {-1, -1, -1}; int primSlot = -1, hitInst = -1;
    float transMat[16] = {0}, transform[16] = {0}, invMat[16];
    vec3 bary = V3(0.f);
    vec4 vert0{}, vert1{}, vert2{};
    Ray rTrans = r;
    const float* N = s->nodes.data();
    const bool cull = c.o->cullBoxes != 0;

    while (index != -1)
    {
        if (c.cnt) c.cnt->nodeVisits++;
        int leftIndex = (int)N[index * 9 + 6];                     // :112 ivec3(texelFetch(BVH, index*3+2).xyz)
        int rightIndex = (int)N[index * 9 + 7];
        int leaf = (int)N[index * 9 + 8];

        if (leaf > 0)                                                // :118 BLAS leaf
        {
            for (int i = 0; i < rightIndex; i++)
            {
                if (c.cnt) c.cnt->triTests++;
                const int32_t* vi = &s->vertIndices[(size_t)(leftIndex + i) * 3];
                const float* p0 = &s->verticesUVX[(size_t)vi[0] * 4]; const float* p1 = &s->verticesUVX[(size_t)vi[1] * 4];
                const float* p2 = &s->verticesUVX[(size_t)vi[2] * 4];
                vec3 v0 = {p0[0], p0[1], p0[2]}, v1 = {p1[0], p1[1], p1[2]}, v2 = {p2[0], p2[1], p2[2]};
                vec3 e0 = v1 - v0;
                vec3 e1 = v2 - v0;
                vec3 pv = cross(rTrans.direction, e1);
                float det = dot(e0, pv);
                vec3 tv = rTrans.origin - v0;
                vec3 qv = cross(tv, e0);
                vec4 uvt;
                uvt.x = dot(tv, pv);
                uvt.y = dot(rTrans.direction, qv);
                uvt.z = dot(e1, qv);
                uvt.x = uvt.x / det; uvt.y = uvt.y / det; uvt.z = uvt.z / det;
                uvt.w = 1.0f - uvt.x - uvt.y;
                if (uvt.x >= 0.0f && uvt.y >= 0.0f && uvt.z >= 0.0f && uvt.w >= 0.0f && uvt.z < t)   // :143
                {
                    t = uvt.z;
                    triID[0] = vi[0]; triID[1] = vi[1]; triID[2] = vi[2];
                    state.matID = currMatID;
                    bary = {uvt.w, uvt.x, uvt.y};
                    vert0 = {p0[0], p0[1], p0[2], p0[3]}; vert1 = {p1[0], p1[1], p1[2], p1[3]}; vert2 = {p2[0], p2[1], p2[2], p2[3]};
                    memcpy(transform, transMat, sizeof(transform));
                    primSlot = leftIndex + i; hitInst = currInst;
                }
            }
        }
        else if (leaf < 0)                                           // :154 TLAS leaf
        {
            if (c.cnt) c.cnt->tlasLeaves++;
            memcpy(transMat, &s->transforms[(size_t)(-leaf - 1) * 16], sizeof(transMat));
            inverse4(transMat, invMat);
            rTrans.origin = xformPoint(invMat, r.origin, 1.0f);
            rTrans.direction = xformPoint(invMat, r.direction, 0.0f);
            stack[ptr++] = -1;                                       // marker
            index = leftIndex;
            BLAS = true;
            currMatID = rightIndex; currInst = -leaf - 1;
            continue;
        }
        else
        {
            if (c.cnt) c.cnt->internalSteps++;
            float e0, e1;
            leftHit = AABBIntersect(texel3(s->nodes, leftIndex * 3 + 0), texel3(s->nodes, leftIndex * 3 + 1), rTrans, &e0);
            rightHit = AABBIntersect(texel3(s->nodes, rightIndex * 3 + 0), texel3(s->nodes, rightIndex * 3 + 1), rTrans, &e1);
            if (cull) { const float tc = t * 1.00001f; if (leftHit > 0.0f && e0 > tc) leftHit = -1.0f; if (rightHit > 0.0f && e1 > tc) rightHit = -1.0f; }

            if (leftHit > 0.0f && rightHit > 0.0f)
            {
                int deferred = -1;
                if (leftHit > rightHit) { index = rightIndex; deferred = leftIndex; }
                else { index = leftIndex; deferred = rightIndex; }
                stack[ptr++] = deferred;
                continue;
            }
            else if (leftHit > 0.f) { index = leftIndex; continue; }
            else if (rightHit > 0.f) { index = rightIndex; continue; }
        }
        index = stack[--ptr];
        if (BLAS && index == -1)                                     // :208-216
        {
            BLAS = false;
            index = stack[--ptr];
            rTrans = r;
        }
    }

    if (info)
    {
        info->kind = (t == INF) ? 0 : (triID[0] != -1 ? 1 : 2);
        info->instance = triID[0] != -1 ? hitInst : -1; info->matID = triID[0] != -1 ? state.matID : -1;
        info->primSlot = primSlot; info->triID[0] = triID[0]; info->triID[1] = triID[1]; info->triID[2] = triID[2];
        info->bary = bary; info->lightIdx = (triID[0] == -1 && t != INF) ? lightIdx : -1;
    }

    if (t == INF) return false;                                      // :221
    state.hitDist = t;
    state.fhp = r.origin + r.direction * t;

    if (triID[0] != -1)                                              // :227
    {
        if (c.cnt) c.cnt->surfaceHits++;
        state.isEmitter = false;
        const float* q0 = &s->normalsUVY[(size_t)triID[0] * 4]; const float* q1 = &s->normalsUVY[(size_t)triID[1] * 4];
        const float* q2 = &s->normalsUVY[(size_t)triID[2] * 4];
        vec3 n0 = {q0[0], q0[1], q0[2]}, n1 = {q1[0], q1[1], q1[2]}, n2 = {q2[0], q2[1], q2[2]};
        vec2 t0 = {vert0.w, q0[3]}, t1 = {vert1.w, q1[3]}, t2 = {vert2.w, q2[3]};
        state.texCoord = t0 * bary.x + t1 * bary.y + t2 * bary.z;
        vec3 normal = normalize(n0 * bary.x + n1 * bary.y + n2 * bary.z);

        // :244 normalize(transpose(inverse(mat3(transform))) * normal) == (D3^-1) * normal, D3 row-major upper 3x3
        float inv3[9]; inverse3(transform, inv3);
        vec3 nw = {inv3[0] * normal.x + inv3[1] * normal.y + inv3[2] * normal.z, inv3[3] * normal.x + inv3[4] * normal.y + inv3[5] * normal.z,
                   inv3[6] * normal.x + inv3[7] * normal.y + inv3[8] * normal.z};
        state.normal = normalize(nw);
        state.ffnormal = dot(state.normal, r.direction) <= 0.0f ? state.normal : -state.normal;

        vec3 deltaPos1 = V3(vert1.x, vert1.y, vert1.z) - V3(vert0.x, vert0.y, vert0.z);
        vec3 deltaPos2 = V3(vert2.x, vert2.y, vert2.z) - V3(vert0.x, vert0.y, vert0.z);
        vec2 deltaUV1 = t1 - t0, deltaUV2 = t2 - t0;
        float invdet = 1.0f / (deltaUV1.x * deltaUV2.y - deltaUV1.y * deltaUV2.x);
        vec3 tg = (deltaPos1 * deltaUV2.y - deltaPos2 * deltaUV1.y) * invdet;
        vec3 bt = (deltaPos2 * deltaUV1.x - deltaPos1 * deltaUV2.x) * invdet;
        // mat3(transform) * v : result[j] = v.x*D[0][j] + v.y*D[1][j] + v.z*D[2][j]
        state.tangent = normalize(xformPoint(transform, tg, 0.0f));
        state.bitangent = normalize(xformPoint(transform, bt, 0.0f));
    }
    return true;
}

// texture(textureMapsArrayTex, vec3(uv, layer)): RGBA8 unorm, LINEAR, REPEAT (Renderer.cpp:200-208)
vec4 sampleTexArray(const OrcCtx* s, vec2 uv, float layerf)
{
    int layer = (int)floorf(layerf + 0.5f); layer = std::max(0, std::min(layer, s->numTextures - 1));
    int W = s->texW, H = s->texH;
    float x = uv.x * (float)W - 0.5f, y = uv.y * (float)H - 0.5f;
    float fx0 = floorf(x), fy0 = floorf(y);
    float ax = x - fx0, ay = y - fy0;
    auto wrap = [](float f, int n) { int i = (int)fmodf(f, (float)n); if (i < 0) i += n; return i; };
    int x0 = wrap(fx0, W), x1 = wrap(fx0 + 1.0f, W), y0 = wrap(fy0, H), y1 = wrap(fy0 + 1.0f, H);
    const uint8_t* base = &s->textures[(size_t)layer * W * H * 4];
    auto tx = [&](int xx, int yy, int ch) { return (float)base[((size_t)yy * W + xx) * 4 + ch] / 255.0f; };
    float out[4];
    for (int ch = 0; ch < 4; ch++)
        out[ch] = mixf(mixf(tx(x0, y0, ch), tx(x1, y0, ch), ax), mixf(tx(x0, y1, ch), tx(x1, y1, ch), ax), ay);
    return {out[0], out[1], out[2], out[3]};
}

// pathtrace.glsl:25-115
void GetMaterial(Ctx& c, State& state, const Ray& r)
{
    const OrcCtx* s = c.s;
    const float* P = &s->materials[(size_t)state.matID * 32];
    Material mat;
    mat.baseColor = {P[0], P[1], P[2]};
    mat.anisotropic = P[3];
    mat.emission = {P[4], P[5], P[6]};
    mat.metallic = P[8];
    mat.roughness = fmaxf(P[9], 0.001f);
    mat.subsurface = P[10];
    mat.specularTint = P[11];
    mat.sheen = P[12];
    mat.sheenTint = P[13];
    mat.clearcoat = P[14];
    mat.clearcoatRoughness = mixf(0.1f, 0.001f, P[15]);
    mat.specTrans = P[16];
    mat.ior = P[17];
    mat.medium.type = (int)P[18];
    mat.medium.density = P[19];
    mat.medium.color = {P[20], P[21], P[22]};
    mat.medium.anisotropy = clampf(P[23], -0.9f, 0.9f);
    int texIDs[4] = {(int)P[24], (int)P[25], (int)P[26], (int)P[27]};
    mat.opacity = P[28];
    mat.alphaMode = (int)P[29];
    mat.alphaCutoff = P[30];

    if (texIDs[0] >= 0)
    {
        vec4 col = sampleTexArray(s, state.texCoord, (float)texIDs[0]);
        mat.baseColor *= vpow(V3(col.x, col.y, col.z), 2.2f);
        mat.opacity *= col.w;
    }
    if (texIDs[1] >= 0)
    {
        vec4 mr = sampleTexArray(s, state.texCoord, (float)texIDs[1]);
        float m = mr.z, rg = mr.y;   // .bg
        mat.metallic = m;
        mat.roughness = fmaxf(rg * rg, 0.001f);
    }
    if (texIDs[2] >= 0)
    {
        vec4 tn = sampleTexArray(s, state.texCoord, (float)texIDs[2]);
        vec3 texNormal = {tn.x, tn.y, tn.z};
        if (c.o->optOpenglNormalMap) texNormal.y = 1.0f - texNormal.y;
        texNormal = normalize(texNormal * 2.0f - V3(1.0f));
        vec3 origNormal = state.normal;
        state.normal = normalize(state.tangent * texNormal.x + state.bitangent * texNormal.y + state.normal * texNormal.z);
        state.ffnormal = dot(origNormal, r.direction) <= 0.0f ? state.normal : -state.normal;
    }
    if (c.o->optRoughnessMollification)
        if (state.depth > 0)
            mat.roughness = fmaxf(mixf(0.0f, state.mat.roughness, c.o->roughnessMollificationAmt), mat.roughness);
    if (texIDs[3] >= 0)
    {
        vec4 e = sampleTexArray(s, state.texCoord, (float)texIDs[3]);
        mat.emission = vpow(V3(e.x, e.y, e.z), 2.2f);
    }
    float aspect = sqrtf(1.0f - mat.anisotropic * 0.9f);
    mat.ax = fmaxf(0.001f, mat.roughness / aspect);
    mat.ay = fmaxf(0.001f, mat.roughness * aspect);
    state.mat = mat;
    state.eta = dot(r.direction, state.normal) < 0.0f ? (1.0f / mat.ior) : mat.ior;
}

// anyhit.glsl:25-216
bool AnyHit(Ctx& c, Ray r, float maxDist, bool allowAlpha = true)
{
    const OrcCtx* s = c.s;
    if (c.cnt) c.cnt->anyRays++;
    if (c.o->optLights)
        for (int i = 0; i < s->numLights; i++)
        {
            Light L = fetchLight(s, i);
            vec3 u = L.u, v = L.v;
            if (L.type == (float)QUAD_LIGHT)
            {
                vec3 normal = normalize(cross(u, v));
                vec4 plane = {normal.x, normal.y, normal.z, dot(normal, L.position)};
                u *= 1.0f / dot(u, u);
                v *= 1.0f / dot(v, v);
                float d = RectIntersect(L.position, u, v, plane, r);
                if (d > 0.0f && d < maxDist) return true;
            }
            if (L.type == (float)SPHERE_LIGHT)
            {
                float d = SphereIntersect(L.radius, L.position, r);
                if (d > 0.0f && d < maxDist) return true;
            }
        }

    const bool alphaTest = allowAlpha && c.o->optAlphaTest && !c.o->optMedium;    // :74,118
    int stack[64];
    int ptr = 0;
    stack[ptr++] = -1;
    int index = s->topLevelIndex;
    float leftHit = 0.0f, rightHit = 0.0f;
    int currMatID = 0;
    bool BLAS = false;
    Ray rTrans = r;
    float invMat[16];
    const float* N = s->nodes.data();
    const bool cull = c.o->cullBoxes != 0;

    while (index != -1)
    {
        if (c.cnt) c.cnt->anyNodeVisits++;
        int leftIndex = (int)N[index * 9 + 6], rightIndex = (int)N[index * 9 + 7], leaf = (int)N[index * 9 + 8];
        if (leaf > 0)
        {
            for (int i = 0; i < rightIndex; i++)
            {
                if (c.cnt) c.cnt->anyTriTests++;
                const int32_t* vi = &s->vertIndices[(size_t)(leftIndex + i) * 3];
                const float* p0 = &s->verticesUVX[(size_t)vi[0] * 4]; const float* p1 = &s->verticesUVX[(size_t)vi[1] * 4];
                const float* p2 = &s->verticesUVX[(size_t)vi[2] * 4];
                vec3 v0 = {p0[0], p0[1], p0[2]}, v1 = {p1[0], p1[1], p1[2]}, v2 = {p2[0], p2[1], p2[2]};
                vec3 e0 = v1 - v0;
                vec3 e1 = v2 - v0;
                vec3 pv = cross(rTrans.direction, e1);
                float det = dot(e0, pv);
                vec3 tv = rTrans.origin - v0;
                vec3 qv = cross(tv, e0);
                vec4 uvt;
                uvt.x = dot(tv, pv);
                uvt.y = dot(rTrans.direction, qv);
                uvt.z = dot(e1, qv);
                uvt.x = uvt.x / det; uvt.y = uvt.y / det; uvt.z = uvt.z / det;
                uvt.w = 1.0f - uvt.x - uvt.y;
                if (uvt.x >= 0.0f && uvt.y >= 0.0f && uvt.z >= 0.0f && uvt.w >= 0.0f && uvt.z < maxDist)
                {
                    if (alphaTest)
                    {
                        vec2 t0 = {p0[3], s->normalsUVY[(size_t)vi[0] * 4 + 3]}, t1 = {p1[3], s->normalsUVY[(size_t)vi[1] * 4 + 3]},
                             t2 = {p2[3], s->normalsUVY[(size_t)vi[2] * 4 + 3]};
                        vec2 texCoord = t0 * uvt.w + t1 * uvt.x + t2 * uvt.y;
                        const float* P = &s->materials[(size_t)currMatID * 32];
                        // :129 texture() with texIDs.x even when it is -1: GL clamps the layer to [0, layers-1]; with no
                        // textures bound the sample returns (0,0,0,1) (incomplete texture) -> alpha 1.
                        float alpha = s->numTextures > 0 ? sampleTexArray(s, texCoord, P[24]).w : 1.0f;
                        float opacity = P[28];
                        int alphaMode = (int)P[29];
                        float alphaCutoff = P[30];
                        opacity *= alpha;
                        if (!((alphaMode == ALPHA_MODE_MASK && opacity < alphaCutoff) || (alphaMode == ALPHA_MODE_BLEND && c.rand() > opacity)))
                            return true;
                    }
                    else
                        return true;
                }
            }
        }
        else if (leaf < 0)
        {
            if (c.cnt) c.cnt->anyTlasLeaves++;
            inverse4(&s->transforms[(size_t)(-leaf - 1) * 16], invMat);
            rTrans.origin = xformPoint(invMat, r.origin, 1.0f);
            rTrans.direction = xformPoint(invMat, r.direction, 0.0f);
            stack[ptr++] = -1;
            index = leftIndex;
            BLAS = true;
            currMatID = rightIndex;
            continue;
        }
        else
        {
            if (c.cnt) c.cnt->anyInternalSteps++;
            float e0, e1;
            leftHit = AABBIntersect(texel3(s->nodes, leftIndex * 3 + 0), texel3(s->nodes, leftIndex * 3 + 1), rTrans, &e0);
            rightHit = AABBIntersect(texel3(s->nodes, rightIndex * 3 + 0), texel3(s->nodes, rightIndex * 3 + 1), rTrans, &e1);
            if (cull) { const float tc = maxDist * 1.00001f; if (leftHit > 0.0f && e0 > tc) leftHit = -1.0f; if (rightHit > 0.0f && e1 > tc) rightHit = -1.0f; }
            if (leftHit > 0.0f && rightHit > 0.0f)
            {
                int deferred = -1;
                if (leftHit > rightHit) { index = rightIndex; deferred = leftIndex; }
                else { index = leftIndex; deferred = rightIndex; }
                stack[ptr++] = deferred;
                continue;
            }
            else if (leftHit > 0.f) { index = leftIndex; continue; }
            else if (rightHit > 0.f) { index = rightIndex; continue; }
        }
        index = stack[--ptr];
        if (BLAS && index == -1)
        {
            BLAS = false;
            index = stack[--ptr];
            rTrans = r;
        }
    }
    return false;
}

// ---------------------------------------------------------------- sampling.glsl ----------------------------------
float GTR1(float NDotH, float a)   // :25-32
{
    if (a >= 1.0f) return INV_PI;
    float a2 = a * a;
    float t = 1.0f + (a2 - 1.0f) * NDotH * NDotH;
    return (a2 - 1.0f) / (PI * logf(a2) * t);
}
vec3 SampleGTR1(float rgh, float r1, float r2)   // :34-47
{
    float a = fmaxf(0.001f, rgh);
    float a2 = a * a;
    float phi = r1 * TWO_PI;
    float cosTheta = sqrtf((1.0f - powf(a2, 1.0f - r2)) / (1.0f - a2));
    float sinTheta = clampf(sqrtf(1.0f - (cosTheta * cosTheta)), 0.0f, 1.0f);
    float sinPhi = sinf(phi), cosPhi = cosf(phi);
    return {sinTheta * cosPhi, sinTheta * sinPhi, cosTheta};
}
vec3 SampleGGXVNDF(vec3 V, float ax, float ay, float r1, float r2)   // :70-88
{
    vec3 Vh = normalize(V3(ax * V.x, ay * V.y, V.z));
    float lensq = Vh.x * Vh.x + Vh.y * Vh.y;
    vec3 T1 = lensq > 0 ? V3(-Vh.y, Vh.x, 0) * (1.0f / sqrtf(lensq)) : V3(1, 0, 0);
    vec3 T2 = cross(Vh, T1);
    float r = sqrtf(r1);
    float phi = 2.0f * PI * r2;
    float t1 = r * cosf(phi);
    float t2 = r * sinf(phi);
    float s = 0.5f * (1.0f + Vh.z);
    t2 = (1.0f - s) * sqrtf(1.0f - t1 * t1) + s * t2;
    vec3 Nh = t1 * T1 + t2 * T2 + sqrtf(fmaxf(0.0f, 1.0f - t1 * t1 - t2 * t2)) * Vh;
    return normalize(V3(ax * Nh.x, ay * Nh.y, fmaxf(0.0f, Nh.z)));
}
float GTR2Aniso(float NDotH, float HDotX, float HDotY, float ax, float ay)   // :90-96
{
    float a = HDotX / ax;
    float b = HDotY / ay;
    float c = a * a + b * b + NDotH * NDotH;
    return 1.0f / (PI * ax * ay * c * c);
}
float SmithG(float NDotV, float alphaG)   // :109-114
{
    float a = alphaG * alphaG;
    float b = NDotV * NDotV;
    return (2.0f * NDotV) / (NDotV + sqrtf(a + b - a * b));
}
float SmithGAniso(float NDotV, float VDotX, float VDotY, float ax, float ay)   // :116-122
{
    float a = VDotX * ax;
    float b = VDotY * ay;
    float c = NDotV;
    return (2.0f * NDotV) / (NDotV + sqrtf(a * a + b * b + c * c));
}
float SchlickWeight(float u)   // :124-129
{
    float m = clampf(1.0f - u, 0.0f, 1.0f);
    float m2 = m * m;
    return m2 * m2 * m;
}
float DielectricFresnel(float cosThetaI, float eta)   // :131-145
{
    float sinThetaTSq = eta * eta * (1.0f - cosThetaI * cosThetaI);
    if (sinThetaTSq > 1.0f) return 1.0f;
    float cosThetaT = sqrtf(fmaxf(1.0f - sinThetaTSq, 0.0f));
    float rs = (eta * cosThetaT - cosThetaI) / (eta * cosThetaT + cosThetaI);
    float rp = (eta * cosThetaI - cosThetaT) / (eta * cosThetaI + cosThetaT);
    return 0.5f * (rs * rs + rp * rp);
}
vec3 CosineSampleHemisphere(float r1, float r2)   // :147-156
{
    vec3 dir;
    float r = sqrtf(r1);
    float phi = TWO_PI * r2;
    dir.x = r * cosf(phi);
    dir.y = r * sinf(phi);
    dir.z = sqrtf(fmaxf(0.0f, 1.0f - dir.x * dir.x - dir.y * dir.y));
    return dir;
}
vec3 UniformSampleHemisphere(float r1, float r2)   // :158-163
{
    float r = sqrtf(fmaxf(0.0f, 1.0f - r1 * r1));
    float phi = TWO_PI * r2;
    return {r * cosf(phi), r * sinf(phi), r1};
}
float PowerHeuristic(float a, float b)   // :173-177
{
    float t = a * a;
    return t / (b * b + t);
}
void Onb(vec3 N, vec3& T, vec3& B)   // :179-184
{
    vec3 up = fabsf(N.z) < 0.9999999f ? V3(0, 0, 1) : V3(1, 0, 0);
    T = normalize(cross(up, N));
    B = cross(N, T);
}
void SampleSphereLight(Ctx& c, const Light& light, vec3 scatterPos, LightSampleRec& ls)   // :186-212
{
    float r1 = c.rand();
    float r2 = c.rand();
    vec3 sphereCentertoSurface = scatterPos - light.position;
    float distToSphereCenter = length(sphereCentertoSurface);
    vec3 sampledDir;
    sphereCentertoSurface /= distToSphereCenter;
    sampledDir = UniformSampleHemisphere(r1, r2);
    vec3 T, B;
    Onb(sphereCentertoSurface, T, B);
    sampledDir = T * sampledDir.x + B * sampledDir.y + sphereCentertoSurface * sampledDir.z;
    vec3 lightSurfacePos = light.position + sampledDir * light.radius;
    ls.direction = lightSurfacePos - scatterPos;
    ls.dist = length(ls.direction);
    float distSq = ls.dist * ls.dist;
    ls.direction /= ls.dist;
    ls.normal = normalize(lightSurfacePos - light.position);
    ls.emission = light.emission * (float)c.s->numLights;
    ls.pdf = distSq / (light.area * 0.5f * fabsf(dot(ls.normal, ls.direction)));
}
void SampleRectLight(Ctx& c, const Light& light, vec3 scatterPos, LightSampleRec& ls)   // :214-227
{
    float r1 = c.rand();
    float r2 = c.rand();
    vec3 lightSurfacePos = light.position + light.u * r1 + light.v * r2;
    ls.direction = lightSurfacePos - scatterPos;
    ls.dist = length(ls.direction);
    float distSq = ls.dist * ls.dist;
    ls.direction /= ls.dist;
    ls.normal = normalize(cross(light.u, light.v));
    ls.emission = light.emission * (float)c.s->numLights;
    ls.pdf = distSq / (light.area * fabsf(dot(ls.normal, ls.direction)));
}
void SampleDistantLight(Ctx& c, const Light& light, vec3 scatterPos, LightSampleRec& ls)   // :229-236
{
    ls.direction = normalize(light.position - V3(0.0f));
    ls.normal = normalize(scatterPos - light.position);
    ls.emission = light.emission * (float)c.s->numLights;
    ls.dist = INF;
    ls.pdf = 1.0f;
}
void SampleOneLight(Ctx& c, const Light& light, vec3 scatterPos, LightSampleRec& ls)   // :238-248
{
    int type = (int)light.type;
    if (type == QUAD_LIGHT) SampleRectLight(c, light, scatterPos, ls);
    else if (type == SPHERE_LIGHT) SampleSphereLight(c, light, scatterPos, ls);
    else SampleDistantLight(c, light, scatterPos, ls);
}
vec3 SampleHG(vec3 V, float g, float r1, float r2)   // :250-270
{
    float cosTheta;
    if (fabsf(g) < 0.001f) cosTheta = 1 - 2 * r2;
    else
    {
        float sqrTerm = (1 - g * g) / (1 + g - 2 * g * r2);
        cosTheta = -(1 + g * g - sqrTerm * sqrTerm) / (2 * g);
    }
    float phi = r1 * TWO_PI;
    float sinTheta = clampf(sqrtf(1.0f - (cosTheta * cosTheta)), 0.0f, 1.0f);
    float sinPhi = sinf(phi), cosPhi = cosf(phi);
    vec3 v1, v2;
    Onb(V, v1, v2);
    return sinTheta * cosPhi * v1 + sinTheta * sinPhi * v2 + cosTheta * V;
}
float PhaseHG(float cosTheta, float g)   // :272-276
{
    float denom = 1 + g * g + 2 * g * cosTheta;
    return INV_4_PI * (1 - g * g) / (denom * sqrtf(denom));
}

// ---------------------------------------------------------------- envmap.glsl ------------------------------------
// texture(envMapTex, uv): RGB32F, LINEAR, REPEAT (GL default wrap; Renderer.cpp:213-218)
vec3 sampleEnv(const OrcCtx* s, vec2 uv)
{
    int W = s->envW, H = s->envH;
    float x = uv.x * (float)W - 0.5f, y = uv.y * (float)H - 0.5f;
    float fx0 = floorf(x), fy0 = floorf(y);
    float ax = x - fx0, ay = y - fy0;
    auto wrap = [](float f, int n) { int i = (int)fmodf(f, (float)n); if (i < 0) i += n; return i; };
    int x0 = wrap(fx0, W), x1 = wrap(fx0 + 1.0f, W), y0 = wrap(fy0, H), y1 = wrap(fy0 + 1.0f, H);
    auto tx = [&](int xx, int yy) { const float* p = &s->envImg[((size_t)yy * W + xx) * 3]; return V3(p[0], p[1], p[2]); };
    return mix(mix(tx(x0, y0), tx(x1, y0), ax), mix(tx(x0, y1), tx(x1, y1), ax), ay);
}
vec2 BinarySearch(const OrcCtx* s, float value)   // :28-55
{
    int W = s->envW, H = s->envH;
    int lower = 0, upper = H - 1;
    while (lower < upper)
    {
        int mid = (lower + upper) >> 1;
        if (value < s->envCdf[(size_t)mid * W + (W - 1)]) upper = mid;
        else lower = mid + 1;
    }
    int y = std::max(0, std::min(lower, H - 1));
    lower = 0; upper = W - 1;
    while (lower < upper)
    {
        int mid = (lower + upper) >> 1;
        if (value < s->envCdf[(size_t)y * W + mid]) upper = mid;
        else lower = mid + 1;
    }
    int x = std::max(0, std::min(lower, W - 1));
    return {(float)x / (float)W, (float)y / (float)H};
}
vec4 EvalEnvMap(Ctx& c, const Ray& r)   // :57-66
{
    const OrcCtx* s = c.s;
    float theta = acosf(clampf(r.direction.y, -1.0f, 1.0f));
    vec2 uv = {(PI + atan2f(r.direction.z, r.direction.x)) * INV_TWO_PI + c.o->envMapRot, theta * INV_PI + 0.0f};
    vec3 color = sampleEnv(s, uv);
    float pdf = Luminance(color) / s->envTotalSum;
    return {color.x, color.y, color.z, (pdf * (float)s->envW * (float)s->envH) / (TWO_PI * PI * sinf(theta))};
}
vec4 SampleEnvMap(Ctx& c, vec3& color)   // :68-83
{
    const OrcCtx* s = c.s;
    vec2 uv = BinarySearch(s, c.rand() * s->envTotalSum);
    color = sampleEnv(s, uv);
    float pdf = Luminance(color) / s->envTotalSum;
    uv.x -= c.o->envMapRot;
    float phi = uv.x * TWO_PI;
    float theta = uv.y * PI;
    if (sinf(theta) == 0.0f) pdf = 0.0f;
    return {-sinf(theta) * cosf(phi), cosf(theta), -sinf(theta) * sinf(phi), (pdf * (float)s->envW * (float)s->envH) / (TWO_PI * PI * sinf(theta))};
}

// ---------------------------------------------------------------- disney.glsl ------------------------------------
inline vec3 ToWorld(vec3 X, vec3 Y, vec3 Z, vec3 V) { return V.x * X + V.y * Y + V.z * Z; }         // :39-42
inline vec3 ToLocal(vec3 X, vec3 Y, vec3 Z, vec3 V) { return {dot(V, X), dot(V, Y), dot(V, Z)}; }   // :44-47

void TintColors(const Material& mat, float eta, float& F0, vec3& Csheen, vec3& Cspec0)   // :49-59
{
    float lum = Luminance(mat.baseColor);
    vec3 ctint = lum > 0.0f ? mat.baseColor / lum : V3(1.0f);
    F0 = (1.0f - eta) / (1.0f + eta);
    F0 *= F0;
    Cspec0 = F0 * mix(V3(1.0f), ctint, mat.specularTint);
    Csheen = mix(V3(1.0f), ctint, mat.sheenTint);
}
vec3 EvalDisneyDiffuse(const Material& mat, vec3 Csheen, vec3 V, vec3 L, vec3 H, float& pdf)   // :61-87
{
    pdf = 0.0f;
    if (L.z <= 0.0f) return V3(0.0f);
    float LDotH = dot(L, H);
    float Rr = 2.0f * mat.roughness * LDotH * LDotH;
    float FL = SchlickWeight(L.z);
    float FV = SchlickWeight(V.z);
    float Fretro = Rr * (FL + FV + FL * FV * (Rr - 1.0f));
    float Fd = (1.0f - 0.5f * FL) * (1.0f - 0.5f * FV);
    float Fss90 = 0.5f * Rr;
    float Fss = mixf(1.0f, Fss90, FL) * mixf(1.0f, Fss90, FV);
    float ss = 1.25f * (Fss * (1.0f / (L.z + V.z) - 0.5f) + 0.5f);
    float FH = SchlickWeight(LDotH);
    vec3 Fsheen = FH * mat.sheen * Csheen;
    pdf = L.z * INV_PI;
    return INV_PI * mat.baseColor * mixf(Fd + Fretro, ss, mat.subsurface) + Fsheen;
}
vec3 EvalMicrofacetReflection(const Material& mat, vec3 V, vec3 L, vec3 H, vec3 F, float& pdf)   // :89-101
{
    pdf = 0.0f;
    if (L.z <= 0.0f) return V3(0.0f);
    float D = GTR2Aniso(H.z, H.x, H.y, mat.ax, mat.ay);
    float G1 = SmithGAniso(fabsf(V.z), V.x, V.y, mat.ax, mat.ay);
    float G2 = G1 * SmithGAniso(fabsf(L.z), L.x, L.y, mat.ax, mat.ay);
    pdf = G1 * D / (4.0f * V.z);
    return F * D * G2 / (4.0f * L.z * V.z);
}
vec3 EvalMicrofacetRefraction(const Material& mat, float eta, vec3 V, vec3 L, vec3 H, vec3 F, float& pdf)   // :103-123
{
    pdf = 0.0f;
    if (L.z >= 0.0f) return V3(0.0f);
    float LDotH = dot(L, H);
    float VDotH = dot(V, H);
    float D = GTR2Aniso(H.z, H.x, H.y, mat.ax, mat.ay);
    float G1 = SmithGAniso(fabsf(V.z), V.x, V.y, mat.ax, mat.ay);
    float G2 = G1 * SmithGAniso(fabsf(L.z), L.x, L.y, mat.ax, mat.ay);
    float denom = LDotH + VDotH * eta;
    denom *= denom;
    float eta2 = eta * eta;
    float jacobian = fabsf(LDotH) / denom;
    pdf = G1 * fmaxf(0.0f, VDotH) * D * jacobian / V.z;
    return vpow(mat.baseColor, 0.5f) * (V3(1.0f) - F) * D * G2 * fabsf(VDotH) * jacobian * eta2 / fabsf(L.z * V.z);
}
vec3 EvalClearcoat(const Material& mat, vec3 V, vec3 L, vec3 H, float& pdf)   // :125-140
{
    pdf = 0.0f;
    if (L.z <= 0.0f) return V3(0.0f);
    float VDotH = dot(V, H);
    float F = mixf(0.04f, 1.0f, SchlickWeight(VDotH));
    float D = GTR1(H.z, mat.clearcoatRoughness);
    float G = SmithG(L.z, 0.25f) * SmithG(V.z, 0.25f);
    float jacobian = 1.0f / (4.0f * VDotH);
    pdf = D * H.z * jacobian;
    return V3(F) * D * G;
}

struct LobePr { float diffPr, dielectricPr, metalPr, glassPr, clearCtPr, dielectricWt, metalWt, glassWt; };
LobePr lobeProbabilities(const Material& m, vec3 Cspec0, float Vz)   // :155-179 / :272-294
{
    LobePr p;
    p.dielectricWt = (1.0f - m.metallic) * (1.0f - m.specTrans);
    p.metalWt = m.metallic;
    p.glassWt = (1.0f - m.metallic) * m.specTrans;
    float schlickWt = SchlickWeight(Vz);
    p.diffPr = p.dielectricWt * Luminance(m.baseColor);
    p.dielectricPr = p.dielectricWt * Luminance(mix(Cspec0, V3(1.0f), schlickWt));
    p.metalPr = p.metalWt * Luminance(mix(m.baseColor, V3(1.0f), schlickWt));
    p.glassPr = p.glassWt;
    p.clearCtPr = 0.25f * m.clearcoat;
    float invTotalWt = 1.0f / (p.diffPr + p.dielectricPr + p.metalPr + p.glassPr + p.clearCtPr);
    p.diffPr *= invTotalWt; p.dielectricPr *= invTotalWt; p.metalPr *= invTotalWt; p.glassPr *= invTotalWt; p.clearCtPr *= invTotalWt;
    return p;
}

vec3 DisneyEval(const State& state, vec3 V, vec3 N, vec3 L, float& pdf)   // :244-351
{
    pdf = 0.0f;
    vec3 f = V3(0.0f);
    vec3 T, B;
    Onb(N, T, B);
    V = ToLocal(T, B, N, V);
    L = ToLocal(T, B, N, L);
    vec3 H;
    if (L.z > 0.0f) H = normalize(L + V);
    else H = normalize(L + V * state.eta);
    if (H.z < 0.0f) H = -H;
    vec3 Csheen, Cspec0;
    float F0;
    TintColors(state.mat, state.eta, F0, Csheen, Cspec0);
    LobePr p = lobeProbabilities(state.mat, Cspec0, V.z);
    bool reflect = L.z * V.z > 0;
    float tmpPdf = 0.0f;
    float VDotH = fabsf(dot(V, H));

    if (p.diffPr > 0.0f && reflect)
    {
        f += EvalDisneyDiffuse(state.mat, Csheen, V, L, H, tmpPdf) * p.dielectricWt;
        pdf += tmpPdf * p.diffPr;
    }
    if (p.dielectricPr > 0.0f && reflect)
    {
        float F = (DielectricFresnel(VDotH, 1.0f / state.mat.ior) - F0) / (1.0f - F0);
        f += EvalMicrofacetReflection(state.mat, V, L, H, mix(Cspec0, V3(1.0f), F), tmpPdf) * p.dielectricWt;
        pdf += tmpPdf * p.dielectricPr;
    }
    if (p.metalPr > 0.0f && reflect)
    {
        vec3 F = mix(state.mat.baseColor, V3(1.0f), SchlickWeight(VDotH));
        f += EvalMicrofacetReflection(state.mat, V, L, H, F, tmpPdf) * p.metalWt;
        pdf += tmpPdf * p.metalPr;
    }
    if (p.glassPr > 0.0f)
    {
        float F = DielectricFresnel(VDotH, state.eta);
        if (reflect)
        {
            f += EvalMicrofacetReflection(state.mat, V, L, H, V3(F), tmpPdf) * p.glassWt;
            pdf += tmpPdf * p.glassPr * F;
        }
        else
        {
            f += EvalMicrofacetRefraction(state.mat, state.eta, V, L, H, V3(F), tmpPdf) * p.glassWt;
            pdf += tmpPdf * p.glassPr * (1.0f - F);
        }
    }
    if (p.clearCtPr > 0.0f && reflect)
    {
        f += EvalClearcoat(state.mat, V, L, H, tmpPdf) * 0.25f * state.mat.clearcoat;
        pdf += tmpPdf * p.clearCtPr;
    }
    return f * fabsf(L.z);
}

vec3 DisneySampleR(const State& state, vec3 V, vec3 N, vec3& L, float& pdf, float r1, float r2, float r3)   // :142-242 (rands passed in)
{
    pdf = 0.0f;
    vec3 T, B;
    Onb(N, T, B);
    V = ToLocal(T, B, N, V);
    vec3 Csheen, Cspec0;
    float F0;
    TintColors(state.mat, state.eta, F0, Csheen, Cspec0);
    LobePr p = lobeProbabilities(state.mat, Cspec0, V.z);
    float cdf[5];
    cdf[0] = p.diffPr;
    cdf[1] = cdf[0] + p.dielectricPr;
    cdf[2] = cdf[1] + p.metalPr;
    cdf[3] = cdf[2] + p.glassPr;
    cdf[4] = cdf[3] + p.clearCtPr;

    if (r3 < cdf[0]) L = CosineSampleHemisphere(r1, r2);
    else if (r3 < cdf[2])
    {
        vec3 H = SampleGGXVNDF(V, state.mat.ax, state.mat.ay, r1, r2);
        if (H.z < 0.0f) H = -H;
        L = normalize(reflect(-V, H));
    }
    else if (r3 < cdf[3])
    {
        vec3 H = SampleGGXVNDF(V, state.mat.ax, state.mat.ay, r1, r2);
        float F = DielectricFresnel(fabsf(dot(V, H)), state.eta);
        if (H.z < 0.0f) H = -H;
        r3 = (r3 - cdf[2]) / (cdf[3] - cdf[2]);
        if (r3 < F) L = normalize(reflect(-V, H));
        else L = normalize(refract(-V, H, state.eta));
    }
    else
    {
        vec3 H = SampleGTR1(state.mat.clearcoatRoughness, r1, r2);
        if (H.z < 0.0f) H = -H;
        L = normalize(reflect(-V, H));
    }
    L = ToWorld(T, B, N, L);
    V = ToWorld(T, B, N, V);
    return DisneyEval(state, V, N, L, pdf);
}
vec3 DisneySample(Ctx& c, const State& state, vec3 V, vec3 N, vec3& L, float& pdf)
{
    float r1 = c.rand();      // disney.glsl:146-147
    float r2 = c.rand();
    float r3 = c.rand();      // :192 (no rand() call between -> same stream order)
    return DisneySampleR(state, V, N, L, pdf, r1, r2, r3);
}

// ---------------------------------------------------------------- pathtrace.glsl ---------------------------------
vec3 EvalTransmittance(Ctx& c, Ray r)   // :119-155 (OPT_MEDIUM && OPT_VOL_MIS)
{
    LightSampleRec lightSample{};
    State state{};
    vec3 transmittance = V3(1.0f);
    for (int depth = 0; depth < c.o->maxDepth; depth++)
    {
        bool hit = ClosestHit(c, r, state, lightSample);
        if (!hit || state.isEmitter) break;
        GetMaterial(c, state, r);
        bool alphatest = (state.mat.alphaMode == ALPHA_MODE_MASK && state.mat.opacity < state.mat.alphaCutoff) ||
                         (state.mat.alphaMode == ALPHA_MODE_BLEND && c.rand() > state.mat.opacity);
        bool refractive = (1.0f - state.mat.metallic) * state.mat.specTrans > 0.0f;
        if (hit && !(alphatest || refractive)) return V3(0.0f);
        if (dot(r.direction, state.normal) > 0 && state.mat.medium.type != MEDIUM_NONE)
        {
            vec3 color = state.mat.medium.type == MEDIUM_ABSORB ? V3(1.0f) - state.mat.medium.color : V3(1.0f);
            transmittance *= vexp(-color * state.mat.medium.density * state.hitDist);
        }
        r.origin = state.fhp + r.direction * EPS;
    }
    return transmittance;
}

vec3 DirectLight(Ctx& c, const Ray& r, const State& state, bool isSurface)   // :158-283
{
    vec3 Ld = V3(0.0f);
    vec3 Li = V3(0.0f);
    vec3 scatterPos = state.fhp + state.normal * EPS;
    ScatterSampleRec scatterSample{};
    const bool volMis = c.o->optMedium && c.o->optVolMis;

    if (c.o->optEnvMap && !c.o->optUniformLight)   // :166-167
    {
        vec4 dirPdf = SampleEnvMap(c, Li);
        vec3 lightDir = {dirPdf.x, dirPdf.y, dirPdf.z};
        float lightPdf = dirPdf.w;
        Ray shadowRay = {scatterPos, lightDir};
        if (volMis)
        {
            Li *= EvalTransmittance(c, shadowRay);
            if (isSurface) scatterSample.f = DisneyEval(state, -r.direction, state.ffnormal, lightDir, scatterSample.pdf);
            else
            {
                float p = PhaseHG(dot(-r.direction, lightDir), state.medium.anisotropy);
                scatterSample.f = V3(p);
                scatterSample.pdf = p;
            }
            if (scatterSample.pdf > 0.0f)
            {
                float misWeight = PowerHeuristic(lightPdf, scatterSample.pdf);
                if (misWeight > 0.0f) Ld += misWeight * Li * scatterSample.f * c.o->envMapIntensity / lightPdf;
            }
        }
        else
        {
            bool inShadow = AnyHit(c, shadowRay, INF - EPS);
            if (!inShadow)
            {
                scatterSample.f = DisneyEval(state, -r.direction, state.ffnormal, lightDir, scatterSample.pdf);
                if (scatterSample.pdf > 0.0f)
                {
                    float misWeight = PowerHeuristic(lightPdf, scatterSample.pdf);
                    if (misWeight > 0.0f) Ld += misWeight * Li * scatterSample.f * c.o->envMapIntensity / lightPdf;
                }
            }
        }
    }

    if (c.o->optLights)   // :217
    {
        LightSampleRec lightSample{};
        int idx = (int)(c.rand() * (float)c.s->numLights);            // :223
        // Q6: rand()==1.0 indexes one past the end in GLSL (undefined texel); clamp (probability ~2^-24).
        if (idx >= c.s->numLights) idx = c.s->numLights - 1;
        Light light = fetchLight(c.s, idx);
        SampleOneLight(c, light, scatterPos, lightSample);
        Li = lightSample.emission;
        if (dot(lightSample.direction, lightSample.normal) < 0.0f)
        {
            Ray shadowRay = {scatterPos, lightSample.direction};
            if (volMis)
            {
                Li *= EvalTransmittance(c, shadowRay);
                if (isSurface) scatterSample.f = DisneyEval(state, -r.direction, state.ffnormal, lightSample.direction, scatterSample.pdf);
                else
                {
                    float p = PhaseHG(dot(-r.direction, lightSample.direction), state.medium.anisotropy);
                    scatterSample.f = V3(p);
                    scatterSample.pdf = p;
                }
                float misWeight = 1.0f;
                if (light.area > 0.0f) misWeight = PowerHeuristic(lightSample.pdf, scatterSample.pdf);
                if (scatterSample.pdf > 0.0f) Ld += misWeight * scatterSample.f * Li / lightSample.pdf;
            }
            else
            {
                if (c.capShadow && !c.capShadowDone && state.depth == c.capDepth)
                {
                    float* q = c.capShadow;
                    q[0] = shadowRay.origin.x; q[1] = shadowRay.origin.y; q[2] = shadowRay.origin.z;
                    q[3] = shadowRay.direction.x; q[4] = shadowRay.direction.y; q[5] = shadowRay.direction.z;
                    q[6] = lightSample.dist - EPS; q[7] = (float)idx; c.capShadowDone = true;
                }
                bool inShadow = AnyHit(c, shadowRay, lightSample.dist - EPS);
                if (!inShadow)
                {
                    scatterSample.f = DisneyEval(state, -r.direction, state.ffnormal, lightSample.direction, scatterSample.pdf);
                    float misWeight = 1.0f;
                    if (light.area > 0.0f) misWeight = PowerHeuristic(lightSample.pdf, scatterSample.pdf);
                    if (scatterSample.pdf > 0.0f) Ld += misWeight * Li * scatterSample.f / lightSample.pdf;
                }
            }
        }
    }
    return Ld;
}

vec4 PathTrace(Ctx& c, Ray r)   // :285-476
{
    const OrcOptions& o = *c.o;
    vec3 radiance = V3(0.0f);
    vec3 throughput = V3(1.0f);
    State state{};                     // H4/Q2: GLSL leaves these undefined; pinned to zero (matID 0 = loader's default material)
    LightSampleRec lightSample{};
    ScatterSampleRec scatterSample{};
    float alpha = 1.0f;
    bool inMedium = false, mediumSampled = false, surfaceScatter = false;

    for (state.depth = 0;; state.depth++)
    {
        if (c.capRay && !c.capDone && state.depth == c.capDepth)
        {
            c.capRay[0] = r.origin.x; c.capRay[1] = r.origin.y; c.capRay[2] = r.origin.z;
            c.capRay[3] = r.direction.x; c.capRay[4] = r.direction.y; c.capRay[5] = r.direction.z; c.capDone = true;
        }
        bool hit = ClosestHit(c, r, state, lightSample);
        if (!hit)
        {
            if (o.optBackground || o.optTransparentBackground)
                if (state.depth == 0) alpha = 0.0f;
            if (!o.optHideEmitters || state.depth > 0)
            {
                if (o.optUniformLight) radiance += V3(o.uniformLightCol[0], o.uniformLightCol[1], o.uniformLightCol[2]) * throughput;
                else if (o.optEnvMap)
                {
                    vec4 envMapColPdf = EvalEnvMap(c, r);
                    float misWeight = 1.0f;
                    if (state.depth > 0) misWeight = PowerHeuristic(scatterSample.pdf, envMapColPdf.w);
                    if (o.optMedium && !o.optVolMis)
                        if (!surfaceScatter) misWeight = 1.0f;
                    if (misWeight > 0) radiance += misWeight * V3(envMapColPdf.x, envMapColPdf.y, envMapColPdf.z) * throughput * o.envMapIntensity;
                }
            }
            break;
        }

        GetMaterial(c, state, r);
        radiance += state.mat.emission * throughput;

        if (o.optLights)
        {
            if (state.isEmitter)
            {
                float misWeight = 1.0f;
                if (state.depth > 0) misWeight = PowerHeuristic(scatterSample.pdf, lightSample.pdf);
                if (o.optMedium && !o.optVolMis)
                    if (!surfaceScatter) misWeight = 1.0f;
                radiance += misWeight * lightSample.emission * throughput;
                break;
            }
        }
        if (state.depth == o.maxDepth) break;

        if (o.optMedium)
        {
            mediumSampled = false;
            surfaceScatter = false;
            if (inMedium)
            {
                if (state.medium.type == MEDIUM_ABSORB)
                    throughput *= vexp(-(V3(1.0f) - state.medium.color) * state.hitDist * state.medium.density);
                else if (state.medium.type == MEDIUM_EMISSIVE)
                    radiance += state.medium.color * state.hitDist * state.medium.density * throughput;
                else
                {
                    float scatterDist = fminf(-logf(c.rand()) / state.medium.density, state.hitDist);
                    mediumSampled = scatterDist < state.hitDist;
                    if (mediumSampled)
                    {
                        throughput *= state.medium.color;
                        r.origin += r.direction * scatterDist;
                        state.fhp = r.origin;
                        radiance += DirectLight(c, r, state, false) * throughput;
                        float hr1 = c.rand(), hr2 = c.rand();    // :405 argument evaluation left to right
                        vec3 scatterDir = SampleHG(-r.direction, state.medium.anisotropy, hr1, hr2);
                        scatterSample.pdf = PhaseHG(dot(-r.direction, scatterDir), state.medium.anisotropy);
                        r.direction = scatterDir;
                    }
                }
            }
        }
        if (!o.optMedium || !mediumSampled)
        {
            bool skipped = false;
            if (o.optAlphaTest)
            {
                if ((state.mat.alphaMode == ALPHA_MODE_MASK && state.mat.opacity < state.mat.alphaCutoff) ||
                    (state.mat.alphaMode == ALPHA_MODE_BLEND && c.rand() > state.mat.opacity))
                {
                    scatterSample.L = r.direction;
                    state.depth--;
                    skipped = true;
                }
            }
            if (!skipped)
            {
                surfaceScatter = true;
                radiance += DirectLight(c, r, state, true) * throughput;
                scatterSample.f = DisneySample(c, state, -r.direction, state.ffnormal, scatterSample.L, scatterSample.pdf);
                if (scatterSample.pdf > 0.0f) throughput *= scatterSample.f / scatterSample.pdf;
                else break;
            }
            r.direction = scatterSample.L;
            r.origin = state.fhp + r.direction * EPS;

            if (o.optMedium)
            {
                if (dot(r.direction, state.normal) < 0 && state.mat.medium.type != MEDIUM_NONE)
                {
                    inMedium = true;
                    state.medium = state.mat.medium;
                }
                else if (state.mat.medium.type != MEDIUM_NONE)
                    inMedium = false;
            }
        }

        if (o.optRR)
        {
            if (state.depth >= o.rrDepth)
            {
                float q = fminf(fmaxf(throughput.x, fmaxf(throughput.y, throughput.z)) + 0.001f, 0.95f);
                if (c.rand() > q) break;
                throughput /= q;
            }
        }
    }
    return {radiance.x, radiance.y, radiance.z, alpha};
}

// ---------------------------------------------------------------- tile.glsl --------------------------------------
struct TileGrid { int numTilesX, numTilesY; float invNumTilesX, invNumTilesY; };
TileGrid tileGrid(const OrcOptions& o)   // Renderer.cpp:293-297
{
    TileGrid g;
    g.invNumTilesX = (float)o.tileW / o.renderW;
    g.invNumTilesY = (float)o.tileH / o.renderH;
    g.numTilesX = (int)ceilf((float)o.renderW / o.tileW);
    g.numTilesY = (int)ceilf((float)o.renderH / o.tileH);
    return g;
}

// tile.glsl:41-68 for tile-local pixel (lx,ly) of tile (tx,ty); leaves c.rng advanced by 4 draws.
Ray cameraRay(Ctx& c, const TileGrid& g, int tx, int ty, int lx, int ly, int frameNum)
{
    const OrcOptions& o = *c.o;
    vec2 TexCoords = {((float)lx + 0.5f) / (float)o.tileW, ((float)ly + 0.5f) / (float)o.tileH};
    vec2 tileOffset = {(float)tx * g.invNumTilesX, (float)ty * g.invNumTilesY};      // Renderer.cpp:780
    vec2 coordsTile = {mixf(tileOffset.x, tileOffset.x + g.invNumTilesX, TexCoords.x), mixf(tileOffset.y, tileOffset.y + g.invNumTilesY, TexCoords.y)};
    c.rng.init((float)lx + 0.5f, (float)ly + 0.5f, frameNum);                       // gl_FragCoord.xy is tile-local (A.5)
    float r1 = 2.0f * c.rand();
    float r2 = 2.0f * c.rand();
    vec2 jitter;
    jitter.x = r1 < 1.0f ? sqrtf(r1) - 1.0f : 1.0f - sqrtf(2.0f - r1);
    jitter.y = r2 < 1.0f ? sqrtf(r2) - 1.0f : 1.0f - sqrtf(2.0f - r2);
    jitter.x /= ((float)o.renderW * 0.5f); jitter.y /= ((float)o.renderH * 0.5f);
    vec2 d = {(coordsTile.x * 2.0f - 1.0f) + jitter.x, (coordsTile.y * 2.0f - 1.0f) + jitter.y};
    float scale = tanf(o.camFov * 0.5f);
    d.y *= (float)o.renderH / (float)o.renderW * scale;
    d.x *= scale;
    vec3 right = {o.camRight[0], o.camRight[1], o.camRight[2]}, up = {o.camUp[0], o.camUp[1], o.camUp[2]},
         fwd = {o.camForward[0], o.camForward[1], o.camForward[2]}, pos = {o.camPosition[0], o.camPosition[1], o.camPosition[2]};
    vec3 rayDir = normalize(d.x * right + d.y * up + fwd);
    vec3 focalPoint = o.camFocalDist * rayDir;
    float cam_r1 = c.rand() * TWO_PI;
    float cam_r2 = c.rand() * o.camAperture;
    vec3 randomAperturePos = (cosf(cam_r1) * right + sinf(cam_r1) * up) * sqrtf(cam_r2);
    vec3 finalRayDir = normalize(focalPoint - randomAperturePos);
    return {pos + randomAperturePos, finalRayDir};
}

// frameNum of (1-based) sample pass s, tile (tx,ty): Renderer.cpp:733-762 — first Update is the dirty one (frame 1),
// tiles advance x fastest from the TOP row down.
inline int frameNumOf(const TileGrid& g, int s, int tx, int ty)
{
    int T = g.numTilesX * g.numTilesY;
    int j = (g.numTilesY - 1 - ty) * g.numTilesX + tx;
    return 2 + (s - 1) * T + j;
}

void mergeCounters(Counters& dst, const Counters& a)
{
    dst.closestRays += a.closestRays; dst.anyRays += a.anyRays; dst.nodeVisits += a.nodeVisits; dst.internalSteps += a.internalSteps;
    dst.triTests += a.triTests; dst.tlasLeaves += a.tlasLeaves; dst.surfaceHits += a.surfaceHits;
    dst.anyNodeVisits += a.anyNodeVisits; dst.anyInternalSteps += a.anyInternalSteps; dst.anyTriTests += a.anyTriTests; dst.anyTlasLeaves += a.anyTlasLeaves;
}

void renderRect(OrcCtx* h, int firstSample, int nSamples, int x0, int y0, int x1, int y1, float* accum, int fixedFrame, int onlyTx, int onlyTy)
{
    const OrcOptions& o = h->o;
    TileGrid g = tileGrid(o);
    const int W = o.renderW;
    for (int s = firstSample; s < firstSample + nSamples; s++)
    {
#pragma omp parallel
        {
            Counters local;
#pragma omp for schedule(dynamic, 1) collapse(1)
            for (int y = y0; y < y1; y++)
                for (int x = x0; x < x1; x++)
                {
                    int tx = x / o.tileW, ty = y / o.tileH, lx = x % o.tileW, ly = y % o.tileH;
                    if (onlyTx >= 0 && (tx != onlyTx || ty != onlyTy)) continue;
                    Ctx c{h, &h->o, Rng{}, &local};
                    int frame = fixedFrame >= 0 ? fixedFrame : frameNumOf(g, s, tx, ty);
                    Ray ray = cameraRay(c, g, tx, ty, lx, ly, frame);
                    vec4 px = PathTrace(c, ray);
                    float* a = &accum[((size_t)y * W + x) * 4];
                    a[0] = px.x + a[0]; a[1] = px.y + a[1]; a[2] = px.z + a[2]; a[3] = px.w + a[3];   // tile.glsl:74
                }
#pragma omp critical
            mergeCounters(h->total, local);
        }
    }
}

} // namespace

// =================================================================== C API ========================================
extern "C" {

OrcCtx* orc_create(const OrcSceneDesc* d, const OrcOptions* opts)
{
    OrcCtx* h = new OrcCtx();
    h->nodes.assign(d->nodes, d->nodes + (size_t)d->numNodes * 9);
    h->topLevelIndex = d->topLevelIndex;
    h->vertIndices.assign(d->vertIndices, d->vertIndices + (size_t)d->numIndices * 3);
    h->verticesUVX.assign(d->verticesUVX, d->verticesUVX + (size_t)d->numVertices * 4);
    h->normalsUVY.assign(d->normalsUVY, d->normalsUVY + (size_t)d->numVertices * 4);
    h->materials.assign(d->materials, d->materials + (size_t)d->numMaterials * 32);
    h->transforms.assign(d->transforms, d->transforms + (size_t)d->numInstances * 16);
    if (d->numLights) h->lights.assign(d->lights, d->lights + (size_t)d->numLights * 15);
    h->numLights = d->numLights; h->numMaterials = d->numMaterials; h->numInstances = d->numInstances;
    h->numTextures = d->numTextures; h->texW = d->texW; h->texH = d->texH;
    if (d->numTextures) h->textures.assign(d->textures, d->textures + (size_t)d->numTextures * d->texW * d->texH * 4);
    h->envW = d->envW; h->envH = d->envH; h->envTotalSum = d->envTotalSum;
    if (d->envImg && d->envW > 0)
    {
        h->envImg.assign(d->envImg, d->envImg + (size_t)d->envW * d->envH * 3);
        h->envCdf.assign(d->envCdf, d->envCdf + (size_t)d->envW * d->envH);
    }
    h->o = *opts;
    return h;
}
void orc_destroy(OrcCtx* h) { delete h; }
void orc_set_options(OrcCtx* h, const OrcOptions* opts) { h->o = *opts; }
void orc_set_num_threads(int n)
{
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);      // overrides OMP_NUM_THREADS (torchrun exports 1) for every OpenMP region of this process
#else
    (void)n;
#endif
}
int orc_num_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static void fillHit(OrcHit& o, float t, const HitInfo& info)
{
    o.t = t; o.kind = info.kind; o.instance = info.instance; o.matID = info.matID; o.primSlot = info.primSlot; o.triIDx = info.triID[0];
    o.bary[0] = info.bary.x; o.bary[1] = info.bary.y; o.bary[2] = info.bary.z; o.lightIdx = info.lightIdx;
}

void orc_trace_closest(OrcCtx* h, const float* rays, int64_t n, int32_t depth, OrcHit* out)
{
#pragma omp parallel
    {
        Counters local;
#pragma omp for schedule(dynamic, 4096)
        for (int64_t i = 0; i < n; i++)
        {
            Ctx c{h, &h->o, Rng{}, &local};
            Ray r = {{rays[i * 6 + 0], rays[i * 6 + 1], rays[i * 6 + 2]}, {rays[i * 6 + 3], rays[i * 6 + 4], rays[i * 6 + 5]}};
            State st{}; st.depth = depth; LightSampleRec ls{}; HitInfo info;
            bool hit = ClosestHit(c, r, st, ls, &info);
            fillHit(out[i], hit ? st.hitDist : INF, info);
        }
#pragma omp critical
        mergeCounters(h->total, local);
    }
}

void orc_trace_any(OrcCtx* h, const float* rays, const float* maxDist, int64_t n, int32_t* out)
{
#pragma omp parallel
    {
        Counters local;
#pragma omp for schedule(dynamic, 4096)
        for (int64_t i = 0; i < n; i++)
        {
            Ctx c{h, &h->o, Rng{}, &local};
            Ray r = {{rays[i * 6 + 0], rays[i * 6 + 1], rays[i * 6 + 2]}, {rays[i * 6 + 3], rays[i * 6 + 4], rays[i * 6 + 5]}};
            out[i] = AnyHit(c, r, maxDist[i], false) ? 1 : 0;
        }
#pragma omp critical
        mergeCounters(h->total, local);
    }
}

// Brute force: same light loop and the same per-triangle arithmetic, but every instance's every leaf slot is tested in
// (TLAS-independent) order; ties may resolve differently, so tests compare t and accept differing IDs only when t is equal.
void orc_trace_closest_brute(OrcCtx* h, const float* rays, int64_t n, int32_t depth, OrcHit* out)
{
    // collect (instance, matID, blasRoot) from the TLAS leaves
    struct Inst { int inst, mat, root; };
    std::vector<Inst> insts;
    for (int i = h->topLevelIndex; i < (int)(h->nodes.size() / 9); i++)
    {
        int leaf = (int)h->nodes[(size_t)i * 9 + 8];
        if (leaf < 0) insts.push_back({-leaf - 1, (int)h->nodes[(size_t)i * 9 + 7], (int)h->nodes[(size_t)i * 9 + 6]});
    }
    // leaf slots per BLAS root: walk the subtree
    std::vector<std::vector<int>> slots(insts.size());
    for (size_t k = 0; k < insts.size(); k++)
    {
        std::vector<int> st{insts[k].root};
        while (!st.empty())
        {
            int i = st.back(); st.pop_back();
            int l = (int)h->nodes[(size_t)i * 9 + 6], r = (int)h->nodes[(size_t)i * 9 + 7], leaf = (int)h->nodes[(size_t)i * 9 + 8];
            if (leaf > 0) for (int j = 0; j < r; j++) slots[k].push_back(l + j);
            else { st.push_back(l); st.push_back(r); }
        }
    }
#pragma omp parallel for schedule(dynamic, 64)
    for (int64_t i = 0; i < n; i++)
    {
        const OrcOptions& o2 = h->o;
        Ray r = {{rays[i * 6 + 0], rays[i * 6 + 1], rays[i * 6 + 2]}, {rays[i * 6 + 3], rays[i * 6 + 4], rays[i * 6 + 5]}};
        // lights via a scene-less copy of ClosestHit's light loop: run ClosestHit on an empty traversal by temporarily using t only
        float t = INF; HitInfo info; int lightIdx = -1;
        if (o2.optLights && (!o2.optHideEmitters || depth > 0))
            for (int li = 0; li < h->numLights; li++)
            {
                Light L = fetchLight(h, li); vec3 u = L.u, v = L.v; float d = INF;
                if (L.type == (float)QUAD_LIGHT)
                {
                    vec3 normal = normalize(cross(u, v));
                    if (dot(normal, r.direction) > 0.f) continue;
                    vec4 plane = {normal.x, normal.y, normal.z, dot(normal, L.position)};
                    u *= 1.0f / dot(u, u); v *= 1.0f / dot(v, v);
                    d = RectIntersect(L.position, u, v, plane, r);
                }
                if (L.type == (float)SPHERE_LIGHT) d = SphereIntersect(L.radius, L.position, r);
                if (d < 0.f) d = INF;
                if (d < t) { t = d; lightIdx = li; }
            }
        int prim = -1, inst = -1, mat = -1, tri0 = -1; vec3 bary = V3(0.f);
        for (size_t k = 0; k < insts.size(); k++)
        {
            float invMat[16]; inverse4(&h->transforms[(size_t)insts[k].inst * 16], invMat);
            Ray rt = {xformPoint(invMat, r.origin, 1.0f), xformPoint(invMat, r.direction, 0.0f)};
            for (int slot : slots[k])
            {
                const int32_t* vi = &h->vertIndices[(size_t)slot * 3];
                const float* p0 = &h->verticesUVX[(size_t)vi[0] * 4]; const float* p1 = &h->verticesUVX[(size_t)vi[1] * 4];
                const float* p2 = &h->verticesUVX[(size_t)vi[2] * 4];
                vec3 v0 = {p0[0], p0[1], p0[2]}, v1 = {p1[0], p1[1], p1[2]}, v2 = {p2[0], p2[1], p2[2]};
                vec3 e0 = v1 - v0, e1 = v2 - v0;
                vec3 pv = cross(rt.direction, e1);
                float det = dot(e0, pv);
                vec3 tv = rt.origin - v0;
                vec3 qv = cross(tv, e0);
                float ux = dot(tv, pv) / det, uy = dot(rt.direction, qv) / det, uz = dot(e1, qv) / det, uw = 1.0f - ux - uy;
                if (ux >= 0.0f && uy >= 0.0f && uz >= 0.0f && uw >= 0.0f && uz < t)
                { t = uz; prim = slot; inst = insts[k].inst; mat = insts[k].mat; tri0 = vi[0]; bary = {uw, ux, uy}; }
            }
        }
        info.kind = (t == INF) ? 0 : (prim >= 0 ? 1 : 2);
        info.instance = inst; info.matID = mat; info.primSlot = prim; info.triID[0] = tri0; info.bary = bary;
        info.lightIdx = (prim < 0 && t != INF) ? lightIdx : -1;
        fillHit(out[i], t, info);
    }
}

static State stateFromQuery(OrcCtx* h, const OrcBsdfQuery& q)
{
    // Build state.mat exactly as GetMaterial does for an untextured material (pathtrace.glsl:31-67,109-114)
    const float* P = q.mat;
    State st{};
    Material& mat = st.mat;
    mat.baseColor = {P[0], P[1], P[2]}; mat.anisotropic = P[3]; mat.emission = {P[4], P[5], P[6]};
    mat.metallic = P[8]; mat.roughness = fmaxf(P[9], 0.001f); mat.subsurface = P[10]; mat.specularTint = P[11];
    mat.sheen = P[12]; mat.sheenTint = P[13]; mat.clearcoat = P[14]; mat.clearcoatRoughness = mixf(0.1f, 0.001f, P[15]);
    mat.specTrans = P[16]; mat.ior = P[17];
    float aspect = sqrtf(1.0f - mat.anisotropic * 0.9f);
    mat.ax = fmaxf(0.001f, mat.roughness / aspect);
    mat.ay = fmaxf(0.001f, mat.roughness * aspect);
    st.eta = q.eta;
    (void)h;
    return st;
}

void orc_bsdf_eval(OrcCtx* h, const OrcBsdfQuery* q, int64_t n, OrcBsdfResult* out)
{
    for (int64_t i = 0; i < n; i++)
    {
        State st = stateFromQuery(h, q[i]);
        float pdf; vec3 f = DisneyEval(st, {q[i].V[0], q[i].V[1], q[i].V[2]}, {q[i].N[0], q[i].N[1], q[i].N[2]}, {q[i].L[0], q[i].L[1], q[i].L[2]}, pdf);
        out[i].f[0] = f.x; out[i].f[1] = f.y; out[i].f[2] = f.z; out[i].pdf = pdf; out[i].L[0] = q[i].L[0]; out[i].L[1] = q[i].L[1]; out[i].L[2] = q[i].L[2];
    }
}
void orc_bsdf_sample(OrcCtx* h, const OrcBsdfQuery* q, int64_t n, OrcBsdfResult* out)
{
    for (int64_t i = 0; i < n; i++)
    {
        State st = stateFromQuery(h, q[i]);
        float pdf; vec3 L;
        vec3 f = DisneySampleR(st, {q[i].V[0], q[i].V[1], q[i].V[2]}, {q[i].N[0], q[i].N[1], q[i].N[2]}, L, pdf, q[i].r1, q[i].r2, q[i].r3);
        out[i].f[0] = f.x; out[i].f[1] = f.y; out[i].f[2] = f.z; out[i].pdf = pdf; out[i].L[0] = L.x; out[i].L[1] = L.y; out[i].L[2] = L.z;
    }
}

// lambert.glsl:25-46 — included by tile.glsl, never called by PathTrace
static vec3 LambertEval(const State& state, vec3 V, vec3 N, vec3 L, float& pdf)   // :41-46
{
    (void)V;
    pdf = dot(N, L) * (1.0f / PI);
    return (1.0f / PI) * state.mat.baseColor * dot(N, L);
}
static vec3 LambertSampleR(const State& state, vec3 V, vec3 N, vec3& L, float& pdf, float r1, float r2)   // :25-39 with the two rand() draws passed in
{
    (void)V;
    vec3 T, B;
    Onb(N, T, B);
    L = CosineSampleHemisphere(r1, r2);
    L = T * L.x + B * L.y + N * L.z;
    pdf = dot(N, L) * (1.0f / PI);
    return (1.0f / PI) * state.mat.baseColor * dot(N, L);
}
void orc_lambert(OrcCtx* h, const OrcBsdfQuery* q, int64_t n, int32_t sample, OrcBsdfResult* out)
{
    for (int64_t i = 0; i < n; i++)
    {
        State st = stateFromQuery(h, q[i]);
        vec3 V = {q[i].V[0], q[i].V[1], q[i].V[2]}, N = {q[i].N[0], q[i].N[1], q[i].N[2]}, L = {q[i].L[0], q[i].L[1], q[i].L[2]};
        float pdf;
        vec3 f = sample ? LambertSampleR(st, V, N, L, pdf, q[i].r1, q[i].r2) : LambertEval(st, V, N, L, pdf);
        out[i].f[0] = f.x; out[i].f[1] = f.y; out[i].f[2] = f.z; out[i].pdf = pdf; out[i].L[0] = L.x; out[i].L[1] = L.y; out[i].L[2] = L.z;
    }
}

// Analysis aid (scripts/simd_sim.py): the closest-hit ray every pixel's path traces at loop depth `depth` of sample pass `sample`
// (valid[i] = 0 where the path ended earlier).  rays: w*h*6 floats, row 0 = bottom.
void orc_capture_rays(OrcCtx* h, int32_t sample, int32_t depth, float* rays, uint8_t* valid)
{
    const OrcOptions& o = h->o;
    TileGrid g = tileGrid(o);
#pragma omp parallel for schedule(dynamic, 1)
    for (int y = 0; y < o.renderH; y++)
        for (int x = 0; x < o.renderW; x++)
        {
            int tx = x / o.tileW, ty = y / o.tileH, lx = x % o.tileW, ly = y % o.tileH;
            Ctx c{h, &h->o, Rng{}, nullptr};
            c.capRay = &rays[((size_t)y * o.renderW + x) * 6]; c.capDepth = depth;
            Ray ray = cameraRay(c, g, tx, ty, lx, ly, frameNumOf(g, sample, tx, ty));
            PathTrace(c, ray);
            valid[(size_t)y * o.renderW + x] = c.capDone ? 1 : 0;
        }
}

// Analysis aid: the light-NEE shadow ray (origin, direction, maxDist, light index: 8 floats) each pixel's path traces while shading at loop
// depth `depth`; valid[i] = 0 where none is traced.
void orc_capture_shadow_rays(OrcCtx* h, int32_t sample, int32_t depth, float* rays8, uint8_t* valid)
{
    const OrcOptions& o = h->o;
    TileGrid g = tileGrid(o);
#pragma omp parallel for schedule(dynamic, 1)
    for (int y = 0; y < o.renderH; y++)
        for (int x = 0; x < o.renderW; x++)
        {
            int tx = x / o.tileW, ty = y / o.tileH, lx = x % o.tileW, ly = y % o.tileH;
            Ctx c{h, &h->o, Rng{}, nullptr};
            c.capShadow = &rays8[((size_t)y * o.renderW + x) * 8]; c.capDepth = depth;
            Ray ray = cameraRay(c, g, tx, ty, lx, ly, frameNumOf(g, sample, tx, ty));
            PathTrace(c, ray);
            valid[(size_t)y * o.renderW + x] = c.capShadowDone ? 1 : 0;
        }
}

void orc_camera_rays(OrcCtx* h, int32_t sample, float* rays)
{
    const OrcOptions& o = h->o;
    TileGrid g = tileGrid(o);
#pragma omp parallel for schedule(static)
    for (int y = 0; y < o.renderH; y++)
        for (int x = 0; x < o.renderW; x++)
        {
            int tx = x / o.tileW, ty = y / o.tileH, lx = x % o.tileW, ly = y % o.tileH;
            Ctx c{h, &h->o, Rng{}, nullptr};
            Ray r = cameraRay(c, g, tx, ty, lx, ly, frameNumOf(g, sample, tx, ty));
            float* p = &rays[((size_t)y * o.renderW + x) * 6];
            p[0] = r.origin.x; p[1] = r.origin.y; p[2] = r.origin.z; p[3] = r.direction.x; p[4] = r.direction.y; p[5] = r.direction.z;
        }
}

// preview.glsl:41-71 (+ Renderer.cpp:555-560,798): low-resolution 1-spp image, InitRNG(gl_FragCoord.xy, 1), TexCoords over the whole
// target, `resolution` uniform = the full render size, maxDepth forced to 2 by the caller of the shader; no accumulation.
void orc_render_preview(OrcCtx* h, int32_t w, int32_t hgt, float* out)
{
    OrcOptions o2 = h->o;
    o2.maxDepth = 2;
#pragma omp parallel for schedule(dynamic, 1)
    for (int y = 0; y < hgt; y++)
        for (int x = 0; x < w; x++)
        {
            Ctx c{h, &o2, Rng{}, nullptr};
            const OrcOptions& o = o2;
            c.rng.init((float)x + 0.5f, (float)y + 0.5f, 1);
            float r1 = 2.0f * c.rand();
            float r2 = 2.0f * c.rand();
            vec2 jitter;
            jitter.x = r1 < 1.0f ? sqrtf(r1) - 1.0f : 1.0f - sqrtf(2.0f - r1);
            jitter.y = r2 < 1.0f ? sqrtf(r2) - 1.0f : 1.0f - sqrtf(2.0f - r2);
            jitter.x /= ((float)o.renderW * 0.5f); jitter.y /= ((float)o.renderH * 0.5f);
            vec2 tc = {((float)x + 0.5f) / (float)w, ((float)y + 0.5f) / (float)hgt};
            vec2 d = {(2.0f * tc.x - 1.0f) + jitter.x, (2.0f * tc.y - 1.0f) + jitter.y};
            float scale = tanf(o.camFov * 0.5f);
            d.y *= (float)o.renderH / (float)o.renderW * scale;
            d.x *= scale;
            vec3 right = {o.camRight[0], o.camRight[1], o.camRight[2]}, up = {o.camUp[0], o.camUp[1], o.camUp[2]},
                 fwd = {o.camForward[0], o.camForward[1], o.camForward[2]}, pos = {o.camPosition[0], o.camPosition[1], o.camPosition[2]};
            vec3 rayDir = normalize(d.x * right + d.y * up + fwd);
            vec3 focalPoint = o.camFocalDist * rayDir;
            float cam_r1 = c.rand() * TWO_PI;
            float cam_r2 = c.rand() * o.camAperture;
            vec3 randomAperturePos = (cosf(cam_r1) * right + sinf(cam_r1) * up) * sqrtf(cam_r2);
            vec3 finalRayDir = normalize(focalPoint - randomAperturePos);
            vec4 px = PathTrace(c, Ray{pos + randomAperturePos, finalRayDir});
            float* a = &out[((size_t)y * w + x) * 4];
            a[0] = px.x; a[1] = px.y; a[2] = px.z; a[3] = px.w;
        }
}

void orc_render_samples(OrcCtx* h, int32_t firstSample, int32_t nSamples, float* accum)
{
    renderRect(h, firstSample, nSamples, 0, 0, h->o.renderW, h->o.renderH, accum, -1, -1, -1);
}
void orc_render_samples_rect(OrcCtx* h, int32_t firstSample, int32_t nSamples, int32_t x0, int32_t y0, int32_t x1, int32_t y1, float* accum)
{
    renderRect(h, firstSample, nSamples, x0, y0, x1, y1, accum, -1, -1, -1);
}
void orc_render_tile(OrcCtx* h, int32_t tx, int32_t ty, int32_t frameNum, float* accum)
{
    const OrcOptions& o = h->o;
    int x0 = tx * o.tileW, y0 = ty * o.tileH;
    int x1 = std::min(x0 + o.tileW, o.renderW), y1 = std::min(y0 + o.tileH, o.renderH);   // overhang is clipped on copy (Q15)
    renderRect(h, 1, 1, x0, y0, x1, y1, accum, frameNum, tx, ty);
}

// tonemap.glsl:44-133
static vec3 mulMat3(const float m[9], vec3 c)
{   // GLSL `color * M` with M = mat3(col0, col1, col2): result[j] = dot(color, col_j)
    return {c.x * m[0] + c.y * m[1] + c.z * m[2], c.x * m[3] + c.y * m[4] + c.z * m[5], c.x * m[6] + c.y * m[7] + c.z * m[8]};
}
void orc_tonemap(const float* accum, int32_t w, int32_t h, float invSampleCounter, int32_t enableTonemap, int32_t enableAces,
                 int32_t simpleAcesFit, const float* backgroundCol, int32_t optBackground, int32_t optTransparentBackground, uint8_t* out)
{
    static const float ACESInputMat[9] = {0.59719f, 0.35458f, 0.04823f, 0.07600f, 0.90834f, 0.01566f, 0.02840f, 0.13383f, 0.83777f};
    static const float ACESOutputMat[9] = {1.60475f, -0.53108f, -0.07367f, -0.10208f, 1.10813f, -0.00605f, -0.00327f, -0.07276f, 1.07602f};
#pragma omp parallel for schedule(static)
    for (int y = 0; y < h; y++)
        for (int x = 0; x < w; x++)
        {
            const float* a = &accum[((size_t)y * w + x) * 4];
            vec3 color = {a[0] * invSampleCounter, a[1] * invSampleCounter, a[2] * invSampleCounter};
            float alpha = a[3] * invSampleCounter;
            if (enableTonemap)
            {
                if (enableAces)
                {
                    if (simpleAcesFit)
                    {
                        float A = 2.51f, B = 0.03f, Y = 2.43f, D = 0.59f, E = 0.14f;
                        vec3 num = color * (A * color + B), den = color * (Y * color + D) + E;
                        color = {clampf(num.x / den.x, 0.f, 1.f), clampf(num.y / den.y, 0.f, 1.f), clampf(num.z / den.z, 0.f, 1.f)};
                    }
                    else
                    {
                        color = mulMat3(ACESInputMat, color);
                        vec3 va = color * (color + 0.0245786f) + (-0.000090537f);
                        vec3 vb = color * (0.983729f * color + 0.4329510f) + 0.238081f;
                        color = va / vb;
                        color = mulMat3(ACESOutputMat, color);
                        color = {clampf(color.x, 0.f, 1.f), clampf(color.y, 0.f, 1.f), clampf(color.z, 0.f, 1.f)};
                    }
                }
                else
                    color = color * 1.0f / (1.0f + Luminance(color) / 1.5f);
            }
            color = vpow(color, 1.0f / 2.2f);
            float outAlpha = 1.0f;
            vec3 bgCol = {backgroundCol[0], backgroundCol[1], backgroundCol[2]};
            if (optTransparentBackground)
            {
                outAlpha = alpha;
                float checkerSize = 10.0f;
                float fx = (float)x + 0.5f, fy = (float)y + 0.5f;
                float m = fmodf(floorf(fx / checkerSize) + floorf(fy / checkerSize), 2.0f);
                float sgn = m > 0.f ? 1.f : (m < 0.f ? -1.f : 0.f);
                float res = fmaxf(sgn, 0.0f);
                bgCol = mix(V3(0.1f), V3(0.2f), res);
            }
            vec4 o4;
            if (optBackground || optTransparentBackground) { vec3 m = mix(bgCol, color, alpha); o4 = {m.x, m.y, m.z, outAlpha}; }
            else o4 = {color.x, color.y, color.z, 1.0f};
            // glGetTexImage(GL_RGBA, GL_UNSIGNED_BYTE) of an RGBA32F texture: clamp to [0,1], round(c*255) (Renderer.cpp:633); NaN -> 0
            float v[4] = {o4.x, o4.y, o4.z, o4.w};
            uint8_t* px = &out[((size_t)y * w + x) * 4];
            for (int k = 0; k < 4; k++)
            {
                float f = v[k]; if (!(f == f)) f = 0.f;
                f = clampf(f, 0.f, 1.f);
                px[k] = (uint8_t)floorf(f * 255.0f + 0.5f);
            }
        }
}

void orc_get_stats(OrcCtx* h, OrcStats* out)
{
    out->closestRays = h->total.closestRays; out->anyRays = h->total.anyRays; out->nodeVisits = h->total.nodeVisits;
    out->internalSteps = h->total.internalSteps; out->triTests = h->total.triTests; out->tlasLeaves = h->total.tlasLeaves;
    out->surfaceHits = h->total.surfaceHits;
    out->anyNodeVisits = h->total.anyNodeVisits; out->anyInternalSteps = h->total.anyInternalSteps; out->anyTriTests = h->total.anyTriTests;
    out->anyTlasLeaves = h->total.anyTlasLeaves;
}
void orc_reset_stats(OrcCtx* h) { h->total = Counters(); }

} // extern "C"
