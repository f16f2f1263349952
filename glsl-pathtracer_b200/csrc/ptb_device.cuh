// ptb_device.cuh — device-side building blocks of the sm_100a wavefront path tracer: exact-arithmetic two-level BVH
// traversal over the derived 64-byte node layout, analytic-light tests, Disney BSDF, light / env-map sampling.
//
// What each block computes is defined by the reference shaders (cited per function, paths relative to
// /root/reference/src/shaders/common); how it is computed (wavefront stages, packed nodes, shared-memory stacks,
// pre-inverted instance transforms) is specific to this implementation.
#pragma once
#include "ptb_internal.h"
#ifdef PTB_HOST_HARNESS
// tests/host_harness compiles this header for the host (g++ -ffp-contract=off) to check the traversal and the camera rays against the
// oracle without a GPU; ptb_host_shim.h maps the device intrinsics used here to their IEEE host equivalents.  Not part of the product.
#include "ptb_host_shim.h"
#else
#include <cuda_runtime.h>
#include <math_constants.h>
#endif

#define PTB_PI         3.14159265358979323f
#define PTB_INV_PI     0.31830988618379067f
#define PTB_TWO_PI     6.28318530717958648f
#define PTB_INV_TWO_PI 0.15915494309189533f
#define PTB_INV_4_PI   0.07957747154594766f
#define PTB_EPS        0.0003f
#define PTB_INF        1000000.0f

// Traversal variants kept switchable for same-box A/B measurements (results are bit-identical either way)
#ifndef PTB_TRI_EARLYOUT
#define PTB_TRI_EARLYOUT 1      // decide triangle misses from the numerators' signs before dividing
#endif
#ifndef PTB_FUSED_POP
#define PTB_FUSED_POP 0         // consume BLAS markers at the pop instead of in a round of their own (measured: -3.5 %, the restore inside the inner loop)
#endif
#ifndef PTB_FAST_INST
#define PTB_FAST_INST 1         // short instance entry for translation-only transforms
#endif
#ifndef PTB_CHEAP_MATH
#define PTB_CHEAP_MATH 1        // shading math (never hit-deciding): bare MUFU reciprocal / rsqrt / sqrt without the range fix-ups of div.approx & co
#endif

namespace ptb {

// ------------------------------------------------------------------ float3 helpers (contractable math) ----------
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 f3(float a) { return make_float3(a, a, a); }
__device__ __forceinline__ float3 f3(const float4& v) { return make_float3(v.x, v.y, v.z); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
// Shading-math division: div.approx (MUFU.RCP + FMUL, 2 ulp) instead of the ~8-instruction full-range sequence; operands here are
// BSDF terms far from the 2^126 range limit.  Hit-deciding code never uses these (it uses the x* exact intrinsics below).
#if PTB_CHEAP_MATH && !defined(PTB_HOST_HARNESS)
// One MUFU each (max error 1 ulp, denormals flushed): the operands here are BSDF / pdf terms far from the ends of the exponent range, so the scaling
// and special-case code div.approx.f32, rsqrtf() and sqrtf() carry around the same MUFU instruction (4-5 extra instructions each, ~100 divisions in
// DisneyEval/Sample) buys nothing.  G2 (1e-4 relative) is unaffected: measured in tests/test_gpu_render.py::test_bsdf_eval_matches_oracle.
__device__ __forceinline__ float rcpApprox(float b) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b)); return r; }
__device__ __forceinline__ float rsqrtApprox(float b) { float r; asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b)); return r; }
__device__ __forceinline__ float sqrtApprox(float b) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(b)); return r; }
__device__ __forceinline__ float fdiv(float a, float b) { return a * rcpApprox(b); }
#else
__device__ __forceinline__ float rsqrtApprox(float b) { return rsqrtf(b); }
__device__ __forceinline__ float sqrtApprox(float b) { return sqrtf(b); }
__device__ __forceinline__ float fdiv(float a, float b) { return __fdividef(a, b); }
#endif
__device__ __forceinline__ float3 operator/(float3 a, float3 b) { return f3(fdiv(a.x, b.x), fdiv(a.y, b.y), fdiv(a.z, b.z)); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 operator*(float s, float3 a) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 operator/(float3 a, float s) { const float r = fdiv(1.0f, s); return f3(a.x * r, a.y * r, a.z * r); }
__device__ __forceinline__ float3 operator+(float3 a, float s) { return f3(a.x + s, a.y + s, a.z + s); }
__device__ __forceinline__ float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float3& operator+=(float3& a, float3 b) { a = a + b; return a; }
__device__ __forceinline__ float3& operator*=(float3& a, float3 b) { a = a * b; return a; }
__device__ __forceinline__ float3& operator*=(float3& a, float s) { a = a * s; return a; }
__device__ __forceinline__ float3& operator/=(float3& a, float s) { a = a / s; return a; }
__device__ __forceinline__ float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float3 cross(float3 a, float3 b) { return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
__device__ __forceinline__ float length(float3 a) { return sqrtApprox(dot(a, a)); }
__device__ __forceinline__ float3 normalize(float3 a) { const float r = rsqrtApprox(dot(a, a)); return f3(a.x * r, a.y * r, a.z * r); }
__device__ __forceinline__ float mixf(float a, float b, float t) { return a * (1.0f - t) + b * t; }
__device__ __forceinline__ float3 mix(float3 a, float3 b, float t) { return a * (1.0f - t) + b * t; }
__device__ __forceinline__ float clampf(float x, float lo, float hi) { return fminf(fmaxf(x, lo), hi); }
__device__ __forceinline__ float3 vpow(float3 a, float e) { return f3(powf(a.x, e), powf(a.y, e), powf(a.z, e)); }
__device__ __forceinline__ float3 vexp(float3 a) { return f3(expf(a.x), expf(a.y), expf(a.z)); }
__device__ __forceinline__ float3 reflect(float3 I, float3 N) { return I - N * (2.0f * dot(N, I)); }
__device__ __forceinline__ float3 refract(float3 I, float3 N, float eta)
{
    float d = dot(N, I);
    float k = 1.0f - eta * eta * (1.0f - d * d);
    if (k < 0.0f) return f3(0.0f);
    return I * eta - N * (eta * d + sqrtApprox(k));
}
__device__ __forceinline__ float Luminance(float3 c) { return 0.212671f * c.x + 0.715160f * c.y + 0.072169f * c.z; }   // globals.glsl:173

// ------------------------------------------------------------------ exact (never contracted) arithmetic ---------
// The hit decisions (which primitive / instance / light, and t) must agree bit-for-bit with a host traversal compiled
// without FMA contraction, so the intersection math uses round-to-nearest intrinsics that ptxas never fuses.
__device__ __forceinline__ float xm(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xa(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float xs(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xd(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ float xdot(float3 a, float3 b) { return xa(xa(xm(a.x, b.x), xm(a.y, b.y)), xm(a.z, b.z)); }
__device__ __forceinline__ float3 xcross(float3 a, float3 b)
{
    return f3(xs(xm(a.y, b.z), xm(a.z, b.y)), xs(xm(a.z, b.x), xm(a.x, b.z)), xs(xm(a.x, b.y), xm(a.y, b.x)));
}
__device__ __forceinline__ float3 xsub(float3 a, float3 b) { return f3(xs(a.x, b.x), xs(a.y, b.y), xs(a.z, b.z)); }
__device__ __forceinline__ float3 xmad(float3 o, float3 d, float t) { return f3(xa(o.x, xm(d.x, t)), xa(o.y, xm(d.y, t)), xa(o.z, xm(d.z, t))); }

// ------------------------------------------------------------------ RNG: globals.glsl:144-166 -------------------
struct Rng
{
    uint4 s;
    __device__ __forceinline__ void init(uint32_t px, uint32_t py, uint32_t frame) { s = make_uint4(px, py, frame, px + py); }
    __device__ __forceinline__ float rand()
    {
        uint32_t x = s.x * 1664525u + 1013904223u, y = s.y * 1664525u + 1013904223u, z = s.z * 1664525u + 1013904223u,
                 w = s.w * 1664525u + 1013904223u;
        x += y * w; y += z * x; z += x * y; w += y * z;
        x ^= x >> 16; y ^= y >> 16; z ^= z >> 16; w ^= w >> 16;
        x += y * w; y += z * x; z += x * y; w += y * z;
        s = make_uint4(x, y, z, w);
        // float(seed.x) / float(0xffffffffu), in [0,1]: float(0xffffffff) rounds to 2^32 and dividing by a power of two is the
        // same IEEE result as multiplying by 2^-32
        return __fmul_rn(__uint2float_rn(x), 2.3283064365386963e-10f);
    }
};

// ------------------------------------------------------------------ traversal -----------------------------------
struct HitRec
{
    float t;        // PTB_INF on miss
    float bu, bv;   // uvt.x, uvt.y of the winning triangle
    int prim;       // leaf-ref slot or -1
    int inst;       // instance (triangle hit) / -1
    int light;      // analytic light index (light hit) / -1
};

// intersection.glsl:68-82 with the ray's reciprocal direction hoisted (1.0/dir is the same IEEE division the shader performs).
__device__ __forceinline__ float aabbHit(float minx, float miny, float minz, float maxx, float maxy, float maxz, float3 o, float3 inv, float& entry)
{
    float fx = xm(xs(maxx, o.x), inv.x), fy = xm(xs(maxy, o.y), inv.y), fz = xm(xs(maxz, o.z), inv.z);
    float nx = xm(xs(minx, o.x), inv.x), ny = xm(xs(miny, o.y), inv.y), nz = xm(xs(minz, o.z), inv.z);
    float t1 = fminf(fmaxf(fx, nx), fminf(fmaxf(fy, ny), fmaxf(fz, nz)));
    float t0 = fmaxf(fminf(fx, nx), fmaxf(fminf(fy, ny), fminf(fz, nz)));
    entry = t0;
    return (t1 >= t0) ? (t0 > 0.f ? t0 : t1) : -1.0f;
}

// Packed variant: two IEEE fp32 operations per instruction (sub.rn.f32x2 / mul.rn.f32x2, sm_100+).  Each half is rounded exactly as
// the scalar __fsub_rn / __fmul_rn, so hits stay bit-identical; the traversal kernels are issue-bound, the slab test is 24 of the
// ~80 instructions of an internal-node step and becomes 12.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 xs2(f32x2 a, f32x2 b) { f32x2 r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ f32x2 xm2(f32x2 a, f32x2 b) { f32x2 r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
struct RayPk { f32x2 oXY, oZZ, iXY, iZZ; };
__device__ __forceinline__ RayPk packRay(float3 o, float3 inv) { RayPk r; r.oXY = pk2(o.x, o.y); r.oZZ = pk2(o.z, o.z); r.iXY = pk2(inv.x, inv.y); r.iZZ = pk2(inv.z, inv.z); return r; }
// Both children of one inner node: minXY/maxXY = {min.x,min.y}/{max.x,max.y}, zz = {min.z,max.z}
__device__ __forceinline__ float aabbHitPk(f32x2 minXY, f32x2 maxXY, f32x2 zz, const RayPk& r, float& entry)
{
    float nx, ny, fx, fy, nz, fz;
    upk2(xm2(xs2(minXY, r.oXY), r.iXY), nx, ny);
    upk2(xm2(xs2(maxXY, r.oXY), r.iXY), fx, fy);
    upk2(xm2(xs2(zz, r.oZZ), r.iZZ), nz, fz);
    float t1 = fminf(fmaxf(fx, nx), fminf(fmaxf(fy, ny), fmaxf(fz, nz)));
    float t0 = fmaxf(fminf(fx, nx), fmaxf(fminf(fy, ny), fminf(fz, nz)));
    entry = t0;
    return (t1 >= t0) ? (t0 > 0.f ? t0 : t1) : -1.0f;
}

// intersection.glsl:47-66 (RectIntersect) split into its plane part and its rectangle part.  Lights that share a plane
// bit-for-bit (same normal, same plane.w — e.g. the 17 ceiling strips of hyperion_rect_lights) reuse dt, t and p: the same
// inputs give the same IEEE results, so the outcome per light is unchanged (closest_hit.glsl:49-53 hoisted to upload time).
struct PlaneHit { float dt, t; float3 p; bool valid; };
__device__ __forceinline__ void planeEval(float3 n, float planeW, float3 o, float3 d, PlaneHit& ph)
{
    ph.dt = xdot(d, n);
    ph.t = xd(xs(planeW, xdot(n, o)), ph.dt);
    ph.valid = ph.t > PTB_EPS;
    ph.p = xmad(o, d, ph.t);
}
__device__ __forceinline__ float rectInside(const PlaneHit& ph, float3 pos, float3 uS, float3 vS)
{
    if (ph.valid)
    {
        float3 vi = xsub(ph.p, pos);
        float a1 = xdot(uS, vi);
        if (a1 >= 0.0f && a1 <= 1.0f)
        {
            float a2 = xdot(vS, vi);
            if (a2 >= 0.0f && a2 <= 1.0f) return ph.t;
        }
    }
    return PTB_INF;
}
// intersection.glsl:25-45
__device__ __forceinline__ float sphereHit(float rad, float3 pos, float3 o, float3 d)
{
    float3 op = xsub(pos, o);
    float b = xdot(op, d);
    float det = xa(xs(xm(b, b), xdot(op, op)), xm(rad, rad));
    if (det < 0.0f) return PTB_INF;
    det = __fsqrt_rn(det);
    float t1 = xs(b, det);
    if (t1 > 0.001f) return t1;
    float t2 = xa(b, det);
    if (t2 > 0.001f) return t2;
    return PTB_INF;
}

// lightsPre rows (8 float4 per light, built at upload by buildLightsPre in ptb_api.cpp):
//   0: position, type | 1: emission, area | 2: u, radius | 3: v, samePlaneAsPrevious | 4: normal, plane.w | 5: u/dot(u,u) | 6: v/dot(v,v)
// lightGroups rows (3 float4 per group of CONSECUTIVE lights): {first, count, kind (0 = quads sharing one plane, 1 = single other light)},
//   boxMin, boxMax = padded bounds of the group's quads.  A plane hit outside the padded box cannot pass any member's inside test
//   (the pad is ~1e3 x larger than the rounding of a1/a2), so skipping the members is exact; index order is preserved.
__device__ __forceinline__ bool outsideBox(float3 p, float4 bmin, float4 bmax)
{
    return !(p.x >= bmin.x && p.x <= bmax.x && p.y >= bmin.y && p.y <= bmax.y && p.z >= bmin.z && p.z <= bmax.z);
}
// Member mask of the grid cell that holds plane hit p (p lies inside the group's padded box [bmin, bmax]); gk = group word (bit 1 gridded, bits 4-5 / 8-9 the two
// grid axes), gridOff = bits(offset into S.lightGrid), bmin.w / bmax.w = cells per unit length along the two axes.
__device__ __forceinline__ uint32_t lightGridMask(const DevScene& S, uint32_t gk, float gridOff, float4 bmin, float4 bmax, float3 p)
{
    const int A = (int)((gk >> 4) & 3u), B = (int)((gk >> 8) & 3u);
    const float pa = A == 0 ? p.x : (A == 1 ? p.y : p.z), pb = B == 0 ? p.x : (B == 1 ? p.y : p.z);
    const float la = A == 0 ? bmin.x : (A == 1 ? bmin.y : bmin.z), lb = B == 0 ? bmin.x : (B == 1 ? bmin.y : bmin.z);
    int ia = (int)((pa - la) * bmin.w), ib = (int)((pb - lb) * bmax.w);
    ia = ia < 0 ? 0 : (ia > PTB_LIGHT_GRID - 1 ? PTB_LIGHT_GRID - 1 : ia);
    ib = ib < 0 ? 0 : (ib > PTB_LIGHT_GRID - 1 ? PTB_LIGHT_GRID - 1 : ib);
    return __ldg(S.lightGrid + __float_as_uint(gridOff) + (uint32_t)(ib * PTB_LIGHT_GRID + ia));
}
// Light loop of ClosestHit (closest_hit.glsl:28-86): nearest light, first index wins ties (strict <).
__device__ __forceinline__ void closestLights(const DevScene& S, float3 o, float3 d, float& t, int& light)
{
    for (int g = 0; g < S.numLightGroups; g++)
    {
        const float4* gp = S.lightGroups + (size_t)g * 3;
        const float4 g0 = __ldg(gp);
        const int first = __float_as_int(g0.x), count = __float_as_int(g0.y);
        const float4* p = S.lightsPre + (size_t)first * 8;
        const uint32_t gk = __float_as_uint(g0.z);
        if ((gk & 1u) == 0u)
        {
            const float4 e = __ldg(p + 4);
            PlaneHit ph;
            planeEval(f3(e), e.w, o, d, ph);
            if (ph.dt > 0.f) continue;                       // hide backfacing quad light (closest_hit.glsl:50)
            if (!ph.valid) continue;                         // RectIntersect returns INF for every member
            const float4 bmin = __ldg(gp + 1), bmax = __ldg(gp + 2);
            if (count > 1 && outsideBox(ph.p, bmin, bmax)) continue;
            if (gk & 2u)
            {   // gridded group: only the quads whose padded box touches the cell of the plane hit, in index order
                uint32_t mask = lightGridMask(S, gk, g0.w, bmin, bmax, ph.p);
                while (mask)
                {
                    const int i = __ffs((int)mask) - 1; mask &= mask - 1u;
                    const float4* q = p + (size_t)i * 8;
                    float dist = rectInside(ph, f3(__ldg(q)), f3(__ldg(q + 5)), f3(__ldg(q + 6)));
                    if (dist < t) { t = dist; light = first + i; }
                }
                continue;
            }
            for (int i = 0; i < count; i++, p += 8)
            {
                float dist = rectInside(ph, f3(__ldg(p)), f3(__ldg(p + 5)), f3(__ldg(p + 6)));
                if (dist < t) { t = dist; light = first + i; }
            }
        }
        else
        {
            const float4 a = __ldg(p);
            if (a.w != 1.0f) continue;                       // distant lights are never intersected
            float dist = sphereHit(__ldg(p + 2).w, f3(a), o, d);
            if (dist < 0.f) dist = PTB_INF;
            if (dist < t) { t = dist; light = first; }
        }
    }
}
// Light loop of AnyHit (anyhit.glsl:28-63): two-sided quads.  (maxDist <= INF in every pipeline call; for larger values the
// reference's `INF < maxDist` quirk makes any quad/sphere light an occluder.)
__device__ __forceinline__ bool anyLights(const DevScene& S, float3 o, float3 d, float maxDist)
{
    if (maxDist > PTB_INF)
    {
        for (int i = 0; i < S.numLights; i++) { float ty = __ldg(S.lightsPre + (size_t)i * 8).w; if (ty == 0.0f || ty == 1.0f) return true; }
        return false;
    }
    for (int g = 0; g < S.numLightGroups; g++)
    {
        const float4* gp = S.lightGroups + (size_t)g * 3;
        const float4 g0 = __ldg(gp);
        const int first = __float_as_int(g0.x), count = __float_as_int(g0.y);
        const float4* p = S.lightsPre + (size_t)first * 8;
        const uint32_t gk = __float_as_uint(g0.z);
        if ((gk & 1u) == 0u)
        {
            const float4 e = __ldg(p + 4);
            PlaneHit ph;
            planeEval(f3(e), e.w, o, d, ph);
            if (!(ph.valid && ph.t < maxDist)) continue;     // d > 0 && d < maxDist can only hold for d == t
            const float4 bmin = __ldg(gp + 1), bmax = __ldg(gp + 2);
            if (count > 1 && outsideBox(ph.p, bmin, bmax)) continue;
            if (gk & 2u)
            {
                uint32_t mask = lightGridMask(S, gk, g0.w, bmin, bmax, ph.p);
                while (mask)
                {
                    const int i = __ffs((int)mask) - 1; mask &= mask - 1u;
                    const float4* q = p + (size_t)i * 8;
                    if (rectInside(ph, f3(__ldg(q)), f3(__ldg(q + 5)), f3(__ldg(q + 6))) < maxDist) return true;
                }
                continue;
            }
            for (int i = 0; i < count; i++, p += 8)
                if (rectInside(ph, f3(__ldg(p)), f3(__ldg(p + 5)), f3(__ldg(p + 6))) < maxDist) return true;
        }
        else
        {
            const float4 a = __ldg(p);
            if (a.w != 1.0f) continue;
            float dist = sphereHit(__ldg(p + 2).w, f3(a), o, d);
            if (dist > 0.0f && dist < maxDist) return true;
        }
    }
    return false;
}

// One triangle of a leaf (closest_hit.glsl:118-153 / anyhit.glsl:88-117): true when the reference's acceptance test passes for the current distance bound t;
// ux, uy, uz = uvt.xyz are valid then.  slot = leaf-ref slot; ro / rd = ray in the space of the leaf.
__device__ __forceinline__ bool triangleHit(const DevScene& S, uint32_t slot, float3 ro, float3 rd, float t, float& ux, float& uy, float& uz)
{
    const float4* tp = S.tris + (size_t)slot * 3;
    const float4 a = __ldg(tp), b = __ldg(tp + 1), c = __ldg(tp + 2);
    const float3 v0 = f3(a.x, a.y, a.z), e0 = f3(a.w, b.x, b.y), e1 = f3(b.z, b.w, c.x);
    // Moeller-Trumbore exactly as closest_hit.glsl:127-141 (three IEEE divisions by det, no det==0 guard).  The acceptance test is
    //   ux >= 0 && uy >= 0 && uz >= 0 && uw >= 0 && uz < t   with   u* = numerator / det,  uw = (1 - ux) - uy.
    // Most tests are misses, and most misses can be decided from the numerators without dividing, with the SAME outcome:
    //  * numerator and det of strictly opposite sign, |numerator| >= 1e-18 and |det| <= 1e18: the IEEE quotient is a negative number
    //    of magnitude >= 1e-36 (it cannot round to -0, which would pass `>= 0`), or -inf for det = +-0;
    //  * both numerators share det's sign and |a| + |b| > 1.001 |det|: ux + uy > 1.0007, so (1 - ux) - uy is below -7e-4 whatever the
    //    rounding (or -inf when a quotient overflows).
    // NaNs and everything near a boundary take the full path below, which is the reference's arithmetic unchanged.
    const float3 pv = xcross(rd, e1);
    const float det = xdot(e0, pv);
    const float3 tv = xsub(ro, v0);
    const float na = xdot(tv, pv);
    const int sd = __float_as_int(det);
    const bool detSmall = fabsf(det) <= 1e18f;
#if PTB_TRI_EARLYOUT
    if (((__float_as_int(na) ^ sd) < 0) && fabsf(na) >= 1e-18f && detSmall) return false;
#endif
    const float3 qv = xcross(tv, e0);
    const float nb = xdot(rd, qv);
#if PTB_TRI_EARLYOUT
    if (((__float_as_int(nb) ^ sd) < 0) && fabsf(nb) >= 1e-18f && detSmall) return false;
    if (((__float_as_int(na) ^ sd) >= 0) && ((__float_as_int(nb) ^ sd) >= 0) && xa(fabsf(na), fabsf(nb)) > xm(1.001f, fabsf(det))) return false;
#endif
    const float nc = xdot(e1, qv);
#if PTB_TRI_EARLYOUT
    if (((__float_as_int(nc) ^ sd) < 0) && fabsf(nc) >= 1e-18f && detSmall) return false;
#endif
    ux = xd(na, det);
    uy = xd(nb, det);
    uz = xd(nc, det);
    const float uw = xs(xs(1.0f, ux), uy);
    return ux >= 0.0f && uy >= 0.0f && uz >= 0.0f && uw >= 0.0f && uz < t;
}

// Stack policies: shared memory (one column per thread, conflict-free) for the wavefront trace kernels,
// thread-local array for the rarely used inline traversal inside the shade kernel.
struct SmemStack
{
    uint32_t* base; uint32_t* top; int stride;
    __device__ __forceinline__ SmemStack(uint32_t* b, int strideWords) : base(b), top(b), stride(strideWords) {}
    __device__ __forceinline__ void reset() { top = base; }
    __device__ __forceinline__ void push(uint32_t v) { *top = v; top += stride; }
    __device__ __forceinline__ uint32_t pop() { top -= stride; return *top; }
};
struct LocalStack
{
    uint32_t a[64]; int sp = 0;
    __device__ __forceinline__ void reset() { sp = 0; }
    __device__ __forceinline__ void push(uint32_t v) { a[sp++] = v; }
    __device__ __forceinline__ uint32_t pop() { return a[--sp]; }
};

// Alpha test hook of AnyHit (anyhit.glsl:118-141), only in inline mode / MASK materials.
struct NoAlpha { };

// Two-level traversal (closest_hit.glsl:88-218 / anyhit.glsl:65-213) over the derived layout:
//   inner[i]   = { lmin.xyz lmax.x | lmax.yz rmin.xy | rmin.z rmax.xyz | lmeta rmeta - - }   (one 64-byte fetch per step;
//                the reference reads LRLeaf of the node and then both children's boxes from two other nodes)
//   tris[slot] = { v0.xyz e0.x | e0.yz e1.xy | e1.z vertIndex.x - - }
//   instTrav[k]= rows of inverse(transform) (xyz) + .w = {-, matID, -, rootMeta}; translation-only instances are flagged in the TLAS-leaf meta
// Visiting order, near/far rule (left first on ties), strict-< acceptance and the un-normalised transformed direction are the
// reference's, so primitive/instance IDs and t are identical to a host traversal of the canonical array.
// Control flow is "while-while": every lane keeps descending internal nodes until it holds a leaf / instance / sentinel, then
// the warp processes those together — incoherent warps spend fewer issue slots with most lanes masked off.
// The traversal is a resumable per-lane state machine (Trav) so the wavefront kernels can hand a finished lane a new ray
// while its neighbours are still traversing.
// ANY: finish at the first accepted hit with t < tmax.  alphaFn(slot, inst, u, v) -> accept? (only when ALPHA)
template <bool ANY, bool ALPHA, bool CULL>
struct Trav
{
#if PTB_PACKED_SLAB
    float3 o, d, rd, invW;
    RayPk rp;           // current-space origin and reciprocal direction in the register pairs the packed slab test consumes
    __device__ __forceinline__ float3 roNow() const { float3 r; float z2; upk2(rp.oXY, r.x, r.y); upk2(rp.oZZ, r.z, z2); return r; }
#else
    float3 o, d, ro, rd, inv, invW;
    __device__ __forceinline__ float3 roNow() const { return ro; }
#endif
    float t;
    uint32_t cur;
    int curInst;
    bool inBlas;
    bool generic;       // no zero / non-finite component in origin or direction: translation-only instances may take the short entry
    bool occluded;      // ANY result
    HitRec h;           // closest result (t filled by finish())

    template <class Stack>
    __device__ __forceinline__ void begin(const DevScene& S, float3 o_, float3 d_, float tmax, Stack& stk)
    {
#if PTB_PACKED_SLAB
        o = o_; d = d_; rd = d_;
        invW = f3(xd(1.0f, d_.x), xd(1.0f, d_.y), xd(1.0f, d_.z));
        rp = packRay(o_, invW);
#else
        o = o_; d = d_; ro = o_; rd = d_;
        inv = f3(xd(1.0f, d_.x), xd(1.0f, d_.y), xd(1.0f, d_.z)); invW = inv;
#endif
        t = tmax; cur = S.rootMeta; curInst = -1; inBlas = false; occluded = false;
        // |x| in (0, inf) for all six components (false for NaN too)
        generic = fabsf(o_.x) > 0.f && fabsf(o_.y) > 0.f && fabsf(o_.z) > 0.f && fabsf(d_.x) > 0.f && fabsf(d_.y) > 0.f && fabsf(d_.z) > 0.f &&
                  fabsf(o_.x) < CUDART_INF_F && fabsf(o_.y) < CUDART_INF_F && fabsf(o_.z) < CUDART_INF_F &&
                  fabsf(d_.x) < CUDART_INF_F && fabsf(d_.y) < CUDART_INF_F && fabsf(d_.z) < CUDART_INF_F;
        stk.reset();
        stk.push(PTB_META_NONE);
    }

    // back in the TLAS: the world-space ray again (closest_hit.glsl:206-216); 1/direction of the world ray was computed once in begin()
    __device__ __forceinline__ void leaveBlas()
    {
        inBlas = false;
#if PTB_PACKED_SLAB
        rd = d; rp = packRay(o, invW);
#else
        ro = o; rd = d; inv = invW;
#endif
    }
    // Pop the next item; a BLAS marker is consumed on the spot (restore the world ray, pop again) instead of costing the warp another round.
    template <class Stack>
    __device__ __forceinline__ uint32_t popNext(Stack& stk)
    {
        uint32_t v = stk.pop();
#if PTB_FUSED_POP
        if (v == PTB_META_NONE && inBlas) { leaveBlas(); v = stk.pop(); }
#endif
        return v;
    }

    // One round: descend to the next non-internal item and process it.  Returns true when the traversal is finished.
    template <class Stack, class AlphaFn>
    __device__ __forceinline__ bool round(const DevScene& S, Stack& stk, AlphaFn alphaFn)
    {
        const float4* __restrict__ innerBase = S.inner;
        while (cur < (1u << 30))                              // PTB_K_INNER: closest_hit.glsl:173-205
        {
            const float4* n = innerBase + (size_t)cur * 4;
            float e0, e1;
#if PTB_PACKED_SLAB
            const ulonglong2 p0 = __ldg(reinterpret_cast<const ulonglong2*>(n)), p1 = __ldg(reinterpret_cast<const ulonglong2*>(n + 1)),
                             p2 = __ldg(reinterpret_cast<const ulonglong2*>(n + 2));
            const float4 q3 = __ldg(n + 3);
            float lh = aabbHitPk(p0.x, p0.y, p1.x, rp, e0);
            float rh = aabbHitPk(p2.x, p2.y, p1.y, rp, e1);
#else
            const float4 q0 = __ldg(n), q1 = __ldg(n + 1), q2 = __ldg(n + 2), q3 = __ldg(n + 3);
            float lh = aabbHit(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, ro, inv, e0);
            float rh = aabbHit(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, ro, inv, e1);
#endif
            if (CULL) { const float tc = t * 1.00001f; if (e0 > tc) lh = -1.0f; if (e1 > tc) rh = -1.0f; }
            const uint32_t lm = __float_as_uint(q3.x), rm = __float_as_uint(q3.y);
            const bool hl = lh > 0.0f, hr = rh > 0.0f;
            if (hl && hr)
            {
                const bool rightFirst = lh > rh;              // near child first; left on ties (closest_hit.glsl:181-190)
                stk.push(rightFirst ? lm : rm);
                cur = rightFirst ? rm : lm;
            }
            else if (hl) cur = lm;
            else if (hr) cur = rm;
            else cur = popNext(stk);
        }
        const uint32_t kind = cur >> 30;
        if (kind == PTB_K_LEAF)                               // closest_hit.glsl:118-153
        {
            const uint32_t first = cur & PTB_MAX_LEAF_SLOT, cnt = (cur >> 26) & 15u;
            for (uint32_t i = 0; i < cnt; i++)
            {
                float ux, uy, uz;
                if (!triangleHit(S, first + i, roNow(), rd, t, ux, uy, uz)) continue;
                if constexpr (ANY)
                {
                    if constexpr (ALPHA) { if (alphaFn((int)(first + i), curInst, ux, uy)) { occluded = true; return true; } }
                    else { occluded = true; return true; }
                }
                else { t = uz; h.prim = (int)(first + i); h.inst = curInst; h.bu = ux; h.bv = uy; h.light = -1; }
            }
            cur = popNext(stk);
        }
        else if (kind == PTB_K_INST)                          // closest_hit.glsl:154-172
        {
            curInst = (int)(cur & PTB_INST_INDEX_MASK);
            const float4* ip = S.instTrav + (size_t)curInst * 4;
            const float4 r3 = __ldg(ip + 3);
            if (PTB_FAST_INST && (cur & PTB_INST_TRANSLATION_ONLY) && generic)
            {   // inverse(transform) is [I | -translation] bit for bit (checked at upload): the general formula below then reduces, for a ray without
                // zero components, to origin + r3 (x*1 = x, x + (+-0) = x for x != 0) and an unchanged direction, whose reciprocal is invW
#if PTB_PACKED_SLAB
                rp = packRay(f3(xa(o.x, r3.x), xa(o.y, r3.y), xa(o.z, r3.z)), invW);
#else
                ro = f3(xa(o.x, r3.x), xa(o.y, r3.y), xa(o.z, r3.z));
#endif
            }
            else
            {
                const float4 r0 = __ldg(ip), r1 = __ldg(ip + 1), r2 = __ldg(ip + 2);
                // rTrans = inverse(transform) * (origin,1) / (direction,0); the inverse is precomputed at upload
                const float3 roI = f3(xa(xa(xa(xm(o.x, r0.x), xm(o.y, r1.x)), xm(o.z, r2.x)), xm(1.0f, r3.x)),
                        xa(xa(xa(xm(o.x, r0.y), xm(o.y, r1.y)), xm(o.z, r2.y)), xm(1.0f, r3.y)),
                        xa(xa(xa(xm(o.x, r0.z), xm(o.y, r1.z)), xm(o.z, r2.z)), xm(1.0f, r3.z)));
                rd = f3(xa(xa(xa(xm(d.x, r0.x), xm(d.y, r1.x)), xm(d.z, r2.x)), xm(0.0f, r3.x)),
                        xa(xa(xa(xm(d.x, r0.y), xm(d.y, r1.y)), xm(d.z, r2.y)), xm(0.0f, r3.y)),
                        xa(xa(xa(xm(d.x, r0.z), xm(d.y, r1.z)), xm(d.z, r2.z)), xm(0.0f, r3.z)));
#if PTB_PACKED_SLAB
                rp = packRay(roI, f3(xd(1.0f, rd.x), xd(1.0f, rd.y), xd(1.0f, rd.z)));
#else
                ro = roI;
                inv = f3(xd(1.0f, rd.x), xd(1.0f, rd.y), xd(1.0f, rd.z));
#endif
            }
            stk.push(PTB_META_NONE);                          // marker: back to the TLAS when it is popped
            cur = __float_as_uint(r3.w);                      // BLAS root meta
            inBlas = true;
        }
        else                                                  // sentinel / marker (closest_hit.glsl:206-216)
        {
            if (!inBlas) { h.t = t; return true; }
            leaveBlas();                                      // (markers are normally consumed by popNext; this handles a marker reached directly)
            cur = stk.pop();
        }
        return false;
    }
};

template <bool ANY, bool ALPHA, bool CULL, class Stack, class AlphaFn>
__device__ __forceinline__ bool traverse(const DevScene& S, float3 o, float3 d, float tmax, Stack& stk, HitRec& h, AlphaFn alphaFn)
{
    Trav<ANY, ALPHA, CULL> tr;
    tr.h = h;
    tr.begin(S, o, d, tmax, stk);
    while (!tr.round(S, stk, alphaFn)) { }
    if (!ANY) h = tr.h;
    return tr.occluded;
}

// ------------------------------------------------------------------ 4-wide any-hit traversal ------------------------
// AnyHit (anyhit.glsl:65-213) over the collapsed 4-wide hierarchy of ptbd_build_wide.  The result is the SAME boolean as the binary traversal's:
// the distance bound never shrinks in any-hit, so the set of leaves reached does not depend on the visiting order, and a leaf is reached exactly when
// its own box passes (every dropped box contains the boxes that replace it and the slab test is monotone in the bounds — for rays whose reciprocal
// direction is finite, so that no 0 * inf NaN exists; other rays, and rays whose direction degenerates inside an instance, use the binary traversal).
// Boxes and arithmetic are the reference's: each child box is tested with the IEEE operations of AABBIntersect (intersection.glsl:68-82).
__device__ __forceinline__ bool wideRayOk(float3 d)
{   // 1/d finite and d finite for all three components (false for NaN)
    return fabsf(d.x) >= 1e-30f && fabsf(d.y) >= 1e-30f && fabsf(d.z) >= 1e-30f && fabsf(d.x) <= 1e30f && fabsf(d.y) <= 1e30f && fabsf(d.z) <= 1e30f;
}

// 0 = not occluded, 1 = occluded, 2 = undecided: trace this ray with the binary traversal
template <bool ALPHA, bool CULL, class Stack, class AlphaFn>
__device__ __forceinline__ int traverseWideAny(const DevScene& S, float3 o, float3 d, float tmax, Stack& stk, AlphaFn alphaFn)
{
    const float3 invW = f3(xd(1.0f, d.x), xd(1.0f, d.y), xd(1.0f, d.z));
    float3 ro = o, rd = d, inv = invW;
    const float tc = tmax * 1.00001f;
    uint32_t cur = S.rootMetaWide;
    int curInst = -1;
    bool inBlas = false;
    stk.reset();
    stk.push(PTB_META_NONE);
    const float4* __restrict__ wideBase = S.wide;
    while (true)
    {
        while (cur < (1u << 30))
        {
            const float4* n = wideBase + (size_t)cur * 8;
            const float4 q0 = __ldg(n), q1 = __ldg(n + 1), q2 = __ldg(n + 2), q3 = __ldg(n + 3), q4 = __ldg(n + 4), q5 = __ldg(n + 5);
            const uint4 m = __ldg(reinterpret_cast<const uint4*>(n + 6));
            float e0, e1, e2, e3;
            const float h0 = aabbHit(q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, ro, inv, e0);
            const float h1 = aabbHit(q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, ro, inv, e1);
            const float h2 = aabbHit(q3.x, q3.y, q3.z, q3.w, q4.x, q4.y, ro, inv, e2);
            const float h3 = aabbHit(q4.z, q4.w, q5.x, q5.y, q5.z, q5.w, ro, inv, e3);
            // the reference's `hit > 0` (and, in the culled variant, entry <= tmax * 1.00001); empty slots hold NaN boxes: every comparison fails
            const bool p0 = h0 > 0.0f && !(CULL && e0 > tc), p1 = h1 > 0.0f && !(CULL && e1 > tc);
            const bool p2 = h2 > 0.0f && !(CULL && e2 > tc), p3 = h3 > 0.0f && !(CULL && e3 > tc);
            uint32_t c = PTB_META_NONE; bool have = false;
            if (p3) { c = m.w; have = true; }
            if (p2) { if (have) stk.push(c); c = m.z; have = true; }
            if (p1) { if (have) stk.push(c); c = m.y; have = true; }
            if (p0) { if (have) stk.push(c); c = m.x; have = true; }
            cur = have ? c : stk.pop();
        }
        const uint32_t kind = cur >> 30;
        if (kind == PTB_K_LEAF)
        {
            const uint32_t first = cur & PTB_MAX_LEAF_SLOT, cnt = (cur >> 26) & 15u;
            for (uint32_t i = 0; i < cnt; i++)
            {
                float ux, uy, uz;
                if (!triangleHit(S, first + i, ro, rd, tmax, ux, uy, uz)) continue;
                if constexpr (ALPHA) { if (alphaFn((int)(first + i), curInst, ux, uy)) return 1; }
                else return 1;
            }
            cur = stk.pop();
        }
        else if (kind == PTB_K_INST)
        {
            curInst = (int)(cur & PTB_INST_INDEX_MASK);
            const float4* ip = S.instTrav + (size_t)curInst * 4;
            const float4 r2 = __ldg(ip + 2), r3 = __ldg(ip + 3);
            const bool generic = fabsf(o.x) > 0.f && fabsf(o.y) > 0.f && fabsf(o.z) > 0.f;      // (d is nonzero and finite here, o finite unless NaN/inf: then the products below are not exact shortcuts)
            if (PTB_FAST_INST && (cur & PTB_INST_TRANSLATION_ONLY) && generic && fabsf(o.x) < CUDART_INF_F && fabsf(o.y) < CUDART_INF_F && fabsf(o.z) < CUDART_INF_F)
                ro = f3(xa(o.x, r3.x), xa(o.y, r3.y), xa(o.z, r3.z));         // see Trav::round: exact for rays without zero components
            else
            {
                const float4 r0 = __ldg(ip), r1 = __ldg(ip + 1);
                ro = f3(xa(xa(xa(xm(o.x, r0.x), xm(o.y, r1.x)), xm(o.z, r2.x)), xm(1.0f, r3.x)),
                        xa(xa(xa(xm(o.x, r0.y), xm(o.y, r1.y)), xm(o.z, r2.y)), xm(1.0f, r3.y)),
                        xa(xa(xa(xm(o.x, r0.z), xm(o.y, r1.z)), xm(o.z, r2.z)), xm(1.0f, r3.z)));
                rd = f3(xa(xa(xa(xm(d.x, r0.x), xm(d.y, r1.x)), xm(d.z, r2.x)), xm(0.0f, r3.x)),
                        xa(xa(xa(xm(d.x, r0.y), xm(d.y, r1.y)), xm(d.z, r2.y)), xm(0.0f, r3.y)),
                        xa(xa(xa(xm(d.x, r0.z), xm(d.y, r1.z)), xm(d.z, r2.z)), xm(0.0f, r3.z)));
                if (!wideRayOk(rd)) return 2;                                  // degenerate direction in object space: the equivalence argument needs finite 1/d
                inv = f3(xd(1.0f, rd.x), xd(1.0f, rd.y), xd(1.0f, rd.z));
            }
            stk.push(PTB_META_NONE);                                           // marker: back to the TLAS when it is popped
            cur = __float_as_uint(r2.w);                                       // BLAS root in the wide hierarchy
            inBlas = true;
        }
        else
        {
            if (!inBlas) return 0;
            inBlas = false; ro = o; rd = d; inv = invW;
            cur = stk.pop();
        }
    }
}

// ------------------------------------------------------------------ slot <-> pixel mapping ----------------------
// A wave holds nSamples sample passes of a pixel rectangle padded to 8x4 pixel blocks; 32 consecutive slots ("group") are the primary rays of one warp.
//   sample-major (blockMajor = 0): group = one 8x4 block of one pass; all blocks of pass 0, then pass 1, ...
//   block-major  (blockMajor = 1): the 32 * nSamples slots of an 8x4 block are consecutive (a 2048-slot tile then holds the paths of a few neighbouring pixels,
//     which is what lets the tile-local grouping by direction / light build warps that share origin AND direction).  Inside the block a group is
//     2^lps consecutive passes of a (2^lpw x 2^lph)-pixel sub-block (lpw + lph + lps = 5): lps = 0 -> the 8x4 block of one pass, lps = 5 -> 32 passes of ONE pixel —
//     rays through one pixel differ by the sub-pixel jitter only, so the warp stays converged in the traversal and hits one material in the first shade pass.
// group -> (sample pass, pixel) of lane `lane`
__device__ __forceinline__ void groupToPixel(const WaveParams& W, uint32_t g, uint32_t lane, uint32_t blocksX, uint32_t blocksPerSample, int& s, int& px, int& py)
{
    uint32_t b, pxin = lane & 7u, pyin = lane >> 3;
    if (W.blockMajor)
    {
        const uint32_t nS = (uint32_t)W.nSamples, lps = (uint32_t)W.lps, lpw = (uint32_t)W.lpw, lpp = 5u - lps /* log2(pixels per sub-block) */;
        b = g / nS;
        const uint32_t gi = g - b * nS;                  // group inside the block: sub-block major, then sample group
        const uint32_t nsg = nS >> lps;                  // sample groups per sub-block
        const uint32_t pb = gi / nsg, sg = gi - pb * nsg;
        const uint32_t ls = lane >> lpp, pi = lane & ((1u << lpp) - 1u);
        s = (int)((sg << lps) + ls);
        const uint32_t pbx = pb & ((8u >> lpw) - 1u), pby = pb >> (3u - lpw);
        pxin = (pbx << lpw) + (pi & ((1u << lpw) - 1u));
        pyin = (pby << (lpp - lpw)) + (pi >> lpw);
    }
    else { s = (int)(g / blocksPerSample); b = g - (uint32_t)s * blocksPerSample; }
    const uint32_t by = b / blocksX, bx = b - by * blocksX;
    px = (int)(bx * 8u + pxin);
    py = (int)(by * 4u + pyin);
}
__device__ __forceinline__ bool slotToPixel(const WaveParams& W, uint32_t slot, int& s, int& px, int& py)
{
    const uint32_t blocksX = (uint32_t)W.vw >> 3, blocksPerSample = blocksX * ((uint32_t)W.vh >> 2);
    groupToPixel(W, slot >> 5, slot & 31u, blocksX, blocksPerSample, s, px, py);
    return px < W.rw && py < W.rh;
}

// slot of sample pass k for the pixel with index idx (= 8x4 block * 32 + row-major pixel inside the block)
__device__ __forceinline__ uint32_t slotOfSample(const WaveParams& W, uint32_t idx, uint32_t k)
{
    if (!W.blockMajor) return idx + k * (uint32_t)W.vw * (uint32_t)W.vh;
    const uint32_t nS = (uint32_t)W.nSamples, lps = (uint32_t)W.lps, lpw = (uint32_t)W.lpw, lpp = 5u - lps, lph = lpp - lpw;
    const uint32_t pxin = idx & 7u, pyin = (idx >> 3) & 3u;
    const uint32_t pb = ((pyin >> lph) << (3u - lpw)) + (pxin >> lpw);
    const uint32_t pi = ((pyin & ((1u << lph) - 1u)) << lpw) + (pxin & ((1u << lpw) - 1u));
    const uint32_t gi = pb * (nS >> lps) + (k >> lps);
    return ((((idx >> 5) * nS + gi) << 5) | ((k & ((1u << lps) - 1u)) << lpp)) + pi;
}

// ------------------------------------------------------------------ camera ray --------------------------------------
// tile.glsl:41-45 / preview.glsl:41-43: what a pixel contributes to its ray before any random number is drawn — the texture coordinate of the
// pixel centre in the full frame (mix(tileOffset, tileOffset + invNumTiles, TexCoords)) and the RNG seed (tile-local gl_FragCoord, frameNum).
struct PixelSeed { float cx, cy; uint32_t sx, sy, frame; };

// frameNum of pass s, tile (tx,ty): first Update is the dirty one, tiles run x-fastest from the top row (Renderer.cpp:745-762)
__device__ __forceinline__ uint32_t frameOf(const FrameParams& F, const WaveParams& W, int tx, int ty, int samplePass)
{
    if (W.fixedFrame >= 0) return (uint32_t)W.fixedFrame;
    const int T = F.numTilesX * F.numTilesY;
    const int j = (F.numTilesY - 1 - ty) * F.numTilesX + tx;
    return (uint32_t)(2 + (samplePass - 1) * T + j);
}

// evaluated per pixel (preview target, parity entry points, host harness)
__device__ __forceinline__ PixelSeed pixelSeed(const FrameParams& F, const WaveParams& W, int x, int y, int samplePass)
{
    PixelSeed p;
    if (W.previewMode)
    {
        p.cx = __fdiv_rn((float)x + 0.5f, (float)W.rw); p.cy = __fdiv_rn((float)y + 0.5f, (float)W.rh);     // TexCoords over the whole low-res target
        p.sx = (uint32_t)x; p.sy = (uint32_t)y; p.frame = 1u;                              // preview.glsl:43
        return p;
    }
    int tx = x / F.tileW, ty = y / F.tileH, lx = x - tx * F.tileW, ly = y - ty * F.tileH;
    float tcx = __fdiv_rn((float)lx + 0.5f, (float)F.tileW), tcy = __fdiv_rn((float)ly + 0.5f, (float)F.tileH);
    float offx = __fmul_rn((float)tx, F.invNumTilesX), offy = __fmul_rn((float)ty, F.invNumTilesY);        // Renderer.cpp:780
    // mix(tileOffset, tileOffset + invNumTiles, TexCoords)  (tile.glsl:43)
    p.cx = __fadd_rn(__fmul_rn(offx, __fsub_rn(1.0f, tcx)), __fmul_rn(__fadd_rn(offx, F.invNumTilesX), tcx));
    p.cy = __fadd_rn(__fmul_rn(offy, __fsub_rn(1.0f, tcy)), __fmul_rn(__fadd_rn(offy, F.invNumTilesY), tcy));
    p.sx = (uint32_t)lx; p.sy = (uint32_t)ly;                                              // gl_FragCoord is tile-local (tile.glsl:45)
    p.frame = frameOf(F, W, tx, ty, samplePass);
    return p;
}

// The same from the per-column / per-row tables the host builds once per (resolution, tile size) with the arithmetic above
// (ptbd_build_pixel_tables: {coordinate, bits(local | tile << 16)}): two 8-byte loads instead of two integer and two IEEE divisions per pixel.
__device__ __forceinline__ PixelSeed pixelSeedFromTables(const FrameParams& F, const WaveParams& W, int x, int y, int samplePass)
{
    const float2 ex = __ldg(F.pixTabX + x), ey = __ldg(F.pixTabY + y);
    const uint32_t ux = __float_as_uint(ex.y), uy = __float_as_uint(ey.y);
    PixelSeed p;
    p.cx = ex.x; p.cy = ey.x; p.sx = ux & 0xffffu; p.sy = uy & 0xffffu;
    p.frame = frameOf(F, W, (int)(ux >> 16), (int)(uy >> 16), samplePass);
    return p;
}

// tile.glsl:45-68 / preview.glsl:43-67: RNG seeding, tent-filter jitter, pinhole / thin-lens ray.  Leaves rng advanced by 4 draws.
__device__ __forceinline__ void cameraRayFromSeed(const FrameParams& F, const PixelSeed& ps, Rng& rng, float3& ro, float3& rd)
{
    const float cx = ps.cx, cy = ps.cy;
    rng.init(ps.sx, ps.sy, ps.frame);
    float r1 = __fmul_rn(2.0f, rng.rand());
    float r2 = __fmul_rn(2.0f, rng.rand());
    float jx = r1 < 1.0f ? __fsub_rn(__fsqrt_rn(r1), 1.0f) : __fsub_rn(1.0f, __fsqrt_rn(__fsub_rn(2.0f, r1)));
    float jy = r2 < 1.0f ? __fsub_rn(__fsqrt_rn(r2), 1.0f) : __fsub_rn(1.0f, __fsqrt_rn(__fsub_rn(2.0f, r2)));
    jx = __fdiv_rn(jx, __fmul_rn((float)F.renderW, 0.5f));
    jy = __fdiv_rn(jy, __fmul_rn((float)F.renderH, 0.5f));
    float dx = __fadd_rn(__fsub_rn(__fmul_rn(cx, 2.0f), 1.0f), jx);
    float dy = __fadd_rn(__fsub_rn(__fmul_rn(cy, 2.0f), 1.0f), jy);
    float scale = F.camScale;
    dy = __fmul_rn(dy, __fmul_rn(F.aspect, scale));     // aspect = float(renderH) / float(renderW), one IEEE division on the host
    dx = __fmul_rn(dx, scale);
    float3 right = f3(F.camRight[0], F.camRight[1], F.camRight[2]), up = f3(F.camUp[0], F.camUp[1], F.camUp[2]),
           fwd = f3(F.camFwd[0], F.camFwd[1], F.camFwd[2]), pos = f3(F.camPos[0], F.camPos[1], F.camPos[2]);
    // exact-op evaluation keeps pinhole primary rays bit-identical to the oracle's (aperture 0 => no sin/cos influence)
    float3 v = f3(xa(xa(xm(dx, right.x), xm(dy, up.x)), fwd.x), xa(xa(xm(dx, right.y), xm(dy, up.y)), fwd.y), xa(xa(xm(dx, right.z), xm(dy, up.z)), fwd.z));
    float vl = __fsqrt_rn(xdot(v, v));
    float3 rayDir = f3(xd(v.x, vl), xd(v.y, vl), xd(v.z, vl));
    float3 focalPoint = f3(xm(F.camFocalDist, rayDir.x), xm(F.camFocalDist, rayDir.y), xm(F.camFocalDist, rayDir.z));
    float cam_r1 = __fmul_rn(rng.rand(), PTB_TWO_PI);
    float cam_r2 = __fmul_rn(rng.rand(), F.camAperture);
    float sr = __fsqrt_rn(cam_r2);
    float3 ap = f3(0.f);
    if (F.camAperture != 0.0f)       // pinhole: the lens offset is (cos,sin)*sqrt(0) = 0, skip the sincos (the two draws above are still consumed)
    {
        float s1, c1; sincosf(cam_r1, &s1, &c1);
        ap = f3(xm(xa(xm(c1, right.x), xm(s1, up.x)), sr), xm(xa(xm(c1, right.y), xm(s1, up.y)), sr), xm(xa(xm(c1, right.z), xm(s1, up.z)), sr));
    }
    float3 fd = xsub(focalPoint, ap);
    float fl = __fsqrt_rn(xdot(fd, fd));
    rd = f3(xd(fd.x, fl), xd(fd.y, fl), xd(fd.z, fl));
    ro = f3(xa(pos.x, ap.x), xa(pos.y, ap.y), xa(pos.z, ap.z));
}
// (x,y) absolute pixel
__device__ __forceinline__ void cameraRay(const FrameParams& F, const WaveParams& W, int x, int y, int samplePass, Rng& rng, float3& ro, float3& rd)
{
    const PixelSeed ps = (F.pixTabX && !W.previewMode) ? pixelSeedFromTables(F, W, x, y, samplePass) : pixelSeed(F, W, x, y, samplePass);
    cameraRayFromSeed(F, ps, rng, ro, rd);
}


// ------------------------------------------------------------------ sampling.glsl -------------------------------
__device__ __forceinline__ float GTR1(float NDotH, float a)   // :25-32
{
    if (a >= 1.0f) return PTB_INV_PI;
    float a2 = a * a;
    float t = 1.0f + (a2 - 1.0f) * NDotH * NDotH;
    return fdiv(a2 - 1.0f, PTB_PI * logf(a2) * t);
}
__device__ __forceinline__ float3 SampleGTR1(float rgh, float r1, float r2)   // :34-47
{
    float a = fmaxf(0.001f, rgh);
    float a2 = a * a;
    float phi = r1 * PTB_TWO_PI;
    float cosTheta = sqrtApprox(fdiv(1.0f - powf(a2, 1.0f - r2), 1.0f - a2));
    float sinTheta = clampf(sqrtApprox(1.0f - (cosTheta * cosTheta)), 0.0f, 1.0f);
    float sinPhi, cosPhi; sincosf(phi, &sinPhi, &cosPhi);
    return f3(sinTheta * cosPhi, sinTheta * sinPhi, cosTheta);
}
__device__ __forceinline__ float3 SampleGGXVNDF(float3 V, float ax, float ay, float r1, float r2)   // :70-88
{
    float3 Vh = normalize(f3(ax * V.x, ay * V.y, V.z));
    float lensq = Vh.x * Vh.x + Vh.y * Vh.y;
    float3 T1 = lensq > 0 ? f3(-Vh.y, Vh.x, 0) * rsqrtApprox(lensq) : f3(1, 0, 0);
    float3 T2 = cross(Vh, T1);
    float r = sqrtApprox(r1);
    float phi = 2.0f * PTB_PI * r2;
    float sp, cp; sincosf(phi, &sp, &cp);
    float t1 = r * cp;
    float t2 = r * sp;
    float s = 0.5f * (1.0f + Vh.z);
    t2 = (1.0f - s) * sqrtApprox(1.0f - t1 * t1) + s * t2;
    float3 Nh = t1 * T1 + t2 * T2 + sqrtApprox(fmaxf(0.0f, 1.0f - t1 * t1 - t2 * t2)) * Vh;
    return normalize(f3(ax * Nh.x, ay * Nh.y, fmaxf(0.0f, Nh.z)));
}
__device__ __forceinline__ float GTR2Aniso(float NDotH, float HDotX, float HDotY, float ax, float ay)   // :90-96
{
    float a = fdiv(HDotX, ax);
    float b = fdiv(HDotY, ay);
    float c = a * a + b * b + NDotH * NDotH;
    return fdiv(1.0f, PTB_PI * ax * ay * c * c);
}
__device__ __forceinline__ float SmithG(float NDotV, float alphaG)   // :109-114
{
    float a = alphaG * alphaG;
    float b = NDotV * NDotV;
    return fdiv(2.0f * NDotV, NDotV + sqrtApprox(a + b - a * b));
}
__device__ __forceinline__ float SmithGAniso(float NDotV, float VDotX, float VDotY, float ax, float ay)   // :116-122
{
    float a = VDotX * ax;
    float b = VDotY * ay;
    float c = NDotV;
    return fdiv(2.0f * NDotV, NDotV + sqrtApprox(a * a + b * b + c * c));
}
__device__ __forceinline__ float SchlickWeight(float u)   // :124-129
{
    float m = clampf(1.0f - u, 0.0f, 1.0f);
    float m2 = m * m;
    return m2 * m2 * m;
}
__device__ __forceinline__ float DielectricFresnel(float cosThetaI, float eta)   // :131-145
{
    float sinThetaTSq = eta * eta * (1.0f - cosThetaI * cosThetaI);
    if (sinThetaTSq > 1.0f) return 1.0f;
    float cosThetaT = sqrtApprox(fmaxf(1.0f - sinThetaTSq, 0.0f));
    float rs = fdiv(eta * cosThetaT - cosThetaI, eta * cosThetaT + cosThetaI);
    float rp = fdiv(eta * cosThetaI - cosThetaT, eta * cosThetaI + cosThetaT);
    return 0.5f * (rs * rs + rp * rp);
}
__device__ __forceinline__ float3 CosineSampleHemisphere(float r1, float r2)   // :147-156
{
    float3 dir;
    float r = sqrtApprox(r1);
    float phi = PTB_TWO_PI * r2;
    float sp, cp; sincosf(phi, &sp, &cp);
    dir.x = r * cp;
    dir.y = r * sp;
    dir.z = sqrtApprox(fmaxf(0.0f, 1.0f - dir.x * dir.x - dir.y * dir.y));
    return dir;
}
__device__ __forceinline__ float3 UniformSampleHemisphere(float r1, float r2)   // :158-163
{
    float r = sqrtApprox(fmaxf(0.0f, 1.0f - r1 * r1));
    float phi = PTB_TWO_PI * r2;
    float sp, cp; sincosf(phi, &sp, &cp);
    return f3(r * cp, r * sp, r1);
}
__device__ __forceinline__ float PowerHeuristic(float a, float b)   // :173-177
{
    float t = a * a;
    return fdiv(t, b * b + t);
}
__device__ __forceinline__ void Onb(float3 N, float3& T, float3& B)   // :179-184
{
    float3 up = fabsf(N.z) < 0.9999999f ? f3(0, 0, 1) : f3(1, 0, 0);
    T = normalize(cross(up, N));
    B = cross(N, T);
}
__device__ __forceinline__ float3 SampleHG(float3 V, float g, float r1, float r2)   // :250-270
{
    float cosTheta;
    if (fabsf(g) < 0.001f) cosTheta = 1 - 2 * r2;
    else
    {
        float sqrTerm = (1 - g * g) / (1 + g - 2 * g * r2);
        cosTheta = -(1 + g * g - sqrTerm * sqrTerm) / (2 * g);
    }
    float phi = r1 * PTB_TWO_PI;
    float sinTheta = clampf(sqrtApprox(1.0f - (cosTheta * cosTheta)), 0.0f, 1.0f);
    float sinPhi, cosPhi; sincosf(phi, &sinPhi, &cosPhi);
    float3 v1, v2;
    Onb(V, v1, v2);
    return sinTheta * cosPhi * v1 + sinTheta * sinPhi * v2 + cosTheta * V;
}
__device__ __forceinline__ float PhaseHG(float cosTheta, float g)   // :272-276
{
    float denom = 1 + g * g + 2 * g * cosTheta;
    return PTB_INV_4_PI * (1 - g * g) / (denom * sqrtApprox(denom));
}

// ------------------------------------------------------------------ material / surface state --------------------
struct Material   // globals.glsl:60-83
{
    float3 baseColor; float opacity; int alphaMode; float alphaCutoff; float3 emission; float anisotropic, metallic, roughness,
        subsurface, specularTint, sheen, sheenTint, clearcoat, clearcoatRoughness, specTrans, ior, ax, ay;
    int medType; float medDensity; float3 medColor; float medAniso;
};

struct LightSample { float3 normal, emission, direction; float dist, pdf; };

// sampling.glsl:186-248 (r1,r2 drawn by the caller in the reference order; distant lights draw none)
__device__ __forceinline__ void SampleOneLight(const DevScene& S, int idx, float3 scatterPos, Rng& rng, LightSample& ls, float& area)
{
    const float4* p = S.lightsPre + (size_t)idx * 8;
    float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2), dd = __ldg(p + 3);
    float3 position = f3(a), emission = f3(b), u = f3(c), v = f3(dd);
    (void)v;
    int type = (int)a.w; area = b.w; float radius = c.w;
    float nl = (float)S.numLights;
    if (type == 0)
    {
        float r1 = rng.rand(), r2 = rng.rand();
        float3 lightSurfacePos = position + u * r1 + v * r2;
        ls.direction = lightSurfacePos - scatterPos;
        ls.dist = length(ls.direction);
        float distSq = ls.dist * ls.dist;
        ls.direction /= ls.dist;
        ls.normal = f3(__ldg(p + 4));                 // normalize(cross(u, v)), evaluated once at upload
        ls.emission = emission * nl;
        ls.pdf = fdiv(distSq, area * fabsf(dot(ls.normal, ls.direction)));
    }
    else if (type == 1)
    {
        // The shadow ray towards a sphere light is tested against the light itself (anyhit.glsl:56-61) with maxDist = dist - EPS; for a grazing
        // sample SphereIntersect's b*b - |op|^2 + r^2 cancels to ~1e-2 at |op| ~ 40, so a systematic half-ulp shift of |direction| (div.approx /
        // rsqrt normalisation) moves t by more than EPS and biases the visibility of ~1 % of the NEE samples (measured: -0.5 % mean radiance on
        // hyperion_sphere_light at 256 spp).  The geometry of this sample therefore uses IEEE operations in the reference's order; what is left
        // between the two implementations is unbiased rounding noise of sin/cos and of the shading point.
        float r1 = rng.rand(), r2 = rng.rand();
        float3 c2s = xsub(scatterPos, position);
        const float distToCenter = __fsqrt_rn(xdot(c2s, c2s));
        c2s = f3(xd(c2s.x, distToCenter), xd(c2s.y, distToCenter), xd(c2s.z, distToCenter));
        const float hr = __fsqrt_rn(fmaxf(0.0f, xs(1.0f, xm(r1, r1))));          // UniformSampleHemisphere (sampling.glsl:158-163)
        float sp, cp; sincosf(xm(PTB_TWO_PI, r2), &sp, &cp);
        const float3 sd = f3(xm(hr, cp), xm(hr, sp), r1);
        const float3 up = fabsf(c2s.z) < 0.9999999f ? f3(0, 0, 1) : f3(1, 0, 0);   // Onb (:179-184)
        float3 T = xcross(up, c2s);
        const float tl = __fsqrt_rn(xdot(T, T));
        T = f3(xd(T.x, tl), xd(T.y, tl), xd(T.z, tl));
        const float3 B = xcross(c2s, T);
        const float3 sampledDir = f3(xa(xa(xm(T.x, sd.x), xm(B.x, sd.y)), xm(c2s.x, sd.z)), xa(xa(xm(T.y, sd.x), xm(B.y, sd.y)), xm(c2s.y, sd.z)),
                                     xa(xa(xm(T.z, sd.x), xm(B.z, sd.y)), xm(c2s.z, sd.z)));
        const float3 lightSurfacePos = f3(xa(position.x, xm(sampledDir.x, radius)), xa(position.y, xm(sampledDir.y, radius)), xa(position.z, xm(sampledDir.z, radius)));
        ls.direction = xsub(lightSurfacePos, scatterPos);
        ls.dist = __fsqrt_rn(xdot(ls.direction, ls.direction));
        const float distSq = xm(ls.dist, ls.dist);
        ls.direction = f3(xd(ls.direction.x, ls.dist), xd(ls.direction.y, ls.dist), xd(ls.direction.z, ls.dist));
        ls.normal = normalize(lightSurfacePos - position);
        ls.emission = emission * nl;
        ls.pdf = fdiv(distSq, area * 0.5f * fabsf(dot(ls.normal, ls.direction)));
    }
    else
    {
        ls.direction = normalize(position);
        ls.normal = normalize(scatterPos - position);
        ls.emission = emission * nl;
        ls.dist = PTB_INF;
        ls.pdf = 1.0f;
    }
}

// ------------------------------------------------------------------ disney.glsl ---------------------------------
__device__ __forceinline__ float3 ToWorld(float3 X, float3 Y, float3 Z, float3 V) { return V.x * X + V.y * Y + V.z * Z; }
__device__ __forceinline__ float3 ToLocal(float3 X, float3 Y, float3 Z, float3 V) { return f3(dot(V, X), dot(V, Y), dot(V, Z)); }

struct Lobes { float diffPr, dielectricPr, metalPr, glassPr, clearCtPr, dielectricWt, metalWt, glassWt; float F0; float3 Csheen, Cspec0; };

// TintColors (disney.glsl:49-59) + model weights / lobe probabilities (:155-179, :272-294)
__device__ __forceinline__ void lobeSetup(const Material& m, float eta, float Vz, Lobes& p)
{
    float lum = Luminance(m.baseColor);
    float3 ctint = lum > 0.0f ? m.baseColor / lum : f3(1.0f);
    float F0 = fdiv(1.0f - eta, 1.0f + eta);
    F0 *= F0;
    p.F0 = F0;
    p.Cspec0 = F0 * mix(f3(1.0f), ctint, m.specularTint);
    p.Csheen = mix(f3(1.0f), ctint, m.sheenTint);
    p.dielectricWt = (1.0f - m.metallic) * (1.0f - m.specTrans);
    p.metalWt = m.metallic;
    p.glassWt = (1.0f - m.metallic) * m.specTrans;
    float schlickWt = SchlickWeight(Vz);
    p.diffPr = p.dielectricWt * Luminance(m.baseColor);
    p.dielectricPr = p.dielectricWt * Luminance(mix(p.Cspec0, f3(1.0f), schlickWt));
    p.metalPr = p.metalWt * Luminance(mix(m.baseColor, f3(1.0f), schlickWt));
    p.glassPr = p.glassWt;
    p.clearCtPr = 0.25f * m.clearcoat;
    float invTotalWt = fdiv(1.0f, p.diffPr + p.dielectricPr + p.metalPr + p.glassPr + p.clearCtPr);
    p.diffPr *= invTotalWt; p.dielectricPr *= invTotalWt; p.metalPr *= invTotalWt; p.glassPr *= invTotalWt; p.clearCtPr *= invTotalWt;
}

__device__ __forceinline__ float3 EvalMicrofacetReflection(const Material& mat, float3 V, float3 L, float3 H, float3 F, float& pdf)   // :89-101
{
    pdf = 0.0f;
    if (L.z <= 0.0f) return f3(0.0f);
    float D = GTR2Aniso(H.z, H.x, H.y, mat.ax, mat.ay);
    float G1 = SmithGAniso(fabsf(V.z), V.x, V.y, mat.ax, mat.ay);
    float G2 = G1 * SmithGAniso(fabsf(L.z), L.x, L.y, mat.ax, mat.ay);
    pdf = fdiv(G1 * D, 4.0f * V.z);
    return F * fdiv(D * G2, 4.0f * L.z * V.z);
}

// DisneyEval (disney.glsl:244-351) in the local frame (T,B,N); V,L already local; p = lobeSetup(mat, eta, V.z) (depends on V only,
// so the NEE evaluation and the evaluation of the sampled direction of one hit share it).
__device__ __forceinline__ float3 DisneyEvalLocal(const Material& mat, float eta, const Lobes& p, float3 V, float3 L, float& pdf)
{
    pdf = 0.0f;
    float3 f = f3(0.0f);
    float3 H;
    if (L.z > 0.0f) H = normalize(L + V);
    else H = normalize(L + V * eta);
    if (H.z < 0.0f) H = -H;
    bool reflect = L.z * V.z > 0;
    float tmpPdf = 0.0f;
    float VDotH = fabsf(dot(V, H));

    if (p.diffPr > 0.0f && reflect)   // EvalDisneyDiffuse :61-87
    {
        tmpPdf = 0.0f;
        float3 fd = f3(0.0f);
        if (L.z > 0.0f)
        {
            float LDotH = dot(L, H);
            float Rr = 2.0f * mat.roughness * LDotH * LDotH;
            float FL = SchlickWeight(L.z);
            float FV = SchlickWeight(V.z);
            float Fretro = Rr * (FL + FV + FL * FV * (Rr - 1.0f));
            float Fd = (1.0f - 0.5f * FL) * (1.0f - 0.5f * FV);
            float Fss90 = 0.5f * Rr;
            float Fss = mixf(1.0f, Fss90, FL) * mixf(1.0f, Fss90, FV);
            float ss = 1.25f * (Fss * (fdiv(1.0f, L.z + V.z) - 0.5f) + 0.5f);
            float FH = SchlickWeight(LDotH);
            float3 Fsheen = FH * mat.sheen * p.Csheen;
            tmpPdf = L.z * PTB_INV_PI;
            fd = PTB_INV_PI * mat.baseColor * mixf(Fd + Fretro, ss, mat.subsurface) + Fsheen;
        }
        f += fd * p.dielectricWt;
        pdf += tmpPdf * p.diffPr;
    }
    if (p.dielectricPr > 0.0f && reflect)
    {
        float F = fdiv(DielectricFresnel(VDotH, fdiv(1.0f, mat.ior)) - p.F0, 1.0f - p.F0);
        f += EvalMicrofacetReflection(mat, V, L, H, mix(p.Cspec0, f3(1.0f), F), tmpPdf) * p.dielectricWt;
        pdf += tmpPdf * p.dielectricPr;
    }
    if (p.metalPr > 0.0f && reflect)
    {
        float3 F = mix(mat.baseColor, f3(1.0f), SchlickWeight(VDotH));
        f += EvalMicrofacetReflection(mat, V, L, H, F, tmpPdf) * p.metalWt;
        pdf += tmpPdf * p.metalPr;
    }
    if (p.glassPr > 0.0f)
    {
        float F = DielectricFresnel(VDotH, eta);
        if (reflect)
        {
            f += EvalMicrofacetReflection(mat, V, L, H, f3(F), tmpPdf) * p.glassWt;
            pdf += tmpPdf * p.glassPr * F;
        }
        else   // EvalMicrofacetRefraction :103-123
        {
            tmpPdf = 0.0f;
            float3 fr = f3(0.0f);
            if (!(L.z >= 0.0f))
            {
                float LDotH = dot(L, H);
                float VDotH2 = dot(V, H);
                float D = GTR2Aniso(H.z, H.x, H.y, mat.ax, mat.ay);
                float G1 = SmithGAniso(fabsf(V.z), V.x, V.y, mat.ax, mat.ay);
                float G2 = G1 * SmithGAniso(fabsf(L.z), L.x, L.y, mat.ax, mat.ay);
                float denom = LDotH + VDotH2 * eta;
                denom *= denom;
                float eta2 = eta * eta;
                float jacobian = fdiv(fabsf(LDotH), denom);
                tmpPdf = fdiv(G1 * fmaxf(0.0f, VDotH2) * D * jacobian, V.z);
                // pow(mat.baseColor, vec3(0.5)) of disney.glsl:122 as a square root
                fr = f3(sqrtApprox(mat.baseColor.x), sqrtApprox(mat.baseColor.y), sqrtApprox(mat.baseColor.z)) * (f3(1.0f) - f3(F)) * fdiv(D * G2 * fabsf(VDotH2) * jacobian * eta2, fabsf(L.z * V.z));
            }
            f += fr * p.glassWt;
            pdf += tmpPdf * p.glassPr * (1.0f - F);
        }
    }
    if (p.clearCtPr > 0.0f && reflect)   // EvalClearcoat :125-140
    {
        tmpPdf = 0.0f;
        float3 fc = f3(0.0f);
        if (L.z > 0.0f)
        {
            float VDotH2 = dot(V, H);
            float F = mixf(0.04f, 1.0f, SchlickWeight(VDotH2));
            float D = GTR1(H.z, mat.clearcoatRoughness);
            float G = SmithG(L.z, 0.25f) * SmithG(V.z, 0.25f);
            float jacobian = fdiv(1.0f, 4.0f * VDotH2);
            tmpPdf = D * H.z * jacobian;
            fc = f3(F) * D * G;
        }
        f += fc * 0.25f * mat.clearcoat;
        pdf += tmpPdf * p.clearCtPr;
    }
    return f * fabsf(L.z);
}

// Shading frame of one hit: Onb(N), V in that frame, lobe weights — computed once per hit, used by the NEE evaluations and the sample.
struct ShadeFrame { float3 T, B, N, Vl; Lobes lobes; };
__device__ __forceinline__ void frameSetup(const Material& mat, float eta, float3 V, float3 N, ShadeFrame& fr)
{
    fr.N = N;
    Onb(N, fr.T, fr.B);
    fr.Vl = ToLocal(fr.T, fr.B, N, V);
    lobeSetup(mat, eta, fr.Vl.z, fr.lobes);
}
__device__ __forceinline__ float3 DisneyEvalFr(const Material& mat, float eta, const ShadeFrame& fr, float3 L, float& pdf)
{
    return DisneyEvalLocal(mat, eta, fr.lobes, fr.Vl, ToLocal(fr.T, fr.B, fr.N, L), pdf);
}
__device__ __forceinline__ float3 DisneyEval(const Material& mat, float eta, float3 V, float3 N, float3 L, float& pdf)
{
    ShadeFrame fr;
    frameSetup(mat, eta, V, N, fr);
    return DisneyEvalFr(mat, eta, fr, L, pdf);
}

// DisneySample (disney.glsl:142-242) with the three draws passed in and the hit's frame / lobe weights precomputed.
__device__ __forceinline__ float3 DisneySampleFr(const Material& mat, float eta, const ShadeFrame& fr, float3& Lw, float& pdf, float r1, float r2, float r3)
{
    pdf = 0.0f;
    const float3 V = fr.Vl;
    const Lobes& p = fr.lobes;
    float cdf0 = p.diffPr;
    float cdf1 = cdf0 + p.dielectricPr;
    float cdf2 = cdf1 + p.metalPr;
    float cdf3 = cdf2 + p.glassPr;
    float3 L;
    if (r3 < cdf0) L = CosineSampleHemisphere(r1, r2);
    else if (r3 < cdf2)
    {
        float3 H = SampleGGXVNDF(V, mat.ax, mat.ay, r1, r2);
        if (H.z < 0.0f) H = -H;
        L = normalize(reflect(-V, H));
    }
    else if (r3 < cdf3)
    {
        float3 H = SampleGGXVNDF(V, mat.ax, mat.ay, r1, r2);
        float F = DielectricFresnel(fabsf(dot(V, H)), eta);
        if (H.z < 0.0f) H = -H;
        r3 = fdiv(r3 - cdf2, cdf3 - cdf2);
        if (r3 < F) L = normalize(reflect(-V, H));
        else L = normalize(refract(-V, H, eta));
    }
    else
    {
        float3 H = SampleGTR1(mat.clearcoatRoughness, r1, r2);
        if (H.z < 0.0f) H = -H;
        L = normalize(reflect(-V, H));
    }
    Lw = ToWorld(fr.T, fr.B, fr.N, L);
    // the reference converts V and L back to world space and calls DisneyEval (disney.glsl:238-241), which re-derives the same frame
    // and takes both to local space again; L makes that round trip here, V (and with it the lobe weights) is reused.
    return DisneyEvalLocal(mat, eta, p, V, ToLocal(fr.T, fr.B, fr.N, Lw), pdf);
}
__device__ __forceinline__ float3 DisneySample(const Material& mat, float eta, float3 V, float3 N, float3& Lw, float& pdf, float r1, float r2, float r3)
{
    ShadeFrame fr;
    frameSetup(mat, eta, V, N, fr);
    return DisneySampleFr(mat, eta, fr, Lw, pdf, r1, r2, r3);
}

// lambert.glsl:25-46.  Included by tile.glsl but never called from PathTrace; built for the parity entry point only.
__device__ __forceinline__ float3 LambertEval(const Material& mat, float3 V, float3 N, float3 L, float& pdf)
{
    pdf = dot(N, L) * (1.0f / PTB_PI);
    return (1.0f / PTB_PI) * mat.baseColor * dot(N, L);
}
__device__ __forceinline__ float3 LambertSample(const Material& mat, float3 V, float3 N, float3& L, float& pdf, float r1, float r2)
{
    float3 T, B;
    Onb(N, T, B);
    L = CosineSampleHemisphere(r1, r2);
    L = T * L.x + B * L.y + N * L.z;
    pdf = dot(N, L) * (1.0f / PTB_PI);
    return (1.0f / PTB_PI) * mat.baseColor * dot(N, L);
}

// Material row -> Material (pathtrace.glsl:31-67 + :109-114), textures handled by the caller.
__device__ __forceinline__ void materialFromRow(const float4* P, Material& mat, int4& texIDs)
{
    float4 p1 = __ldg(P), p2 = __ldg(P + 1), p3 = __ldg(P + 2), p4 = __ldg(P + 3), p5 = __ldg(P + 4), p6 = __ldg(P + 5), p7 = __ldg(P + 6), p8 = __ldg(P + 7);
    mat.baseColor = f3(p1); mat.anisotropic = p1.w;
    mat.emission = f3(p2);
    mat.metallic = p3.x; mat.roughness = fmaxf(p3.y, 0.001f); mat.subsurface = p3.z; mat.specularTint = p3.w;
    mat.sheen = p4.x; mat.sheenTint = p4.y; mat.clearcoat = p4.z; mat.clearcoatRoughness = mixf(0.1f, 0.001f, p4.w);
    mat.specTrans = p5.x; mat.ior = p5.y; mat.medType = (int)p5.z; mat.medDensity = p5.w;
    mat.medColor = f3(p6); mat.medAniso = clampf(p6.w, -0.9f, 0.9f);
    texIDs = make_int4((int)p7.x, (int)p7.y, (int)p7.z, (int)p7.w);
    mat.opacity = p8.x; mat.alphaMode = (int)p8.y; mat.alphaCutoff = p8.z;
}
__device__ __forceinline__ void materialFinish(Material& mat)
{
    float aspect = sqrtApprox(1.0f - mat.anisotropic * 0.9f);
    mat.ax = fmaxf(0.001f, fdiv(mat.roughness, aspect));
    mat.ay = fmaxf(0.001f, mat.roughness * aspect);
}

// texture(textureMapsArrayTex, vec3(uv, layer)): RGBA8 unorm, LINEAR, REPEAT — manual fp32 bilinear
__device__ __forceinline__ int wrapi(float f, int n) { int i = (int)fmodf(f, (float)n); if (i < 0) i += n; return i; }
__device__ __forceinline__ float4 sampleTexArray(const DevScene& S, float2 uv, float layerf)
{
    int layer = (int)floorf(layerf + 0.5f); layer = max(0, min(layer, S.numTextures - 1));
    int W = S.texW, H = S.texH;
    float x = uv.x * (float)W - 0.5f, y = uv.y * (float)H - 0.5f;
    float fx0 = floorf(x), fy0 = floorf(y);
    float ax = x - fx0, ay = y - fy0;
    int x0 = wrapi(fx0, W), x1 = wrapi(fx0 + 1.0f, W), y0 = wrapi(fy0, H), y1 = wrapi(fy0 + 1.0f, H);
    const uchar4* base = S.textures + (size_t)layer * W * H;
    uchar4 c00 = __ldg(base + (size_t)y0 * W + x0), c10 = __ldg(base + (size_t)y0 * W + x1), c01 = __ldg(base + (size_t)y1 * W + x0),
           c11 = __ldg(base + (size_t)y1 * W + x1);
    float4 o;
    o.x = mixf(mixf(c00.x / 255.0f, c10.x / 255.0f, ax), mixf(c01.x / 255.0f, c11.x / 255.0f, ax), ay);
    o.y = mixf(mixf(c00.y / 255.0f, c10.y / 255.0f, ax), mixf(c01.y / 255.0f, c11.y / 255.0f, ax), ay);
    o.z = mixf(mixf(c00.z / 255.0f, c10.z / 255.0f, ax), mixf(c01.z / 255.0f, c11.z / 255.0f, ax), ay);
    o.w = mixf(mixf(c00.w / 255.0f, c10.w / 255.0f, ax), mixf(c01.w / 255.0f, c11.w / 255.0f, ax), ay);
    return o;
}

// ------------------------------------------------------------------ envmap.glsl ---------------------------------
__device__ __forceinline__ float3 envTexel(const DevScene& S, int x, int y)
{
    const float* p = S.envImg + ((size_t)y * S.envW + x) * 3;
    return f3(__ldg(p), __ldg(p + 1), __ldg(p + 2));
}
__device__ __forceinline__ float3 sampleEnv(const DevScene& S, float2 uv)   // texture(envMapTex, uv): RGB32F LINEAR REPEAT
{
    int W = S.envW, H = S.envH;
    float x = uv.x * (float)W - 0.5f, y = uv.y * (float)H - 0.5f;
    float fx0 = floorf(x), fy0 = floorf(y);
    float ax = x - fx0, ay = y - fy0;
    int x0 = wrapi(fx0, W), x1 = wrapi(fx0 + 1.0f, W), y0 = wrapi(fy0, H), y1 = wrapi(fy0 + 1.0f, H);
    return mix(mix(envTexel(S, x0, y0), envTexel(S, x1, y0), ax), mix(envTexel(S, x0, y1), envTexel(S, x1, y1), ax), ay);
}
__device__ __forceinline__ float2 envBinarySearch(const DevScene& S, float value)   // envmap.glsl:28-55
{
    int W = S.envW, H = S.envH;
    if (S.envGuide)
    {   // Non-decreasing CDF (checked at upload): the two searches below return the row / column of the FIRST texel f (row-major) with value < cdf[f] — the last
        // row / column when there is none — whatever the probing sequence; the guide table brackets f within ~2 texels (see ptbd_build_env_guide), which replaces
        // ~20 dependent loads by ~3.  Same texel, same uv.
        const float bx = fminf(fmaxf(__fmul_rn(value, S.envGuideScale), 0.0f), (float)(S.envGuideN - 1));
        const int b = (int)bx;
        uint32_t lower = __ldg(S.envGuide + b), upper = __ldg(S.envGuide + b + 1);
        while (lower < upper)
        {
            const uint32_t mid = (lower + upper) >> 1;
            if (value < __ldg(S.envCdf + mid)) upper = mid;
            else lower = mid + 1;
        }
        const uint32_t n = (uint32_t)W * (uint32_t)H;
        const int y = min((int)(lower / (uint32_t)W), H - 1);
        const int x = lower >= n ? W - 1 : (int)(lower - (uint32_t)y * (uint32_t)W);
        return make_float2((float)x / (float)W, (float)y / (float)H);
    }
    int lower = 0, upper = H - 1;
    while (lower < upper)
    {
        int mid = (lower + upper) >> 1;
        if (value < __ldg(S.envCdf + (size_t)mid * W + (W - 1))) upper = mid;
        else lower = mid + 1;
    }
    int y = max(0, min(lower, H - 1));
    lower = 0; upper = W - 1;
    while (lower < upper)
    {
        int mid = (lower + upper) >> 1;
        if (value < __ldg(S.envCdf + (size_t)y * W + mid)) upper = mid;
        else lower = mid + 1;
    }
    int x = max(0, min(lower, W - 1));
    return make_float2((float)x / (float)W, (float)y / (float)H);
}
__device__ __forceinline__ float4 EvalEnvMap(const DevScene& S, const FrameParams& F, float3 dir)   // envmap.glsl:57-66
{
    float theta = acosf(clampf(dir.y, -1.0f, 1.0f));
    float2 uv = make_float2((PTB_PI + atan2f(dir.z, dir.x)) * PTB_INV_TWO_PI + F.envMapRot, theta * PTB_INV_PI);
    float3 color = sampleEnv(S, uv);
    float pdf = Luminance(color) / S.envTotalSum;
    return make_float4(color.x, color.y, color.z, (pdf * (float)S.envW * (float)S.envH) / (PTB_TWO_PI * PTB_PI * sinf(theta)));
}
__device__ __forceinline__ float4 SampleEnvMap(const DevScene& S, const FrameParams& F, Rng& rng, float3& color)   // envmap.glsl:68-83
{
    float2 uv = envBinarySearch(S, rng.rand() * S.envTotalSum);
    color = sampleEnv(S, uv);
    float pdf = Luminance(color) / S.envTotalSum;
    uv.x -= F.envMapRot;
    float phi = uv.x * PTB_TWO_PI;
    float theta = uv.y * PTB_PI;
    float st, ct, sp, cp; sincosf(theta, &st, &ct); sincosf(phi, &sp, &cp);
    if (st == 0.0f) pdf = 0.0f;
    return make_float4(-st * cp, ct, -st * sp, (pdf * (float)S.envW * (float)S.envH) / (PTB_TWO_PI * PTB_PI * st));
}

} // namespace ptb
