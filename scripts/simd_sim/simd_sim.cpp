// simd_sim.cpp — ANALYSIS TOOL (not product, not test infrastructure): a host-side model of how many warp issue slots the
// closest-hit traversal of glsl-pathtracer_b200/csrc/ptb_device.cuh spends on a given ray stream under different per-warp
// schedules.  32 consecutive rays form a warp; every lane runs the reference's traversal (same visiting order, same arithmetic,
// optional t-culling) as an explicit state machine; a schedule decides which lanes advance in which step, and a step costs its
// SASS instruction count (measured with cuobjdump on the shipped kernel) once per warp, however many lanes take part.
//
//   policy 0  while-while (shipped): all lanes descend internal nodes until each holds a leaf / instance / marker, then those are
//             processed together (leaf loop runs max-triangle-count iterations)
//   policy 1  postponed leaf: a lane that reaches a leaf parks it and keeps descending until it reaches a second non-internal item
//   policy 2  if-if: one item of any kind per lane per iteration, the kinds present are issued one after the other
//
// Built and driven by scripts/simd_sim.py.  Results feed DESIGN.md §9; nothing here is linked into libptb200.so.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

namespace {

struct V3 { float x, y, z; };
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline float dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

struct Scene
{
    const float* nodes; int numNodes, top;
    const int32_t* vi; const float* verts; const float* invT;   // invT: 16 floats per instance, rows of inverse(transform)
};

// SASS instruction counts of the shipped k_trace (cuobjdump, round 1)
constexpr int C_INNER = 77, C_TRI = 64, C_LEAF_SETUP = 6, C_INST = 78, C_POP = 10, C_ROUND = 8;

inline float slab(const float* n, V3 o, V3 inv, float& entry)
{
    float fx = (n[3] - o.x) * inv.x, fy = (n[4] - o.y) * inv.y, fz = (n[5] - o.z) * inv.z;
    float nx = (n[0] - o.x) * inv.x, ny = (n[1] - o.y) * inv.y, nz = (n[2] - o.z) * inv.z;
    float t1 = fminf(fmaxf(fx, nx), fminf(fmaxf(fy, ny), fmaxf(fz, nz)));
    float t0 = fmaxf(fminf(fx, nx), fmaxf(fminf(fy, ny), fminf(fz, nz)));
    entry = t0;
    return (t1 >= t0) ? (t0 > 0.f ? t0 : t1) : -1.0f;
}

struct Lane
{
    bool active = false, done = true, inBlas = false, anyHit = false, occluded = false;
    V3 o, d, ro, rd, inv, invW;
    float t; int prim;
    int cur;                 // node index, or -1 marker / sentinel
    int stack[64]; int sp;
    int pending = -2;        // parked leaf node (policy 1), -2 = none
    long steps = 0;
    bool enteredInst = false;
    void begin(const Scene& S, const float* r)
    {
        o = {r[0], r[1], r[2]}; d = {r[3], r[4], r[5]}; ro = o; rd = d;
        inv = {1.0f / d.x, 1.0f / d.y, 1.0f / d.z}; invW = inv;
        t = 1000000.0f; prim = -1; cur = S.top; sp = 0; stack[sp++] = -1; inBlas = false; active = true; done = false; pending = -2;
    }
};

enum Kind { K_INNER, K_LEAF, K_INST, K_MARK };
inline Kind kindOf(const Scene& S, int cur)
{
    if (cur < 0) return K_MARK;
    int leaf = (int)S.nodes[cur * 9 + 8];
    return leaf == 0 ? K_INNER : (leaf > 0 ? K_LEAF : K_INST);
}

inline void stepInner(const Scene& S, Lane& L, bool cull)
{
    const float* n = S.nodes + (size_t)L.cur * 9;
    int l = (int)n[6], r = (int)n[7];
    float e0, e1;
    float lh = slab(S.nodes + (size_t)l * 9, L.ro, L.inv, e0), rh = slab(S.nodes + (size_t)r * 9, L.ro, L.inv, e1);
    if (cull) { float tc = L.t * 1.00001f; if (e0 > tc) lh = -1.0f; if (e1 > tc) rh = -1.0f; }
    bool hl = lh > 0.0f, hr = rh > 0.0f;
    if (hl && hr) { bool rf = lh > rh; L.stack[L.sp++] = rf ? l : r; L.cur = rf ? r : l; }
    else if (hl) L.cur = l;
    else if (hr) L.cur = r;
    else L.cur = L.stack[--L.sp];
    L.steps++;
}
inline int leafCount(const Scene& S, int node) { return (int)S.nodes[node * 9 + 7]; }
inline void testTri(const Scene& S, Lane& L, int node, int i)
{
    int first = (int)S.nodes[node * 9 + 6];
    const int32_t* vi = S.vi + (size_t)(first + i) * 3;
    const float* p0 = S.verts + (size_t)vi[0] * 4; const float* p1 = S.verts + (size_t)vi[1] * 4; const float* p2 = S.verts + (size_t)vi[2] * 4;
    V3 v0 = {p0[0], p0[1], p0[2]}, e0 = V3{p1[0], p1[1], p1[2]} - v0, e1 = V3{p2[0], p2[1], p2[2]} - v0;
    V3 pv = cross(L.rd, e1); float det = dot(e0, pv);
    V3 tv = L.ro - v0; V3 qv = cross(tv, e0);
    float ux = dot(tv, pv) / det, uy = dot(L.rd, qv) / det, uz = dot(e1, qv) / det, uw = 1.0f - ux - uy;
    if (ux >= 0.0f && uy >= 0.0f && uz >= 0.0f && uw >= 0.0f && uz < L.t)
    {
        if (L.anyHit) { L.occluded = true; L.done = true; L.prim = first + i; }      // any-hit: first accepted hit ends the ray (t stays maxDist)
        else { L.t = uz; L.prim = first + i; }
    }
}
inline void enterInst(const Scene& S, Lane& L)
{
    const float* n = S.nodes + (size_t)L.cur * 9;
    int inst = -(int)n[8] - 1;
    const float* M = S.invT + (size_t)inst * 16;
    auto xf = [&](V3 p, float w) { return V3{p.x * M[0] + p.y * M[4] + p.z * M[8] + w * M[12], p.x * M[1] + p.y * M[5] + p.z * M[9] + w * M[13],
                                           p.x * M[2] + p.y * M[6] + p.z * M[10] + w * M[14]}; };
    L.ro = xf(L.o, 1.0f); L.rd = xf(L.d, 0.0f);
    L.inv = {1.0f / L.rd.x, 1.0f / L.rd.y, 1.0f / L.rd.z};
    L.stack[L.sp++] = -1;
    L.cur = (int)n[6];
    L.inBlas = true;
    L.enteredInst = true;
}
inline void popMarker(Lane& L)
{
    if (!L.inBlas) { L.done = true; return; }
    L.inBlas = false;
    L.cur = L.stack[--L.sp];
    L.ro = L.o; L.rd = L.d; L.inv = L.invW;
}

struct Stats { double slots = 0, laneSlots = 0, innerSlots = 0, innerLane = 0, triSlots = 0, triLane = 0, otherSlots = 0, wavefronts = 0; long rays = 0, innerSteps = 0; };

// L1 wavefronts of one warp-wide fetch: distinct 128-byte lines among the lanes' addresses (a 64-byte node is 4 x LDG.128, a 48-byte
// triangle record 3 x LDG.128; each of those instructions is replayed once per distinct line).
inline int distinctLines(const long* addr, int n)
{
    long lines[32]; int m = 0;
    for (int i = 0; i < n; i++) { long l = addr[i] >> 7; bool seen = false; for (int j = 0; j < m; j++) if (lines[j] == l) { seen = true; break; } if (!seen) lines[m++] = l; }
    return m;
}

bool g_stopAtInst = false;      // phase-1 model: a ray that reaches an instance leaf stops there (it is handed to phase 2)
void runWarp(const Scene& S, const float* rays, const float* maxDist, int n, int policy, bool cull, float* outT, int32_t* outPrim, Stats& st)
{
    Lane L[32];
    for (int i = 0; i < n; i++) { L[i].begin(S, rays + (size_t)i * 6); if (maxDist) { L[i].anyHit = true; L[i].t = maxDist[i]; } }
    auto anyLive = [&] { for (int i = 0; i < n; i++) if (!L[i].done) return true; return false; };
    while (anyLive())
    {
        st.slots += C_ROUND; st.otherSlots += C_ROUND;
        if (policy == 2)
        {   // if-if: one item per lane; kinds issued one after the other
            int nInner = 0, nInst = 0, nMark = 0, maxTri = 0, triLane = 0;
            Kind k[32];
            for (int i = 0; i < n; i++) if (!L[i].done) k[i] = kindOf(S, L[i].cur);
            for (int i = 0; i < n; i++)
            {
                if (L[i].done) continue;
                if (k[i] == K_INNER) { stepInner(S, L[i], cull); nInner++; }
                else if (k[i] == K_LEAF) { int c = leafCount(S, L[i].cur); for (int j = 0; j < c; j++) testTri(S, L[i], L[i].cur, j); maxTri = std::max(maxTri, c); triLane += c; L[i].cur = L[i].stack[--L[i].sp]; }
                else if (k[i] == K_INST) { enterInst(S, L[i]); nInst++; }
                else { popMarker(L[i]); nMark++; }
            }
            if (nInner) { st.slots += C_INNER; st.innerSlots += C_INNER; st.innerLane += (double)C_INNER * nInner; st.laneSlots += (double)C_INNER * nInner; }
            if (maxTri) { double c = C_LEAF_SETUP + (double)C_TRI * maxTri; st.slots += c; st.triSlots += c; st.triLane += (double)C_TRI * triLane; st.laneSlots += (double)C_TRI * triLane; }
            if (nInst) { st.slots += C_INST; st.otherSlots += C_INST; st.laneSlots += (double)C_INST * nInst; }
            if (nMark) { st.slots += C_POP; st.otherSlots += C_POP; st.laneSlots += (double)C_POP * nMark; }
            continue;
        }
        // inner phase
        while (true)
        {
            int cnt = 0; long addr[32];
            for (int i = 0; i < n; i++)
            {
                if (L[i].done) continue;
                Kind k = kindOf(S, L[i].cur);
                if (policy == 1 && k == K_LEAF && L[i].pending == -2)
                {   // park the leaf, continue with the next stack entry (free: folded into the step that found the leaf)
                    L[i].pending = L[i].cur; L[i].cur = L[i].stack[--L[i].sp];
                    k = kindOf(S, L[i].cur);
                }
                if (k == K_INNER) { addr[cnt] = (long)L[i].cur * 64; stepInner(S, L[i], cull); cnt++; }
            }
            if (!cnt) break;
            st.wavefronts += 4.0 * distinctLines(addr, cnt);
            st.slots += C_INNER; st.innerSlots += C_INNER; st.innerLane += (double)C_INNER * cnt; st.laneSlots += (double)C_INNER * cnt; st.innerSteps += cnt;
        }
        // leaf phase: parked leaf first, then the current one (reference order)
        auto leafPass = [&](bool parked) {
            int maxTri = 0, triLane = 0; long triAddr[4][32]; int triN[4] = {0, 0, 0, 0};
            for (int i = 0; i < n; i++)
            {
                if (L[i].done) continue;
                int node = -2;
                if (parked) { node = L[i].pending; L[i].pending = -2; }
                else if (kindOf(S, L[i].cur) == K_LEAF) node = L[i].cur;
                if (node < 0) continue;
                int c = leafCount(S, node);
                for (int j = 0; j < c && !L[i].done; j++) testTri(S, L[i], node, j);
                for (int j = 0; j < c && j < 4; j++) triAddr[j][triN[j]++] = ((long)S.nodes[node * 9 + 6] + j) * 48;
                maxTri = std::max(maxTri, c); triLane += c;
                if (!parked && !L[i].done) L[i].cur = L[i].stack[--L[i].sp];
            }
            for (int j = 0; j < 4; j++) if (triN[j]) st.wavefronts += 3.0 * distinctLines(triAddr[j], triN[j]);
            if (maxTri) { double c = C_LEAF_SETUP + (double)C_TRI * maxTri; st.slots += c; st.triSlots += c; st.triLane += (double)C_TRI * triLane; st.laneSlots += (double)C_TRI * triLane; }
        };
        if (policy == 1) leafPass(true);
        leafPass(false);
        // instance / marker phase (items that were current when the inner phase ended and are not leaves)
        int nInst = 0, nMark = 0;
        for (int i = 0; i < n; i++)
        {
            if (L[i].done) continue;
            Kind k = kindOf(S, L[i].cur);
            if (k == K_INST) { if (g_stopAtInst) { L[i].done = true; L[i].enteredInst = true; continue; } enterInst(S, L[i]); nInst++; }
            else if (k == K_MARK) { popMarker(L[i]); nMark++; }
        }
        if (nInst) { st.slots += C_INST; st.otherSlots += C_INST; st.laneSlots += (double)C_INST * nInst; }
        if (nMark) { st.slots += C_POP; st.otherSlots += C_POP; st.laneSlots += (double)C_POP * nMark; }
    }
    for (int i = 0; i < n; i++) { outT[i] = L[i].t; outPrim[i] = L[i].anyHit ? ((L[i].occluded ? 1 : 0) | (L[i].enteredInst ? 2 : 0) | (int)(L[i].steps << 2)) : L[i].prim; }
    st.rays += n;
}

}  // namespace

extern "C" void simd_sim_stop_at_inst(int on) { g_stopAtInst = on != 0; }

extern "C" void simd_sim(const float* nodes, int numNodes, int top, const int32_t* vi, const float* verts, const float* invT,
                         const float* rays, int64_t n, int policy, int cull, int warp, float* outT, int32_t* outPrim, double* out9, const float* maxDist)
{
    Scene S{nodes, numNodes, top, vi, verts, invT};
    Stats total;
#pragma omp parallel
    {
        Stats st;
#pragma omp for schedule(dynamic, 64)
        for (int64_t b = 0; b < n; b += warp)
            runWarp(S, rays + (size_t)b * 6, maxDist ? maxDist + b : nullptr, (int)std::min<int64_t>(warp, n - b), policy, cull != 0, outT + b, outPrim + b, st);
#pragma omp critical
        {
            total.slots += st.slots; total.laneSlots += st.laneSlots; total.innerSlots += st.innerSlots; total.innerLane += st.innerLane;
            total.triSlots += st.triSlots; total.triLane += st.triLane; total.otherSlots += st.otherSlots; total.rays += st.rays; total.innerSteps += st.innerSteps; total.wavefronts += st.wavefronts;
        }
    }
    out9[0] = total.slots; out9[1] = total.laneSlots; out9[2] = total.innerSlots; out9[3] = total.innerLane; out9[4] = total.triSlots;
    out9[5] = total.triLane; out9[6] = total.otherSlots; out9[7] = (double)total.innerSteps; out9[8] = total.wavefronts;
}
