// ptb_kernels.cu — sm_100a kernels of the wavefront path tracer and their launchers.
//
// Pipeline per wave (S sample passes of a pixel rectangle, all paths resident in HBM as SoA float4 state, slots in block-major order: ptb_device.cuh):
//   k_trace_primary  tile.glsl:41-68 + closest_hit.glsl   camera rays generated in registers and traced at once (k_camera: preview target only)
//   loop over bounces:
//     k_sort_tile_local                   tile-local grouping of queue entries: by direction cell before the bounce-1 trace, by material before a shade pass,
//                                         by sampled light before the first k_shadow
//     k_trace     closest_hit.glsl     persistent warps, dynamic 32-ray fetch, shared-memory stacks, 64-byte node fetches; finishes paths that end at the hit
//     k_shade     pathtrace.glsl       hit attributes, GetMaterial, emission/MIS, media, alpha, NEE sample + DisneyEval,
//                                      DisneySample, Russian roulette; next / shadow queues (warp-ballot compaction, or records by path slot)
//     k_shadow    anyhit.glsl          any-hit for the queued NEE rays (4-wide hierarchy), adds the unoccluded contributions
//     k_transmit  pathtrace.glsl:119   EvalTransmittance for the queued NEE rays of volume-MIS scenes without BLEND materials
//   k_accumulate  tile.glsl:70-74      sum of the wave's samples into the running-sum buffer
//   k_tonemap     tonemap.glsl         on readback
#include "ptb_device.cuh"
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>

using namespace ptb;

// Launch bookkeeping is per context (LaunchCfg::launches) and the occupancy / shared-memory attributes are per device
// (ptbk_configure_device): one process may hold one context per GPU.
#define COUNT_LAUNCH(c, n) do { if ((c).launches) *(c).launches += (n); } while (0)

#define OPT(F, bit) (((F).features & (bit)) != 0u)
enum {
    O_ENVMAP = 1 << 0, O_LIGHTS = 1 << 1, O_RR = 1 << 2, O_UNIFORM = 1 << 3, O_GLNORMAL = 1 << 4, O_HIDE = 1 << 5, O_BG = 1 << 6,
    O_TRANSPBG = 1 << 7, O_ALPHA = 1 << 8, O_MOLLIFY = 1 << 9, O_MEDIUM = 1 << 10, O_VOLMIS = 1 << 11
};

constexpr int TRACE_THREADS = 256;
#ifndef PTB_TRACE_MINB
#define PTB_TRACE_MINB 5
#endif
constexpr int TRACE_MIN_BLOCKS = PTB_TRACE_MINB;     // 48 registers: measured 772 spp/s vs 720 (3 blocks, 72 regs), 767 (4), 753 (6)
constexpr int SHADE_THREADS = 128;
#ifndef PTB_SHADE_LATER_BLOCKS
#define PTB_SHADE_LATER_BLOCKS 4      // resident k_shade<0> blocks per SM for the bounces after the first (latency-bound: scattered, material-sorted state)
#endif

// slot <-> pixel mapping: groupToPixel / slotToPixel / slotOfSample in ptb_device.cuh (compiled for the host by the test harness, too)

__global__ void __launch_bounds__(256) k_camera(DevScene S, FrameParams F, WaveParams W, PathState P, uint32_t* ctr0)
{
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t base = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; base < W.nSlots; base += gridDim.x * blockDim.x)
    {
        uint32_t slot = base + lane;
        bool live = false;
        if (slot < W.nSlots)
        {
            int s, px, py;
            live = slotToPixel(W, slot, s, px, py);
            if (live)
            {
                Rng rng; float3 ro, rd;
                cameraRay(F, W, W.x0 + px, W.y0 + py, W.firstSample + s * W.sampleStride, rng, ro, rd);
                P.rayO[slot] = make_float4(ro.x, ro.y, ro.z, 0.0f);
                P.rayD[slot] = make_float4(rd.x, rd.y, rd.z, __uint_as_float(0u));
                P.rng[slot] = rng.s;          // throughput (1) / radiance (0) / alpha (1) are implied in the first shade iteration
                if (F.general)
                {
                    P.med[slot] = make_float4(0.f, 0.f, __int_as_float(0), __int_as_float(0));
                    P.medCol[slot] = make_float4(0.f, 0.f, 0.f, 0.f);
                    P.prevUV[slot] = make_float2(0.f, 0.f);
                }
            }
        }
        unsigned m = __ballot_sync(0xffffffffu, live);
        if (m)
        {
            uint32_t b = 0;
            if (lane == 0) b = atomicAdd(&ctr0[CTR_NPATHS], (uint32_t)__popc(m));
            b = __shfl_sync(0xffffffffu, b, 0);
            if (live) P.queue[0][b + __popc(m & ((1u << lane) - 1u))] = slot;
        }
    }
}

// ------------------------------------------------------------------ closest-hit trace ---------------------------
extern __shared__ uint32_t g_stackSmem[];

// Shading sort key: 0 = miss, 1 = analytic light, 2 + matID = triangle of that material (material-sorted shading).
__device__ __forceinline__ uint32_t shadeKey(const DevScene& S, int hitInst)
{
    if (hitInst == -1) return 0u;
    if (hitInst <= -2) return 1u;
    return 2u + (uint32_t)__float_as_int(__ldg(S.instTrav + (size_t)hitInst * 4 + 1).w);
}

// lightSample.pdf / emission of a light hit (closest_hit.glsl:56-62, 72-83)
__device__ __forceinline__ void lightHitInfo(const DevScene& S, int idx, float3 ro, float3 rd, float t, float& pdf, float3& emission)
{
    const float4* p = S.lightsPre + (size_t)idx * 8;
    float4 a = __ldg(p), b = __ldg(p + 1), e = __ldg(p + 4);
    emission = f3(b);
    float area = b.w;
    if (a.w == 0.0f)
    {
        float cosTheta = dot(-rd, f3(e));
        pdf = (t * t) / (area * cosTheta);
    }
    else
    {
        float3 hitPt = ro + t * rd;
        float cosTheta = dot(-rd, normalize(hitPt - f3(a)));
        pdf = (t * t) / (area * cosTheta * 0.5f);
    }
}

// Paths that end at this closest hit without shading work — a miss (pathtrace.glsl:305-339) or an analytic-light hit (:341-364) — are finished by the trace
// kernel itself in the specialisations without media (F.general <= 1): radiance += MIS weight x emission x throughput, same expressions as shadePath.  On hyperion
// 43 % of the closest-hit rays end this way (ibl_spheres 43 %, instancing 35 %); they never enter a shade queue (key = hole).
enum { TRACE_FINISH = 1, TRACE_FIRST_ITER = 2, TRACE_MARK_DEAD = 4 };     // MARK_DEAD: the next shade pass is not fed by the sorter (which drops finished paths)
template <bool GEN>
__device__ __forceinline__ void finishPath(const DevScene& S, const FrameParams& F, const PathState& P, uint32_t p, int hi, float3 ro, float3 rd, float t, int depth,
                                           float prevPdf, bool firstIter)
{
    if (hi == -1 && !firstIter && !OPT(F, O_UNIFORM) && !(GEN && OPT(F, O_ENVMAP))) return;      // a later-bounce miss into nothing: the radiance sum is already in place
    const float4 thr4 = firstIter ? make_float4(1.f, 1.f, 1.f, 0.f) : P.thr[p], rad4 = firstIter ? make_float4(0.f, 0.f, 0.f, 1.f) : P.rad[p];
    float3 thr = f3(thr4), rad = f3(rad4);
    float alpha = rad4.w;
    if (hi == -1)
    {
        if (OPT(F, O_BG) || OPT(F, O_TRANSPBG)) { if (depth == 0) alpha = 0.0f; }
        if (!OPT(F, O_HIDE) || depth > 0)
        {
            if (OPT(F, O_UNIFORM)) rad += f3(F.uniformLightCol[0], F.uniformLightCol[1], F.uniformLightCol[2]) * thr;
            else if (GEN && OPT(F, O_ENVMAP))
            {
                float4 e = EvalEnvMap(S, F, rd);
                float misWeight = 1.0f;
                if (depth > 0) misWeight = PowerHeuristic(prevPdf, e.w);
                if (misWeight > 0) rad += misWeight * f3(e) * thr * F.envMapIntensity;
            }
        }
    }
    else
    {
        if (GEN)
        {   // emission of the stale matID / texCoord: material 0 and uv (0,0) before the first surface hit
            int prevMat = 0; float2 prevUV = make_float2(0.f, 0.f);
            if (!firstIter) { prevMat = __float_as_int(P.med[p].w); prevUV = P.prevUV[p]; }
            float3 em = f3(__ldg(S.materials + (size_t)prevMat * 8 + 1));
            float etex = __ldg(S.materials + (size_t)prevMat * 8 + 6).w;
            if (etex >= 0.f) em = vpow(f3(sampleTexArray(S, prevUV, (float)(int)etex)), 2.2f);
            rad += em * thr;
        }
        float lpdf; float3 lem;
        lightHitInfo(S, -(hi + 2), ro, rd, t, lpdf, lem);
        float misWeight = 1.0f;
        if (depth > 0) misWeight = PowerHeuristic(prevPdf, lpdf);
        rad += misWeight * lem * thr;
    }
    P.rad[p] = make_float4(rad.x, rad.y, rad.z, alpha);
}
// the env-map / texture evaluation stays out of line: the traversal loop keeps its 48 registers
__device__ __noinline__ void finishPathGeneral(const DevScene& S, const FrameParams& F, const PathState& P, uint32_t p, int hi, float3 ro, float3 rd, float t, int depth,
                                               float prevPdf, bool firstIter)
{
    finishPath<true>(S, F, P, p, hi, ro, rd, t, depth, prevPdf, firstIter);
}
__device__ __forceinline__ void finishInTrace(const DevScene& S, const FrameParams& F, const PathState& P, uint32_t p, int hi, float3 ro, float3 rd, float t, int depth,
                                              float prevPdf, bool firstIter)
{
    if (F.general) finishPathGeneral(S, F, P, p, hi, ro, rd, t, depth, prevPdf, firstIter);
    else finishPath<false>(S, F, P, p, hi, ro, rd, t, depth, prevPdf, firstIter);
}

// Persistent warps fetch 32 consecutive queue entries at a time (one atomic per fetch).  A lane-granular refill of finished
// lanes was measured and rejected: it breaks the screen-space coherence of the warp (primary rays 0.68 -> 1.17 ms per 8.3 M rays)
// and did not help the incoherent bounces (1.33 -> 1.35 ms), see DESIGN.md.
// CAM: first bounce of a wave.  There is no queue yet: the 32 entries of a fetch are the path slots of one 8x4 pixel block, the camera
// ray (tile.glsl:41-68) is generated in registers and traced at once; the ray, the RNG state and the initial queue (slot, or a hole
// for the off-image pixels of a padded block) are written once from here, for the first shade pass.
template <bool CULL, bool CAM>
__device__ __forceinline__ void traceLoop(const DevScene& S, const FrameParams& F, const WaveParams& W, const PathState& P, const uint32_t* __restrict__ queue, uint32_t n,
                                          uint32_t* fetchCtr, int lightsFromDepth, uint32_t* __restrict__ keys, uint32_t* hist, uint32_t holeKey, uint32_t tflags)
{
    const uint32_t lane = threadIdx.x & 31u;
    SmemStack stk(g_stackSmem + threadIdx.x, (int)blockDim.x);
    const bool lights = OPT(F, O_LIGHTS);
    uint32_t blocksX = 1, blocksPerSample = 1;
    if (CAM) { blocksX = (uint32_t)W.vw >> 3; blocksPerSample = blocksX * ((uint32_t)W.vh >> 2); }
    while (true)
    {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(fetchCtr, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) break;
        const uint32_t i = base + lane;
        uint32_t pq;
        float3 o, d; int depth = 0; float prevPdf = 0.0f;
        if (CAM)
        {   // slot -> (sample, 8x4 block, pixel): the divisions are per fetch, not per pixel (n is a multiple of 32)
            int s, px, py;
            groupToPixel(W, base >> 5, lane, blocksX, blocksPerSample, s, px, py);
            const bool live = px < W.rw && py < W.rh;
            pq = live ? i : 0xffffffffu;
            P.queue[0][i] = pq;
            if (!live) P.hitInst[i] = PTB_HIT_DEAD;        // off-image pixel of a padded block: the streaming first shade pass skips it without reading the queue
            if (live)
            {
                Rng rng;
                cameraRay(F, W, W.x0 + px, W.y0 + py, W.firstSample + s * W.sampleStride, rng, o, d);
                P.rayO[i] = make_float4(o.x, o.y, o.z, 0.0f);
                P.rayD[i] = make_float4(d.x, d.y, d.z, __uint_as_float(0u));
                P.rng[i] = rng.s;             // throughput (1) / radiance (0) / alpha (1) are implied in the first shade iteration
                if (F.general)
                {
                    P.med[i] = make_float4(0.f, 0.f, __int_as_float(0), __int_as_float(0));
                    P.medCol[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                    P.prevUV[i] = make_float2(0.f, 0.f);
                }
            }
        }
        else
        {
            pq = i < n ? queue[i] : 0xffffffffu;
            if (pq != 0xffffffffu)
            {
                const float4 o4 = P.rayO[pq], d4 = P.rayD[pq];
                o = f3(o4); d = f3(d4); prevPdf = o4.w;
                depth = (int)(short)(__float_as_uint(d4.w) & 0xffffu);
            }
        }
        if (i < n && pq == 0xffffffffu) { if (keys) keys[i] = holeKey; }      // hole (dead path of a slot-ordered queue / off-image pixel): nothing to trace
        else if (i < n)
        {
            const uint32_t p = pq;
            HitRec h; h.t = PTB_INF; h.prim = -1; h.inst = -1; h.light = -1; h.bu = h.bv = 0.f;
            float t = PTB_INF;
            if (lights && depth >= lightsFromDepth) closestLights(S, o, d, t, h.light);         // OPT_HIDE_EMITTERS: lights only at depth > 0
            traverse<false, false, CULL>(S, o, d, t, stk, h, NoAlpha());
            const int hi = (h.inst >= 0) ? h.inst : (h.light >= 0 && h.t < PTB_INF ? -(h.light + 2) : -1);
            const bool finished = (tflags & TRACE_FINISH) && hi < 0;
            if (finished)
            {
                const bool firstIter = CAM || (tflags & TRACE_FIRST_ITER);
                finishInTrace(S, F, P, p, hi, o, d, h.t, depth, prevPdf, firstIter);
                if (tflags & TRACE_MARK_DEAD) P.hitInst[p] = PTB_HIT_DEAD;       // an unsorted shade pass (identity queue of the first bounce) skips the slot on this mark
            }
            else
            {
                P.hit[p] = make_float4(h.t, h.bu, h.bv, __int_as_float(h.prim));
                P.hitInst[p] = hi;
            }
            if (keys)
            {   // histogram of shading keys, one atomic per distinct key per (converged part of the) warp
                const uint32_t key = finished ? holeKey : shadeKey(S, hi);
                keys[i] = key;
                if (hist)
                {   // global counting sort only (PTB_SORT=1/2); the tile-local sorter builds its own histograms
                    const unsigned peers = __match_any_sync(__activemask(), key);
                    if (lane == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&hist[key], (uint32_t)__popc(peers));
                }
            }
        }
    }
}

__global__ void __launch_bounds__(TRACE_THREADS, TRACE_MIN_BLOCKS) k_trace(DevScene S, FrameParams F, PathState P, const uint32_t* __restrict__ queue,
                                                          const uint32_t* __restrict__ countPtr, uint32_t* fetchCtr, int lightsFromDepth, DevStats* stats,
                                                          uint32_t* __restrict__ keys, uint32_t* hist, uint32_t nOverride, uint32_t holeKey, uint32_t tflags)
{
    const uint32_t live = *countPtr;                     // rays actually traced (statistics)
    const uint32_t n = nOverride ? nOverride : live;     // queue length: the slot count when the queue is slot-ordered with holes
    if (blockIdx.x == 0 && threadIdx.x == 0 && live) atomicAdd(&stats->pathSegments, (unsigned long long)live);
    const WaveParams W{};
    if (F.cullBoxes) traceLoop<true, false>(S, F, W, P, queue, n, fetchCtr, lightsFromDepth, keys, hist, holeKey, tflags);
    else traceLoop<false, false>(S, F, W, P, queue, n, fetchCtr, lightsFromDepth, keys, hist, holeKey, tflags);
}

// First bounce: camera-ray generation fused into the closest-hit trace (see traceLoop<., true>).  liveCount = rays generated (host-known).
__global__ void __launch_bounds__(TRACE_THREADS, TRACE_MIN_BLOCKS) k_trace_primary(DevScene S, FrameParams F, WaveParams W, PathState P, uint32_t* ctr0, int lightsFromDepth,
                                                                  DevStats* stats, uint32_t* __restrict__ keys, uint32_t* hist, uint32_t liveCount, uint32_t holeKey, uint32_t tflags)
{
    if (blockIdx.x == 0 && threadIdx.x == 0) { atomicAdd(&stats->pathSegments, (unsigned long long)liveCount); ctr0[CTR_NPATHS] = liveCount; }
    if (F.cullBoxes) traceLoop<true, true>(S, F, W, P, nullptr, W.nSlots, &ctr0[CTR_FETCH_TRACE], lightsFromDepth, keys, hist, holeKey, tflags);
    else traceLoop<false, true>(S, F, W, P, nullptr, W.nSlots, &ctr0[CTR_FETCH_TRACE], lightsFromDepth, keys, hist, holeKey, tflags);
}

// Exclusive scan of the key histogram into bucket cursors (one warp); clears the histogram for the next bounce.
__global__ void k_sort_scan(uint32_t* hist, uint32_t* cursor, int numKeys)
{
    const int lane = threadIdx.x;
    const int per = (numKeys + 31) / 32;
    uint32_t sum = 0;
    for (int k = lane * per; k < min(numKeys, (lane + 1) * per); k++) sum += hist[k];
    uint32_t incl = sum;
    for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    uint32_t run = incl - sum;
    for (int k = lane * per; k < min(numKeys, (lane + 1) * per); k++) { uint32_t h = hist[k]; cursor[k] = run; run += h; hist[k] = 0; }
}

// Counting-sort scatter: queue entries -> key buckets.  Each 256-thread block ranks a tile of 2048 entries in shared memory
// (warp match + shared atomics) and reserves its bucket ranges with ONE global atomic per key per tile, so the popular keys
// ("miss", the floor material) do not serialise on a single L2 address.
constexpr int SORT_ITEMS = 8;
__global__ void __launch_bounds__(256) k_sort_scatter(const uint32_t* __restrict__ queue, const uint32_t* __restrict__ keys, const uint32_t* __restrict__ countPtr,
                                                       uint32_t* cursor, uint32_t* __restrict__ sorted, int numKeys)
{
    extern __shared__ uint32_t sm[];
    uint32_t* hist = sm; uint32_t* base = sm + numKeys;
    const uint32_t n = *countPtr;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t tileSize = 256u * SORT_ITEMS;
    for (uint32_t tile = blockIdx.x * tileSize; tile < n; tile += gridDim.x * tileSize)
    {
        for (int k = threadIdx.x; k < numKeys; k += 256) hist[k] = 0;
        __syncthreads();
        uint32_t key[SORT_ITEMS], rank[SORT_ITEMS];
#pragma unroll
        for (int j = 0; j < SORT_ITEMS; j++)
        {
            const uint32_t i = tile + j * 256u + threadIdx.x;
            key[j] = i < n ? keys[i] : 0xffffffffu;
            const unsigned peers = __match_any_sync(0xffffffffu, key[j]);
            const int leader = __ffs(peers) - 1;
            uint32_t off = 0;
            if ((int)lane == leader && key[j] != 0xffffffffu) off = atomicAdd(&hist[key[j]], (uint32_t)__popc(peers));
            rank[j] = __shfl_sync(0xffffffffu, off, leader) + __popc(peers & ((1u << lane) - 1u));
        }
        __syncthreads();
        for (int k = threadIdx.x; k < numKeys; k += 256) { uint32_t c = hist[k]; if (c) base[k] = atomicAdd(&cursor[k], c); }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < SORT_ITEMS; j++)
        {
            const uint32_t i = tile + j * 256u + threadIdx.x;
            if (key[j] != 0xffffffffu) sorted[base[key[j]] + rank[j]] = queue[i];
        }
        __syncthreads();
    }
}
// Tile-local variant: entries are grouped by key only inside their own tile of 2048 queue entries.  Warps become material-coherent
// while every warp still touches path slots from one 2048-entry window, so the scattered state fetches stay sector/page local.
// queue == nullptr: the entries are the indices themselves (slot order).  Entries whose key is holeKey come out as 0xffffffff ("hole") at the
// end of their tile; keys above numKeys-1 are clamped.  nOverride != 0 replaces *countPtr as the number of entries.
#ifndef PTB_SORT_MINB
#define PTB_SORT_MINB 6      // resident sorter blocks per SM (40 registers): 1.26 -> 1.07 ms of sorting per hyperion step against 4
#endif
#ifndef PTB_SORT_MANY_KEYS
#define PTB_SORT_MANY_KEYS 100
#endif
__global__ void __launch_bounds__(256, PTB_SORT_MINB) k_sort_tile_local(const uint32_t* __restrict__ queue, const uint32_t* __restrict__ keys, const uint32_t* __restrict__ countPtr,
                                                          uint32_t* __restrict__ sorted, int numKeys, int holeKey, uint32_t nOverride)
{
    extern __shared__ uint32_t sm[];
    uint32_t* hist = sm; uint32_t* base = sm + numKeys;
    const uint32_t n = nOverride ? nOverride : *countPtr;
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t tileSize = 256u * SORT_ITEMS;
    const bool manyKeys = numKeys > PTB_SORT_MANY_KEYS;
    for (uint32_t tile = blockIdx.x * tileSize; tile < n; tile += gridDim.x * tileSize)
    {
        for (int k = threadIdx.x; k < numKeys; k += 256) hist[k] = 0;
        __syncthreads();
        uint32_t key[SORT_ITEMS], rank[SORT_ITEMS];
#pragma unroll
        for (int j = 0; j < SORT_ITEMS; j++)
        {
            const uint32_t i = tile + j * 256u + threadIdx.x;
            key[j] = i < n ? min(keys[i], (uint32_t)(numKeys - 1)) : 0xffffffffu;
            if (manyKeys)
            {   // many classes (16x16 direction cells): a warp's 32 keys are mostly distinct — one shared-memory atomic each beats the warp match; only the hole
                // key (ended paths, up to half of a tile) is aggregated
                const bool hole = (int)key[j] == holeKey;
                const unsigned hm = __ballot_sync(0xffffffffu, hole);
                uint32_t off = 0;
                if (hole) { const int leader = __ffs(hm) - 1; if ((int)lane == leader) off = atomicAdd(&hist[key[j]], (uint32_t)__popc(hm)); off = __shfl_sync(hm, off, leader) + __popc(hm & ((1u << lane) - 1u)); }
                else if (key[j] != 0xffffffffu) off = atomicAdd(&hist[key[j]], 1u);
                rank[j] = off;
                continue;
            }
            const unsigned peers = __match_any_sync(0xffffffffu, key[j]);
            const int leader = __ffs(peers) - 1;
            uint32_t off = 0;
            if ((int)lane == leader && key[j] != 0xffffffffu) off = atomicAdd(&hist[key[j]], (uint32_t)__popc(peers));
            rank[j] = __shfl_sync(0xffffffffu, off, leader) + __popc(peers & ((1u << lane) - 1u));
        }
        __syncthreads();
        if (threadIdx.x < 32)
        {   // exclusive scan of the tile histogram by one warp
            const int per = (numKeys + 31) / 32;
            uint32_t sum = 0;
            for (int k = lane * per; k < min(numKeys, (int)(lane + 1) * per); k++) sum += hist[k];
            uint32_t incl = sum;
            for (int o = 1; o < 32; o <<= 1) { uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= (uint32_t)o) incl += v; }
            uint32_t run = incl - sum;
            for (int k = lane * per; k < min(numKeys, (int)(lane + 1) * per); k++) { base[k] = run; run += hist[k]; }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < SORT_ITEMS; j++)
        {
            const uint32_t i = tile + j * 256u + threadIdx.x;
            if (key[j] != 0xffffffffu) sorted[tile + base[key[j]] + rank[j]] = ((int)key[j] == holeKey) ? 0xffffffffu : (queue ? queue[i] : i);
        }
        __syncthreads();
    }
}
// Fallback for very large key ranges (more materials than fit the shared histogram): direct global atomics.
__global__ void __launch_bounds__(256) k_sort_scatter_global(const uint32_t* __restrict__ queue, const uint32_t* __restrict__ keys,
                                                              const uint32_t* __restrict__ countPtr, uint32_t* cursor, uint32_t* __restrict__ sorted)
{
    const uint32_t n = *countPtr;
    const uint32_t lane = threadIdx.x & 31u;
    for (uint32_t b0 = (blockIdx.x * blockDim.x + threadIdx.x) & ~31u; b0 < n; b0 += gridDim.x * blockDim.x)
    {
        const uint32_t i = b0 + lane;
        const bool live = i < n;
        const uint32_t key = live ? keys[i] : 0xffffffffu;
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        uint32_t b = 0;
        const int leader = __ffs(peers) - 1;
        if (live && (int)lane == leader) b = atomicAdd(&cursor[key], (uint32_t)__popc(peers));
        b = __shfl_sync(0xffffffffu, b, leader);
        if (live) sorted[b + __popc(peers & ((1u << lane) - 1u))] = queue[i];
    }
}

// ------------------------------------------------------------------ surface state -------------------------------
struct Surf
{
    float3 fhp, normal, ffnormal, tangent, bitangent; float2 uv; int matID; float hitDist;
};

// closest_hit.glsl:220-263 for a triangle hit (slot = leaf-ref slot, inst = instance)
__device__ __forceinline__ void triangleSurface(const DevScene& S, int slot, int inst, float bu, float bv, float3 ro, float3 rd, float t, bool needUV, bool needTangents, Surf& sf)
{
    const float bw = 1.0f - bu - bv;       // uvt.w; bary = uvt.wxy
    // the three vertex normals (and texture coordinates) of this leaf-ref slot, gathered at upload (closest_hit.glsl:226-235 reads them through vertexIndicesTex)
    const float4* ts = S.triShade + (size_t)slot * 4;
    const float4 s0 = __ldg(ts), s1 = __ldg(ts + 1), s2 = __ldg(ts + 2);
    const float3 nn0 = f3(s0.x, s0.y, s0.z), nn1 = f3(s0.w, s1.x, s1.y), nn2 = f3(s1.z, s1.w, s2.x);
    float3 normal = normalize(nn0 * bw + nn1 * bu + nn2 * bv);
    const float4* is = S.instShade + (size_t)inst * 8;
    const float4 m0 = __ldg(is + 4), m1 = __ldg(is + 5), m2 = __ldg(is + 6);       // inverse(mat3(transform)) rows
    float3 nw = f3(m0.x * normal.x + m0.y * normal.y + m0.z * normal.z, m1.x * normal.x + m1.y * normal.y + m1.z * normal.z,
                   m2.x * normal.x + m2.y * normal.y + m2.z * normal.z);
    sf.normal = normalize(nw);
    sf.ffnormal = dot(sf.normal, rd) <= 0.0f ? sf.normal : -sf.normal;
    sf.hitDist = t;
    sf.fhp = ro + rd * t;
    sf.matID = __float_as_int(__ldg(S.instTrav + (size_t)inst * 4 + 1).w);
    sf.uv = make_float2(0.f, 0.f);
    sf.tangent = sf.bitangent = f3(0.f);
    if (needUV)
    {
        const float4 s3 = __ldg(ts + 3);
        float2 t0 = make_float2(s2.y, s2.z), t1 = make_float2(s2.w, s3.x), t2 = make_float2(s3.y, s3.z);
        sf.uv = make_float2(t0.x * bw + t1.x * bu + t2.x * bv, t0.y * bw + t1.y * bu + t2.y * bv);
        if (needTangents)
        {
            const int i0 = __ldg(S.vertIndices + (size_t)slot * 3), i1 = __ldg(S.vertIndices + (size_t)slot * 3 + 1), i2 = __ldg(S.vertIndices + (size_t)slot * 3 + 2);
            const float4 v0 = __ldg(S.verticesUVX + i0), v1 = __ldg(S.verticesUVX + i1), v2 = __ldg(S.verticesUVX + i2);
            float3 dp1 = f3(v1) - f3(v0), dp2 = f3(v2) - f3(v0);
            float2 duv1 = make_float2(t1.x - t0.x, t1.y - t0.y), duv2 = make_float2(t2.x - t0.x, t2.y - t0.y);
            float invdet = 1.0f / (duv1.x * duv2.y - duv1.y * duv2.x);
            float3 tg = (dp1 * duv2.y - dp2 * duv1.y) * invdet;
            float3 bt = (dp2 * duv1.x - dp1 * duv2.x) * invdet;
            const float4 d0 = __ldg(is), d1 = __ldg(is + 1), d2 = __ldg(is + 2);    // transform rows: mat3(transform) * v
            sf.tangent = normalize(f3(tg.x * d0.x + tg.y * d1.x + tg.z * d2.x, tg.x * d0.y + tg.y * d1.y + tg.z * d2.y, tg.x * d0.z + tg.y * d1.z + tg.z * d2.z));
            sf.bitangent = normalize(f3(bt.x * d0.x + bt.y * d1.x + bt.z * d2.x, bt.x * d0.y + bt.y * d1.y + bt.z * d2.y, bt.x * d0.z + bt.y * d1.z + bt.z * d2.z));
        }
    }
}

// GetMaterial (pathtrace.glsl:25-115).  prevRoughness = state.mat.roughness of the previous call (mollification).
template <int MODE>
__device__ __forceinline__ void getMaterial(const DevScene& S, const FrameParams& F, Surf& sf, float3 rd, int depth, float prevRoughness, Material& mat, float& eta)
{
    int4 tex;
    materialFromRow(S.materials + (size_t)sf.matID * 8, mat, tex);
    constexpr bool GEN = MODE >= 1;
    if (GEN)
    {
        if (tex.x >= 0)
        {
            float4 col = sampleTexArray(S, sf.uv, (float)tex.x);
            mat.baseColor *= vpow(f3(col), 2.2f);
            mat.opacity *= col.w;
        }
        if (tex.y >= 0)
        {
            float4 mr = sampleTexArray(S, sf.uv, (float)tex.y);
            mat.metallic = mr.z;
            mat.roughness = fmaxf(mr.y * mr.y, 0.001f);
        }
        if (tex.z >= 0)
        {
            float4 tn = sampleTexArray(S, sf.uv, (float)tex.z);
            float3 texNormal = f3(tn);
            if (OPT(F, O_GLNORMAL)) texNormal.y = 1.0f - texNormal.y;
            texNormal = normalize(texNormal * 2.0f - f3(1.0f));
            float3 origNormal = sf.normal;
            sf.normal = normalize(sf.tangent * texNormal.x + sf.bitangent * texNormal.y + sf.normal * texNormal.z);
            sf.ffnormal = dot(origNormal, rd) <= 0.0f ? sf.normal : -sf.normal;
        }
        if (OPT(F, O_MOLLIFY) && depth > 0)
            mat.roughness = fmaxf(mixf(0.0f, prevRoughness, F.roughnessMollificationAmt), mat.roughness);
        if (tex.w >= 0)
            mat.emission = vpow(f3(sampleTexArray(S, sf.uv, (float)tex.w)), 2.2f);
    }
    materialFinish(mat);
    eta = dot(rd, sf.normal) < 0.0f ? (1.0f / mat.ior) : mat.ior;
}

__device__ __forceinline__ bool materialNeedsTangents(const DevScene& S, int matID) { return __ldg(S.materials + (size_t)matID * 8 + 6).z >= 0.f; }

// ------------------------------------------------------------------ inline traversal (RNG-consuming shadow rays) -
struct InlineCounters { unsigned segs, shadows; };

// ClosestHit + hit attributes for EvalTransmittance steps (pathtrace.glsl:128).  Returns hit kind: 0 miss, 1 tri, 2 light.
__device__ __noinline__ int closestFull(const DevScene& S, const FrameParams& F, float3 ro, float3 rd, int depthForLights, Surf& sf, InlineCounters& ic)
{
    LocalStack stk;
    HitRec h; h.t = PTB_INF; h.prim = -1; h.inst = -1; h.light = -1; h.bu = h.bv = 0.f;
    float t = PTB_INF;
    ic.segs++;
    if (OPT(F, O_LIGHTS) && (!OPT(F, O_HIDE) || depthForLights > 0)) closestLights(S, ro, rd, t, h.light);
    traverse<false, false, true>(S, ro, rd, t, stk, h, NoAlpha());
    if (h.t == PTB_INF) return 0;
    if (h.inst < 0) { sf.hitDist = h.t; sf.fhp = ro + rd * h.t; return 2; }
    int matID = __float_as_int(__ldg(S.instTrav + (size_t)h.inst * 4 + 1).w);
    triangleSurface(S, h.prim, h.inst, h.bu, h.bv, ro, rd, h.t, true, materialNeedsTangents(S, matID), sf);
    return 1;
}

struct AlphaRng   // AnyHit's alpha test (anyhit.glsl:118-141) with the path RNG
{
    const DevScene* S; Rng* rng;
    __device__ __forceinline__ bool operator()(int slot, int inst, float ux, float uy) const
    {
        const DevScene& sc = *S;
        int matID = __float_as_int(__ldg(sc.instTrav + (size_t)inst * 4 + 1).w);
        const float4 texIDs = __ldg(sc.materials + (size_t)matID * 8 + 6), ap = __ldg(sc.materials + (size_t)matID * 8 + 7);
        float alpha = 1.0f;
        if (sc.numTextures > 0)
        {
            const int i0 = __ldg(sc.vertIndices + (size_t)slot * 3), i1 = __ldg(sc.vertIndices + (size_t)slot * 3 + 1), i2 = __ldg(sc.vertIndices + (size_t)slot * 3 + 2);
            float uw = 1.0f - ux - uy;
            float2 t0 = make_float2(__ldg(sc.verticesUVX + i0).w, __ldg(sc.normalsUVY + i0).w), t1 = make_float2(__ldg(sc.verticesUVX + i1).w, __ldg(sc.normalsUVY + i1).w),
                   t2 = make_float2(__ldg(sc.verticesUVX + i2).w, __ldg(sc.normalsUVY + i2).w);
            float2 uv = make_float2(t0.x * uw + t1.x * ux + t2.x * uy, t0.y * uw + t1.y * ux + t2.y * uy);
            alpha = sampleTexArray(sc, uv, texIDs.x).w;
        }
        float opacity = ap.x * alpha;
        int alphaMode = (int)ap.y;
        float alphaCutoff = ap.z;
        return !((alphaMode == 2 && opacity < alphaCutoff) || (alphaMode == 1 && rng->rand() > opacity));
    }
};

__device__ __noinline__ bool anyHitInline(const DevScene& S, const FrameParams& F, float3 ro, float3 rd, float maxDist, Rng& rng, InlineCounters& ic)
{
    ic.shadows++;
    if (OPT(F, O_LIGHTS) && anyLights(S, ro, rd, maxDist)) return true;
    LocalStack stk; HitRec h;
    if (OPT(F, O_ALPHA) && !OPT(F, O_MEDIUM))
        return traverse<true, true, true>(S, ro, rd, maxDist, stk, h, AlphaRng{&S, &rng});
    return traverse<true, false, true>(S, ro, rd, maxDist, stk, h, NoAlpha());
}

// EvalTransmittance (pathtrace.glsl:119-155)
__device__ __noinline__ float3 evalTransmittance(const DevScene& S, const FrameParams& F, float3 ro, float3 rd, Rng& rng, InlineCounters& ic)
{
    float3 transmittance = f3(1.0f);
    for (int depth = 0; depth < F.maxDepth; depth++)
    {
        Surf sf;
        int kind = closestFull(S, F, ro, rd, 0, sf, ic);
        if (kind != 1) break;                       // miss or emitter
        Material mat; float eta;
        getMaterial<2>(S, F, sf, rd, 0, 0.f, mat, eta);
        bool alphatest = (mat.alphaMode == 2 && mat.opacity < mat.alphaCutoff) || (mat.alphaMode == 1 && rng.rand() > mat.opacity);
        bool refractive = (1.0f - mat.metallic) * mat.specTrans > 0.0f;
        if (!(alphatest || refractive)) return f3(0.0f);
        if (dot(rd, sf.normal) > 0 && mat.medType != 0)
        {
            float3 color = mat.medType == 1 ? f3(1.0f) - mat.medColor : f3(1.0f);
            transmittance *= vexp(-color * mat.medDensity * sf.hitDist);
        }
        ro = sf.fhp + rd * PTB_EPS;
    }
    return transmittance;
}

// ------------------------------------------------------------------ shade ----------------------------------------
// Cell of direction d in a GxG octahedral map, rows walked boustrophedon so that consecutive cells are neighbours on the sphere.
template <int G = 8>
__device__ __forceinline__ uint32_t octCell(float dx, float dy, float dz)
{
    const float inv = rcpApprox(fabsf(dx) + fabsf(dy) + fabsf(dz));
    float u = dx * inv, v = dy * inv;
    if (dz < 0.f)
    {
        const float uu = (1.0f - fabsf(v)) * (u < 0.f ? -1.0f : 1.0f), vv = (1.0f - fabsf(u)) * (v < 0.f ? -1.0f : 1.0f);
        u = uu; v = vv;
    }
    int iu = (int)((u * 0.5f + 0.5f) * (float)G), iv = (int)((v * 0.5f + 0.5f) * (float)G);
    iu = min(max(iu, 0), G - 1); iv = min(max(iv, 0), G - 1);
    return (uint32_t)(iv * G + ((iv & 1) ? G - 1 - iu : iu));
}

struct ShadowOut { bool valid; float3 o, d, c; float maxDist; uint32_t light; };

// Immediate push (k_shade<1>: env map + lights, two pending records = 20 registers): where a deferred shadow ray is produced it is appended
// to its queue at once by the lanes that are converged there (one atomic per group), instead of being carried to a warp-converged push at
// the end of the iteration.  Queue order does not matter for the result: every path has at most one entry per queue and k_shadow adds it
// to that path's radiance.  Measured on one box: ibl_spheres 944 -> 976 spp/s (with 5 blocks/SM); k_shade<0> (one record, hyperion) loses 2 %
// with it (846 -> 828: the light-NEE queue of bounce 0 is no longer in pixel order), so modes 0 and 2 keep the converged push.
template <int MODE> struct ImmediatePush { static constexpr bool value = (MODE == 1); };
// Slot-ordered shadow queues (first shade pass of a wave, block-major slot order): the record of path slot p is written AT index p and a sort key
// per slot says what it is — the sampled light (queue B, 2..255 lights) or the octahedral cell of the ray direction, or "no ray" (the arrays are preset to the
// hole key) — and the tile-local sorter turns that into an index queue in which the rays of a few neighbouring pixels towards one light / in one direction
// are adjacent: k_shadow's warps then share origin and direction.  keys[which] == nullptr: compacted arrival-order queue (later bounces).
struct SlotShadow { uint32_t* keys[2]; uint32_t lightKeys; };
__device__ __forceinline__ uint32_t shadowKey(const SlotShadow& ss, int which, float3 d, uint32_t light)
{
    return (which == 1 && ss.lightKeys) ? light : octCell(d.x, d.y, d.z);
}
struct ShadowSink { const PathState* P; uint32_t* ctrA; uint32_t* ctrB; uint32_t path; SlotShadow ss; };
__device__ __forceinline__ void pushShadowNow(const ShadowSink& sk, int which, float3 o, float3 d, float maxDist, float3 c, uint32_t light)
{
    if (sk.ss.keys[which])
    {
        const uint32_t k = sk.path;
        sk.P->shO[which][k] = make_float4(o.x, o.y, o.z, maxDist);
        sk.P->shD[which][k] = make_float4(d.x, d.y, d.z, __uint_as_float(k));
        sk.P->shC[which][k] = make_float4(c.x, c.y, c.z, 0.f);
        sk.ss.keys[which][k] = shadowKey(sk.ss, which, d, light);
        return;
    }
    const unsigned m = __activemask();
    const uint32_t lane = threadIdx.x & 31u;
    const int leader = __ffs(m) - 1;
    uint32_t b = 0;
    if ((int)lane == leader) b = atomicAdd(which ? sk.ctrB : sk.ctrA, (uint32_t)__popc(m));
    b = __shfl_sync(m, b, leader);
    const uint32_t k = b + __popc(m & ((1u << lane) - 1u));
    sk.P->shO[which][k] = make_float4(o.x, o.y, o.z, maxDist);
    sk.P->shD[which][k] = make_float4(d.x, d.y, d.z, __uint_as_float(sk.path));
    sk.P->shC[which][k] = make_float4(c.x, c.y, c.z, 0.f);
}

// DirectLight (pathtrace.glsl:158-283).  Deferred mode: fills sa (env) / sb (light) with contribution*throughput.
// Inline mode (shadow rays draw from the path RNG): traces here and returns Ld.
template <int MODE>
__device__ __forceinline__ float3 directLight(const DevScene& S, const FrameParams& F, float3 rd, const Surf& sf, const Material& mat, float eta, const ShadeFrame& fr,
                                              bool isSurface, float medAniso, float3 thr, Rng& rng, ShadowOut& sa, ShadowOut& sb, InlineCounters& ic, const ShadowSink& sink)
{
    float3 Ld = f3(0.0f);
    const float3 scatterPos = sf.fhp + sf.normal * PTB_EPS;
    constexpr bool GEN = MODE >= 1, FULL = MODE >= 2;       // FULL: media, alpha test; MODE 2 also traces the RNG-consuming shadow rays inline
    const bool volMis = FULL && OPT(F, O_MEDIUM) && OPT(F, O_VOLMIS);
    // MODE 3: every NEE ray is deferred — to k_shadow (binary AnyHit) or, under volMis, to k_transmit, which multiplies the queued contribution
    // by EvalTransmittance (pathtrace.glsl:176, 242: Li *= EvalTransmittance(shadowRay); no random number is drawn without BLEND materials)
    const bool inl = MODE == 2 && F.inlineShadow;
    const bool inlT = MODE == 2 && volMis;                  // transmittance evaluated here

    if (GEN && OPT(F, O_ENVMAP) && !OPT(F, O_UNIFORM))
    {
        float3 Li;
        float4 dirPdf = SampleEnvMap(S, F, rng, Li);
        float3 lightDir = f3(dirPdf);
        float lightPdf = dirPdf.w;
        if constexpr (MODE == 2) { if (inlT) Li *= evalTransmittance(S, F, scatterPos, lightDir, rng, ic); }
        bool visible = true;
        if constexpr (MODE == 2) { if (!volMis && inl) visible = !anyHitInline(S, F, scatterPos, lightDir, PTB_INF - PTB_EPS, rng, ic); }
        if (visible)
        {
            float3 f; float pdf;
            // the phase function is only used by the OPT_MEDIUM && OPT_VOL_MIS branch (pathtrace.glsl:178-186); the binary-AnyHit
            // branch evaluates DisneyEval with the boundary surface's material even for a medium scatter (:200)
            if (isSurface || !volMis) f = DisneyEvalFr(mat, eta, fr, lightDir, pdf);
            else { float ph = PhaseHG(dot(-rd, lightDir), medAniso); f = f3(ph); pdf = ph; }
            if (pdf > 0.0f)
            {
                float misWeight = PowerHeuristic(lightPdf, pdf);
                if (misWeight > 0.0f)
                {
                    float3 c = misWeight * Li * f * F.envMapIntensity / lightPdf;
                    if (inlT || inl) Ld += c;
                    else if constexpr (ImmediatePush<MODE>::value) pushShadowNow(sink, 0, scatterPos, lightDir, PTB_INF - PTB_EPS, c * thr, 0u);
                    else { sa.valid = true; sa.o = scatterPos; sa.d = lightDir; sa.maxDist = PTB_INF - PTB_EPS; sa.c = c * thr; sa.light = 0u; }
                }
            }
        }
    }

    if (OPT(F, O_LIGHTS))
    {
        int idx = (int)(rng.rand() * (float)S.numLights);      // pathtrace.glsl:223
        if (idx >= S.numLights) idx = S.numLights - 1;          // Q6: rand()==1 would index past the end; clamp
        LightSample ls; float area;
        SampleOneLight(S, idx, scatterPos, rng, ls, area);
        float3 Li = ls.emission;
        if (dot(ls.direction, ls.normal) < 0.0f)
        {
            if constexpr (MODE == 2) { if (inlT) Li *= evalTransmittance(S, F, scatterPos, ls.direction, rng, ic); }
            bool visible = true;
            if constexpr (MODE == 2) { if (!volMis && inl) visible = !anyHitInline(S, F, scatterPos, ls.direction, ls.dist - PTB_EPS, rng, ic); }
            if (visible)
            {
                float3 f; float pdf;
                if (isSurface || !volMis) f = DisneyEvalFr(mat, eta, fr, ls.direction, pdf);        // pathtrace.glsl:246-254 vs :268
                else { float ph = PhaseHG(dot(-rd, ls.direction), medAniso); f = f3(ph); pdf = ph; }
                float misWeight = 1.0f;
                if (area > 0.0f) misWeight = PowerHeuristic(ls.pdf, pdf);
                if (pdf > 0.0f)
                {
                    float3 c = misWeight * Li * f / ls.pdf;
                    if (inlT || inl) Ld += c;
                    else if constexpr (ImmediatePush<MODE>::value) pushShadowNow(sink, 1, scatterPos, ls.direction, ls.dist - PTB_EPS, c * thr, (uint32_t)idx);
                    else { sb.valid = true; sb.o = scatterPos; sb.d = ls.direction; sb.maxDist = ls.dist - PTB_EPS; sb.c = c * thr; sb.light = (uint32_t)idx; }
                }
            }
        }
    }
    return Ld;
}

// One iteration of the PathTrace loop body after ClosestHit (pathtrace.glsl:303-471) for path slot p.
template <int MODE>
__device__ __forceinline__ void shadePath(const DevScene& S, const FrameParams& F, const PathState& P, uint32_t p, bool firstIter, bool& cont, ShadowOut& sa, ShadowOut& sb, InlineCounters& ic,
                                          uint32_t* ctrThis, const SlotShadow& ss)
{
    const ShadowSink sink{&P, &ctrThis[CTR_NSHA], &ctrThis[CTR_NSHB], p, ss};
    const float4 ro4 = P.rayO[p], rd4 = P.rayD[p], hit4 = P.hit[p];
    const float4 thr4 = firstIter ? make_float4(1.f, 1.f, 1.f, 0.f) : P.thr[p], rad4 = firstIter ? make_float4(0.f, 0.f, 0.f, 1.f) : P.rad[p];
    const int hitInst = P.hitInst[p];
    Rng rng; rng.s = P.rng[p];
    const uint32_t fl = __float_as_uint(rd4.w);
    int depth = (int)(short)(fl & 0xffffu);
    bool inMedium = (fl & PTB_FLAG_INMEDIUM) != 0u, surfaceScatter = (fl & PTB_FLAG_SURFSCAT) != 0u;
    float3 ro = f3(ro4), rd = f3(rd4), thr = f3(thr4), rad = f3(rad4);
    float alpha = rad4.w, prevPdf = ro4.w, prevRough = thr4.w;
    constexpr bool GEN = MODE >= 1, FULL = MODE >= 2;
    cont = false;
    if (hitInst == PTB_HIT_DEAD) return;          // finished by the trace kernel (finishInTrace)

    if (hitInst == -1)      // miss: pathtrace.glsl:305-339
    {
        if (OPT(F, O_BG) || OPT(F, O_TRANSPBG)) { if (depth == 0) alpha = 0.0f; }
        if (!OPT(F, O_HIDE) || depth > 0)
        {
            if (OPT(F, O_UNIFORM)) rad += f3(F.uniformLightCol[0], F.uniformLightCol[1], F.uniformLightCol[2]) * thr;
            else if (GEN && OPT(F, O_ENVMAP))
            {
                float4 e = EvalEnvMap(S, F, rd);
                float misWeight = 1.0f;
                if (depth > 0) misWeight = PowerHeuristic(prevPdf, e.w);
                if (FULL && OPT(F, O_MEDIUM) && !OPT(F, O_VOLMIS)) { if (!surfaceScatter) misWeight = 1.0f; }
                if (misWeight > 0) rad += misWeight * f3(e) * thr * F.envMapIntensity;
            }
        }
        P.rad[p] = make_float4(rad.x, rad.y, rad.z, alpha);
        return;
    }

    const float t = hit4.x;
    Surf sf;
    Material mat; float eta = 1.0f;
    float4 med = make_float4(0.f, 0.f, 0.f, 0.f), medCol = make_float4(0.f, 0.f, 0.f, 0.f);
    if (GEN) med = P.med[p];            // .w carries the previous matID
    if (FULL) medCol = P.medCol[p];

    if (hitInst <= -2)      // analytic light hit: pathtrace.glsl:341-364 with the stale matID/texCoord of SURVEY Q2
    {
        const int li = -(hitInst + 2);
        if (GEN)
        {
            int prevMat = __float_as_int(med.w);
            float3 em = f3(__ldg(S.materials + (size_t)prevMat * 8 + 1));
            float etex = __ldg(S.materials + (size_t)prevMat * 8 + 6).w;
            if (etex >= 0.f) em = vpow(f3(sampleTexArray(S, P.prevUV[p], (float)(int)etex)), 2.2f);
            rad += em * thr;
        }
        float lpdf; float3 lem;
        lightHitInfo(S, li, ro, rd, t, lpdf, lem);
        float misWeight = 1.0f;
        if (depth > 0) misWeight = PowerHeuristic(prevPdf, lpdf);
        if (FULL && OPT(F, O_MEDIUM) && !OPT(F, O_VOLMIS)) { if (!surfaceScatter) misWeight = 1.0f; }
        rad += misWeight * lem * thr;
        P.rad[p] = make_float4(rad.x, rad.y, rad.z, alpha);
        return;
    }

    {
        const int slot = __float_as_int(hit4.w);
        int matID = __float_as_int(__ldg(S.instTrav + (size_t)hitInst * 4 + 1).w);
        triangleSurface(S, slot, hitInst, hit4.y, hit4.z, ro, rd, t, GEN, GEN && materialNeedsTangents(S, matID), sf);
        getMaterial<MODE>(S, F, sf, rd, depth, prevRough, mat, eta);
    }
    if (GEN) rad += mat.emission * thr;                      // pathtrace.glsl:344

    bool terminated = (depth == F.maxDepth);                 // :367
    bool mediumSampled = false;

    if (!terminated)
    {
        if (FULL && OPT(F, O_MEDIUM))                         // :370-410
        {
            surfaceScatter = false;
            if (inMedium)
            {
                const int mtype = __float_as_int(med.z);
                const float density = med.x, aniso = med.y;
                const float3 mcol = f3(medCol);
                if (mtype == 1) thr *= vexp(-(f3(1.0f) - mcol) * sf.hitDist * density);
                else if (mtype == 3) rad += mcol * sf.hitDist * density * thr;
                else
                {
                    float scatterDist = fminf(-logf(rng.rand()) / density, sf.hitDist);
                    mediumSampled = scatterDist < sf.hitDist;
                    if (mediumSampled)
                    {
                        thr *= mcol;
                        ro += rd * scatterDist;
                        sf.fhp = ro;
                        ShadeFrame mfr;
                        if (!OPT(F, O_VOLMIS)) frameSetup(mat, eta, -rd, sf.ffnormal, mfr);      // DisneyEval(state, ...) of the non-VOL_MIS branch
                        rad += directLight<MODE>(S, F, rd, sf, mat, eta, mfr, false, aniso, thr, rng, sa, sb, ic, sink) * thr;
                        float hr1 = rng.rand(), hr2 = rng.rand();
                        float3 scatterDir = SampleHG(-rd, aniso, hr1, hr2);
                        prevPdf = PhaseHG(dot(-rd, scatterDir), aniso);
                        rd = scatterDir;
                    }
                }
            }
        }
        if (!mediumSampled)
        {
            bool skipped = false;
            float3 L = rd;
            if (FULL && OPT(F, O_ALPHA))                      // :416-426
            {
                if ((mat.alphaMode == 2 && mat.opacity < mat.alphaCutoff) || (mat.alphaMode == 1 && rng.rand() > mat.opacity))
                {
                    depth--;
                    skipped = true;
                }
            }
            if (!skipped)
            {
                surfaceScatter = true;
                ShadeFrame fr;
                frameSetup(mat, eta, -rd, sf.ffnormal, fr);
                rad += directLight<MODE>(S, F, rd, sf, mat, eta, fr, true, 0.f, thr, rng, sa, sb, ic, sink) * thr;  // :431
                float r1 = rng.rand(), r2 = rng.rand(), r3 = rng.rand();
                float pdf;
                float3 f = DisneySampleFr(mat, eta, fr, L, pdf, r1, r2, r3);                                 // :434
                if (pdf > 0.0f) { thr *= f / pdf; prevPdf = pdf; }
                else terminated = true;
            }
            if (!terminated)
            {
                rd = L;
                ro = sf.fhp + rd * PTB_EPS;                   // :442-443
                if (FULL && OPT(F, O_MEDIUM))                 // :447-458
                {
                    if (dot(rd, sf.normal) < 0 && mat.medType != 0)
                    {
                        inMedium = true;
                        med.x = mat.medDensity; med.y = mat.medAniso; med.z = __int_as_float(mat.medType);
                        medCol = make_float4(mat.medColor.x, mat.medColor.y, mat.medColor.z, 0.f);
                    }
                    else if (mat.medType != 0) inMedium = false;
                }
            }
        }
        if (!terminated && OPT(F, O_RR))                     // :464-470
        {
            if (depth >= F.rrDepth)
            {
                float q = fminf(fmaxf(thr.x, fmaxf(thr.y, thr.z)) + 0.001f, 0.95f);
                if (rng.rand() > q) terminated = true;
                else thr /= q;
            }
        }
    }

    P.rad[p] = make_float4(rad.x, rad.y, rad.z, alpha);
    if (!terminated)
    {
        depth++;
        uint32_t nf = ((uint32_t)depth & 0xffffu) | (inMedium ? PTB_FLAG_INMEDIUM : 0u) | (surfaceScatter ? PTB_FLAG_SURFSCAT : 0u);
        P.rayO[p] = make_float4(ro.x, ro.y, ro.z, prevPdf);
        P.rayD[p] = make_float4(rd.x, rd.y, rd.z, __uint_as_float(nf));
        P.thr[p] = make_float4(thr.x, thr.y, thr.z, mat.roughness);
        P.rng[p] = rng.s;
        if (GEN)
        {
            med.w = __int_as_float(sf.matID);
            P.med[p] = med; P.prevUV[p] = sf.uv;
            if (FULL) P.medCol[p] = medCol;
        }
        cont = true;
    }
}

__device__ __forceinline__ void pushShadow(const PathState& P, int which, uint32_t* ctr, uint32_t lane, const ShadowOut& s, uint32_t p, const SlotShadow& ss)
{
    if (ss.keys[which])
    {   // slot-ordered queue: the record goes to index p (see SlotShadow)
        if (s.valid)
        {
            P.shO[which][p] = make_float4(s.o.x, s.o.y, s.o.z, s.maxDist);
            P.shD[which][p] = make_float4(s.d.x, s.d.y, s.d.z, __uint_as_float(p));
            P.shC[which][p] = make_float4(s.c.x, s.c.y, s.c.z, 0.f);
            ss.keys[which][p] = shadowKey(ss, which, s.d, s.light);
        }
        return;
    }
    unsigned m = __ballot_sync(0xffffffffu, s.valid);
    if (!m) return;
    uint32_t b = 0;
    if (lane == 0) b = atomicAdd(ctr, (uint32_t)__popc(m));
    b = __shfl_sync(0xffffffffu, b, 0);
    if (s.valid)
    {
        uint32_t k = b + __popc(m & ((1u << lane) - 1u));
        P.shO[which][k] = make_float4(s.o.x, s.o.y, s.o.z, s.maxDist);
        P.shD[which][k] = make_float4(s.d.x, s.d.y, s.d.z, __uint_as_float(p));
        P.shC[which][k] = make_float4(s.c.x, s.c.y, s.c.z, 0.f);
    }
}

// flags: first shade pass of a wave whose camera rays were generated by k_trace_primary — the queue is the identity over the path slots, so it streams:
//   SHADE_IDENTITY   entry i is path slot i (no queue read in front of the state loads; off-image slots of padded pixel blocks carry PTB_HIT_DEAD)
//   SHADE_STATIC     chunks are dealt to the warps round-robin (uniform work per chunk: no fetch atomic to wait for)
//   SHADE_COUNT_ONLY the continuing paths are counted, not queued (the next trace runs over the slots in screen order: slot-ordered bounce 1)
//   SHADE_OCT_KEYS   direction class of the continuation ray = cell of an 8x8 octahedral map (64 classes, 64 = ended) instead of dominant axis + sign (6 classes, 7 = ended)
enum { SHADE_IDENTITY = 1, SHADE_STATIC = 2, SHADE_COUNT_ONLY = 4, SHADE_OCT_KEYS = 8, SHADE_OCT16_KEYS = 16 /* 16x16 cells, 256 = ended */ };


template <int MODE, int MINB>
__global__ void __launch_bounds__(SHADE_THREADS, MINB) k_shade(DevScene S, FrameParams F, PathState P, const uint32_t* __restrict__ queue, uint32_t* ctrThis,
                                                          uint32_t* ctrNext, uint32_t* nextQueue, DevStats* stats, int firstIter, uint32_t* __restrict__ slotKeys, uint32_t nOverride,
                                                          uint32_t flags, SlotShadow ss)
{
    const uint32_t n = nOverride ? nOverride : ctrThis[CTR_NPATHS];
    const uint32_t lane = threadIdx.x & 31u;
    InlineCounters ic{0u, 0u};
    const bool staticChunks = (flags & SHADE_STATIC) != 0u;
    const uint32_t chunkStride = gridDim.x * (SHADE_THREADS / 32) * 32u;
    uint32_t nextBase = (blockIdx.x * (SHADE_THREADS / 32) + (threadIdx.x >> 5)) * 32u;
    if (!staticChunks && lane == 0) nextBase = atomicAdd(&ctrThis[CTR_FETCH_SHADE], 32u);
    uint32_t continued = 0;
    while (true)
    {
        const uint32_t base = staticChunks ? nextBase : __shfl_sync(0xffffffffu, nextBase, 0);
        if (base >= n) break;
        if (staticChunks) nextBase += chunkStride;
        else if (lane == 0) nextBase = atomicAdd(&ctrThis[CTR_FETCH_SHADE], 32u);      // prefetch the next chunk index
        const uint32_t i = base + lane;
        bool cont = false;
        ShadowOut sa, sb; sa.valid = false; sb.valid = false;
        uint32_t p = 0;
        if (i < n)
        {
            p = (flags & SHADE_IDENTITY) ? (P.hitInst[i] == PTB_HIT_DEAD ? 0xffffffffu : i) : queue[i];
            if (p != 0xffffffffu)                     // (hole of a slot-ordered queue)
            {
                shadePath<MODE>(S, F, P, p, firstIter != 0, cont, sa, sb, ic, ctrThis, ss);
                if (slotKeys)
                {   // direction class of the continuation ray (dominant axis, sign) per path SLOT, 7 = path ended: the bounce-1 trace runs over
                    // the slots in screen order, grouped by this class inside tiles
                    uint32_t k = (flags & SHADE_OCT16_KEYS) ? 256u : (flags & SHADE_OCT_KEYS) ? 64u : 7u;
                    if (cont)
                    {
                        const float4 d4 = P.rayD[p];
                        const float ax = fabsf(d4.x), ay = fabsf(d4.y), az = fabsf(d4.z);
                        if (flags & SHADE_OCT16_KEYS) k = octCell<16>(d4.x, d4.y, d4.z);
                        else if (flags & SHADE_OCT_KEYS) k = octCell(d4.x, d4.y, d4.z);
                        else k = (ax >= ay && ax >= az) ? (d4.x < 0.f ? 1u : 0u) : (ay >= az ? (d4.y < 0.f ? 3u : 2u) : (d4.z < 0.f ? 5u : 4u));
                    }
                    slotKeys[p] = k;
                }
            }
        }
        unsigned m = __ballot_sync(0xffffffffu, cont);
        if (flags & SHADE_COUNT_ONLY) continued += (uint32_t)__popc(m);
        else if (m)
        {
            uint32_t b = 0;
            if (lane == 0) b = atomicAdd(&ctrNext[CTR_NPATHS], (uint32_t)__popc(m));
            b = __shfl_sync(0xffffffffu, b, 0);
            if (cont) nextQueue[b + __popc(m & ((1u << lane) - 1u))] = p;
        }
        if constexpr (!ImmediatePush<MODE>::value)
        {
            if (MODE >= 1) pushShadow(P, 0, &ctrThis[CTR_NSHA], lane, sa, p, ss);
            pushShadow(P, 1, &ctrThis[CTR_NSHB], lane, sb, p, ss);
        }
    }
    if ((flags & SHADE_COUNT_ONLY) && lane == 0 && continued) atomicAdd(&ctrNext[CTR_NPATHS], continued);       // one reduction per warp, nothing waits for it
    if (MODE == 2 && (ic.segs | ic.shadows))
    {
        atomicAdd(&stats->pathSegments, (unsigned long long)ic.segs);
        atomicAdd(&stats->shadowRays, (unsigned long long)ic.shadows);
    }
}

// ------------------------------------------------------------------ shadow (any-hit) trace ----------------------
struct AlphaMask   // deferred AnyHit alpha test: MASK only (BLEND needs the path RNG and runs inline in k_shade)
{
    const DevScene* S;
    __device__ __forceinline__ bool operator()(int slot, int inst, float ux, float uy) const
    {
        Rng dummy; dummy.s = make_uint4(0, 0, 0, 0);
        AlphaRng a{S, &dummy};
        return a(slot, inst, ux, uy);
    }
};

template <bool ALPHA, bool CULL, class AlphaFn>
__device__ __forceinline__ void shadowLoop(const DevScene& S, const FrameParams& F, const PathState& P, int which, uint32_t n, uint32_t* fetchCtr,
                                           AlphaFn alphaFn, const uint32_t* __restrict__ idxQueue, DevStats* stats)
{
    const uint32_t lane = threadIdx.x & 31u;
    SmemStack stk(g_stackSmem + threadIdx.x, (int)blockDim.x);
    const bool lights = OPT(F, O_LIGHTS);
    uint32_t traced = 0;
    while (true)
    {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(fetchCtr, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) break;
        uint32_t i = base + lane;
        bool live = i < n;
        if (idxQueue)
        {   // slot-ordered queue: entry = index of the record (the path slot), 0xffffffff = hole
            i = live ? idxQueue[i] : 0xffffffffu;
            live = i != 0xffffffffu;
            traced += live ? 1u : 0u;
        }
        if (live)
        {
            const float4 o4 = P.shO[which][i], d4 = P.shD[which][i];
            const float3 o = f3(o4), d = f3(d4);
            const float maxDist = o4.w;
            bool occluded = lights && anyLights(S, o, d, maxDist);
            if (!occluded)
            {
                int r = 2;
                if (S.wide && wideRayOk(d)) r = traverseWideAny<ALPHA, CULL>(S, o, d, maxDist, stk, alphaFn);       // same boolean, fewer and wider steps
                if (r == 2) { HitRec h; occluded = traverse<true, ALPHA, CULL>(S, o, d, maxDist, stk, h, alphaFn); }
                else occluded = r != 0;
            }
            if (!occluded)
            {
                const uint32_t p = __float_as_uint(d4.w);
                const float4 c = P.shC[which][i];
                float4 r = P.rad[p];
                r.x += c.x; r.y += c.y; r.z += c.z;
                P.rad[p] = r;
            }
        }
    }
    if (idxQueue)
    {
        traced = __reduce_add_sync(0xffffffffu, traced);
        if (lane == 0 && traced) atomicAdd(&stats->shadowRays, (unsigned long long)traced);
    }
}

// idxQueue != nullptr: slot-ordered queue of nOverride entries (record indices with holes, see SlotShadow); else the compacted queue of *countPtr records
__global__ void __launch_bounds__(TRACE_THREADS, TRACE_MIN_BLOCKS) k_shadow(DevScene S, FrameParams F, PathState P, int which, const uint32_t* __restrict__ countPtr,
                                                           uint32_t* fetchCtr, DevStats* stats, const uint32_t* __restrict__ idxQueue, uint32_t nOverride)
{
    const uint32_t n = idxQueue ? nOverride : *countPtr;
    if (!idxQueue && blockIdx.x == 0 && threadIdx.x == 0 && n) atomicAdd(&stats->shadowRays, (unsigned long long)n);
    const bool alpha = OPT(F, O_ALPHA) && !OPT(F, O_MEDIUM);
    if (alpha) shadowLoop<true, true>(S, F, P, which, n, fetchCtr, AlphaMask{&S}, idxQueue, stats);
    else if (F.cullBoxes) shadowLoop<false, true>(S, F, P, which, n, fetchCtr, NoAlpha(), idxQueue, stats);
    else shadowLoop<false, false>(S, F, P, which, n, fetchCtr, NoAlpha(), idxQueue, stats);
}

// ------------------------------------------------------------------ transmittance (NEE under OPT_MEDIUM + OPT_VOL_MIS) -------
// EvalTransmittance (pathtrace.glsl:119-155) for the queued NEE rays of a scene without BLEND materials (no random number is drawn then, so the
// evaluation can leave the shading kernel): up to maxDepth closest hits along the ray — through refractive / alpha-masked surfaces, attenuated inside
// media — with the wavefront kernels' shared-memory stack; the queued contribution (already misWeight * Li * f / pdf * throughput) times the
// transmittance is added to the path's radiance.  One entry per path per queue: no atomics, deterministic.
constexpr int TRANSMIT_MIN_BLOCKS = 3;
__global__ void __launch_bounds__(TRACE_THREADS, TRANSMIT_MIN_BLOCKS) k_transmit(DevScene S, FrameParams F, PathState P, int which, const uint32_t* __restrict__ countPtr,
                                                                uint32_t* fetchCtr, DevStats* stats)
{
    const uint32_t n = *countPtr;
    const uint32_t lane = threadIdx.x & 31u;
    SmemStack stk(g_stackSmem + threadIdx.x, (int)blockDim.x);
    const bool lights = OPT(F, O_LIGHTS) && !OPT(F, O_HIDE);      // ClosestHit with a fresh State: depth 0 (closestFull)
    uint32_t segs = 0;
    while (true)
    {
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(fetchCtr, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        if (base >= n) break;
        const uint32_t i = base + lane;
        if (i < n)
        {
            const float4 o4 = P.shO[which][i], d4 = P.shD[which][i];
            float3 ro = f3(o4);
            const float3 rd = f3(d4);
            float3 transmittance = f3(1.0f);
            bool blocked = false;
            for (int depth = 0; depth < F.maxDepth; depth++)
            {
                HitRec h; h.t = PTB_INF; h.prim = -1; h.inst = -1; h.light = -1; h.bu = h.bv = 0.f;
                float t = PTB_INF;
                segs++;
                if (lights) closestLights(S, ro, rd, t, h.light);
                traverse<false, false, true>(S, ro, rd, t, stk, h, NoAlpha());
                if (h.t == PTB_INF || h.inst < 0) break;            // no hit, or an emitter: the transmittance so far
                Surf sf;
                const int matID = __float_as_int(__ldg(S.instTrav + (size_t)h.inst * 4 + 1).w);
                triangleSurface(S, h.prim, h.inst, h.bu, h.bv, ro, rd, h.t, true, materialNeedsTangents(S, matID), sf);
                Material mat; float eta;
                getMaterial<2>(S, F, sf, rd, 0, 0.f, mat, eta);
                const bool alphatest = mat.alphaMode == 2 && mat.opacity < mat.alphaCutoff;       // (no BLEND material in this mode: F.deferTransmit)
                const bool refractive = (1.0f - mat.metallic) * mat.specTrans > 0.0f;
                if (!(alphatest || refractive)) { blocked = true; break; }
                if (dot(rd, sf.normal) > 0 && mat.medType != 0)
                {
                    const float3 color = mat.medType == 1 ? f3(1.0f) - mat.medColor : f3(1.0f);
                    transmittance *= vexp(-color * mat.medDensity * sf.hitDist);
                }
                ro = sf.fhp + rd * PTB_EPS;
            }
            if (!blocked)
            {
                const uint32_t p = __float_as_uint(d4.w);
                const float4 c = P.shC[which][i];
                float4 r = P.rad[p];
                r.x += c.x * transmittance.x; r.y += c.y * transmittance.y; r.z += c.z * transmittance.z;
                P.rad[p] = r;
            }
        }
    }
    segs = __reduce_add_sync(0xffffffffu, segs);
    if (lane == 0 && segs) atomicAdd(&stats->pathSegments, (unsigned long long)segs);
}

// ------------------------------------------------------------------ accumulate / tonemap ------------------------
// tile.glsl:70-74: color = pixelColor + accumColor, one sample pass after the other (deterministic order).
__global__ void __launch_bounds__(256) k_accumulate(FrameParams F, WaveParams W, PathState P, float4* accum, float4* previewOut)
{
    const uint32_t perSample = (uint32_t)W.vw * (uint32_t)W.vh;
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < perSample; idx += gridDim.x * blockDim.x)
    {
        int px, py;
        {   // pixel of index idx inside one pass (8x4 blocks, row-major)
            const uint32_t b = idx >> 5, l = idx & 31u, blocksX = (uint32_t)W.vw >> 3;
            px = (int)((b % blocksX) * 8u + (l & 7u)); py = (int)((b / blocksX) * 4u + (l >> 3));
            if (!(px < W.rw && py < W.rh)) continue;
        }
        if (W.previewMode)
        {
            previewOut[(size_t)py * W.rw + px] = P.rad[slotOfSample(W, idx, 0)];
            continue;
        }
        float4* dst = accum + (size_t)(W.y0 + py) * F.renderW + (W.x0 + px);
        float4 a = *dst;
        const int k0 = W.accCount > 0 ? W.accFirst : 0, k1 = W.accCount > 0 ? W.accFirst + W.accCount : W.nSamples;
        for (int k = k0; k < k1; k++)
        {
            const float4 r = P.rad[slotOfSample(W, idx, (uint32_t)k)];
            a.x = r.x + a.x; a.y = r.y + a.y; a.z = r.z + a.z; a.w = r.w + a.w;
        }
        *dst = a;
    }
}

// tonemap.glsl:44-133 + float->unorm8 of glGetTexImage (Renderer.cpp:633)
__global__ void __launch_bounds__(256) k_tonemap(const float4* __restrict__ accum, int w, int h, float invSampleCounter, int enableTonemap, int enableAces,
                                                  int simpleAcesFit, float bgr, float bgg, float bgb, uint32_t features, uchar4* __restrict__ out, float4* __restrict__ outF)
{
    const int n = w * h;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    {
        const float4 a = accum[i];
        float3 color = f3(a.x * invSampleCounter, a.y * invSampleCounter, a.z * invSampleCounter);
        float alpha = a.w * invSampleCounter;
        if (enableTonemap)
        {
            if (enableAces)
            {
                if (simpleAcesFit)
                {
                    float3 num = color * (2.51f * color + 0.03f), den = color * (2.43f * color + 0.59f) + 0.14f;
                    color = f3(clampf(num.x / den.x, 0.f, 1.f), clampf(num.y / den.y, 0.f, 1.f), clampf(num.z / den.z, 0.f, 1.f));
                }
                else
                {
                    float3 c = f3(color.x * 0.59719f + color.y * 0.35458f + color.z * 0.04823f, color.x * 0.07600f + color.y * 0.90834f + color.z * 0.01566f,
                                  color.x * 0.02840f + color.y * 0.13383f + color.z * 0.83777f);
                    float3 va = c * (c + 0.0245786f) + (-0.000090537f);
                    float3 vb = c * (0.983729f * c + 0.4329510f) + 0.238081f;
                    c = va / vb;
                    c = f3(c.x * 1.60475f + c.y * -0.53108f + c.z * -0.07367f, c.x * -0.10208f + c.y * 1.10813f + c.z * -0.00605f,
                           c.x * -0.00327f + c.y * -0.07276f + c.z * 1.07602f);
                    color = f3(clampf(c.x, 0.f, 1.f), clampf(c.y, 0.f, 1.f), clampf(c.z, 0.f, 1.f));
                }
            }
            else color = color * 1.0f / (1.0f + Luminance(color) / 1.5f);
        }
        color = vpow(color, 1.0f / 2.2f);
        float outAlpha = 1.0f;
        float3 bgCol = f3(bgr, bgg, bgb);
        if (features & O_TRANSPBG)
        {
            outAlpha = alpha;
            int x = i % w, y = i / w;
            float mm = fmodf(floorf(((float)x + 0.5f) / 10.0f) + floorf(((float)y + 0.5f) / 10.0f), 2.0f);
            float res = fmaxf(mm > 0.f ? 1.f : (mm < 0.f ? -1.f : 0.f), 0.0f);
            bgCol = mix(f3(0.1f), f3(0.2f), res);
        }
        float4 o;
        if (features & (O_BG | O_TRANSPBG)) { float3 m = mix(bgCol, color, alpha); o = make_float4(m.x, m.y, m.z, outAlpha); }
        else o = make_float4(color.x, color.y, color.z, 1.0f);
        if (outF) outF[i] = o;                  // the RGBA32F texel of tileOutputTexture before glGetTexImage's unorm8 conversion
        float v[4] = {o.x, o.y, o.z, o.w};
        unsigned char b[4];
#pragma unroll
        for (int k = 0; k < 4; k++)
        {
            float f = v[k]; if (!(f == f)) f = 0.f;
            f = clampf(f, 0.f, 1.f);
            b[k] = (unsigned char)floorf(f * 255.0f + 0.5f);
        }
        out[i] = make_uchar4(b[0], b[1], b[2], b[3]);
    }
}

// ------------------------------------------------------------------ parity / batch kernels ----------------------
struct HitOut { float t; int kind, instance, matID, primSlot, triIDx; float bary[3]; int lightIdx; };

__global__ void __launch_bounds__(TRACE_THREADS) k_trace_batch(DevScene S, FrameParams F, const float* __restrict__ rays, long long n, int depth, HitOut* out)
{
    SmemStack stk(g_stackSmem + threadIdx.x, (int)blockDim.x);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    {
        const float3 o = f3(rays[i * 6 + 0], rays[i * 6 + 1], rays[i * 6 + 2]), d = f3(rays[i * 6 + 3], rays[i * 6 + 4], rays[i * 6 + 5]);
        HitRec h; h.t = PTB_INF; h.prim = -1; h.inst = -1; h.light = -1; h.bu = h.bv = 0.f;
        float t = PTB_INF;
        if (OPT(F, O_LIGHTS) && (!OPT(F, O_HIDE) || depth > 0)) closestLights(S, o, d, t, h.light);
        if (F.cullBoxes) traverse<false, false, true>(S, o, d, t, stk, h, NoAlpha());
        else traverse<false, false, false>(S, o, d, t, stk, h, NoAlpha());
        HitOut r;
        r.t = h.t;
        if (h.t == PTB_INF) { r.kind = 0; r.instance = r.matID = r.primSlot = r.triIDx = r.lightIdx = -1; r.bary[0] = r.bary[1] = r.bary[2] = 0.f; }
        else if (h.inst >= 0)
        {
            r.kind = 1; r.instance = h.inst; r.matID = __float_as_int(__ldg(S.instTrav + (size_t)h.inst * 4 + 1).w); r.primSlot = h.prim;
            r.triIDx = __ldg(S.vertIndices + (size_t)h.prim * 3);
            r.bary[0] = xs(xs(1.0f, h.bu), h.bv); r.bary[1] = h.bu; r.bary[2] = h.bv; r.lightIdx = -1;
        }
        else { r.kind = 2; r.instance = r.matID = r.primSlot = r.triIDx = -1; r.bary[0] = r.bary[1] = r.bary[2] = 0.f; r.lightIdx = h.light; }
        out[i] = r;
    }
}

__global__ void __launch_bounds__(TRACE_THREADS) k_any_batch(DevScene S, FrameParams F, const float* __restrict__ rays, const float* __restrict__ maxDist, long long n, int* out)
{
    SmemStack stk(g_stackSmem + threadIdx.x, (int)blockDim.x);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    {
        const float3 o = f3(rays[i * 6 + 0], rays[i * 6 + 1], rays[i * 6 + 2]), d = f3(rays[i * 6 + 3], rays[i * 6 + 4], rays[i * 6 + 5]);
        bool occ = OPT(F, O_LIGHTS) && anyLights(S, o, d, maxDist[i]);
        if (!occ)
        {   // the production any-hit path of k_shadow
            int r = 2;
            if (S.wide && wideRayOk(d)) r = F.cullBoxes ? traverseWideAny<false, true>(S, o, d, maxDist[i], stk, NoAlpha()) : traverseWideAny<false, false>(S, o, d, maxDist[i], stk, NoAlpha());
            if (r == 2) { HitRec h; occ = F.cullBoxes ? traverse<true, false, true>(S, o, d, maxDist[i], stk, h, NoAlpha()) : traverse<true, false, false>(S, o, d, maxDist[i], stk, h, NoAlpha()); }
            else occ = r != 0;
        }
        out[i] = occ ? 1 : 0;
    }
}

struct BsdfQuery { float mat[32]; float V[3], N[3], L[3]; float eta, r1, r2, r3; };
struct BsdfResult { float f[3]; float pdf; float L[3]; };

__global__ void k_bsdf_batch(const BsdfQuery* __restrict__ q, long long n, BsdfResult* out, int sample)
{
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* P = q[i].mat;
    Material mat;
    mat.baseColor = f3(P[0], P[1], P[2]); mat.anisotropic = P[3]; mat.emission = f3(P[4], P[5], P[6]);
    mat.metallic = P[8]; mat.roughness = fmaxf(P[9], 0.001f); mat.subsurface = P[10]; mat.specularTint = P[11];
    mat.sheen = P[12]; mat.sheenTint = P[13]; mat.clearcoat = P[14]; mat.clearcoatRoughness = mixf(0.1f, 0.001f, P[15]);
    mat.specTrans = P[16]; mat.ior = P[17]; mat.medType = 0; mat.medDensity = 0; mat.medColor = f3(0.f); mat.medAniso = 0;
    mat.opacity = P[28]; mat.alphaMode = (int)P[29]; mat.alphaCutoff = P[30];
    materialFinish(mat);
    float3 V = f3(q[i].V[0], q[i].V[1], q[i].V[2]), N = f3(q[i].N[0], q[i].N[1], q[i].N[2]), L = f3(q[i].L[0], q[i].L[1], q[i].L[2]);
    float pdf; float3 f;
    if (sample == 1) f = DisneySample(mat, q[i].eta, V, N, L, pdf, q[i].r1, q[i].r2, q[i].r3);
    else if (sample == 2) f = LambertEval(mat, V, N, L, pdf);
    else if (sample == 3) f = LambertSample(mat, V, N, L, pdf, q[i].r1, q[i].r2);
    else f = DisneyEval(mat, q[i].eta, V, N, L, pdf);
    out[i].f[0] = f.x; out[i].f[1] = f.y; out[i].f[2] = f.z; out[i].pdf = pdf; out[i].L[0] = L.x; out[i].L[1] = L.y; out[i].L[2] = L.z;
}

__global__ void k_camera_rays(FrameParams F, WaveParams W, float* out)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= F.renderW * F.renderH) return;
    int x = i % F.renderW, y = i / F.renderW;
    Rng rng; float3 ro, rd;
    cameraRay(F, W, x, y, W.firstSample, rng, ro, rd);
    out[i * 6 + 0] = ro.x; out[i * 6 + 1] = ro.y; out[i * 6 + 2] = ro.z; out[i * 6 + 3] = rd.x; out[i * 6 + 4] = rd.y; out[i * 6 + 5] = rd.z;
}

// ------------------------------------------------------------------ launchers -----------------------------------
static inline cudaStream_t st(const LaunchCfg& c) { return (cudaStream_t)c.stream; }
static inline size_t stackBytes(const DevScene& S, int threads) { return (size_t)S.stackDepth * threads * sizeof(uint32_t); }
static inline size_t stackBytesAny(const DevScene& S, int threads) { return (size_t)S.stackDepthAny * threads * sizeof(uint32_t); }

// Per-device kernel attributes for the current device: dynamic shared memory of the stack-carrying kernels (the default limit is 48 KB) and the
// resident blocks per SM the persistent grids are sized with.  Called by the host side after cudaSetDevice whenever a context is created
// or its stack depth changes; the results live in the context, not in process-wide statics.
int ptbk_configure_device(const DevScene& S, int* traceBlocks, int* shadowBlocks, int shadeBlocks[4], int* transmitBlocks)
{
    const size_t smem = stackBytes(S, TRACE_THREADS), smemAny = stackBytesAny(S, TRACE_THREADS);
    cudaError_t e = cudaSuccess;
    auto upd = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
    upd(cudaFuncSetAttribute(k_trace, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    upd(cudaFuncSetAttribute(k_trace_primary, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    upd(cudaFuncSetAttribute(k_shadow, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemAny));
    upd(cudaFuncSetAttribute(k_transmit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    upd(cudaFuncSetAttribute(k_trace_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    upd(cudaFuncSetAttribute(k_any_batch, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smemAny));
    int nb = 0;
    upd(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_trace, TRACE_THREADS, smem));
    *traceBlocks = nb > 0 ? nb : 1;
    nb = 0;
    upd(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_shadow, TRACE_THREADS, smemAny));
    *shadowBlocks = nb > 0 ? nb : 1;
    upd(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&shadeBlocks[0], k_shade<0, 4>, SHADE_THREADS, 0));
    upd(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&shadeBlocks[1], k_shade<1, 5>, SHADE_THREADS, 0));
    upd(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&shadeBlocks[2], k_shade<2, 4>, SHADE_THREADS, 0));
    upd(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&shadeBlocks[3], k_shade<3, 4>, SHADE_THREADS, 0));
    for (int k = 0; k < 4; k++) if (shadeBlocks[k] < 1) shadeBlocks[k] = 1;
    nb = 0;
    upd(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_transmit, TRACE_THREADS, smem));
    *transmitBlocks = nb > 0 ? nb : 1;
    return (int)e;
}

void ptbk_camera(const LaunchCfg& c, const DevScene& S, const FrameParams& F, const WaveParams& W, const PathState& P, uint32_t* ctr0)
{
    int blocks = c.numSMs * 8;
    k_camera<<<blocks, 256, 0, st(c)>>>(S, F, W, P, ctr0);
    COUNT_LAUNCH(c, 1);
}

void ptbk_trace(const LaunchCfg& c, const DevScene& S, const FrameParams& F, const PathState& P, const uint32_t* queue,
                const uint32_t* countPtr, uint32_t* fetchCtr, int depthForLights, DevStats* stats, uint32_t* keys, uint32_t* hist, uint32_t nOverride, uint32_t holeKey, uint32_t tflags)
{
    const int bps = c.traceBlocks;
    k_trace<<<c.numSMs * bps, TRACE_THREADS, stackBytes(S, TRACE_THREADS), st(c)>>>(S, F, P, queue, countPtr, fetchCtr, depthForLights, stats, keys, hist, nOverride, holeKey, tflags);
    COUNT_LAUNCH(c, 1);
}

void ptbk_trace_primary(const LaunchCfg& c, const DevScene& S, const FrameParams& F, const WaveParams& W, const PathState& P, uint32_t* ctr0, int depthForLights,
                        DevStats* stats, uint32_t* keys, uint32_t* hist, uint32_t liveCount, uint32_t holeKey, uint32_t tflags)
{
    const int bps = c.traceBlocks;
    k_trace_primary<<<c.numSMs * bps, TRACE_THREADS, stackBytes(S, TRACE_THREADS), st(c)>>>(S, F, W, P, ctr0, depthForLights, stats, keys, hist, liveCount, holeKey, tflags);
    COUNT_LAUNCH(c, 1);
}

void ptbk_sort_tile_local(const LaunchCfg& c, const uint32_t* queue, const uint32_t* keys, const uint32_t* countPtr, int numKeys, uint32_t* sorted, int holeKey, uint32_t nOverride)
{
    k_sort_tile_local<<<c.numSMs * PTB_SORT_MINB, 256, (size_t)numKeys * 2 * sizeof(uint32_t), st(c)>>>(queue, keys, countPtr, sorted, numKeys, holeKey, nOverride);
    COUNT_LAUNCH(c, 1);
}

void ptbk_sort(const LaunchCfg& c, const uint32_t* queue, const uint32_t* keys, const uint32_t* countPtr, uint32_t* hist, uint32_t* cursor, int numKeys,
               uint32_t* sorted)
{
    k_sort_scan<<<1, 32, 0, st(c)>>>(hist, cursor, numKeys);
    if (numKeys <= 4096) k_sort_scatter<<<c.numSMs * 4, 256, (size_t)numKeys * 2 * sizeof(uint32_t), st(c)>>>(queue, keys, countPtr, cursor, sorted, numKeys);
    else k_sort_scatter_global<<<c.numSMs * 8, 256, 0, st(c)>>>(queue, keys, countPtr, cursor, sorted);
    COUNT_LAUNCH(c, 2);
}

void ptbk_shade(const LaunchCfg& c, const DevScene& S, const FrameParams& F, const PathState& P, const uint32_t* queue,
                uint32_t* ctrThis, uint32_t* ctrNext, uint32_t* nextQueue, DevStats* stats, int firstIter, uint32_t* slotKeys, uint32_t nOverride, uint32_t flags,
                uint32_t* shadowKeysA, uint32_t* shadowKeysB, int lightKeys)
{
    const int* bps = c.shadeBlocks;
    const SlotShadow ss{{shadowKeysA, shadowKeysB}, lightKeys ? 1u : 0u};
    if (F.general == 2 && F.inlineShadow) k_shade<2, 4><<<c.numSMs * bps[2], SHADE_THREADS, 0, st(c)>>>(S, F, P, queue, ctrThis, ctrNext, nextQueue, stats, firstIter, slotKeys, nOverride, flags, ss);
    else if (F.general == 2) k_shade<3, 4><<<c.numSMs * bps[3], SHADE_THREADS, 0, st(c)>>>(S, F, P, queue, ctrThis, ctrNext, nextQueue, stats, firstIter, slotKeys, nOverride, flags, ss);
    else if (F.general == 1) k_shade<1, 5><<<c.numSMs * bps[1], SHADE_THREADS, 0, st(c)>>>(S, F, P, queue, ctrThis, ctrNext, nextQueue, stats, firstIter, slotKeys, nOverride, flags, ss);
    else if (PTB_SHADE_LATER_BLOCKS != 4 && !firstIter)
        k_shade<0, PTB_SHADE_LATER_BLOCKS><<<c.numSMs * PTB_SHADE_LATER_BLOCKS, SHADE_THREADS, 0, st(c)>>>(S, F, P, queue, ctrThis, ctrNext, nextQueue, stats, firstIter, slotKeys, nOverride, flags, ss);
    else k_shade<0, 4><<<c.numSMs * bps[0], SHADE_THREADS, 0, st(c)>>>(S, F, P, queue, ctrThis, ctrNext, nextQueue, stats, firstIter, slotKeys, nOverride, flags, ss);
    COUNT_LAUNCH(c, 1);
}

void ptbk_shadow(const LaunchCfg& c, const DevScene& S, const FrameParams& F, const PathState& P, int which, const uint32_t* countPtr,
                 uint32_t* fetchCtr, DevStats* stats, const uint32_t* idxQueue, uint32_t nOverride)
{
    const int bps = c.shadowBlocks;
    k_shadow<<<c.numSMs * bps, TRACE_THREADS, stackBytesAny(S, TRACE_THREADS), st(c)>>>(S, F, P, which, countPtr, fetchCtr, stats, idxQueue, nOverride);
    COUNT_LAUNCH(c, 1);
}

void ptbk_transmit(const LaunchCfg& c, const DevScene& S, const FrameParams& F, const PathState& P, int which, const uint32_t* countPtr,
                   uint32_t* fetchCtr, DevStats* stats)
{
    k_transmit<<<c.numSMs * c.transmitBlocks, TRACE_THREADS, stackBytes(S, TRACE_THREADS), st(c)>>>(S, F, P, which, countPtr, fetchCtr, stats);
    COUNT_LAUNCH(c, 1);
}

void ptbk_accumulate(const LaunchCfg& c, const FrameParams& F, const WaveParams& W, const PathState& P, float4* accum, float4* previewOut)
{
    k_accumulate<<<c.numSMs * 8, 256, 0, st(c)>>>(F, W, P, accum, previewOut);
    COUNT_LAUNCH(c, 1);
}

void ptbk_tonemap(const LaunchCfg& c, const float4* accum, int w, int h, float invSampleCounter, int enableTonemap, int enableAces,
                  int simpleAcesFit, const float* bg, uint32_t features, uchar4* out, float4* outF)
{
    k_tonemap<<<c.numSMs * 8, 256, 0, st(c)>>>(accum, w, h, invSampleCounter, enableTonemap, enableAces, simpleAcesFit, bg[0], bg[1], bg[2], features, out, outF);
    COUNT_LAUNCH(c, 1);
}

void ptbk_trace_closest_batch(const LaunchCfg& c, const DevScene& S, const FrameParams& F, const float* rays, long long n, int depth, void* hitsOut)
{
    const int bps = c.traceBlocks;
    k_trace_batch<<<c.numSMs * bps, TRACE_THREADS, stackBytes(S, TRACE_THREADS), st(c)>>>(S, F, rays, n, depth, (HitOut*)hitsOut);
    COUNT_LAUNCH(c, 1);
}
void ptbk_trace_any_batch(const LaunchCfg& c, const DevScene& S, const FrameParams& F, const float* rays, const float* maxDist, long long n, int* out)
{
    const int bps = c.shadowBlocks;
    k_any_batch<<<c.numSMs * bps, TRACE_THREADS, stackBytesAny(S, TRACE_THREADS), st(c)>>>(S, F, rays, maxDist, n, out);
    COUNT_LAUNCH(c, 1);
}
void ptbk_bsdf_batch(const LaunchCfg& c, const void* queries, long long n, void* results, int sample)
{
    if (n <= 0) return;
    k_bsdf_batch<<<(unsigned)((n + 127) / 128), 128, 0, st(c)>>>((const BsdfQuery*)queries, n, (BsdfResult*)results, sample);
    COUNT_LAUNCH(c, 1);
}
void ptbk_camera_rays(const LaunchCfg& c, const FrameParams& F, const WaveParams& W, float* outRays)
{
    int n = F.renderW * F.renderH;
    k_camera_rays<<<(n + 255) / 256, 256, 0, st(c)>>>(F, W, outRays);
    COUNT_LAUNCH(c, 1);
}

// ------------------------------------------------------------------ device-side TLAS rebuild (N3) ---------------------------------
// Scene::RebuildInstances on the device: instance world boxes (Scene.cpp:154-184), Bvh(10, 64, usesah = false)::Build (bvh.cpp:68-243) and
// BvhTranslator::ProcessTLASNodes (bvh_translator.cpp:58-86), written straight into the canonical node array — byte-identical to the reference's host code
// (ptbd_build_tlas_host is its exact sequential twin and the checker in the tests).
//
// The reference partitions in place, but what it BUILDS depends only on which instances go left and right at every node: child boxes are unions (min / max
// are exact and order-free as long as no -0.0 is involved), `near2far` depends on (count + start) and start = parent.start (+ left count), and with one
// instance per leaf the pre-order index of a child follows from the leaf counts (left = parent + 1, right = parent + 2 * nLeft).  So no instance is ever moved:
// every instance carries the record id of the node it currently sits in, and one level is  classify -> count + grow the child records with atomics -> emit nodes.
// The cases where the reference's element ORDER matters (a side stays empty: coincident centroids or a centre that rounds onto an end; -0.0 / non-finite
// coordinates) raise a flag and the host rebuilds with the sequential algorithm.
constexpr int TLAS_THREADS = 512;
struct TlasRec { int start, count, dfs, lastPrim; unsigned key[12]; };      // key: bounds min.xyz, max.xyz, centroid min.xyz, max.xyz as order-preserving uints

__device__ __forceinline__ unsigned fkey(float f) { unsigned b = __float_as_uint(f); return b ^ ((b >> 31) ? 0xFFFFFFFFu : 0x80000000u); }
__device__ __forceinline__ float keyf(unsigned k) { return __uint_as_float(k ^ ((k >> 31) ? 0x80000000u : 0xFFFFFFFFu)); }
__device__ __forceinline__ float stdmin(float a, float b) { return b < a ? b : a; }       // std::min / std::max as Vec3::Min / Max use them
__device__ __forceinline__ float stdmax(float a, float b) { return a < b ? b : a; }

// grow record q by one instance's box / centroid keys.  `peers` = lanes of this warp that grow the SAME record (at the top levels whole warps do): they are
// reduced with redux.sync first, one lane issues the 12 atomics — without this the first levels serialise 14 atomics per instance on two records.
__device__ __forceinline__ void tlasGrow(TlasRec& q, unsigned peers, int i, const unsigned (&k)[12])
{
    const unsigned lane = threadIdx.x & 31u;
    const bool leader = lane == (unsigned)(__ffs((int)peers) - 1);
    if (__popc(peers) == 1)
    {
        atomicAdd(&q.count, 1); q.lastPrim = i;
        for (int a = 0; a < 3; a++) { atomicMin(&q.key[a], k[a]); atomicMax(&q.key[3 + a], k[3 + a]); atomicMin(&q.key[6 + a], k[6 + a]); atomicMax(&q.key[9 + a], k[9 + a]); }
        return;
    }
    unsigned r[12];
    for (int a = 0; a < 3; a++)
    {
        r[a] = __reduce_min_sync(peers, k[a]); r[3 + a] = __reduce_max_sync(peers, k[3 + a]);
        r[6 + a] = __reduce_min_sync(peers, k[6 + a]); r[9 + a] = __reduce_max_sync(peers, k[9 + a]);
    }
    if (leader)
    {
        atomicAdd(&q.count, __popc(peers)); q.lastPrim = i;
        for (int a = 0; a < 3; a++) { atomicMin(&q.key[a], r[a]); atomicMax(&q.key[3 + a], r[3 + a]); atomicMin(&q.key[6 + a], r[6 + a]); atomicMax(&q.key[9 + a], r[9 + a]); }
    }
}

__device__ __forceinline__ void tlasEmitLeaf(float* out, int top, const TlasRec& r, const float* __restrict__ instBounds, const int* __restrict__ blasRoot, const int* __restrict__ materialID)
{
    const int inst = r.lastPrim;
    float* n = out + (size_t)(top + r.dfs) * 9;
    const float* b = instBounds + (size_t)inst * 6;
    for (int a = 0; a < 6; a++) n[a] = b[a];
    n[6] = (float)blasRoot[inst]; n[7] = (float)materialID[inst]; n[8] = (float)(-inst - 1);
}

// Cooperative launch, one CTA per SM: the levels (and the phases inside a level) are separated by grid-wide barriers.
__global__ void __launch_bounds__(TLAS_THREADS) k_tlas_build(float* __restrict__ nodes, int top, const float4* __restrict__ transforms, int n, const int* __restrict__ blasRoot,
                                                      const int* __restrict__ materialID, float* __restrict__ instBounds, float* __restrict__ cent, int* __restrict__ nodeOf,
                                                      TlasRec* recA, TlasRec* recB, TlasRec* recC, int* __restrict__ remap, int* result /* [0] fallback flag, [1] height, [2] nodes of the next level; zeroed by the host */)
{
    cooperative_groups::grid_group grid = cooperative_groups::this_grid();
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
    volatile int* vres = result;
#define sFlag (vres[0])
#define sActive (result[2])
    const unsigned EMPTY_LO = fkey(3.402823466e+38f), EMPTY_HI = fkey(-3.402823466e+38f);
    TlasRec* cur = recA; TlasRec* nxt = recB; TlasRec* cmp = recC;      // this level's internal nodes (compact) / their children (2 per node) / next level (compact)
    if (tid == 0) { TlasRec r; r.start = 0; r.count = 0 /* counted by tlasGrow */; r.dfs = 0; r.lastPrim = 0; for (int a = 0; a < 3; a++) { r.key[a] = EMPTY_LO; r.key[3 + a] = EMPTY_HI; r.key[6 + a] = EMPTY_LO; r.key[9 + a] = EMPTY_HI; } cur[0] = r; }
    grid.sync();
    // instance world boxes and centroids (Scene.cpp:154-184, bbox::center), root bounds (Bvh::Build / BuildImpl)
    for (int base = tid & ~31; base < n; base += nt)
    {   // (warp-uniform trip count: the aggregation in tlasGrow uses warp collectives)
        const int i = base + (tid & 31);
        const unsigned live = __ballot_sync(0xffffffffu, i < n);
        if (i >= n) continue;
        const float* b = nodes + (size_t)blasRoot[i] * 9;
        const float4 r0 = transforms[(size_t)i * 4], r1 = transforms[(size_t)i * 4 + 1], r2 = transforms[(size_t)i * 4 + 2], r3 = transforms[(size_t)i * 4 + 3];
        const float R[3] = {r0.x, r0.y, r0.z}, U[3] = {r1.x, r1.y, r1.z}, Fw[3] = {r2.x, r2.y, r2.z}, T[3] = {r3.x, r3.y, r3.z};
        bool bad = false;
        unsigned rk[12];
        for (int c = 0; c < 3; c++)
        {
            const float xa = __fmul_rn(R[c], b[0]), xb = __fmul_rn(R[c], b[3]), ya = __fmul_rn(U[c], b[1]), yb = __fmul_rn(U[c], b[4]), za = __fmul_rn(Fw[c], b[2]), zb = __fmul_rn(Fw[c], b[5]);
            const float lo = __fadd_rn(__fadd_rn(__fadd_rn(stdmin(xa, xb), stdmin(ya, yb)), stdmin(za, zb)), T[c]);
            const float hi = __fadd_rn(__fadd_rn(__fadd_rn(stdmax(xa, xb), stdmax(ya, yb)), stdmax(za, zb)), T[c]);
            const float ce = __fmul_rn(__fadd_rn(hi, lo), 0.5f);
            instBounds[(size_t)i * 6 + c] = lo; instBounds[(size_t)i * 6 + 3 + c] = hi; cent[(size_t)i * 3 + c] = ce;
            // order-free unions need: finite values and no negative zero (std::min keeps the FIRST of +0 / -0 it meets)
            bad |= !(fabsf(lo) <= 3.0e38f) || !(fabsf(hi) <= 3.0e38f) || __float_as_uint(lo) == 0x80000000u || __float_as_uint(hi) == 0x80000000u || __float_as_uint(ce) == 0x80000000u;
            rk[c] = fkey(lo); rk[3 + c] = fkey(hi); rk[6 + c] = fkey(ce); rk[9 + c] = fkey(ce);
        }
        if (bad) vres[0] = 1;
        nodeOf[i] = 0;
        tlasGrow(cur[0], live, i, rk);        // root bounds (n == 1: the root is a leaf holding this instance: lastPrim)
    }
    grid.sync();
    int numCur = 1, level = 0;
    if (n == 1) { if (tid == 0) tlasEmitLeaf(nodes, top, cur[0], instBounds, blasRoot, materialID); numCur = 0; }
    while (numCur > 0 && !sFlag)
    {
        // phase 1: every internal node of this level (cur[0 .. numCur), all with >= 2 instances) gets its two child records nxt[2j], nxt[2j+1]
        for (int j = tid; j < 2 * numCur; j += nt)
        {
            TlasRec l;
            l.start = l.count = l.dfs = l.lastPrim = 0;
            for (int a = 0; a < 3; a++) { l.key[a] = EMPTY_LO; l.key[3 + a] = EMPTY_HI; l.key[6 + a] = EMPTY_LO; l.key[9 + a] = EMPTY_HI; }
            nxt[j] = l;
        }
        if (tid == 0) vres[2] = 0;
        grid.sync();
        // phase 2: classify every instance that sits in an internal node (bvh.cpp:97-98, 150-206) and grow its child's record
        for (int base = tid & ~31; base < n; base += nt)
        {
            const int i = base + (tid & 31);
            const int j = i < n ? nodeOf[i] : -1;                            // -1: already in a leaf
            int child = -1;
            if (j >= 0)
            {
                const TlasRec& p = cur[j];
                float clo[3], chi[3];
                for (int a = 0; a < 3; a++) { clo[a] = keyf(p.key[6 + a]); chi[a] = keyf(p.key[9 + a]); }
                const float ex = __fsub_rn(chi[0], clo[0]), ey = __fsub_rn(chi[1], clo[1]), ez = __fsub_rn(chi[2], clo[2]);
                const int axis = (ex >= ey && ex >= ez) ? 0 : ((ey >= ex && ey >= ez) ? 1 : ((ez >= ex && ez >= ey) ? 2 : 0));      // bbox::maxdim
                const float border = __fmul_rn(__fadd_rn(chi[axis], clo[axis]), 0.5f);                                          // centroid_bounds.center()[axis]
                const float ext = axis == 0 ? ex : (axis == 1 ? ey : ez);
                if (!(ext > 0.f)) vres[0] = 1;                                  // the reference splits by position then: order matters
                else
                {
                    const bool near2far = ((p.count + p.start) & 1) != 0;
                    const float c = cent[(size_t)i * 3 + axis];
                    const bool left = near2far ? (c < border) : (c >= border);
                    child = 2 * j + (left ? 0 : 1);
                }
            }
            const unsigned growing = __ballot_sync(0xffffffffu, child >= 0);
            if (child >= 0)
            {
                unsigned k[12];
                for (int a = 0; a < 3; a++)
                {
                    k[a] = fkey(instBounds[(size_t)i * 6 + a]); k[3 + a] = fkey(instBounds[(size_t)i * 6 + 3 + a]);
                    k[6 + a] = k[9 + a] = fkey(cent[(size_t)i * 3 + a]);
                }
                tlasGrow(nxt[child], __match_any_sync(growing, child), i, k);
                nodeOf[i] = child;
            }
        }
        grid.sync();
        // phase 3: emit the internal nodes of this level and the leaves among their children; number the children in pre-order; the children that are
        // internal nodes themselves form the next level (compacted into cmp, remap = child record -> index there, -1 for leaves)
        for (int j = tid; j < numCur; j += nt)
        {
            const TlasRec p = cur[j];
            TlasRec& l = nxt[2 * j]; TlasRec& r = nxt[2 * j + 1];
            remap[2 * j] = -1; remap[2 * j + 1] = -1;
            if (l.count == 0 || r.count == 0) { vres[0] = 1; continue; }       // a side stayed empty: the reference falls back to a split by position
            l.start = p.start; r.start = p.start + l.count;
            l.dfs = p.dfs + 1; r.dfs = p.dfs + 2 * l.count;                   // a subtree with k single-instance leaves has 2k - 1 nodes
            float* o = nodes + (size_t)(top + p.dfs) * 9;
            for (int a = 0; a < 6; a++) o[a] = keyf(p.key[a]);
            o[6] = (float)(top + l.dfs); o[7] = (float)(top + r.dfs); o[8] = 0.f;
            if (l.count == 1) tlasEmitLeaf(nodes, top, l, instBounds, blasRoot, materialID);
            else { const int k = atomicAdd(&sActive, 1); cmp[k] = l; remap[2 * j] = k; }
            if (r.count == 1) tlasEmitLeaf(nodes, top, r, instBounds, blasRoot, materialID);
            else { const int k = atomicAdd(&sActive, 1); cmp[k] = r; remap[2 * j + 1] = k; }
            atomicMax(&result[1], level + 1);
        }
        grid.sync();
        for (int i = tid; i < n; i += nt) { const int j = nodeOf[i]; if (j >= 0) nodeOf[i] = remap[j]; }
        numCur = vres[2];
        TlasRec* t = cur; cur = cmp; cmp = t;
        level++;
        grid.sync();
    }
    // the slot BvhTranslator reserves but never uses (2n slots, 2n - 1 nodes) stays zero, as nodes.resize() leaves it
    if (tid < 9) nodes[(size_t)(top + 2 * n - 1) * 9 + tid] = 0.f;
#undef sFlag
#undef sActive
}

int ptbk_tlas_build(const LaunchCfg& c, float* nodes, int top, const float4* transforms, int n, const int* blasRoot, const int* materialID, float* instBounds, float* cent,
                    int* nodeOf, void* recA, void* recB, void* recC, int* remap, int* result)
{
    TlasRec* a = (TlasRec*)recA; TlasRec* b = (TlasRec*)recB; TlasRec* cc = (TlasRec*)recC;
    void* args[] = {&nodes, &top, &transforms, &n, &blasRoot, &materialID, &instBounds, &cent, &nodeOf, &a, &b, &cc, &remap, &result};
    // one CTA per SM is always co-resident (32 registers, no shared memory): the grid barrier cannot deadlock
    cudaError_t e = cudaLaunchCooperativeKernel((const void*)k_tlas_build, dim3(c.numSMs), dim3(TLAS_THREADS), args, 0, st(c));
    COUNT_LAUNCH(c, 1);
    return (int)e;
}
int ptbk_tlas_rec_size() { return (int)sizeof(TlasRec); }
