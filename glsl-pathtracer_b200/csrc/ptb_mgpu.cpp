// ptb_mgpu.cpp — ptb_mgpu_*: N contexts (one per GPU of one box) behind one handle, for a host that is a single process and a
// single thread like the reference's Renderer (Renderer.h:169-179 has no notion of a device).
//
// The path shards by sample pass (every pixel-sample is independent, tile.glsl:41-75): context r renders the passes
// {first + r, first + r + N, ...} of the whole frame with the seeds a 1-GPU run would use; there is no data-path collective.
// All launches are asynchronous, so one host thread keeps the N GPUs busy.  On readback the N running sums are combined with ONE
// ncclReduce over NVLink into a SCRATCH buffer on GPU 0 (the running sums themselves stay per-GPU partial sums, so any number of
// progressive readbacks is correct) and the tonemap runs there.
//
// NCCL is bound at run time (dlopen "libnccl.so.2"): a process that already carries an NCCL (torch) shares it, a process
// that never asks for more than one GPU does not need it.
#include "ptb200.h"
#include <cuda_runtime_api.h>
#include <dlfcn.h>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

extern "C" __attribute__((visibility("hidden"))) void ptb_set_last_error_(const char* msg);      // ptb_api.cpp

namespace {

// the slice of nccl.h this file uses (ABI-stable since NCCL 2.0)
typedef struct ncclComm* ncclComm_t;
enum { NCCL_SUCCESS = 0, NCCL_FLOAT = 7, NCCL_SUM = 0 };
struct Nccl
{
    void* lib = nullptr;
    int (*CommInitAll)(ncclComm_t*, int, const int*) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Reduce)(const void*, void*, size_t, int, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load(std::string& err)
    {
        if (lib) return true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"})
            if ((lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL))) break;
        if (!lib) { err = std::string("NCCL not found: ") + dlerror(); return false; }
        auto sym = [&](const char* n) { void* p = dlsym(lib, n); if (!p) err = std::string("NCCL symbol missing: ") + n; return p; };
        CommInitAll = (decltype(CommInitAll))sym("ncclCommInitAll");
        CommDestroy = (decltype(CommDestroy))sym("ncclCommDestroy");
        Reduce = (decltype(Reduce))sym("ncclReduce");
        GroupStart = (decltype(GroupStart))sym("ncclGroupStart");
        GroupEnd = (decltype(GroupEnd))sym("ncclGroupEnd");
        GetErrorString = (decltype(GetErrorString))sym("ncclGetErrorString");
        return err.empty();
    }
};
Nccl g_nccl;

int fail(int code, const std::string& msg) { ptb_set_last_error_(msg.c_str()); return code; }

} // namespace

struct PtbMgpu
{
    std::vector<PtbCtx*> ctx;
    std::vector<int> devices;
    std::vector<ncclComm_t> comms;
    std::vector<cudaStream_t> streams;      // per-device stream the contexts run on (NCCL is enqueued on the same streams)
    void* scratch = nullptr;                // float4[w*h] on devices[0]: reduced sum
    size_t scratchBytes = 0;
    // ptb_mgpu_snapshot_output with several GPUs: every GPU freezes ITS running sum with an asynchronous device-to-device copy (no cross-GPU
    // synchronisation while rendering goes on); the reduce + tonemap happen when the frozen image is asked for
    std::vector<void*> frozen; size_t frozenBytes = 0; float frozenInv = 1.f; bool frozenPending = false;
    int w = 0, h = 0;
    int nextPass = 1;
};

extern "C" {

int ptb_mgpu_destroy(PtbMgpu* m)
{
    if (!m) return PTB_OK;
    for (size_t i = 0; i < m->comms.size(); i++) if (m->comms[i]) g_nccl.CommDestroy(m->comms[i]);
    for (size_t i = 0; i < m->ctx.size(); i++) if (m->ctx[i]) ptb_destroy(m->ctx[i]);
    if (m->scratch) { cudaSetDevice(m->devices[0]); cudaFree(m->scratch); }
    for (size_t i = 0; i < m->frozen.size(); i++) if (m->frozen[i]) { cudaSetDevice(m->devices[i]); cudaFree(m->frozen[i]); }
    for (size_t i = 0; i < m->streams.size(); i++) if (m->streams[i]) { cudaSetDevice(m->devices[i]); cudaStreamDestroy(m->streams[i]); }
    delete m;
    return PTB_OK;
}

int ptb_mgpu_create(const PtbSceneDesc* scene, const PtbOptions* opts, const int32_t* devices, int32_t numDevices, PtbMgpu** out)
{
    if (!scene || !opts || !out || numDevices < 1) return fail(PTB_ERR_INVALID_ARGUMENT, "bad argument");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0) { cudaGetLastError(); return fail(PTB_ERR_NO_DEVICE, "no usable CUDA device (libptb200 has no CPU fallback)"); }
    PtbMgpu* m = new PtbMgpu();
    for (int i = 0; i < numDevices; i++)
    {
        int d = devices ? devices[i] : i;
        if (d < 0 || d >= ndev) { ptb_mgpu_destroy(m); return fail(PTB_ERR_NO_DEVICE, "device ordinal out of range"); }
        for (int j = 0; j < i; j++) if (m->devices[j] == d) { ptb_mgpu_destroy(m); return fail(PTB_ERR_INVALID_ARGUMENT, "duplicate device ordinal"); }
        m->devices.push_back(d);
    }
    m->ctx.assign(numDevices, nullptr); m->streams.assign(numDevices, nullptr);
    for (int i = 0; i < numDevices; i++)
    {
        int rc = ptb_create(scene, opts, m->devices[i], &m->ctx[i]);          // scene replica per GPU (15 MB for hyperion)
        if (rc) { ptb_mgpu_destroy(m); return rc; }
        cudaSetDevice(m->devices[i]);
        if (cudaStreamCreateWithFlags(&m->streams[i], cudaStreamNonBlocking) != cudaSuccess) { ptb_mgpu_destroy(m); return fail(PTB_ERR_CUDA, "cudaStreamCreate failed"); }
        ptb_set_stream(m->ctx[i], m->streams[i]);
    }
    m->w = opts->renderW; m->h = opts->renderH;
    if (numDevices > 1)
    {
        std::string err;
        if (!g_nccl.load(err)) { ptb_mgpu_destroy(m); return fail(PTB_ERR_UNSUPPORTED, err); }
        m->comms.assign(numDevices, nullptr);
        int r = g_nccl.CommInitAll(m->comms.data(), numDevices, m->devices.data());
        if (r != NCCL_SUCCESS) { std::string e = std::string("ncclCommInitAll: ") + g_nccl.GetErrorString(r); m->comms.clear(); ptb_mgpu_destroy(m); return fail(PTB_ERR_CUDA, e); }
    }
    *out = m;
    return PTB_OK;
}

int ptb_mgpu_num_devices(PtbMgpu* m) { return m ? (int)m->ctx.size() : 0; }
PtbCtx* ptb_mgpu_context(PtbMgpu* m, int32_t i) { return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[i] : nullptr; }

#define EACH(call) do { for (size_t i_ = 0; i_ < m->ctx.size(); i_++) { PtbCtx* c = m->ctx[i_]; int rc_ = (call); if (rc_) return rc_; } } while (0)

int ptb_mgpu_set_options(PtbMgpu* m, const PtbOptions* o)
{
    if (!m || !o) return fail(PTB_ERR_INVALID_ARGUMENT, "null argument");
    EACH(ptb_set_options(c, o));
    if (o->renderW != m->w || o->renderH != m->h) { m->w = o->renderW; m->h = o->renderH; }
    return PTB_OK;
}
int ptb_mgpu_set_camera(PtbMgpu* m, const PtbCamera* cam) { if (!m) return fail(PTB_ERR_INVALID_ARGUMENT, "null argument"); EACH(ptb_set_camera(c, cam)); return PTB_OK; }
int ptb_mgpu_set_cull(PtbMgpu* m, int32_t enable) { if (!m) return fail(PTB_ERR_INVALID_ARGUMENT, "null argument"); EACH(ptb_set_cull(c, enable)); return PTB_OK; }
int ptb_mgpu_reset_accum(PtbMgpu* m) { if (!m) return fail(PTB_ERR_INVALID_ARGUMENT, "null argument"); EACH(ptb_reset_accum(c)); return PTB_OK; }
int ptb_mgpu_synchronize(PtbMgpu* m) { if (!m) return fail(PTB_ERR_INVALID_ARGUMENT, "null argument"); EACH(ptb_synchronize(c)); return PTB_OK; }
int ptb_mgpu_update_instances(PtbMgpu* m, const float* transforms, int32_t numInstances, const float* materials, int32_t numMaterials, const float* tlasNodes, int32_t numTlasNodes)
{
    if (!m) return fail(PTB_ERR_INVALID_ARGUMENT, "null argument");
    EACH(ptb_update_instances(c, transforms, numInstances, materials, numMaterials, tlasNodes, numTlasNodes));
    return PTB_OK;
}
int ptb_mgpu_rebuild_instances(PtbMgpu* m, const float* transforms, int32_t numInstances, const float* materials, int32_t numMaterials, const int32_t* instanceMaterialIDs, int32_t onHost)
{
    if (!m) return fail(PTB_ERR_INVALID_ARGUMENT, "null argument");
    EACH(ptb_rebuild_instances(c, transforms, numInstances, materials, numMaterials, instanceMaterialIDs, onHost));
    return PTB_OK;
}
int ptb_mgpu_update_envmap(PtbMgpu* m, const float* img, const float* cdf, int32_t w, int32_t h, float totalSum)
{
    if (!m) return fail(PTB_ERR_INVALID_ARGUMENT, "null argument");
    EACH(ptb_update_envmap(c, img, cdf, w, h, totalSum));
    return PTB_OK;
}

// passes [firstSample, firstSample + nSamples) round-robin over the contexts; every GPU gets its launches before any of them is waited for
int ptb_mgpu_render_samples(PtbMgpu* m, int32_t firstSample, int32_t nSamples)
{
    if (!m || firstSample < 1 || nSamples < 0) return fail(PTB_ERR_INVALID_ARGUMENT, "bad sample range");
    const int N = (int)m->ctx.size();
    for (int r = 0; r < N; r++)
    {
        int cnt = nSamples > r ? (nSamples - r + N - 1) / N : 0;
        if (cnt == 0) continue;
        int rc = ptb_render_samples(m->ctx[r], firstSample + r, cnt, N);
        if (rc) return rc;
    }
    return PTB_OK;
}

int ptb_mgpu_render_pass(PtbMgpu* m, int32_t sample, int32_t maxLookahead)
{
    if (!m || sample < 1) return fail(PTB_ERR_INVALID_ARGUMENT, "bad sample pass");
    const int N = (int)m->ctx.size();
    return ptb_render_pass(m->ctx[(sample - 1) % N], sample, N, maxLookahead);
}

// sum of the per-GPU running sums (or, fromFrozen, of their frozen copies) -> scratch on devices[0]; the sources are left untouched
static int reduceToScratch(PtbMgpu* m, const void** devSum, bool fromFrozen = false)
{
    const int N = (int)m->ctx.size();
    void* p0 = nullptr; uint64_t nbytes = 0;
    int rc = ptb_accum_device_ptr(m->ctx[0], &p0, &nbytes);
    if (rc) return rc;
    if (N == 1) { *devSum = nullptr; return PTB_OK; }            // single context: its own running sum
    cudaSetDevice(m->devices[0]);
    if (m->scratchBytes < nbytes)
    {
        if (m->scratch) cudaFree(m->scratch);
        m->scratch = nullptr; m->scratchBytes = 0;
        if (cudaMalloc(&m->scratch, nbytes) != cudaSuccess) return fail(PTB_ERR_OUT_OF_MEMORY, "scratch allocation failed");
        m->scratchBytes = nbytes;
    }
    int r = g_nccl.GroupStart();
    for (int i = 0; i < N && r == NCCL_SUCCESS; i++)
    {
        void* pi = nullptr;
        if (fromFrozen) pi = m->frozen[i]; else ptb_accum_device_ptr(m->ctx[i], &pi, nullptr);
        cudaSetDevice(m->devices[i]);
        r = g_nccl.Reduce(pi, i == 0 ? m->scratch : nullptr, (size_t)(nbytes / 4), NCCL_FLOAT, NCCL_SUM, 0, m->comms[i], m->streams[i]);
    }
    int r2 = g_nccl.GroupEnd();
    if (r == NCCL_SUCCESS) r = r2;
    if (r != NCCL_SUCCESS) return fail(PTB_ERR_CUDA, std::string("ncclReduce: ") + g_nccl.GetErrorString(r));
    *devSum = m->scratch;
    return PTB_OK;
}

int ptb_mgpu_read_output_rgba8(PtbMgpu* m, float invSampleCounter, uint8_t* out)
{
    if (!m || !out) return fail(PTB_ERR_INVALID_ARGUMENT, "null argument");
    const void* sum = nullptr;
    int rc = reduceToScratch(m, &sum);
    if (rc) return rc;
    rc = ptb_read_output_rgba8_from(m->ctx[0], sum, invSampleCounter, out);     // same stream as the reduce on device 0: ordered
    if (rc) return rc;
    for (size_t i = 1; i < m->ctx.size(); i++) { rc = ptb_synchronize(m->ctx[i]); if (rc) return rc; }   // the ranks' send side has completed
    return PTB_OK;
}

int ptb_mgpu_snapshot_output(PtbMgpu* m, float invSampleCounter)
{
    if (!m) return fail(PTB_ERR_INVALID_ARGUMENT, "null argument");
    const int N = (int)m->ctx.size();
    if (N == 1) return ptb_snapshot_output(m->ctx[0], invSampleCounter);
    // Several GPUs: a reduce here would make every GPU wait for every other one once per pass (the drop-in calls this at each pass end while the
    // GPUs trace waves of their own passes).  Freeze each running sum locally instead; resolveFrozen() combines them when the image is read.
    void* p0 = nullptr; uint64_t nbytes = 0;
    int rc = ptb_accum_device_ptr(m->ctx[0], &p0, &nbytes);
    if (rc) return rc;
    if (m->frozen.size() != (size_t)N || m->frozenBytes < nbytes)
    {
        for (size_t i = 0; i < m->frozen.size(); i++) if (m->frozen[i]) { cudaSetDevice(m->devices[i]); cudaFree(m->frozen[i]); }
        m->frozen.assign(N, nullptr); m->frozenBytes = 0;
        for (int i = 0; i < N; i++)
        {
            cudaSetDevice(m->devices[i]);
            if (cudaMalloc(&m->frozen[i], nbytes) != cudaSuccess) return fail(PTB_ERR_OUT_OF_MEMORY, "frozen-sum allocation failed");
        }
        m->frozenBytes = nbytes;
    }
    for (int i = 0; i < N; i++)
    {
        void* pi = nullptr;
        ptb_accum_device_ptr(m->ctx[i], &pi, nullptr);
        cudaSetDevice(m->devices[i]);
        if (cudaMemcpyAsync(m->frozen[i], pi, nbytes, cudaMemcpyDeviceToDevice, m->streams[i]) != cudaSuccess) return fail(PTB_ERR_CUDA, "freezing a running sum failed");
    }
    m->frozenInv = invSampleCounter; m->frozenPending = true;
    return PTB_OK;
}

// combine the frozen sums into the tonemapped snapshot on devices[0] (one ncclReduce), if a newer freeze is pending
static int resolveFrozen(PtbMgpu* m)
{
    if (m->ctx.size() == 1 || !m->frozenPending) return PTB_OK;
    const void* sum = nullptr;
    int rc = reduceToScratch(m, &sum, true);
    if (rc) return rc;
    rc = ptb_snapshot_output_from(m->ctx[0], sum, m->frozenInv);
    if (rc) return rc;
    m->frozenPending = false;
    return PTB_OK;
}

int ptb_mgpu_read_snapshot_rgba8(PtbMgpu* m, uint8_t* out)
{
    if (!m || !out) return fail(PTB_ERR_INVALID_ARGUMENT, "null argument");
    int rc = resolveFrozen(m);
    if (rc) return rc;
    rc = ptb_read_snapshot_rgba8(m->ctx[0], out);
    if (rc) return rc;
    for (size_t i = 1; i < m->ctx.size(); i++) { rc = ptb_synchronize(m->ctx[i]); if (rc) return rc; }     // the other GPUs' send side has completed
    return PTB_OK;
}

int ptb_mgpu_read_snapshot_rgb32f(PtbMgpu* m, float* outRgb)
{
    if (!m || !outRgb) return fail(PTB_ERR_INVALID_ARGUMENT, "null argument");
    int rc = resolveFrozen(m);
    if (rc) return rc;
    rc = ptb_read_snapshot_rgb32f(m->ctx[0], outRgb);
    if (rc) return rc;
    for (size_t i = 1; i < m->ctx.size(); i++) { rc = ptb_synchronize(m->ctx[i]); if (rc) return rc; }
    return PTB_OK;
}

int ptb_mgpu_read_accum_f32(PtbMgpu* m, float* out)
{
    if (!m || !out) return fail(PTB_ERR_INVALID_ARGUMENT, "null argument");
    const void* sum = nullptr;
    int rc = reduceToScratch(m, &sum);
    if (rc) return rc;
    if (!sum) return ptb_read_accum_f32(m->ctx[0], out);
    cudaSetDevice(m->devices[0]);
    cudaError_t e = cudaMemcpyAsync(out, sum, m->scratchBytes < (size_t)m->w * m->h * 16 ? m->scratchBytes : (size_t)m->w * m->h * 16, cudaMemcpyDeviceToHost, m->streams[0]);
    if (e == cudaSuccess) e = cudaStreamSynchronize(m->streams[0]);
    if (e != cudaSuccess) return fail(PTB_ERR_CUDA, cudaGetErrorString(e));
    for (size_t i = 1; i < m->ctx.size(); i++) { rc = ptb_synchronize(m->ctx[i]); if (rc) return rc; }
    return PTB_OK;
}

int ptb_mgpu_get_stats(PtbMgpu* m, PtbStats* out)
{
    if (!m || !out) return fail(PTB_ERR_INVALID_ARGUMENT, "null argument");
    memset(out, 0, sizeof(*out));
    for (size_t i = 0; i < m->ctx.size(); i++)
    {
        PtbStats s; int rc = ptb_get_stats(m->ctx[i], &s);
        if (rc) return rc;
        out->pathSegments += s.pathSegments; out->shadowRays += s.shadowRays; out->samplesRendered += s.samplesRendered; out->kernelLaunches += s.kernelLaunches;
        if (s.lastRenderMs > out->lastRenderMs) out->lastRenderMs = s.lastRenderMs;       // max over GPUs
    }
    return PTB_OK;
}

} // extern "C"
