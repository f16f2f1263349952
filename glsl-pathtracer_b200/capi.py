"""ctypes binding of libptb200.so (include/ptb200.h).  Thin: structs, prototypes, error translation.

The shared library is the product; there is no Python/CPU implementation behind these calls.  If the library or a CUDA
device is missing the calls raise — they never fall back.
"""
from __future__ import annotations
import ctypes as C
import os
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PTB200_LIB") or os.path.join(_HERE, "libptb200.so")     # PTB200_LIB: an A/B build of the same library (scripts/ab_build.sh)

PTB_OK, PTB_ERR_INVALID_ARGUMENT, PTB_ERR_NO_DEVICE, PTB_ERR_CUDA, PTB_ERR_UNSUPPORTED, PTB_ERR_OUT_OF_MEMORY = range(6)


class PtbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"ptb200 error {code}: {msg}")
        self.code = code


class PtbSceneDesc(C.Structure):
    _fields_ = [("nodes", C.c_void_p), ("numNodes", C.c_int32), ("topLevelIndex", C.c_int32),
                ("vertIndices", C.c_void_p), ("numIndices", C.c_int32),
                ("verticesUVX", C.c_void_p), ("numVertices", C.c_int32), ("normalsUVY", C.c_void_p),
                ("materials", C.c_void_p), ("numMaterials", C.c_int32),
                ("transforms", C.c_void_p), ("numInstances", C.c_int32),
                ("lights", C.c_void_p), ("numLights", C.c_int32),
                ("textures", C.c_void_p), ("numTextures", C.c_int32), ("texW", C.c_int32), ("texH", C.c_int32),
                ("envImg", C.c_void_p), ("envCdf", C.c_void_p), ("envW", C.c_int32), ("envH", C.c_int32), ("envTotalSum", C.c_float)]


class PtbOptions(C.Structure):
    _fields_ = [("renderW", C.c_int32), ("renderH", C.c_int32), ("tileW", C.c_int32), ("tileH", C.c_int32),
                ("maxDepth", C.c_int32), ("rrDepth", C.c_int32), ("features", C.c_uint32),
                ("envMapIntensity", C.c_float), ("envMapRot", C.c_float), ("roughnessMollificationAmt", C.c_float),
                ("uniformLightCol", C.c_float * 3), ("backgroundCol", C.c_float * 3),
                ("enableTonemap", C.c_int32), ("enableAces", C.c_int32), ("simpleAcesFit", C.c_int32), ("samplesPerWave", C.c_int32)]


class PtbCamera(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("right", C.c_float * 3), ("up", C.c_float * 3), ("forward", C.c_float * 3),
                ("fov", C.c_float), ("focalDist", C.c_float), ("aperture", C.c_float)]


class PtbStats(C.Structure):
    _fields_ = [("pathSegments", C.c_uint64), ("shadowRays", C.c_uint64), ("samplesRendered", C.c_uint64), ("kernelLaunches", C.c_uint64),
                ("lastRenderMs", C.c_float), ("lastTraceMs", C.c_float), ("lastTraceRays", C.c_uint64),
                ("lastCameraMs", C.c_float), ("lastSortMs", C.c_float), ("lastShadeMs", C.c_float), ("lastShadowMs", C.c_float), ("lastAccumMs", C.c_float),
                ("reserved_", C.c_float)]


HIT_DTYPE = np.dtype([("t", "<f4"), ("kind", "<i4"), ("instance", "<i4"), ("matID", "<i4"), ("primSlot", "<i4"),
                      ("triIDx", "<i4"), ("bary", "<f4", 3), ("lightIdx", "<i4")])
BSDF_QUERY_DTYPE = np.dtype([("mat", "<f4", 32), ("V", "<f4", 3), ("N", "<f4", 3), ("L", "<f4", 3), ("eta", "<f4"),
                             ("r1", "<f4"), ("r2", "<f4"), ("r3", "<f4")])
BSDF_RESULT_DTYPE = np.dtype([("f", "<f4", 3), ("pdf", "<f4"), ("L", "<f4", 3)])

# every symbol include/ptb200.h declares (checked by tests/test_capi_symbols.py against the header text)
SYMBOLS = ["ptb_create", "ptb_destroy", "ptb_last_error", "ptb_derive_features", "ptb_set_options", "ptb_resize", "ptb_set_camera",
           "ptb_update_instances", "ptb_update_envmap", "ptb_reset_accum", "ptb_render_tile", "ptb_render_samples", "ptb_render_preview",
           "ptb_read_accum_f32", "ptb_write_accum_f32", "ptb_accum_device_ptr", "ptb_read_output_rgba8", "ptb_get_stats", "ptb_reset_stats",
           "ptb_set_profiling", "ptb_set_stream", "ptb_synchronize", "ptb_set_cull", "ptb_trace_closest", "ptb_trace_any", "ptb_bsdf_eval",
           "ptb_bsdf_sample", "ptb_lambert_eval", "ptb_lambert_sample", "ptb_camera_rays", "ptb_trace_closest_device", "ptb_read_nodes", "ptb_stack_depth",
           "ptb_render_pass", "ptb_read_output_rgba8_from", "ptb_snapshot_output", "ptb_read_snapshot_rgba8", "ptb_host_alloc", "ptb_host_free",
           "ptb_mgpu_create", "ptb_mgpu_destroy", "ptb_mgpu_num_devices", "ptb_mgpu_context", "ptb_mgpu_set_options", "ptb_mgpu_set_camera", "ptb_mgpu_set_cull",
           "ptb_mgpu_update_instances", "ptb_mgpu_update_envmap", "ptb_mgpu_reset_accum", "ptb_mgpu_render_samples", "ptb_mgpu_read_output_rgba8",
           "ptb_mgpu_read_accum_f32", "ptb_mgpu_get_stats", "ptb_mgpu_synchronize", "ptb_mgpu_render_pass", "ptb_mgpu_snapshot_output", "ptb_mgpu_read_snapshot_rgba8",
           "ptb_snapshot_output_from", "ptb_set_snapshot_float", "ptb_read_snapshot_rgb32f", "ptb_mgpu_read_snapshot_rgb32f", "ptb_rebuild_instances", "ptb_last_rebuild_info", "ptb_mgpu_rebuild_instances"]

_lib = None


def load():
    """dlopen libptb200.so (built in-tree by `make -C glsl-pathtracer_b200/csrc` / __graft_entry__.build())."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise PtbError(-1, f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` (no CPU fallback exists)")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, u32, f32 = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_float
    L.ptb_last_error.restype = C.c_char_p
    L.ptb_derive_features.restype = u32
    L.ptb_derive_features.argtypes = [C.POINTER(PtbSceneDesc), u32]
    protos = {
        "ptb_create": [C.POINTER(PtbSceneDesc), C.POINTER(PtbOptions), C.c_int, C.POINTER(vp)],
        "ptb_destroy": [vp], "ptb_set_options": [vp, C.POINTER(PtbOptions)], "ptb_resize": [vp, i32, i32, i32, i32],
        "ptb_set_camera": [vp, C.POINTER(PtbCamera)],
        "ptb_update_instances": [vp, vp, i32, vp, i32, vp, i32], "ptb_update_envmap": [vp, vp, vp, i32, i32, f32],
        "ptb_reset_accum": [vp], "ptb_render_tile": [vp, i32, i32, i32], "ptb_render_samples": [vp, i32, i32, i32],
        "ptb_render_preview": [vp, i32, i32, vp], "ptb_read_accum_f32": [vp, vp], "ptb_write_accum_f32": [vp, vp],
        "ptb_accum_device_ptr": [vp, C.POINTER(vp), C.POINTER(C.c_uint64)], "ptb_read_output_rgba8": [vp, f32, vp],
        "ptb_get_stats": [vp, C.POINTER(PtbStats)], "ptb_reset_stats": [vp], "ptb_set_profiling": [vp, i32], "ptb_set_stream": [vp, vp],
        "ptb_synchronize": [vp], "ptb_set_cull": [vp, i32], "ptb_trace_closest": [vp, vp, i64, i32, vp], "ptb_trace_any": [vp, vp, vp, i64, vp],
        "ptb_bsdf_eval": [vp, vp, i64, vp], "ptb_bsdf_sample": [vp, vp, i64, vp], "ptb_lambert_eval": [vp, vp, i64, vp], "ptb_lambert_sample": [vp, vp, i64, vp], "ptb_camera_rays": [vp, i32, vp],
        "ptb_trace_closest_device": [vp, vp, i64, i32, vp], "ptb_read_nodes": [vp, vp, i32], "ptb_stack_depth": [vp, C.POINTER(i32)],
        "ptb_render_pass": [vp, i32, i32, i32], "ptb_snapshot_output_from": [vp, vp, f32], "ptb_set_snapshot_float": [vp, i32], "ptb_read_snapshot_rgb32f": [vp, vp], "ptb_mgpu_render_pass": [vp, i32, i32],
        "ptb_mgpu_snapshot_output": [vp, f32], "ptb_mgpu_read_snapshot_rgba8": [vp, vp], "ptb_mgpu_read_snapshot_rgb32f": [vp, vp], "ptb_rebuild_instances": [vp, vp, i32, vp, i32, vp, i32], "ptb_last_rebuild_info": [vp, C.POINTER(i32)],
        "ptb_mgpu_rebuild_instances": [vp, vp, i32, vp, i32, vp, i32], "ptb_read_output_rgba8_from": [vp, vp, f32, vp], "ptb_snapshot_output": [vp, f32], "ptb_read_snapshot_rgba8": [vp, vp],
        "ptb_host_alloc": [C.c_uint64, C.POINTER(vp)], "ptb_host_free": [vp],
        "ptb_mgpu_create": [C.POINTER(PtbSceneDesc), C.POINTER(PtbOptions), C.POINTER(i32), i32, C.POINTER(vp)], "ptb_mgpu_destroy": [vp],
        "ptb_mgpu_num_devices": [vp], "ptb_mgpu_set_options": [vp, C.POINTER(PtbOptions)], "ptb_mgpu_set_camera": [vp, C.POINTER(PtbCamera)],
        "ptb_mgpu_set_cull": [vp, i32], "ptb_mgpu_update_instances": [vp, vp, i32, vp, i32, vp, i32], "ptb_mgpu_update_envmap": [vp, vp, vp, i32, i32, f32],
        "ptb_mgpu_reset_accum": [vp], "ptb_mgpu_render_samples": [vp, i32, i32], "ptb_mgpu_read_output_rgba8": [vp, f32, vp],
        "ptb_mgpu_read_accum_f32": [vp, vp], "ptb_mgpu_get_stats": [vp, C.POINTER(PtbStats)], "ptb_mgpu_synchronize": [vp],
    }
    L.ptb_mgpu_context.restype = vp
    L.ptb_mgpu_context.argtypes = [vp, i32]
    for name, args in protos.items():
        fn = getattr(L, name)
        fn.restype = C.c_int
        fn.argtypes = args
    _lib = L
    return L


def check(rc):
    if rc != PTB_OK:
        raise PtbError(rc, load().ptb_last_error().decode(errors="replace"))


def _ptr(a):
    return a.ctypes.data if a is not None and a.size else None


def scene_desc(scene):
    """PtbSceneDesc over a scene_io.Scene; returns (desc, keepalive)."""
    keep = [np.ascontiguousarray(scene.nodes, np.float32), np.ascontiguousarray(scene.vertIndices, np.int32),
            np.ascontiguousarray(scene.verticesUVX, np.float32), np.ascontiguousarray(scene.normalsUVY, np.float32),
            np.ascontiguousarray(scene.materials, np.float32), np.ascontiguousarray(scene.transforms, np.float32),
            np.ascontiguousarray(scene.lights, np.float32), np.ascontiguousarray(scene.textures, np.uint8)]
    n, vi, vx, nm, mt, tr, lt, tx = keep
    d = PtbSceneDesc()
    d.nodes, d.numNodes, d.topLevelIndex = _ptr(n), len(n), scene.topLevelIndex
    d.vertIndices, d.numIndices = _ptr(vi), len(vi)
    d.verticesUVX, d.numVertices, d.normalsUVY = _ptr(vx), len(vx), _ptr(nm)
    d.materials, d.numMaterials = _ptr(mt), len(mt)
    d.transforms, d.numInstances = _ptr(tr), len(tr)
    d.lights, d.numLights = _ptr(lt), len(lt)
    d.textures, d.numTextures = _ptr(tx), (tx.shape[0] if tx.size else 0)
    d.texW, d.texH = (tx.shape[2], tx.shape[1]) if tx.size else (0, 0)
    if scene.envImg is not None:
        env = (np.ascontiguousarray(scene.envImg, np.float32), np.ascontiguousarray(scene.envCdf, np.float32))
        keep.append(env)
        d.envImg, d.envCdf = _ptr(env[0]), _ptr(env[1])
        d.envH, d.envW = scene.envCdf.shape
        d.envTotalSum = scene.envTotalSum
    return d, keep


def option_bools(ro) -> int:
    """bit order documented at ptb_derive_features in include/ptb200.h"""
    flags = [ro.enableEnvMap, ro.enableRR, ro.enableUniformLight, ro.openglNormalMap, ro.hideEmitters, ro.enableBackground,
             ro.transparentBackground, ro.enableRoughnessMollification, ro.enableVolumeMIS]
    return sum(1 << i for i, f in enumerate(flags) if f)


def make_options(scene, features=None, samples_per_wave=0) -> PtbOptions:
    ro = scene.renderOptions
    o = PtbOptions()
    o.renderW, o.renderH = ro.renderResolution
    o.tileW, o.tileH = ro.tileWidth, ro.tileHeight
    o.maxDepth, o.rrDepth = ro.maxDepth, ro.RRDepth
    if features is None:
        d, keep = scene_desc(scene)
        features = load().ptb_derive_features(C.byref(d), option_bools(ro))
    o.features = features
    o.envMapIntensity, o.envMapRot, o.roughnessMollificationAmt = ro.envMapIntensity, ro.envMapRot, ro.roughnessMollificationAmt
    o.uniformLightCol[:] = ro.uniformLightCol
    o.backgroundCol[:] = ro.backgroundCol
    o.enableTonemap, o.enableAces, o.simpleAcesFit = int(ro.enableTonemap), int(ro.enableAces), int(ro.simpleAcesFit)
    o.samplesPerWave = samples_per_wave
    return o


def make_camera(cam) -> PtbCamera:
    c = PtbCamera()
    c.position[:] = np.asarray(cam.position, np.float32).tolist()
    c.right[:] = np.asarray(cam.right, np.float32).tolist()
    c.up[:] = np.asarray(cam.up, np.float32).tolist()
    c.forward[:] = np.asarray(cam.forward, np.float32).tolist()
    c.fov, c.focalDist, c.aperture = cam.fov, cam.focalDist, cam.aperture
    return c


class Context:
    """One PtbCtx (one GPU).  Methods map 1:1 onto the C ABI."""

    def __init__(self, scene, device=0, features=None, samples_per_wave=0, _borrowed=None):
        L = load()
        self.scene = scene
        self._pinned = None
        self._owned = _borrowed is None
        self.opts = make_options(scene, features, samples_per_wave)
        if _borrowed is not None:           # a per-GPU context owned by an Mgpu handle
            self.h = C.c_void_p(_borrowed)
            return
        d, self._keep = scene_desc(scene)
        h = C.c_void_p()
        check(L.ptb_create(C.byref(d), C.byref(self.opts), device, C.byref(h)))
        self.h = h
        self.set_camera(scene.camera)

    # -- lifetime
    def close(self):
        if getattr(self, "_pinned", None) is not None:
            load().ptb_host_free(self._pinned[0])
            self._pinned = None
        if getattr(self, "h", None):
            if self._owned:
                load().ptb_destroy(self.h)
            self.h = None

    def _pinned_out(self, nbytes):
        """page-locked readback target (one per context, reused): numpy view over ptb_host_alloc memory"""
        if self._pinned is None or self._pinned[1] < nbytes:
            if self._pinned is not None:
                load().ptb_host_free(self._pinned[0])
            p = C.c_void_p()
            check(load().ptb_host_alloc(nbytes, C.byref(p)))
            self._pinned = (p, nbytes, (C.c_uint8 * nbytes).from_address(p.value))
        return self._pinned

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def size(self):
        return self.opts.renderW, self.opts.renderH

    # -- state
    def set_camera(self, cam):
        self._cam = make_camera(cam)
        check(load().ptb_set_camera(self.h, C.byref(self._cam)))

    def set_options(self, opts: PtbOptions):
        self.opts = opts
        check(load().ptb_set_options(self.h, C.byref(opts)))

    def resize(self, w, h, tw, th):
        check(load().ptb_resize(self.h, w, h, tw, th))
        self.opts.renderW, self.opts.renderH, self.opts.tileW, self.opts.tileH = w, h, tw, th

    def update_instances(self, transforms, materials, tlas_nodes):
        t = np.ascontiguousarray(transforms, np.float32); m = np.ascontiguousarray(materials, np.float32)
        n = np.ascontiguousarray(tlas_nodes, np.float32)
        check(load().ptb_update_instances(self.h, _ptr(t), len(t.reshape(-1, 16)), _ptr(m), len(m.reshape(-1, 32)), _ptr(n), len(n.reshape(-1, 9))))

    def rebuild_instances(self, transforms, materials, material_ids=None, on_host=False):
        """Scene::RebuildInstances from the transforms alone: the library rebuilds the TLAS (on the device unless on_host).  Returns where it was built:
        0 device, 1 host as asked, 2 host as the device builder's fallback."""
        t = np.ascontiguousarray(transforms, np.float32); m = np.ascontiguousarray(materials, np.float32)
        ids = None if material_ids is None else np.ascontiguousarray(material_ids, np.int32)
        check(load().ptb_rebuild_instances(self.h, _ptr(t), len(t.reshape(-1, 16)), _ptr(m), len(m.reshape(-1, 32)), None if ids is None else ids.ctypes.data, int(on_host)))
        w, b, t = C.c_int32(), C.c_float(), C.c_float()
        check(load().ptb_last_rebuild_info(self.h, C.byref(w), C.byref(b), C.byref(t)))
        self.last_rebuild = {"where": w.value, "build_ms": b.value, "total_ms": t.value}
        return w.value

    def update_envmap(self, img, cdf, total):
        img = np.ascontiguousarray(img, np.float32); cdf = np.ascontiguousarray(cdf, np.float32)
        check(load().ptb_update_envmap(self.h, _ptr(img), _ptr(cdf), cdf.shape[1], cdf.shape[0], total))

    def set_cull(self, on):
        check(load().ptb_set_cull(self.h, int(on)))

    def set_profiling(self, on):
        check(load().ptb_set_profiling(self.h, int(on)))

    def set_stream(self, stream_ptr):
        check(load().ptb_set_stream(self.h, stream_ptr))

    def synchronize(self):
        check(load().ptb_synchronize(self.h))

    # -- rendering
    def reset_accum(self):
        check(load().ptb_reset_accum(self.h))

    def render_tile(self, tx, ty, frame):
        check(load().ptb_render_tile(self.h, tx, ty, frame))

    def render_samples(self, first, n, stride=1):
        check(load().ptb_render_samples(self.h, first, n, stride))

    def render_preview(self, w, h):
        out = np.zeros((h, w, 4), np.float32)
        check(load().ptb_render_preview(self.h, w, h, out.ctypes.data))
        return out

    def read_accum(self):
        w, h = self.size
        out = np.zeros((h, w, 4), np.float32)
        check(load().ptb_read_accum_f32(self.h, out.ctypes.data))
        return out

    def write_accum(self, a):
        a = np.ascontiguousarray(a, np.float32)
        check(load().ptb_write_accum_f32(self.h, a.ctypes.data))

    def accum_device_ptr(self):
        p = C.c_void_p(); n = C.c_uint64()
        check(load().ptb_accum_device_ptr(self.h, C.byref(p), C.byref(n)))
        return p.value, n.value

    def render_pass(self, sample, max_lookahead=0, stride=1):
        check(load().ptb_render_pass(self.h, sample, stride, max_lookahead))

    def read_output(self, inv_sample_counter, dev_accum=None, pinned=False):
        """tonemapped RGBA8 [h, w, 4], bottom row first.  dev_accum: device pointer of another float4[w*h] sum to tonemap (reduced
        multi-GPU sum).  pinned=True: the copy lands in the context's page-locked buffer and a VIEW of it is returned (valid until
        the next pinned read); default is a fresh pageable array."""
        w, h = self.size
        if pinned:
            p, _, raw = self._pinned_out(w * h * 4)
            check(load().ptb_read_output_rgba8_from(self.h, dev_accum, inv_sample_counter, p))
            return np.frombuffer(raw, np.uint8, w * h * 4).reshape(h, w, 4)
        out = np.zeros((h, w, 4), np.uint8)
        check(load().ptb_read_output_rgba8_from(self.h, dev_accum, inv_sample_counter, out.ctypes.data))
        return out

    def snapshot_output(self, inv_sample_counter):
        check(load().ptb_snapshot_output(self.h, inv_sample_counter))

    def read_snapshot(self):
        w, h = self.size
        out = np.zeros((h, w, 4), np.uint8)
        check(load().ptb_read_snapshot_rgba8(self.h, out.ctypes.data))
        return out

    def stats(self):
        s = PtbStats()
        check(load().ptb_get_stats(self.h, C.byref(s)))
        return {n: getattr(s, n) for n, _ in PtbStats._fields_}

    def reset_stats(self):
        check(load().ptb_reset_stats(self.h))

    # -- parity entry points
    def trace_closest(self, rays, depth=0):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        out = np.zeros(len(rays), HIT_DTYPE)
        check(load().ptb_trace_closest(self.h, _ptr(rays), len(rays), depth, _ptr(out)))
        return out

    def trace_any(self, rays, max_dist):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        md = np.ascontiguousarray(np.broadcast_to(np.asarray(max_dist, np.float32), (len(rays),)))
        out = np.zeros(len(rays), np.int32)
        check(load().ptb_trace_any(self.h, _ptr(rays), _ptr(md), len(rays), _ptr(out)))
        return out

    def bsdf(self, queries, sample=False):
        q = np.ascontiguousarray(queries, BSDF_QUERY_DTYPE)
        out = np.zeros(len(q), BSDF_RESULT_DTYPE)
        fn = load().ptb_bsdf_sample if sample else load().ptb_bsdf_eval
        check(fn(self.h, _ptr(q), len(q), _ptr(out)))
        return out

    def lambert(self, queries, sample=False):
        """lambert.glsl (dead code in the reference's PathTrace): eval, or sample with the query's r1, r2."""
        q = np.ascontiguousarray(queries, BSDF_QUERY_DTYPE)
        out = np.zeros(len(q), BSDF_RESULT_DTYPE)
        fn = load().ptb_lambert_sample if sample else load().ptb_lambert_eval
        check(fn(self.h, _ptr(q), len(q), _ptr(out)))
        return out

    def camera_rays(self, sample=1):
        w, h = self.size
        out = np.zeros((h * w, 6), np.float32)
        check(load().ptb_camera_rays(self.h, sample, out.ctypes.data))
        return out

    def read_nodes(self):
        out = np.zeros_like(np.ascontiguousarray(self.scene.nodes, np.float32))
        check(load().ptb_read_nodes(self.h, out.ctypes.data, len(out)))
        return out

    def stack_depth(self):
        v = C.c_int32()
        check(load().ptb_stack_depth(self.h, C.byref(v)))
        return v.value


class Mgpu:
    """PtbMgpu: N per-GPU contexts of one box behind one handle (one host thread), NCCL reduce into a scratch buffer per readback."""

    def __init__(self, scene, devices=None, num_devices=None, features=None, samples_per_wave=0):
        L = load()
        self.scene = scene
        d, self._keep = scene_desc(scene)
        self.opts = make_options(scene, features, samples_per_wave)
        if devices is None:
            devices = list(range(num_devices or 1))
        arr = (C.c_int32 * len(devices))(*devices)
        h = C.c_void_p()
        check(L.ptb_mgpu_create(C.byref(d), C.byref(self.opts), arr, len(devices), C.byref(h)))
        self.h = h
        self.n = len(devices)
        self.contexts = [Context(scene, features=self.opts.features, samples_per_wave=samples_per_wave, _borrowed=L.ptb_mgpu_context(h, i)) for i in range(self.n)]
        self.set_camera(scene.camera)

    def close(self):
        if getattr(self, "h", None):
            for c in self.contexts:
                c.close()
            load().ptb_mgpu_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def size(self):
        return self.opts.renderW, self.opts.renderH

    def set_camera(self, cam):
        self._cam = make_camera(cam)
        check(load().ptb_mgpu_set_camera(self.h, C.byref(self._cam)))

    def set_options(self, opts):
        self.opts = opts
        check(load().ptb_mgpu_set_options(self.h, C.byref(opts)))

    def set_cull(self, on):
        check(load().ptb_mgpu_set_cull(self.h, int(on)))

    def reset_accum(self):
        check(load().ptb_mgpu_reset_accum(self.h))

    def render_samples(self, first, n):
        check(load().ptb_mgpu_render_samples(self.h, first, n))

    def read_output(self, inv_sample_counter):
        w, h = self.size
        out = np.zeros((h, w, 4), np.uint8)
        check(load().ptb_mgpu_read_output_rgba8(self.h, inv_sample_counter, out.ctypes.data))
        return out

    def read_accum(self):
        w, h = self.size
        out = np.zeros((h, w, 4), np.float32)
        check(load().ptb_mgpu_read_accum_f32(self.h, out.ctypes.data))
        return out

    def render_pass(self, sample, max_lookahead=0):
        check(load().ptb_mgpu_render_pass(self.h, sample, max_lookahead))

    def snapshot_output(self, inv_sample_counter):
        check(load().ptb_mgpu_snapshot_output(self.h, inv_sample_counter))

    def read_snapshot(self):
        w, h = self.size
        out = np.zeros((h, w, 4), np.uint8)
        check(load().ptb_mgpu_read_snapshot_rgba8(self.h, out.ctypes.data))
        return out

    def stats(self):
        s = PtbStats()
        check(load().ptb_mgpu_get_stats(self.h, C.byref(s)))
        return {n: getattr(s, n) for n, _ in PtbStats._fields_}

    def synchronize(self):
        check(load().ptb_mgpu_synchronize(self.h))
