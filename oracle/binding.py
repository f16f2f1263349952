"""ctypes binding of the CPU oracle (oracle/libpt_oracle.so).  TEST INFRASTRUCTURE: import only from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs."""
from __future__ import annotations
import ctypes as C, os, subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class OrcSceneDesc(C.Structure):
    _fields_ = [("nodes", C.c_void_p), ("numNodes", C.c_int32), ("topLevelIndex", C.c_int32),
                ("vertIndices", C.c_void_p), ("numIndices", C.c_int32),
                ("verticesUVX", C.c_void_p), ("numVertices", C.c_int32), ("normalsUVY", C.c_void_p),
                ("materials", C.c_void_p), ("numMaterials", C.c_int32),
                ("transforms", C.c_void_p), ("numInstances", C.c_int32),
                ("lights", C.c_void_p), ("numLights", C.c_int32),
                ("textures", C.c_void_p), ("numTextures", C.c_int32), ("texW", C.c_int32), ("texH", C.c_int32),
                ("envImg", C.c_void_p), ("envCdf", C.c_void_p), ("envW", C.c_int32), ("envH", C.c_int32), ("envTotalSum", C.c_float)]


class OrcOptions(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("optEnvMap", "optLights", "optRR", "rrDepth", "optUniformLight", "optOpenglNormalMap",
                                         "optHideEmitters", "optBackground", "optTransparentBackground", "optAlphaTest",
                                         "optRoughnessMollification", "optMedium", "optVolMis", "maxDepth")] + \
               [("envMapIntensity", C.c_float), ("envMapRot", C.c_float), ("roughnessMollificationAmt", C.c_float),
                ("uniformLightCol", C.c_float * 3),
                ("renderW", C.c_int32), ("renderH", C.c_int32), ("tileW", C.c_int32), ("tileH", C.c_int32),
                ("camPosition", C.c_float * 3), ("camRight", C.c_float * 3), ("camUp", C.c_float * 3), ("camForward", C.c_float * 3),
                ("camFov", C.c_float), ("camFocalDist", C.c_float), ("camAperture", C.c_float), ("cullBoxes", C.c_int32)]


HIT_DTYPE = np.dtype([("t", "<f4"), ("kind", "<i4"), ("instance", "<i4"), ("matID", "<i4"), ("primSlot", "<i4"),
                      ("triIDx", "<i4"), ("bary", "<f4", 3), ("lightIdx", "<i4")])
BSDF_QUERY_DTYPE = np.dtype([("mat", "<f4", 32), ("V", "<f4", 3), ("N", "<f4", 3), ("L", "<f4", 3), ("eta", "<f4"),
                             ("r1", "<f4"), ("r2", "<f4"), ("r3", "<f4")])
BSDF_RESULT_DTYPE = np.dtype([("f", "<f4", 3), ("pdf", "<f4"), ("L", "<f4", 3)])


class OrcStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("closestRays", "anyRays", "nodeVisits", "internalSteps", "triTests", "tlasLeaves", "surfaceHits",
                                         "anyNodeVisits", "anyInternalSteps", "anyTriTests", "anyTlasLeaves")]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "libpt_oracle.so")
    src = os.path.join(_HERE, "pt_oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(os.path.join(_HERE, "pt_oracle.h"))):
        subprocess.check_call(["make", "-C", _HERE, "libpt_oracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        L.orc_create.restype = C.c_void_p
        L.orc_create.argtypes = [C.POINTER(OrcSceneDesc), C.POINTER(OrcOptions)]
        L.orc_destroy.argtypes = [C.c_void_p]
        L.orc_set_options.argtypes = [C.c_void_p, C.POINTER(OrcOptions)]
        L.orc_trace_closest.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]
        L.orc_trace_closest_brute.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]
        L.orc_trace_any.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_bsdf_eval.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_bsdf_sample.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.orc_lambert.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]
        L.orc_capture_rays.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        L.orc_capture_shadow_rays.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        L.orc_camera_rays.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.orc_render_samples.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.orc_render_samples_rect.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
        L.orc_render_tile.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
        L.orc_render_preview.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.orc_tonemap.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.orc_get_stats.argtypes = [C.c_void_p, C.POINTER(OrcStats)]
        L.orc_reset_stats.argtypes = [C.c_void_p]
        L.orc_num_threads.restype = C.c_int
        L.orc_set_num_threads.argtypes = [C.c_int]; L.orc_set_num_threads.restype = None
    return _LIB


def _ptr(a):
    return a.ctypes.data if a is not None and a.size else None


def make_options(scene, features: int | None = None, cull: bool = False, **over) -> OrcOptions:
    """OrcOptions from a scene_io.Scene: feature defines per Renderer.cpp:401-459, uniforms per Renderer.cpp:766-811."""
    from glsl_pathtracer_b200 import scene_io as sio
    ro, cam = scene.renderOptions, scene.camera
    f = sio.derive_features(scene) if features is None else features
    o = OrcOptions()
    o.optEnvMap = int(bool(f & sio.OPT_ENVMAP)); o.optLights = int(bool(f & sio.OPT_LIGHTS)); o.optRR = int(bool(f & sio.OPT_RR))
    o.rrDepth = ro.RRDepth; o.optUniformLight = int(bool(f & sio.OPT_UNIFORM_LIGHT))
    o.optOpenglNormalMap = int(bool(f & sio.OPT_OPENGL_NORMALMAP)); o.optHideEmitters = int(bool(f & sio.OPT_HIDE_EMITTERS))
    o.optBackground = int(bool(f & sio.OPT_BACKGROUND)); o.optTransparentBackground = int(bool(f & sio.OPT_TRANSPARENT_BACKGROUND))
    o.optAlphaTest = int(bool(f & sio.OPT_ALPHA_TEST)); o.optRoughnessMollification = int(bool(f & sio.OPT_ROUGHNESS_MOLLIFICATION))
    o.optMedium = int(bool(f & sio.OPT_MEDIUM)); o.optVolMis = int(bool(f & sio.OPT_VOL_MIS))
    o.maxDepth = ro.maxDepth
    o.envMapIntensity = ro.envMapIntensity; o.envMapRot = np.float32(ro.envMapRot) / np.float32(360.0)
    o.roughnessMollificationAmt = ro.roughnessMollificationAmt
    o.uniformLightCol[:] = ro.uniformLightCol
    o.renderW, o.renderH = ro.renderResolution; o.tileW, o.tileH = ro.tileWidth, ro.tileHeight
    o.camPosition[:] = cam.position.tolist(); o.camRight[:] = cam.right.tolist(); o.camUp[:] = cam.up.tolist(); o.camForward[:] = cam.forward.tolist()
    o.camFov = cam.fov; o.camFocalDist = cam.focalDist; o.camAperture = cam.aperture
    o.cullBoxes = int(cull)
    for k, v in over.items():
        setattr(o, k, v)
    return o


class Oracle:
    def __init__(self, scene, **optkw):
        self.scene = scene
        self._keep = [np.ascontiguousarray(a) for a in (scene.nodes, scene.vertIndices, scene.verticesUVX, scene.normalsUVY,
                                                          scene.materials, scene.transforms, scene.lights, scene.textures)]
        n, vi, vx, nm, mt, tr, lt, tx = self._keep
        d = OrcSceneDesc()
        d.nodes, d.numNodes, d.topLevelIndex = _ptr(n), len(n), scene.topLevelIndex
        d.vertIndices, d.numIndices = _ptr(vi), len(vi)
        d.verticesUVX, d.numVertices, d.normalsUVY = _ptr(vx), len(vx), _ptr(nm)
        d.materials, d.numMaterials = _ptr(mt), len(mt)
        d.transforms, d.numInstances = _ptr(tr), len(tr)
        d.lights, d.numLights = _ptr(lt), len(lt)
        d.textures, d.numTextures = _ptr(tx), (tx.shape[0] if tx.size else 0)
        d.texW, d.texH = (tx.shape[2], tx.shape[1]) if tx.size else (0, 0)
        if scene.envImg is not None:
            self._env = (np.ascontiguousarray(scene.envImg), np.ascontiguousarray(scene.envCdf))
            d.envImg, d.envCdf = _ptr(self._env[0]), _ptr(self._env[1])
            d.envH, d.envW = scene.envCdf.shape; d.envTotalSum = scene.envTotalSum
        self.opts = make_options(scene, **optkw)
        self.h = lib().orc_create(C.byref(d), C.byref(self.opts))

    def close(self):
        if self.h:
            lib().orc_destroy(self.h); self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_options(self, **kw):
        for k, v in kw.items():
            setattr(self.opts, k, v)
        lib().orc_set_options(self.h, C.byref(self.opts))

    @property
    def size(self):
        return self.opts.renderW, self.opts.renderH

    def trace_closest(self, rays, depth=0, brute=False):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        out = np.zeros(len(rays), HIT_DTYPE)
        (lib().orc_trace_closest_brute if brute else lib().orc_trace_closest)(self.h, _ptr(rays), len(rays), depth, out.ctypes.data)
        return out

    def trace_any(self, rays, maxDist):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        md = np.ascontiguousarray(np.broadcast_to(np.asarray(maxDist, np.float32), (len(rays),)))
        out = np.zeros(len(rays), np.int32)
        lib().orc_trace_any(self.h, _ptr(rays), _ptr(md), len(rays), out.ctypes.data)
        return out

    def bsdf(self, queries, sample=False):
        q = np.ascontiguousarray(queries, BSDF_QUERY_DTYPE)
        out = np.zeros(len(q), BSDF_RESULT_DTYPE)
        (lib().orc_bsdf_sample if sample else lib().orc_bsdf_eval)(self.h, q.ctypes.data, len(q), out.ctypes.data)
        return out

    def lambert(self, queries, sample=False):
        q = np.ascontiguousarray(queries, BSDF_QUERY_DTYPE)
        out = np.zeros(len(q), BSDF_RESULT_DTYPE)
        lib().orc_lambert(self.h, q.ctypes.data, len(q), int(sample), out.ctypes.data)
        return out

    def camera_rays(self, sample=1):
        w, h = self.size
        rays = np.zeros((h * w, 6), np.float32)
        lib().orc_camera_rays(self.h, sample, rays.ctypes.data)
        return rays

    def capture_rays(self, sample=1, depth=1):
        """(rays (h, w, 6), valid (h, w)): the ray traced at path-loop depth `depth` (analysis aid for scripts/simd_sim.py)."""
        w, h = self.size
        rays = np.zeros((h, w, 6), np.float32); valid = np.zeros((h, w), np.uint8)
        lib().orc_capture_rays(self.h, sample, depth, rays.ctypes.data, valid.ctypes.data)
        return rays, valid.astype(bool)

    def capture_shadow_rays(self, sample=1, depth=0):
        """(rays (h, w, 8) = origin, direction, maxDist, light index; valid (h, w)): the light-NEE shadow ray traced at loop depth `depth`."""
        w, h = self.size
        rays = np.zeros((h, w, 8), np.float32); valid = np.zeros((h, w), np.uint8)
        lib().orc_capture_shadow_rays(self.h, sample, depth, rays.ctypes.data, valid.ctypes.data)
        return rays, valid.astype(bool)

    def render(self, first_sample=1, n_samples=1, accum=None, rect=None):
        w, h = self.size
        if accum is None:
            accum = np.zeros((h, w, 4), np.float32)
        if rect is None:
            lib().orc_render_samples(self.h, first_sample, n_samples, accum.ctypes.data)
        else:
            lib().orc_render_samples_rect(self.h, first_sample, n_samples, *rect, accum.ctypes.data)
        return accum

    def render_preview(self, w, h):
        out = np.zeros((h, w, 4), np.float32)
        lib().orc_render_preview(self.h, w, h, out.ctypes.data)
        return out

    def render_tile(self, tx, ty, frame, accum):
        lib().orc_render_tile(self.h, tx, ty, frame, accum.ctypes.data)
        return accum

    def stats(self, reset=False):
        s = OrcStats(); lib().orc_get_stats(self.h, C.byref(s))
        if reset:
            lib().orc_reset_stats(self.h)
        return {n: getattr(s, n) for n, _ in OrcStats._fields_}


def tonemap(accum, inv_sample_counter, ro, features=0):
    from glsl_pathtracer_b200 import scene_io as sio
    h, w, _ = accum.shape
    out = np.zeros((h, w, 4), np.uint8)
    bg = np.asarray(ro.backgroundCol, np.float32)
    lib().orc_tonemap(np.ascontiguousarray(accum, np.float32).ctypes.data, w, h, inv_sample_counter, int(ro.enableTonemap), int(ro.enableAces),
                      int(ro.simpleAcesFit), bg.ctypes.data, int(bool(features & sio.OPT_BACKGROUND)),
                      int(bool(features & sio.OPT_TRANSPARENT_BACKGROUND)), out.ctypes.data)
    return out
