"""Build + ctypes binding of the EXECUTING REFERENCE SHADERS (oracle/glsl_ref).  TEST INFRASTRUCTURE: import only from
tests/, scripts/ and tests/golden generators — never from the product package.

`GlslRef(scene)` compiles (once per `#define OPT_*` set, as Renderer::InitShaders does) the reference's own
tile.glsl / preview.glsl / tonemap.glsl, read in place from /root/reference/src/shaders, into
oracle/_ref/glsl_ref/libglslref_<defines>.so and runs them on the host.  The sources never enter the repo; the built
objects are git-ignored but travel to the GPU box with the snapshot, so a variant built here is usable there.
"""
from __future__ import annotations
import ctypes as C, hashlib, os, subprocess, sys, threading
import numpy as np

from oracle.binding import OrcSceneDesc, OrcOptions, make_options, _ptr

_HERE = os.path.dirname(os.path.abspath(__file__))
_OUT = os.path.join(os.path.dirname(_HERE), "_ref", "glsl_ref")
REFERENCE_SHADERS = os.path.join(os.environ.get("PTB_REFERENCE_DIR", "/root/reference"), "src", "shaders")
_SHADERS = ("tile", "preview", "tonemap")
_LIBS = {}


def defines_for(o: OrcOptions) -> list[str]:
    """The define list Renderer::InitShaders would prepend for these options (Renderer.cpp:401-459)."""
    d = []
    if o.optEnvMap: d.append("OPT_ENVMAP")
    if o.optLights: d.append("OPT_LIGHTS")
    if o.optRR: d += ["OPT_RR", "OPT_RR_DEPTH=%d" % o.rrDepth]
    if o.optUniformLight: d.append("OPT_UNIFORM_LIGHT")
    if o.optOpenglNormalMap: d.append("OPT_OPENGL_NORMALMAP")
    if o.optHideEmitters: d.append("OPT_HIDE_EMITTERS")
    if o.optBackground: d.append("OPT_BACKGROUND")
    if o.optTransparentBackground: d.append("OPT_TRANSPARENT_BACKGROUND")
    if o.optAlphaTest: d.append("OPT_ALPHA_TEST")
    if o.optRoughnessMollification: d.append("OPT_ROUGHNESS_MOLLIFICATION")
    if o.optMedium: d.append("OPT_MEDIUM")
    if o.optVolMis: d.append("OPT_VOL_MIS")
    return d


def variant_path(defines: list[str]) -> str:
    tag = hashlib.sha1(" ".join(defines).encode()).hexdigest()[:12]
    return os.path.join(_OUT, "libglslref_%s.so" % tag)


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_SHADERS, "tile.glsl"))


def build_variant(defines: list[str], force: bool = False) -> str:
    so = variant_path(defines)
    own = [os.path.join(_HERE, f) for f in ("glsl_ref.cpp", "glsl_compat.h", "glsl2cpp.py")] + [os.path.join(os.path.dirname(_HERE), "pt_oracle.h")]
    if not reference_available():
        if os.path.exists(so):
            return so                      # prebuilt object that travelled with the snapshot (GPU box)
        raise FileNotFoundError("reference shaders not found under %s and no prebuilt %s" % (REFERENCE_SHADERS, so))
    newest = max(os.path.getmtime(p) for p in own)
    if not force and os.path.exists(so) and os.path.getmtime(so) >= newest:
        return so
    os.makedirs(_OUT, exist_ok=True)
    incs = {}
    for s in _SHADERS:
        incs[s] = os.path.join(_OUT, s + ".inc")
        tmp = "%s.%d.%d.tmp" % (incs[s], os.getpid(), threading.get_ident())      # concurrent builds: readers only ever see a whole file
        subprocess.check_call([sys.executable, os.path.join(_HERE, "glsl2cpp.py"), REFERENCE_SHADERS, s + ".glsl", tmp])
        os.replace(tmp, incs[s])
    cxx = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"
    cmd = [cxx, "-std=c++20", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", "-Wall", "-Wno-unused", "-Wno-deprecated-copy",
           "-I", _HERE] + ['-D%s_INC="%s"' % (s.upper(), incs[s]) for s in _SHADERS] + ["-D" + d for d in defines] + \
          [os.path.join(_HERE, "glsl_ref.cpp"), "-o", so + ".tmp"]
    subprocess.check_call(cmd)
    os.replace(so + ".tmp", so)
    with open(so[:-3] + ".defines", "w") as f:
        f.write(" ".join(defines) + "\n")
    return so


def _load(defines):
    key = tuple(defines)
    if key not in _LIBS:
        L = C.CDLL(build_variant(defines))
        L.gref_create.argtypes = [C.POINTER(OrcSceneDesc), C.POINTER(OrcOptions)]; L.gref_create.restype = C.c_int
        L.gref_set_options.argtypes = [C.POINTER(OrcOptions)]; L.gref_set_options.restype = C.c_int
        L.gref_render_tile.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_void_p]
        L.gref_render_samples.argtypes = [C.c_int32, C.c_int32, C.c_void_p]
        L.gref_render_samples_rect.argtypes = [C.c_int32] * 6 + [C.c_void_p]
        L.gref_render_preview.argtypes = [C.c_int32, C.c_int32, C.c_void_p]
        L.gref_tonemap.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p]
        L.gref_num_threads.restype = C.c_int
        L.gref_trace_closest.argtypes = [C.c_void_p, C.c_int64, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p]
        L.gref_trace_any.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.gref_bsdf_eval.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.gref_lambert_eval.argtypes = [C.c_void_p, C.c_int64, C.c_void_p]
        L.gref_lambert_sample.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p]
        _LIBS[key] = L
    return _LIBS[key]


class GlslRef:
    """The reference shaders bound to one scene.  Same call surface as oracle.binding.Oracle for the render entry points."""

    def __init__(self, scene, **optkw):
        optkw.pop("cull", None)             # the shader text has exactly one traversal: unculled
        self.scene = scene
        self.opts = make_options(scene, **optkw)
        self.L = _load(defines_for(self.opts))
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
            env = (np.ascontiguousarray(scene.envImg), np.ascontiguousarray(scene.envCdf))
            self._keep += list(env)
            d.envImg, d.envCdf = _ptr(env[0]), _ptr(env[1])
            d.envH, d.envW = scene.envCdf.shape; d.envTotalSum = scene.envTotalSum
        if self.L.gref_create(C.byref(d), C.byref(self.opts)) != 0:
            raise RuntimeError("glsl_ref variant does not match the option set")

    def set_options(self, **kw):
        for k, v in kw.items():
            setattr(self.opts, k, v)
        if self.L.gref_set_options(C.byref(self.opts)) != 0:
            raise RuntimeError("option change needs another shader variant (compile-time define): build a new GlslRef")

    def close(self):
        pass

    @property
    def size(self):
        return self.opts.renderW, self.opts.renderH

    def render(self, first_sample=1, n_samples=1, accum=None, rect=None):
        w, h = self.size
        if accum is None:
            accum = np.zeros((h, w, 4), np.float32)
        if rect is None:
            self.L.gref_render_samples(first_sample, n_samples, accum.ctypes.data)
        else:
            self.L.gref_render_samples_rect(first_sample, n_samples, *rect, accum.ctypes.data)
        return accum

    def render_tile(self, tx, ty, frame, accum):
        self.L.gref_render_tile(tx, ty, frame, accum.ctypes.data)
        return accum

    def render_preview(self, w, h):
        out = np.zeros((h, w, 4), np.float32)
        self.L.gref_render_preview(w, h, out.ctypes.data)
        return out

    def tonemap(self, accum, inv_sample_counter, ro):
        h, w, _ = accum.shape
        out = np.zeros((h, w, 4), np.uint8)
        bg = np.asarray(ro.backgroundCol, np.float32)
        a = np.ascontiguousarray(accum, np.float32)
        self.L.gref_tonemap(a.ctypes.data, w, h, inv_sample_counter, int(ro.enableTonemap), int(ro.enableAces), int(ro.simpleAcesFit),
                            bg.ctypes.data, out.ctypes.data)
        return out

    def trace_closest(self, rays, depth=0):
        """The shader's own ClosestHit: (t, kind 0 miss / 1 triangle / 2 light, matID)."""
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        t = np.zeros(len(rays), np.float32); kind = np.zeros(len(rays), np.int32); mat = np.zeros(len(rays), np.int32)
        self.L.gref_trace_closest(rays.ctypes.data, len(rays), depth, t.ctypes.data, kind.ctypes.data, mat.ctypes.data)
        return t, kind, mat

    def trace_any(self, rays, max_dist):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        md = np.ascontiguousarray(np.broadcast_to(np.asarray(max_dist, np.float32), (len(rays),)))
        out = np.zeros(len(rays), np.int32)
        self.L.gref_trace_any(rays.ctypes.data, md.ctypes.data, len(rays), out.ctypes.data)
        return out

    def bsdf_eval(self, queries):
        from oracle.binding import BSDF_QUERY_DTYPE, BSDF_RESULT_DTYPE
        q = np.ascontiguousarray(queries, BSDF_QUERY_DTYPE)
        out = np.zeros(len(q), BSDF_RESULT_DTYPE)
        self.L.gref_bsdf_eval(q.ctypes.data, len(q), out.ctypes.data)
        return out

    def lambert_eval(self, queries):
        from oracle.binding import BSDF_QUERY_DTYPE, BSDF_RESULT_DTYPE
        q = np.ascontiguousarray(queries, BSDF_QUERY_DTYPE)
        out = np.zeros(len(q), BSDF_RESULT_DTYPE)
        self.L.gref_lambert_eval(q.ctypes.data, len(q), out.ctypes.data)
        return out

    def lambert_sample(self, queries, seeds):
        """The shader's LambertSample with its own RNG seeded per query; returns (results, the two rand() draws it consumed)."""
        from oracle.binding import BSDF_QUERY_DTYPE, BSDF_RESULT_DTYPE
        q = np.ascontiguousarray(queries, BSDF_QUERY_DTYPE)
        seeds = np.ascontiguousarray(seeds, np.uint32).reshape(len(q), 4)
        out = np.zeros(len(q), BSDF_RESULT_DTYPE); r12 = np.zeros((len(q), 2), np.float32)
        self.L.gref_lambert_sample(q.ctypes.data, len(q), seeds.ctypes.data, out.ctypes.data, r12.ctypes.data)
        return out, r12

    def num_threads(self):
        return self.L.gref_num_threads()
