"""ctypes binding of tests/host_harness/libtrav_host.so: the product's traversal / camera-ray SOURCE compiled for the host
(TEST INFRASTRUCTURE, see ptb_host_shim.h).  `HostTrav` offers the trace_closest / trace_any / camera_rays calls of capi.Context."""
import ctypes as C
import os, subprocess
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        subprocess.check_call(["make", "-s", "-C", _HERE])          # rebuilds when the product sources changed
        L = C.CDLL(os.path.join(_HERE, "libtrav_host.so"))
        vp, i32, f32 = C.c_void_p, C.c_int32, C.c_float
        L.hh_create.restype = vp
        L.hh_create.argtypes = [vp, i32, i32, vp, i32, vp, i32, vp, i32, i32, vp, i32, C.c_char_p, i32]
        L.hh_destroy.argtypes = [vp]
        L.hh_stack_depth.argtypes = [vp]
        L.hh_trace_closest.argtypes = [vp, vp, C.c_longlong, i32, i32, vp]
        L.hh_trace_any.argtypes = [vp, vp, vp, C.c_longlong, i32, i32, i32, vp, vp]
        L.hh_wide_nodes.argtypes = [vp]
        L.hh_any_stack_high.argtypes = [vp]
        L.hh_any_bytes.argtypes = [vp]; L.hh_any_bytes.restype = C.c_double
        L.hh_build_tlas.argtypes = [vp, i32, i32, vp, i32, vp, vp, vp]
        L.hh_camera_rays.argtypes = [i32, i32, i32, i32, vp, vp, vp, vp, f32, f32, f32, i32, i32, vp]
        L.hh_slot_map.argtypes = [i32, i32, i32, i32, i32, vp, vp]
        L.hh_env_search.argtypes = [vp, i32, i32, f32, vp, i32, vp, vp]; L.hh_env_search.restype = i32
        _LIB = L
    return _LIB


class HostTrav:
    def __init__(self, scene):
        from glsl_pathtracer_b200 import capi
        self.scene = scene
        k = [np.ascontiguousarray(scene.nodes, np.float32), np.ascontiguousarray(scene.vertIndices, np.int32), np.ascontiguousarray(scene.verticesUVX, np.float32),
             np.ascontiguousarray(scene.transforms, np.float32), np.ascontiguousarray(scene.lights, np.float32)]
        self._keep = k
        err = C.create_string_buffer(256)
        p = lambda a: a.ctypes.data if a.size else None
        self.h = lib().hh_create(p(k[0]), len(k[0]), scene.topLevelIndex, p(k[1]), len(k[1]), p(k[2]), len(k[2]), p(k[3]), len(k[3].reshape(-1, 16)),
                                 len(scene.materials), p(k[4]), len(k[4]), err, 256)
        if not self.h:
            raise RuntimeError(err.value.decode())
        self.hit_dtype = capi.HIT_DTYPE
        self.cull = True
        self.lights = len(scene.lights) > 0

    def set_cull(self, on):
        self.cull = bool(on)

    def trace_closest(self, rays, depth=0):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        out = np.zeros(len(rays), self.hit_dtype)
        hide = bool(self.scene.renderOptions.hideEmitters)
        lib().hh_trace_closest(self.h, rays.ctypes.data, len(rays), int(self.lights and (not hide or depth > 0)), int(self.cull), out.ctypes.data)
        return out

    def trace_any(self, rays, max_dist, wide=True):
        """wide=True: the production any-hit path (4-wide hierarchy where admissible); self.fallbacks = rays it handed back to the binary traversal"""
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 6)
        md = np.ascontiguousarray(np.broadcast_to(np.asarray(max_dist, np.float32), (len(rays),)))
        out = np.zeros(len(rays), np.int32)
        fb = C.c_longlong(0)
        lib().hh_trace_any(self.h, rays.ctypes.data, md.ctypes.data, len(rays), int(self.lights), int(self.cull), int(wide), out.ctypes.data, C.byref(fb))
        self.fallbacks = fb.value
        return out

    def wide_nodes(self):
        return lib().hh_wide_nodes(self.h)

    def any_bytes(self):
        """bytes the last trace_any call fetched through the read-only path (work measure)"""
        return lib().hh_any_bytes(self.h)

    def any_stack(self):
        """(deepest any-hit stack seen so far, entries the kernels reserve)"""
        return lib().hh_any_stack_high(self.h), lib().hh_stack_depth(self.h)

    def camera_rays(self, sample=1, tables=True):
        ro, cam = self.scene.renderOptions, self.scene.camera
        w, h = ro.renderResolution
        out = np.zeros((w * h, 6), np.float32)
        v = [np.ascontiguousarray(x, np.float32) for x in (cam.position, cam.right, cam.up, cam.forward)]
        lib().hh_camera_rays(w, h, ro.tileWidth, ro.tileHeight, v[0].ctypes.data, v[1].ctypes.data, v[2].ctypes.data, v[3].ctypes.data,
                             float(cam.fov), float(cam.focalDist), float(cam.aperture), sample, int(tables), out.ctypes.data)
        return out

    def stack_depth(self):
        return lib().hh_stack_depth(self.h)

    def close(self):
        if self.h:
            lib().hh_destroy(self.h); self.h = None


def build_tlas(nodes, top, transforms, material_ids=None):
    """ptbd_build_tlas_host on a scene's arrays: returns (TLAS slice [2 * numInstances, 9] float32, height).  material_ids: per-instance material ids
    (default: those of the current TLAS leaves)."""
    nodes = np.ascontiguousarray(nodes, np.float32).reshape(-1, 9); tr = np.ascontiguousarray(transforms, np.float32).reshape(-1, 16)
    out = np.zeros((2 * len(tr), 9), np.float32); h = C.c_int(0)
    mid = None if material_ids is None else np.ascontiguousarray(material_ids, np.int32)
    rc = lib().hh_build_tlas(nodes.ctypes.data, len(nodes), top, tr.ctypes.data, len(tr), None if mid is None else mid.ctypes.data, out.ctypes.data, C.byref(h))
    if rc:
        raise RuntimeError(f"ptbd_build_tlas_host failed: {rc}")
    return out, h.value


def env_search(cdf, w, h, total_sum, values):
    """(has_guide, uv with the guide table, uv of the reference's two binary searches) for the product's envBinarySearch compiled for the host"""
    cdf = np.ascontiguousarray(cdf, np.float32).ravel(); values = np.ascontiguousarray(values, np.float32)
    fast = np.zeros((len(values), 2), np.float32); ref = np.zeros((len(values), 2), np.float32)
    have = lib().hh_env_search(cdf.ctypes.data, w, h, float(total_sum), values.ctypes.data, len(values), fast.ctypes.data, ref.ctypes.data)
    return bool(have), fast, ref


def slot_map(w, h, n_samples, block_major=True, max_lps=5):
    """(rows of {pass, px, py, slotOfSample(pixel, pass)} per slot of the padded wave, passes per warp group as log2)"""
    vw, vh = (w + 7) & ~7, (h + 3) & ~3
    out = np.zeros((vw * vh * n_samples, 4), np.int32); lps = C.c_int32(0)
    lib().hh_slot_map(w, h, n_samples, int(block_major), max_lps, out.ctypes.data, C.byref(lps))
    return out, lps.value
