"""Debug aid: which feature of gltf_mix makes the CUDA path diverge from the oracle."""
import sys, os, copy, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import glsl_pathtracer_b200
from glsl_pathtracer_b200 import capi
from conftest import scene_at, rel_mse
from oracle import binding as ob

def run(tag, mod, depth=2, spp=4):
    sc = scene_at("gltf_mix", 240, 136, 64, 36, depth)
    sc.materials = sc.materials.copy()
    mod(sc)
    ctx = capi.Context(sc); orc = ob.Oracle(sc)
    ctx.render_samples(1, spp); g = ctx.read_accum(); o = orc.render(1, spp)
    d = np.abs(g[..., :3] - o[..., :3]).max(-1); rel = d / (np.abs(o[..., :3]).max(-1) + 1e-4)
    rays = orc.camera_rays(1); hit = orc.trace_closest(rays, 0)
    mat = np.where(hit["kind"] == 1, hit["matID"], -hit["kind"] - 1).reshape(rel.shape)
    per = {int(m): round(float((rel[mat == m] > 1e-3).mean()), 3) for m in np.unique(mat)}
    np.save(f"gpurun_out/dbgv_{tag.replace(' ', '_').replace(',', '')}_g.npy", g); np.save(f"gpurun_out/dbgv_{tag.replace(' ', '_').replace(',', '')}_o.npy", o)
    print("   per first-hit material:", per)
    print(f"{tag:34s} relMSE {rel_mse(o / spp, g / spp):.3g} bad {(rel > 1e-3).mean():.4f} seg {ctx.stats()['pathSegments']} {orc.stats()['closestRays']} shadow {ctx.stats()['shadowRays']} {orc.stats()['anyRays']}")
    ctx.close(); orc.close()

def noop(sc): pass
def opaque(sc): sc.materials[:, 29] = 0
def no_blend(sc): sc.materials[sc.materials[:, 29] == 1, 29] = 0
def no_mask(sc): sc.materials[sc.materials[:, 29] == 2, 29] = 0
def no_normalmap(sc): sc.materials[:, 26] = -1
def no_mr(sc): sc.materials[:, 25] = -1
def no_emtex(sc): sc.materials[:, 27] = -1
def no_albedo(sc): sc.materials[:, 24] = -1
def no_tex(sc): sc.materials[:, 24:28] = -1
def no_lights(sc): sc.lights = sc.lights[:0]
def no_env(sc): sc.renderOptions.enableEnvMap = False
def no_glass(sc): sc.materials[:, 16] = 0
def no_emission(sc): sc.materials[:, 4:7] = 0
def g2_notex(sc): sc.materials[7:13, 24:28] = -1
def g1_notex(sc): sc.materials[1:7, 24:28] = -1
def only_layer0(sc):
    t = sc.materials[:, 24:28]; t[t >= 0] = 0
def layers_mod4(sc):
    t = sc.materials[:, 24:28]; t[t >= 4] -= 4
def g2_nonormal(sc): sc.materials[7:13, 26] = -1
def only_albedo(sc): sc.materials[:, 25:28] = -1
def only_albedo_opaque(sc): sc.materials[:, 25:28] = -1; sc.materials[:, 29] = 0
def only_albedo_noalpha_tex(sc):
    sc.materials[:, 25:28] = -1
    sc.textures = sc.textures.copy(); sc.textures[..., 3] = 255
def only_normal(sc): sc.materials[:, 24:26] = -1; sc.materials[:, 27] = -1
def only_mr(sc): sc.materials[:, 24] = -1; sc.materials[:, 26:28] = -1
def only_em(sc): sc.materials[:, 24:27] = -1
def const_tex(sc):
    only_albedo_opaque(sc)
    sc.textures = sc.textures.copy(); sc.textures[:] = (200, 120, 60, 255)
def smooth_tex(sc):
    only_albedo_opaque(sc)
    sc.textures = sc.textures.copy()
    h, w = sc.textures.shape[1:3]
    y, x = np.mgrid[0:h, 0:w]
    sc.textures[:, ..., 0] = (127 + 100 * np.sin(x / w * 2 * np.pi)).astype(np.uint8); sc.textures[:, ..., 1] = (127 + 100 * np.cos(y / h * 2 * np.pi)).astype(np.uint8)
    sc.textures[:, ..., 2] = 128; sc.textures[:, ..., 3] = 255
def diffuse_all(sc):
    only_albedo_opaque(sc)
    sc.materials[:, 8] = 0; sc.materials[:, 9] = 0.7; sc.materials[:, 16] = 0
for tag, f in [("as is", noop), ("all opaque", opaque), ("no BLEND", no_blend), ("no MASK", no_mask), ("no normal maps", no_normalmap), ("group 2 no normal map", g2_nonormal),
               ("no metallic-roughness maps", no_mr), ("no albedo maps", no_albedo), ("no textures", no_tex), ("no glass", no_glass), ("no env", no_env), ("no lights", no_lights)]:
    run(tag, f, depth=4)
