"""Scene variants that exercise every feature define / branch of the hot path beyond the in-repo scenes.
Shared by tests/test_gpu_features.py (CUDA vs oracle) and tests/test_glsl_ref.py (oracle vs the executing reference shaders)."""
import numpy as np
from conftest import scene_at


def light(pos, emission, u=(0, 0, 0), v=(0, 0, 0), radius=0.0, area=0.0, type_=0):
    return np.array([*pos, *emission, *u, *v, radius, area, type_], np.float32)


def look_at(sc, pos, target):
    cam = sc.camera
    cam.position = np.asarray(pos, np.float32)
    f = np.asarray(target, np.float32) - cam.position; f /= np.linalg.norm(f)
    r = np.cross(f, np.array([0, 1, 0], np.float32)); r /= np.linalg.norm(r)
    cam.forward, cam.right, cam.up = f.astype(np.float32), r.astype(np.float32), np.cross(r, f).astype(np.float32)


def all_light_types():
    sc = scene_at("cornell_box_orig", 96, 96, 48, 48, 3)
    quad = sc.lights[0].copy()
    sphere = light((0.2, 0.3, 0.2), (30, 20, 10), radius=0.04, area=4 * np.pi * 0.04 ** 2, type_=1)
    distant = light((0.3, 0.2, -1.0), (1.5, 1.5, 2.0), type_=2)          # direction = normalize(position), area 0 (Loader.cpp:198-202)
    sc.lights = np.stack([quad, sphere, distant])
    return sc


def env_rotation_hide_emitters_background():
    sc = scene_at("ibl_spheres", 160, 90, 80, 45)
    ro = sc.renderOptions
    ro.envMapRot = 135.0; ro.hideEmitters = True; ro.enableBackground = True; ro.backgroundCol = (0.2, 0.3, 0.4); ro.envMapIntensity = 1.5
    look_at(sc, (9, 1.0, 0), (-1, 2.5, 0))       # look towards the horizon so that sky (background) pixels exist
    return sc


def mesh_emitter_stale_material():
    """Q2: when a bounce ray hits an analytic light, GetMaterial runs with the previous hit's matID and its emission is added again."""
    sc = scene_at("cornell_box_orig", 96, 96, 48, 48, 4)
    sc.materials = sc.materials.copy()
    sc.materials[4, 4:7] = (4.0, 1.0, 0.5)       # small box becomes a mesh emitter
    sc.materials[1, 4:7] = (0.3, 0.3, 0.6)       # ceiling glows: rays leaving it often hit the quad light next
    return sc


def _alpha_texture(sc):
    tex = sc.textures.copy()
    h, w = tex.shape[1:3]
    y, x = np.mgrid[0:h, 0:w]
    tex[0, ..., 3] = np.where(((x // 16) + (y // 16)) % 3 == 0, 40, 230).astype(np.uint8)     # alpha varies across the checker
    return tex


def albedo_alpha(mode, cutoff):
    """mode 2 = MASK (deferred any-hit alpha), 1 = BLEND (RNG-consuming inline any-hit)."""
    sc = scene_at("ibl_spheres", 160, 90, 80, 45)
    sc.textures = _alpha_texture(sc)
    sc.materials = sc.materials.copy()
    tl = sc.nodes[sc.topLevelIndex:]
    floor_mat = int(tl[tl[:, 8] == -3][0, 7])      # TLAS leaf of instance 2 (the floor): LRLeaf.y = material id
    sc.materials[floor_mat, 29] = mode; sc.materials[floor_mat, 30] = cutoff; sc.materials[floor_mat, 28] = 0.9
    # an analytic light so that light NEE shadow rays cross the cut-out floor from below as well
    sc.lights = np.stack([light((-1, 6, -1), (40, 40, 40), u=(2, 0, 0), v=(0, 0, 2), area=4.0, type_=0)])
    look_at(sc, (9, -3.0, 0), (0, 0.5, 0))       # look at the floor from below: camera, bounce and shadow rays all meet the alpha test
    return sc


def texture_maps(opengl_normal_map=True):
    sc = scene_at("ibl_spheres", 160, 90, 80, 45)
    sc.materials = sc.materials.copy()
    for m in range(len(sc.materials)):
        if sc.materials[m, 24] >= 0:               # the checker material: reuse its texture in every slot
            sc.materials[m, 25] = 0; sc.materials[m, 26] = 0; sc.materials[m, 27] = 0
    sc.materials[1, 26] = 0                        # normal map on the glossy sphere too (UV-derived tangent frame)
    sc.renderOptions.openglNormalMap = opengl_normal_map
    return sc


def thin_lens():
    sc = scene_at("cornell_box_sphere", 96, 96, 48, 48)
    sc.camera.aperture = 0.0004; sc.camera.focalDist = 0.85
    return sc


def media(medium_type, vol_mis):
    """absorb (1) / scatter (2) / emissive (3) media; without volume MIS the shadow rays are binary any-hit tests that ignore alpha
    (anyhit.glsl:74) and light hits after a medium scatter get MIS weight 1 (pathtrace.glsl:356-359)."""
    sc = scene_at("volume_cube", 128, 72, 64, 36)
    sc.materials = sc.materials.copy()
    sc.materials[1, 18] = medium_type
    sc.materials[1, 23] = 0.4                      # anisotropic phase function
    sc.renderOptions.enableVolumeMIS = vol_mis
    if not vol_mis:
        sc.materials[1, 28] = 0.35                 # partly opaque BLEND boundary so both branches of the alpha test occur
    return sc


def media_no_blend(medium_type):
    """volume MIS with a purely refractive medium boundary (no BLEND material anywhere): EvalTransmittance draws no random number, the NEE rays are
    deferred to k_transmit (F.deferTransmit)."""
    sc = media(medium_type, True)
    sc.materials[:, 29] = 0                        # alphaMode OPAQUE; the boundary stays transparent for shadow rays through specTrans = 1
    sc.materials[:, 28] = 1.0
    return sc


def many_bounces_no_rr():
    sc = scene_at("cornell_box_sphere", 96, 96, 48, 48, 12)
    sc.renderOptions.enableRR = False
    return sc


def uniform_light_mollification_transparent():
    sc = scene_at("ibl_spheres", 160, 90, 80, 45, 4)
    look_at(sc, (9, 1.0, 0), (-1, 2.5, 0))       # sky pixels: alpha 0 with OPT_TRANSPARENT_BACKGROUND
    ro = sc.renderOptions
    ro.enableUniformLight = True; ro.uniformLightCol = (0.4, 0.5, 0.7)
    ro.enableRoughnessMollification = True; ro.roughnessMollificationAmt = 0.6
    ro.transparentBackground = True
    return sc


# name -> (builder, kwargs): every variant the reference compiles a distinct shader program for, plus the data-driven branches
VARIANTS = {
    "all_light_types": (all_light_types, {}),
    "env_rot_hide_emitters_background": (env_rotation_hide_emitters_background, {}),
    "mesh_emitter_stale_material": (mesh_emitter_stale_material, {}),
    "alpha_mask": (albedo_alpha, dict(mode=2, cutoff=0.5)),
    "alpha_blend": (albedo_alpha, dict(mode=1, cutoff=0.0)),
    "texture_maps_gl": (texture_maps, dict(opengl_normal_map=True)),
    "texture_maps_dx": (texture_maps, dict(opengl_normal_map=False)),
    "thin_lens": (thin_lens, {}),
    "medium_absorb_volmis": (media, dict(medium_type=1, vol_mis=True)),
    "medium_emissive_volmis": (media, dict(medium_type=3, vol_mis=True)),
    "medium_scatter_volmis": (media, dict(medium_type=2, vol_mis=True)),
    "medium_scatter": (media, dict(medium_type=2, vol_mis=False)),
    "medium_absorb": (media, dict(medium_type=1, vol_mis=False)),
    "many_bounces_no_rr": (many_bounces_no_rr, {}),
    "uniform_light_mollification_transparent": (uniform_light_mollification_transparent, {}),
}


def build(name):
    fn, kw = VARIANTS[name]
    return fn(**kw)


def resized(sc, w, h, tw, th):
    ro = sc.renderOptions
    ro.renderResolution = (w, h); ro.windowResolution = (w, h); ro.tileWidth = tw; ro.tileHeight = th
    return sc
