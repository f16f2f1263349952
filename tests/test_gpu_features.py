"""Feature coverage of the hot path beyond what the in-repo scenes exercise: every light type, env-map
rotation, mesh emitters and the stale-material quirk (Q2), albedo alpha with MASK and BLEND alpha modes in closest-hit and
any-hit, metallic-roughness / normal / emission maps, thin-lens camera, absorbing / emissive / scattering media with and without
volume MIS.  Each variant renders the same modified scene arrays with the CUDA path (C ABI) and the oracle: relMSE <= 1e-3."""
import copy
import numpy as np
import pytest
from conftest import scene_at, rel_mse

pytestmark = pytest.mark.gpu


def compare(sc, spp=8, tol=1e-3, minfrac=0.9, oracle_mod=None):
    from glsl_pathtracer_b200 import capi
    ctx = capi.Context(sc); orc = oracle_mod.Oracle(sc)
    ctx.render_samples(1, spp)
    g = np.nan_to_num(ctx.read_accum() / spp); o = np.nan_to_num(orc.render(1, spp) / spp)
    r = rel_mse(o, g)
    frac = np.isclose(g[..., :3], o[..., :3], rtol=2e-3, atol=2e-4).all(axis=-1).mean()
    ctx.close(); orc.close()
    assert o[..., :3].max() > 0, "degenerate test scene (black image)"
    assert r <= tol, f"relMSE {r}"
    assert frac >= minfrac, f"only {frac:.3f} of pixels agree"
    return g, o


def light(pos, emission, u=(0, 0, 0), v=(0, 0, 0), radius=0.0, area=0.0, type_=0):
    return np.array([*pos, *emission, *u, *v, radius, area, type_], np.float32)


def test_all_light_types_together(oracle_mod):
    sc = scene_at("cornell_box_orig", 96, 96, 48, 48, 3)
    quad = sc.lights[0].copy()
    sphere = light((0.2, 0.3, 0.2), (30, 20, 10), radius=0.04, area=4 * np.pi * 0.04 ** 2, type_=1)
    distant = light((0.3, 0.2, -1.0), (1.5, 1.5, 2.0), type_=2)          # direction = normalize(position), area 0 (Loader.cpp:198-202)
    sc.lights = np.stack([quad, sphere, distant])
    compare(sc, oracle_mod=oracle_mod)


def test_env_rotation_hide_emitters_background(oracle_mod):
    sc = scene_at("ibl_spheres", 160, 90, 80, 45)
    ro = sc.renderOptions
    ro.envMapRot = 135.0; ro.hideEmitters = True; ro.enableBackground = True; ro.backgroundCol = (0.2, 0.3, 0.4); ro.envMapIntensity = 1.5
    cam = sc.camera                                # look towards the horizon so that sky (background) pixels exist
    cam.position = np.array([9, 1.0, 0], np.float32)
    f = np.array([-1, 0.15, 0], np.float32); f /= np.linalg.norm(f)
    r = np.cross(f, np.array([0, 1, 0], np.float32)); r /= np.linalg.norm(r)
    cam.forward, cam.right, cam.up = f.astype(np.float32), r.astype(np.float32), np.cross(r, f).astype(np.float32)
    g, o = compare(sc, minfrac=0.85, oracle_mod=oracle_mod)
    np.testing.assert_allclose(g[..., 3], o[..., 3], atol=1e-6)
    assert (g[..., 3] < 1).any()             # background pixels have alpha 0 (OPT_BACKGROUND)


def test_mesh_emitter_and_stale_material_on_light_hit(oracle_mod):
    """Q2: when a bounce ray hits an analytic light, GetMaterial runs with the previous hit's matID and its emission is added again."""
    sc = scene_at("cornell_box_orig", 96, 96, 48, 48, 4)
    sc.materials = sc.materials.copy()
    sc.materials[4, 4:7] = (4.0, 1.0, 0.5)       # small box becomes a mesh emitter
    sc.materials[1, 4:7] = (0.3, 0.3, 0.6)       # ceiling glows: rays leaving it often hit the quad light next
    compare(sc, spp=16, oracle_mod=oracle_mod)


def _alpha_texture(sc):
    tex = sc.textures.copy()
    h, w = tex.shape[1:3]
    y, x = np.mgrid[0:h, 0:w]
    tex[0, ..., 3] = np.where(((x // 16) + (y // 16)) % 3 == 0, 40, 230).astype(np.uint8)     # alpha varies across the checker
    return tex


@pytest.mark.parametrize("mode,cutoff", [(2, 0.5), (1, 0.0)])       # MASK (deferred any-hit alpha) / BLEND (RNG-consuming inline any-hit)
def test_albedo_alpha_mask_and_blend(mode, cutoff, oracle_mod):
    sc = scene_at("ibl_spheres", 160, 90, 80, 45)
    sc.textures = _alpha_texture(sc)
    sc.materials = sc.materials.copy()
    tl = sc.nodes[sc.topLevelIndex:]
    floor_mat = int(tl[tl[:, 8] == -3][0, 7])      # TLAS leaf of instance 2 (the floor): LRLeaf.y = material id
    sc.materials[floor_mat, 29] = mode; sc.materials[floor_mat, 30] = cutoff; sc.materials[floor_mat, 28] = 0.9
    # an analytic light so that light NEE shadow rays cross the cut-out floor from below as well
    sc.lights = np.stack([light((-1, 6, -1), (40, 40, 40), u=(2, 0, 0), v=(0, 0, 2), area=4.0, type_=0)])
    sc.camera.position = np.array([9, -3.0, 0], np.float32)        # look at the floor from below: shadow rays must pass the alpha test
    compare(sc, minfrac=0.8, oracle_mod=oracle_mod)


def test_metallic_roughness_normal_and_emission_maps(oracle_mod):
    sc = scene_at("ibl_spheres", 160, 90, 80, 45)
    sc.materials = sc.materials.copy()
    for m in range(len(sc.materials)):
        if sc.materials[m, 24] >= 0:               # the checker material: reuse its texture in every slot
            sc.materials[m, 25] = 0; sc.materials[m, 26] = 0; sc.materials[m, 27] = 0
    sc.materials[1, 26] = 0                        # normal map on the glossy sphere too (UV-derived tangent frame)
    compare(sc, minfrac=0.85, oracle_mod=oracle_mod)
    sc.renderOptions.openglNormalMap = False
    compare(sc, minfrac=0.85, oracle_mod=oracle_mod)


def test_thin_lens_camera(oracle_mod):
    sc = scene_at("cornell_box_sphere", 96, 96, 48, 48)
    sc.camera.aperture = 0.0004; sc.camera.focalDist = 0.85
    compare(sc, minfrac=0.85, oracle_mod=oracle_mod)


@pytest.mark.parametrize("medium_type,vol_mis", [(1, True), (3, True), (2, False), (1, False)])
def test_media_variants(medium_type, vol_mis, oracle_mod):
    """absorb / emissive / scatter media; without volume MIS the shadow rays are deferred binary any-hit tests that ignore alpha
    (anyhit.glsl:74) and light hits after a medium scatter get MIS weight 1 (pathtrace.glsl:356-359)."""
    sc = scene_at("volume_cube", 128, 72, 64, 36)
    sc.materials = sc.materials.copy()
    sc.materials[1, 18] = medium_type
    sc.materials[1, 23] = 0.4                      # anisotropic phase function
    sc.renderOptions.enableVolumeMIS = vol_mis
    if not vol_mis:
        sc.materials[1, 28] = 0.35                 # partly opaque BLEND boundary so both branches of the alpha test occur
    compare(sc, spp=8, minfrac=0.8, oracle_mod=oracle_mod)


def test_many_bounces_without_russian_roulette(oracle_mod):
    sc = scene_at("cornell_box_sphere", 96, 96, 48, 48, 12)
    sc.renderOptions.enableRR = False
    compare(sc, spp=4, minfrac=0.85, oracle_mod=oracle_mod)


def test_render_options_update_without_recreate(oracle_mod):
    """ReloadShaders / uniform changes (ptb_set_options) take effect on an existing context."""
    from glsl_pathtracer_b200 import capi
    sc = scene_at("cornell_box_orig", 64, 64, 32, 32, 2)
    ctx = capi.Context(sc)
    ctx.render_samples(1, 2); a = ctx.read_accum()
    sc2 = copy.deepcopy(sc); sc2.renderOptions.maxDepth = 5; sc2.renderOptions.enableUniformLight = True
    ctx.set_options(capi.make_options(sc2)); ctx.reset_accum()
    ctx.render_samples(1, 2); b = ctx.read_accum()
    orc = oracle_mod.Oracle(sc2)
    assert rel_mse(orc.render(1, 2), b) <= 1e-3 and not np.allclose(a, b)
    ctx.close(); orc.close()
