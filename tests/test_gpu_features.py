"""Feature coverage of the hot path beyond what the in-repo scenes exercise: every light type, env-map
rotation, mesh emitters and the stale-material quirk (Q2), albedo alpha with MASK and BLEND alpha modes in closest-hit and
any-hit, metallic-roughness / normal / emission maps, thin-lens camera, absorbing / emissive / scattering media with and without
volume MIS.  Each variant renders the same modified scene arrays with the CUDA path (C ABI) and the oracle: relMSE <= 1e-3."""
import copy
import numpy as np
import pytest
from conftest import scene_at, rel_mse
import feature_scenes as fs

pytestmark = pytest.mark.gpu


def compare(sc, spp=8, tol=1e-3, minfrac=0.9, oracle_mod=None, max_bias=None):
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
    if max_bias is not None:
        # signed bias of the mean image: a wrong estimator (e.g. phase function where the reference evaluates the BSDF) shifts the mean
        # even when relMSE and the matched-pixel fraction still pass
        d = (g[..., :3].astype(np.float64) - o[..., :3].astype(np.float64)).mean(axis=-1).ravel()
        mean, sem, level = d.mean(), d.std(ddof=1) / np.sqrt(d.size), float(o[..., :3].mean())
        print(f"signed bias {mean:+.3e} +- {sem:.3e}, level {level:.4f}, relMSE {r:.2e}, matched {frac:.3f}")
        assert abs(mean) <= max(3.0 * sem, 1e-7) and abs(mean) <= max_bias * level, f"signed bias {mean} (sem {sem}, level {level})"
    return g, o


def test_all_light_types_together(oracle_mod):
    compare(fs.all_light_types(), oracle_mod=oracle_mod)


def test_env_rotation_hide_emitters_background(oracle_mod):
    g, o = compare(fs.env_rotation_hide_emitters_background(), minfrac=0.85, oracle_mod=oracle_mod)
    np.testing.assert_allclose(g[..., 3], o[..., 3], atol=1e-6)
    assert (g[..., 3] < 1).any()             # background pixels have alpha 0 (OPT_BACKGROUND)


def test_mesh_emitter_and_stale_material_on_light_hit(oracle_mod):
    """Q2: when a bounce ray hits an analytic light, GetMaterial runs with the previous hit's matID and its emission is added again."""
    compare(fs.mesh_emitter_stale_material(), spp=16, oracle_mod=oracle_mod)


@pytest.mark.parametrize("mode,cutoff", [(2, 0.5), (1, 0.0)])       # MASK (deferred any-hit alpha) / BLEND (RNG-consuming inline any-hit)
def test_albedo_alpha_mask_and_blend(mode, cutoff, oracle_mod):
    compare(fs.albedo_alpha(mode, cutoff), minfrac=0.8, oracle_mod=oracle_mod)


def test_metallic_roughness_normal_and_emission_maps(oracle_mod):
    compare(fs.texture_maps(True), minfrac=0.85, oracle_mod=oracle_mod)
    compare(fs.texture_maps(False), minfrac=0.85, oracle_mod=oracle_mod)


def test_thin_lens_camera(oracle_mod):
    compare(fs.thin_lens(), minfrac=0.85, oracle_mod=oracle_mod)


@pytest.mark.parametrize("medium_type,vol_mis", [(1, True), (3, True), (2, True), (2, False), (1, False)])
def test_media_variants(medium_type, vol_mis, oracle_mod):
    """absorb / emissive / scatter media; without volume MIS the shadow rays are deferred binary any-hit tests that ignore alpha
    (anyhit.glsl:74) and light hits after a medium scatter get MIS weight 1 (pathtrace.glsl:356-359)."""
    # the NEE of a medium scatter WITHOUT volume MIS evaluates DisneyEval on the boundary material, not the phase function
    # (pathtrace.glsl:200,268): the signed-bias gate is what catches a deviation there
    compare(fs.media(medium_type, vol_mis), spp=16, minfrac=0.9 if (medium_type == 2 and not vol_mis) else 0.8, oracle_mod=oracle_mod, max_bias=5e-3)


def test_uniform_light_mollification_transparent_background(oracle_mod):
    g, o = compare(fs.uniform_light_mollification_transparent(), minfrac=0.85, oracle_mod=oracle_mod)
    np.testing.assert_allclose(g[..., 3], o[..., 3], atol=1e-6)


def test_many_bounces_without_russian_roulette(oracle_mod):
    compare(fs.many_bounces_no_rr(), spp=4, minfrac=0.85, oracle_mod=oracle_mod)


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


def test_malformed_scenes_are_rejected_not_dereferenced():
    """Scene validation at the boundary: a BLAS leaf that references triangles past the end of vertIndices, a TLAS leaf whose material id is out
    of range, and a delta upload that shrinks the material table below an instance's id all return PTB_ERR_INVALID_ARGUMENT instead of reading out of bounds."""
    from glsl_pathtracer_b200 import capi
    sc = scene_at("cornell_box_orig", 32, 32, 16, 16)
    nodes = np.ascontiguousarray(sc.nodes, np.float32).reshape(-1, 9)
    bad = copy.deepcopy(sc); bad.nodes = nodes.copy()
    leaf = np.nonzero(bad.nodes[:, 8] > 0)[0][0]
    bad.nodes[leaf, 6] = len(sc.vertIndices) - 1; bad.nodes[leaf, 7] = 3           # first + count runs past the end
    with pytest.raises(capi.PtbError) as e:
        capi.Context(bad)
    assert e.value.code == capi.PTB_ERR_INVALID_ARGUMENT
    bad = copy.deepcopy(sc); bad.nodes = nodes.copy()
    tl = np.nonzero(bad.nodes[:, 8] < 0)[0][0]
    bad.nodes[tl, 7] = len(sc.materials) + 5                                         # TLAS leaf material id out of range
    with pytest.raises(capi.PtbError) as e:
        capi.Context(bad)
    assert e.value.code == capi.PTB_ERR_INVALID_ARGUMENT
    ctx = capi.Context(sc)
    ctx.render_samples(1, 1)
    before = ctx.read_accum()
    with pytest.raises(capi.PtbError) as e:                                          # fewer materials than the instances reference
        ctx.update_instances(sc.transforms, np.ascontiguousarray(sc.materials, np.float32)[:1], nodes[sc.topLevelIndex:])
    assert e.value.code == capi.PTB_ERR_INVALID_ARGUMENT
    ctx.reset_accum(); ctx.render_samples(1, 1)                                      # the context was left untouched by the rejected update
    assert ctx.read_accum().tobytes() == before.tobytes()
    ctx.close()


@pytest.mark.parametrize("name", ["hyperion_rect_lights", "hyperion_sphere_light", "ibl_spheres", "feature:env_hide_bg", "feature:uniform_transparent",
                                  "feature:stale_material", "feature:all_lights"])
def test_paths_finished_in_the_trace_kernel_equal_paths_finished_in_the_shade_kernel(name, monkeypatch):
    """Misses and light hits are finished by the trace kernel (finishInTrace) and never enter a shade queue; PTB_TRACE_FINISH=0 keeps them in
    the shade kernel.  Same expressions on the same inputs: the two running sums may differ by FMA contraction only."""
    from glsl_pathtracer_b200 import capi
    sc = {"feature:env_hide_bg": fs.env_rotation_hide_emitters_background, "feature:uniform_transparent": fs.uniform_light_mollification_transparent,
          "feature:stale_material": fs.mesh_emitter_stale_material, "feature:all_lights": fs.all_light_types}[name]() if name.startswith("feature:") else scene_at(name, 320, 180)
    imgs = []
    for flag in ("1", "0"):
        monkeypatch.setenv("PTB_TRACE_FINISH", flag)
        ctx = capi.Context(sc)
        ctx.render_samples(1, 6)
        imgs.append(np.nan_to_num(ctx.read_accum().astype(np.float64)))
        st = ctx.stats(); ctx.close()
        imgs.append((st["pathSegments"], st["shadowRays"]))
    a, ca, b, cb = imgs
    assert ca == cb, f"ray counts differ: {ca} vs {cb}"
    assert a[..., :3].max() > 0
    np.testing.assert_array_equal(a[..., 3], b[..., 3])                       # alpha: exact
    scale = np.maximum(np.abs(b[..., :3]), 1e-3)
    assert (np.abs(a[..., :3] - b[..., :3]) / scale).max() <= 2e-5


@pytest.mark.parametrize("medium_type", [1, 2, 3])
def test_deferred_transmittance_equals_inline_transmittance(medium_type, monkeypatch, oracle_mod):
    """Under OPT_MEDIUM + OPT_VOL_MIS the NEE rays carry EvalTransmittance (pathtrace.glsl:119-155).  Without BLEND materials it draws no random number, so the
    rays are queued and evaluated by k_transmit (k_shade<3>); PTB_DEFER_TRANSMIT=0 keeps the evaluation inside k_shade<2>.  Same paths; the sums
    differ by the association of (Li * T) * f vs (Li * f) * T only.  Both equal the oracle."""
    from glsl_pathtracer_b200 import capi
    sc = fs.media_no_blend(medium_type)
    res = []
    for flag in ("1", "0"):
        monkeypatch.setenv("PTB_DEFER_TRANSMIT", flag)
        ctx = capi.Context(sc)
        ctx.render_samples(1, 6)
        img = np.nan_to_num(ctx.read_accum().astype(np.float64))
        st = ctx.stats(); ctx.close()
        res.append((img, st["pathSegments"], st["kernelLaunches"]))
    (a, sa, la), (b, sb, lb) = res
    assert la > lb, "the deferred variant launches k_transmit after every shade pass"
    assert 0.5 * sb < sa <= sb, f"path segments: {sa} deferred vs {sb} inline"      # (a ray whose BSDF pdf is 0 contributes nothing and is not queued)
    assert a[..., :3].max() > 0
    np.testing.assert_array_equal(a[..., 3], b[..., 3])
    scale = np.maximum(np.abs(b[..., :3]), 1e-3)
    assert (np.abs(a[..., :3] - b[..., :3]) / scale).max() <= 5e-5
    monkeypatch.setenv("PTB_DEFER_TRANSMIT", "1")
    compare(sc, spp=16, minfrac=0.8, oracle_mod=oracle_mod, max_bias=5e-3)
