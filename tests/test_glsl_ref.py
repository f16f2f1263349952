"""Pins the oracle to the EXECUTING reference: oracle/glsl_ref compiles the reference's own shader text (tile.glsl, preview.glsl,
tonemap.glsl and everything under common/, read in place from /root/reference) with g++ and runs it on the host.

* `test_oracle_equals_committed_reference_shader_output` — always runs: the oracle reproduces, BIT FOR BIT, the accumulation
  buffers, previews and RGBA8 readbacks the reference shaders produced for 8 scenes + 15 feature variants
  (tests/golden/glslref_golden.npz, written by tests/golden/make_glslref_golden.py).
* the `live` tests re-run the reference shaders here (they need /root/reference, or a prebuilt variant under oracle/_ref/) and
  compare at other sizes, tile layouts, sample ranges and tone-mapping modes, and check that the committed golden is current.
"""
import os
import numpy as np
import pytest
from conftest import scene_at
import feature_scenes as fs

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "glslref_golden.npz")
W, H, TW, TH, SPP = 48, 32, 20, 12, 4
SCENES = ("cornell_box_orig", "cornell_box_sphere", "hyperion_rect_lights", "hyperion_sphere_light", "volume_cube", "ibl_spheres",
          "teapot", "instancing", "gltf_mix")
CASES = list(SCENES) + ["variant_" + v for v in fs.VARIANTS]


def case_scene(case):
    if case.startswith("variant_"):
        return fs.resized(fs.build(case[len("variant_"):]), W, H, TW, TH)
    return scene_at(case, W, H, TW, TH)


def bits_equal(a, b):
    a = np.ascontiguousarray(a, np.float32); b = np.ascontiguousarray(b, np.float32)
    return bool(np.all((a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))))


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.fixture(scope="module")
def glsl_mod():
    from oracle.glsl_ref import binding as gb
    if not gb.reference_available() and not os.path.isdir(gb._OUT):
        pytest.skip("reference shaders (/root/reference) not present and no prebuilt variant")
    return gb


def live(gb, sc, **kw):
    try:
        return gb.GlslRef(sc, **kw)
    except FileNotFoundError as e:
        pytest.skip(str(e))


@pytest.mark.parametrize("case", CASES)
def test_oracle_equals_committed_reference_shader_output(case, golden, oracle_mod):
    from glsl_pathtracer_b200 import scene_io as sio
    sc = case_scene(case)
    orc = oracle_mod.Oracle(sc)                      # reference-faithful traversal (no t-culling), as the shader text
    accum = orc.render(1, SPP)
    assert bits_equal(accum, golden[case + "/accum"])
    assert bits_equal(orc.render_preview(W // 2, H // 2), golden[case + "/preview"])
    rgba = oracle_mod.tonemap(accum, 1.0 / SPP, sc.renderOptions, sio.derive_features(sc))
    assert np.array_equal(rgba, golden[case + "/rgba8"])
    orc.close()


def test_golden_cases_are_not_degenerate(golden):
    lit = [c for c in CASES if np.nanmax(golden[c + "/accum"][..., :3]) > 0]
    assert set(CASES) - set(lit) <= {"teapot"}       # teapot.scene as shipped has neither lights nor its HDR
    assert not bits_equal(golden["variant_alpha_mask/accum"], golden["variant_alpha_blend/accum"])
    assert not bits_equal(golden["variant_texture_maps_gl/accum"], golden["variant_texture_maps_dx/accum"])
    assert (golden["variant_env_rot_hide_emitters_background/accum"][..., 3] < SPP).any()


@pytest.mark.parametrize("case", ["cornell_box_orig", "volume_cube", "variant_alpha_blend", "variant_medium_scatter"])
def test_committed_golden_is_current(case, golden, glsl_mod):
    """The committed vectors are what the reference shader text produces today (guards the generator and the shim)."""
    sc = case_scene(case)
    g = live(glsl_mod, sc)
    accum = g.render(1, SPP)
    assert bits_equal(accum, golden[case + "/accum"])
    assert np.array_equal(g.tonemap(accum, 1.0 / SPP, sc.renderOptions), golden[case + "/rgba8"])


@pytest.mark.parametrize("name,size", [("cornell_box_orig", (96, 64, 40, 24)), ("hyperion_rect_lights", (160, 90, 64, 36)),
                                        ("hyperion_sphere_light", (96, 54, 96, 54)), ("volume_cube", (96, 64, 48, 32)),
                                        ("ibl_spheres", (128, 72, 50, 30)), ("instancing", (96, 54, 32, 32)), ("gltf_mix", (128, 72, 48, 40))])
def test_live_reference_shaders_equal_oracle(name, size, glsl_mod, oracle_mod):
    """Other sizes and tile layouts, a later sample range (frame counter schedule), one explicit tile draw."""
    sc = scene_at(name, *size)
    g = live(glsl_mod, sc); orc = oracle_mod.Oracle(sc)
    a = g.render(3, 2); b = orc.render(3, 2)
    assert bits_equal(a, b)
    ntx = -(-size[0] // size[2]); nty = -(-size[1] // size[3])
    ta = np.zeros_like(a); tb = np.zeros_like(b)
    g.render_tile(ntx - 1, nty - 1, 7, ta); orc.render_tile(ntx - 1, nty - 1, 7, tb)       # over-hanging corner tile
    assert bits_equal(ta, tb) and np.any(ta != 0)
    orc.close()


@pytest.mark.parametrize("variant", list(fs.VARIANTS))
def test_live_feature_variants_equal_oracle(variant, glsl_mod, oracle_mod):
    """Every feature define the reference compiles a distinct program for, at the variants' native sizes."""
    sc = fs.build(variant)
    g = live(glsl_mod, sc); orc = oracle_mod.Oracle(sc)
    assert bits_equal(g.render(1, 2), orc.render(1, 2))
    orc.close()


@pytest.mark.parametrize("tonemap,aces,simple", [(False, False, False), (True, False, False), (True, True, True), (True, True, False)])
@pytest.mark.parametrize("background", ["none", "background", "transparent"])
def test_live_tonemap_modes_equal_oracle(tonemap, aces, simple, background, glsl_mod, oracle_mod):
    from glsl_pathtracer_b200 import scene_io as sio
    sc = scene_at("ibl_spheres", 64, 36, 32, 18)
    fs.look_at(sc, (9, 1.0, 0), (-1, 2.5, 0))
    ro = sc.renderOptions
    ro.enableTonemap, ro.enableAces, ro.simpleAcesFit = tonemap, aces, simple
    ro.enableBackground = background == "background"; ro.transparentBackground = background == "transparent"
    ro.backgroundCol = (0.25, 0.5, 0.75)
    g = live(glsl_mod, sc); orc = oracle_mod.Oracle(sc)
    accum = orc.render(1, 3)
    a = g.tonemap(accum, 1.0 / 3, ro)
    b = oracle_mod.tonemap(accum, 1.0 / 3, ro, sio.derive_features(sc))
    assert np.array_equal(a, b)
    orc.close()


def test_translator_is_lexical_only(glsl_mod):
    """glsl2cpp.py keeps every token of the shader text apart from the documented rewrites: same identifiers in the same order."""
    import re
    from oracle.glsl_ref import glsl2cpp
    if not glsl_mod.reference_available():
        pytest.skip("needs /root/reference")
    src = glsl2cpp.load_with_includes(os.path.join(glsl_mod.REFERENCE_SHADERS, "tile.glsl"), glsl_mod.REFERENCE_SHADERS)
    out = glsl2cpp.translate(src)
    drop = {"in", "out", "inout", "uniform", "thread_local", "main", "glsl_main", "void", "version", "f"}
    def idents(s):
        s = glsl2cpp.strip_comments(s)
        s = re.sub(r"^\s*#version[^\n]*", "", s, flags=re.M)
        s = re.sub(r"glsl_rand_arg\d+", "rand", s)
        s = re.sub(r"float rand = rand\(\);\s*", "", s)
        return [t for t in re.findall(r"[A-Za-z_]\w*", s) if t not in drop and not re.fullmatch(r"[eE]\d*|\d+f", t)]
    a, b = idents(src), idents(out)
    assert a == b


# ---------------------------------------------------------------- G1 / G2 pins against the shader's own functions -----------
@pytest.mark.parametrize("name", ["cornell_box_orig", "cornell_box_sphere", "hyperion_rect_lights", "hyperion_sphere_light",
                                  "volume_cube", "ibl_spheres", "instancing", "gltf_mix"])
def test_closest_hit_and_any_hit_equal_the_shader_functions(name, glsl_mod, oracle_mod):
    """G1 pin: the oracle's host traversal == the reference's ClosestHit / AnyHit text on primary, random (incl. axis-parallel)
    and bounce-like rays: hit kind, material id and t bit-identical; occlusion identical."""
    from test_gpu_trace import random_rays
    sc = scene_at(name, 96, 54, 48, 27)
    g = live(glsl_mod, sc); orc = oracle_mod.Oracle(sc)
    rays = np.concatenate([orc.camera_rays(1), random_rays(sc, 30_000, 11)])
    for depth in (0, 1):                                   # depth matters under OPT_HIDE_EMITTERS only; cheap to cover
        o = orc.trace_closest(rays, depth)
        t, kind, mat = g.trace_closest(rays, depth)
        assert np.array_equal(kind, o["kind"])
        assert np.array_equal(t.view(np.uint32), o["t"].view(np.uint32))
        tri = kind == 1
        assert np.array_equal(mat[tri], o["matID"][tri])
    # bounce-like rays: leave the first hit point in a random direction; shadow-like: bounded by a random distance
    rng = np.random.default_rng(5)
    hit = o["kind"] == 1
    org = rays[hit, :3] + rays[hit, 3:] * o["t"][hit, None]
    d = rng.normal(size=org.shape).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
    b = np.concatenate([org + d * np.float32(3e-4), d], axis=1).astype(np.float32)
    ob2 = orc.trace_closest(b, 1); t2, k2, m2 = g.trace_closest(b, 1)
    assert np.array_equal(k2, ob2["kind"]) and np.array_equal(t2.view(np.uint32), ob2["t"].view(np.uint32))
    md = (rng.random(len(b), dtype=np.float32) * np.float32(2.0) * np.nanmedian(o["t"][hit])).astype(np.float32)
    from glsl_pathtracer_b200 import scene_io as sio
    f = sio.derive_features(sc)
    if not (f & sio.OPT_ALPHA_TEST) or (f & sio.OPT_MEDIUM):
        # (with the alpha test compiled in, AnyHit samples textures and draws from the path RNG: covered by the image tests instead)
        assert np.array_equal(g.trace_any(b, md), orc.trace_any(b, md))
        assert 0 < g.trace_any(b, md).mean() < 1
    orc.close()


def test_disney_eval_equals_the_shader_function(glsl_mod, oracle_mod):
    """G2 pin: the oracle's DisneyEval (f, pdf) == the reference's DisneyEval text, bit for bit, on the G2 query table
    (every in-repo material + sheen / tint / anisotropy / transmission rows; grazing, back-side and refraction configurations)."""
    from test_gpu_render import bsdf_queries
    sc = scene_at("cornell_box_orig", 32, 32, 16, 16)
    g = live(glsl_mod, sc); orc = oracle_mod.Oracle(sc)
    q = bsdf_queries(oracle_mod, seed=1, per_mat=200)
    a, b = g.bsdf_eval(q), orc.bsdf(q)
    assert np.count_nonzero(b["pdf"] > 0) > len(q) // 4
    assert bits_equal(a["pdf"], b["pdf"]) and bits_equal(a["f"], b["f"])
    orc.close()


def test_lambert_equals_the_shader_functions(glsl_mod, oracle_mod):
    """a14: LambertEval / LambertSample (dead code in the reference's PathTrace) — the oracle's copies == the shader text bit for bit;
    the sample comparison feeds the oracle the two draws the shader's own RNG produced."""
    from test_gpu_render import bsdf_queries
    sc = scene_at("cornell_box_orig", 32, 32, 16, 16)
    g = live(glsl_mod, sc); orc = oracle_mod.Oracle(sc)
    q = bsdf_queries(oracle_mod, seed=3, per_mat=50)
    a, b = g.lambert_eval(q), orc.lambert(q)
    assert bits_equal(a["pdf"], b["pdf"]) and bits_equal(a["f"], b["f"])
    seeds = np.random.default_rng(4).integers(0, 2 ** 32, size=(len(q), 4), dtype=np.uint64).astype(np.uint32)
    sa, r12 = g.lambert_sample(q, seeds)
    q2 = q.copy(); q2["r1"], q2["r2"] = r12[:, 0], r12[:, 1]
    sb = orc.lambert(q2, sample=True)
    assert bits_equal(sa["L"], sb["L"]) and bits_equal(sa["pdf"], sb["pdf"]) and bits_equal(sa["f"], sb["f"])
    assert np.all(sb["pdf"] >= 0) and np.all(r12 >= 0) and np.all(r12 <= 1)
    orc.close()


@pytest.mark.parametrize("name", ["cornell_box_orig", "hyperion_rect_lights"])
def test_live_instance_edit_equals_oracle(name, glsl_mod, oracle_mod):
    """After a reference-made instance edit (scaled + translated instance, TLAS rebuilt by Scene::RebuildInstances): the shader's
    inverse(transMat) / normal-matrix path on a non-trivial transform, reference shaders == oracle bit for bit."""
    from conftest import edited_scene
    sc, sc2 = edited_scene(name, 96, 64, 48, 32)
    g = live(glsl_mod, sc2); orc = oracle_mod.Oracle(sc2)
    a, b = g.render(1, 2), orc.render(1, 2)
    assert bits_equal(a, b)
    assert not bits_equal(b, oracle_mod.Oracle(sc).render(1, 2))
    orc.close()


def test_live_full_size_pass_equals_oracle(glsl_mod, oracle_mod):
    """BASELINE size: one full 1920x1080 sample pass of hyperion_rect_lights (64 tiles of 256x144, over-hanging last column / top row):
    all 2 073 600 pixels of the reference shaders' accumulation buffer bit-identical to the oracle's."""
    sc = scene_at("hyperion_rect_lights", 1920, 1080)
    g = live(glsl_mod, sc); orc = oracle_mod.Oracle(sc)
    a, b = g.render(2, 1), orc.render(2, 1)
    assert bits_equal(a, b) and np.isfinite(b).all() and b[..., :3].mean() > 0.01
    orc.close()
