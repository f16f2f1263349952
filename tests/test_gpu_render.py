"""G2 (BSDF eval/pdf within 1e-4 relative) and G3 (image relMSE <= 1e-3 at equal spp, RNG-matched paths) gates,
plus the Renderer state machine, tile-vs-whole-frame equivalence and tonemap byte parity.  CUDA path called through the
C ABI (libptb200.so) and compared with the CPU oracle on the same inputs."""
import numpy as np
import pytest
from conftest import scene_at, rel_mse

pytestmark = pytest.mark.gpu


def _ctx(sc, **kw):
    from glsl_pathtracer_b200 import capi
    return capi.Context(sc, **kw)


# ---------------------------------------------------------------- G2 ------------------------------------------------
def bsdf_queries(oracle_mod, seed=1, per_mat=400):
    from conftest import load_scene_cached
    rng = np.random.default_rng(seed)
    mats = []
    for name in ("hyperion_rect_lights", "hyperion_sphere_light", "cornell_box_orig", "cornell_box_sphere", "volume_cube", "teapot"):
        mats.append(load_scene_cached(name).materials)
    mats = np.unique(np.concatenate(mats), axis=0)
    # synthetic rows that exercise sheen / specular tint / anisotropy / partial transmission
    extra = np.repeat(mats[:1], 4, axis=0).copy()
    extra[0, 12:14] = (1.0, 0.5); extra[1, 11] = 1.0; extra[1, 8] = 0.3; extra[2, 3] = 0.8; extra[2, 8] = 1.0; extra[2, 9] = 0.3
    extra[3, 16] = 0.6; extra[3, 9] = 0.2; extra[3, 14:16] = (0.5, 0.5)
    mats = np.concatenate([mats, extra])
    q = np.zeros(len(mats) * per_mat, oracle_mod.BSDF_QUERY_DTYPE)
    k = 0
    for m in mats:
        for j in range(per_mat):
            N = rng.normal(size=3); N /= np.linalg.norm(N)
            if j % 50 == 0: N = np.array([0, 0, 1.0])                     # Onb() singular branch (sampling.glsl:181)
            V = rng.normal(size=3); V /= np.linalg.norm(V)
            if np.dot(V, N) < 0: V = -V                                     # N is the face-forward normal
            if j % 7 == 0:                                                  # grazing view
                t = np.cross(N, rng.normal(size=3)); t /= np.linalg.norm(t); V = 0.02 * N + t; V /= np.linalg.norm(V)
            L = rng.normal(size=3); L /= np.linalg.norm(L)                  # both hemispheres (refraction configurations)
            inside = (j % 5 == 0)
            ior = m[17]
            q[k]["mat"] = m; q[k]["V"] = V; q[k]["N"] = N; q[k]["L"] = L
            q[k]["eta"] = ior if inside else np.float32(1.0) / np.float32(ior)   # pathtrace.glsl:114
            q[k]["r1"], q[k]["r2"], q[k]["r3"] = rng.random(3)
            k += 1
    return q


def _close(a, b, rtol=1e-4, atol=1e-6):
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    both_nan = np.isnan(a) & np.isnan(b)
    return both_nan | (np.abs(a - b) <= rtol * np.abs(b) + atol)


def test_bsdf_eval_matches_oracle(oracle_mod):
    sc = scene_at("cornell_box_orig", 32, 32, 16, 16)
    ctx = _ctx(sc); orc = oracle_mod.Oracle(sc)
    q = bsdf_queries(oracle_mod)
    g, o = ctx.bsdf(q), orc.bsdf(q)
    assert np.count_nonzero(o["pdf"] > 0) > len(q) // 4
    ok = _close(g["pdf"], o["pdf"]) & _close(g["f"], o["f"]).all(axis=1)
    assert ok.all(), f"{np.count_nonzero(~ok)} of {len(q)} eval mismatches; first {np.nonzero(~ok)[0][:5]}"
    ctx.close(); orc.close()


def test_lambert_matches_oracle(oracle_mod):
    """a14 (lambert.glsl, dead code in the reference's render loop): eval and sample through the C ABI vs the oracle, 1e-4 relative."""
    sc = scene_at("cornell_box_orig", 32, 32, 16, 16)
    ctx = _ctx(sc); orc = oracle_mod.Oracle(sc)
    q = bsdf_queries(oracle_mod, seed=5, per_mat=60)
    for sample in (False, True):
        g, o = ctx.lambert(q, sample=sample), orc.lambert(q, sample=sample)
        # sampled directions near the horizon come from z = sqrt(1 - x^2 - y^2) with full cancellation: compare on the scale of |L| = 1
        atol = 3e-5 if sample else 1e-6
        ok = _close(g["pdf"], o["pdf"], atol=atol) & _close(g["f"], o["f"], atol=atol).all(axis=1)
        assert ok.all(), f"sample={sample}: {np.count_nonzero(~ok)} of {len(q)} differ, e.g. {g['pdf'][~ok][:3]} vs {o['pdf'][~ok][:3]}"
        assert (np.abs(g["L"] - o["L"]) <= 3e-5).all()
    ctx.close(); orc.close()


def test_bsdf_sample_matches_oracle(oracle_mod):
    sc = scene_at("cornell_box_orig", 32, 32, 16, 16)
    ctx = _ctx(sc); orc = oracle_mod.Oracle(sc)
    q = bsdf_queries(oracle_mod, seed=2)
    g, o = ctx.bsdf(q, sample=True), orc.bsdf(q, sample=True)
    # near-singular specular lobes (roughness 0.001) amplify 1-ulp sin/cos differences of L into large f/pdf changes:
    # compare L tightly and f, pdf where the lobe is not a near-delta
    okL = (np.abs(g["L"] - o["L"]) <= 3e-4).all(axis=1) | np.isnan(o["L"]).any(axis=1)
    assert okL.mean() > 0.999, f"sampled directions differ for {np.count_nonzero(~okL)} queries: {[(int(i), g['L'][i].tolist(), o['L'][i].tolist(), q['mat'][i][9]) for i in np.nonzero(~okL)[0][:5]]}"
    smooth = q["mat"][:, 9] >= 0.05
    # sampled directions agree to ~1e-4 (2-ulp division/sqrt in the shading math); peaked lobes turn that into <1% in f and pdf
    ok = _close(g["pdf"], o["pdf"], rtol=1e-2) & _close(g["f"], o["f"], rtol=1e-2).all(axis=1)
    sel = smooth & okL
    bad = np.nonzero(sel & ~ok)[0]
    assert ok[sel].mean() > 0.999, f"{bad.size} of {sel.sum()} differ; e.g. {[(int(i), g['pdf'][i], o['pdf'][i], g['f'][i].tolist(), o['f'][i].tolist()) for i in bad[:4]]}"
    ctx.close(); orc.close()


# ---------------------------------------------------------------- G3 ------------------------------------------------
# last column: minimum fraction of pixels that must agree to float rounding.  hyperion_sphere_light is lower by construction of
# the REFERENCE algorithm: the sphere light occludes its own NEE shadow ray whenever SphereIntersect's t lands below
# dist - EPS (anyhit.glsl:56-61 vs sampling.glsl:203-205); at |p| ~ 40 the fp32 error of t is of the order of EPS, so the
# visibility of the grazing ~4% of NEE samples (sqrt(det) < 0.3, where dt ~ 2e-4 / (2 sqrt(det)) > EPS) is decided by rounding noise
# of the inputs (libm vs CUDA sin/cos, 2-ulp division).  Those samples carry little energy (contribution ~ |cos| at the light).
CASES = [("cornell_box_orig", 128, 128, 64, 64, 4, 8, 0.97), ("cornell_box_sphere", 128, 128, 64, 64, None, 8, 0.97),
         ("hyperion_rect_lights", 240, 136, 64, 36, None, 8, 0.97), ("hyperion_sphere_light", 240, 136, 64, 36, None, 8, 0.5),
         ("volume_cube", 160, 90, 80, 45, None, 8, 0.97), ("teapot", 128, 72, 64, 36, None, 8, 0.97),
         # generated HDR environment (SampleEnvMap / EvalEnvMap / CDF binary search, env NEE shadow queue), checker texture with REPEAT wrap
         ("ibl_spheres", 240, 136, 64, 36, None, 8, 0.95),
         # 10 001 instances, TLAS height 15, rotated + non-uniformly scaled transforms, depth 8, glass/metal/clearcoat/sheen/anisotropic materials
         ("instancing", 240, 136, 64, 36, None, 4, 0.93),
         # glTF input loaded by the reference's GLTFLoader.cpp: all four texture-map kinds, MASK + BLEND alpha, transmission, env map + quad light
         ("gltf_mix", 240, 136, 64, 36, None, 8, 0.9)]


@pytest.mark.parametrize("name,w,h,tw,th,depth,spp,minfrac", CASES)
def test_image_matches_oracle_rng_matched(name, w, h, tw, th, depth, spp, minfrac, oracle_mod):
    sc = scene_at(name, w, h, tw, th, depth)
    ctx = _ctx(sc); orc = oracle_mod.Oracle(sc)
    ctx.render_samples(1, spp)
    g = ctx.read_accum() / spp
    o = orc.render(1, spp) / spp
    assert np.isfinite(g[..., :3]).all() == np.isfinite(o[..., :3]).all()
    g = np.nan_to_num(g); o = np.nan_to_num(o)
    r = rel_mse(o, g)
    assert r <= 1e-3, f"relMSE {r}"
    # RNG-matched paths: the large majority of pixels agree to float rounding
    close = np.isclose(g[..., :3], o[..., :3], rtol=1e-3, atol=1e-4).all(axis=-1)
    assert close.mean() > minfrac, f"only {close.mean():.4f} of pixels match"
    np.testing.assert_allclose(g[..., 3], o[..., 3], atol=1e-6)
    st_g, st_o = ctx.stats(), orc.stats()
    assert abs(st_g["pathSegments"] - st_o["closestRays"]) <= 0.01 * st_o["closestRays"] + 8
    ctx.close(); orc.close()


def test_tile_path_equals_whole_frame_path(oracle_mod):
    """numTiles Render() tile draws with the reference frameNum schedule == one ptb_render_samples pass (bitwise)."""
    sc = scene_at("cornell_box_orig", 100, 72, 48, 32, 3)      # over-hanging last column and top row (Q15)
    a = _ctx(sc); b = _ctx(sc)
    ntx, nty = 3, 3
    for s in (1, 2):
        for j in range(ntx * nty):
            tx, ty = j % ntx, nty - 1 - j // ntx
            a.render_tile(tx, ty, 2 + (s - 1) * ntx * nty + j)
    b.render_samples(1, 2)
    assert a.read_accum().tobytes() == b.read_accum().tobytes()
    orc = oracle_mod.Oracle(sc)
    acc = np.zeros((72, 100, 4), np.float32)
    for j in range(ntx * nty):
        orc.render_tile(j % ntx, nty - 1 - j // ntx, 2 + j, acc)
    assert acc.tobytes() == orc.render(1, 1).tobytes()
    a.close(); b.close(); orc.close()


def test_sample_stride_sharding_sums_to_single(oracle_mod):
    """multi-GPU sample sharding: passes {1,3,5,7} + {2,4,6,8} rendered separately sum to passes 1..8."""
    sc = scene_at("hyperion_sphere_light", 160, 90, 64, 36)
    one = _ctx(sc); one.render_samples(1, 8)
    ref = one.read_accum()
    parts = []
    for r in range(2):
        c = _ctx(sc, samples_per_wave=3); c.render_samples(1 + r, 4, 2); parts.append(c.read_accum()); c.close()
    np.testing.assert_allclose(parts[0] + parts[1], ref, rtol=1e-5, atol=1e-5)
    one.close()


@pytest.mark.parametrize("aces,simple,tm", [(0, 0, 1), (1, 0, 1), (1, 1, 1), (0, 0, 0)])
def test_tonemap_bytes_match_oracle(aces, simple, tm, oracle_mod):
    sc = scene_at("cornell_box_orig", 96, 64, 48, 32)
    sc.renderOptions.enableAces = bool(aces); sc.renderOptions.simpleAcesFit = bool(simple); sc.renderOptions.enableTonemap = bool(tm)
    ctx = _ctx(sc)
    rng = np.random.default_rng(0)
    acc = (rng.random((64, 96, 4), dtype=np.float32) ** 3 * 40).astype(np.float32); acc[..., 3] = 8.0
    acc[0, 0, :3] = 0; acc[0, 1, :3] = np.nan; acc[0, 2, :3] = 1e9
    ctx.write_accum(acc)
    g = ctx.read_output(1.0 / 8).astype(np.int32)
    o = oracle_mod.tonemap(acc, np.float32(1.0 / 8), sc.renderOptions).astype(np.int32)
    d = np.abs(g - o)
    assert d.max() <= 1 and (d > 0).mean() < 0.002, f"max {d.max()} frac {(d > 0).mean()}"
    ctx.close()


def test_background_alpha_and_transparent_checker(oracle_mod):
    sc = scene_at("cornell_box_orig", 96, 64, 48, 32)
    sc.camera.position = np.array([0.276, 0.275, -3.0], np.float32)      # box small in frame: background pixels exist
    sc.renderOptions.transparentBackground = True
    ctx = _ctx(sc); orc = oracle_mod.Oracle(sc)
    ctx.render_samples(1, 2)
    g, o = ctx.read_accum(), orc.render(1, 2)
    np.testing.assert_allclose(g[..., 3], o[..., 3], atol=1e-6)
    assert (g[..., 3] == 0).any() and (g[..., 3] == 2).any()
    from glsl_pathtracer_b200 import scene_io as sio
    f = sio.derive_features(sc)
    gb = ctx.read_output(0.5).astype(np.int32); ob = oracle_mod.tonemap(o, np.float32(0.5), sc.renderOptions, f).astype(np.int32)
    assert np.abs(gb - ob).max() <= 1
    ctx.close(); orc.close()


def test_uniform_light_hide_emitters_and_mollification(oracle_mod):
    sc = scene_at("cornell_box_sphere", 96, 96, 48, 48)
    ro = sc.renderOptions
    ro.enableUniformLight = True; ro.uniformLightCol = (0.5, 0.6, 0.7); ro.hideEmitters = True
    ro.enableRoughnessMollification = True; ro.roughnessMollificationAmt = 0.6; ro.enableRR = False; ro.maxDepth = 4
    ctx = _ctx(sc); orc = oracle_mod.Oracle(sc)
    ctx.render_samples(1, 4)
    r = rel_mse(orc.render(1, 4), ctx.read_accum())
    assert r <= 1e-3, r
    ctx.close(); orc.close()


# ---------------------------------------------------------------- Renderer mirror -----------------------------------
def test_renderer_state_machine_matches_reference_semantics(oracle_mod):
    """Drive the mirror of GLSLPT::Renderer like Main.cpp's loop (Update -> Render) and compare with the oracle."""
    from glsl_pathtracer_b200.renderer import Renderer
    sc = scene_at("cornell_box_orig", 96, 64, 48, 32)
    sc.renderOptions.maxSpp = 3                      # Q1: renders maxSpp-1 = 2 passes
    r = Renderer(sc, "shaders/")
    assert r.GetSampleCount() == 1 and sc.dirty
    n_calls = 0
    for _ in range(64):
        r.Update(0.016); r.Render(); r.Present(); n_calls += 1
    assert r.GetSampleCount() == 3 and abs(r.GetProgress() - 100.0) < 1e-6
    acc = r.ctx.read_accum()
    orc = oracle_mod.Oracle(sc)
    o = orc.render(1, 2)
    assert rel_mse(o, acc) <= 1e-3
    img, w, h = r.GetOutputBuffer()
    assert (w, h) == (96, 64) and img.shape == (64, 96, 4)
    ref = oracle_mod.tonemap(o, np.float32(1.0) / np.float32(2), sc.renderOptions).astype(np.int32)
    assert (np.abs(img.astype(np.int32) - ref) > 2).mean() < 0.03
    # RenderSamples fast path gives the same accumulation as the tile loop
    sc2 = scene_at("cornell_box_orig", 96, 64, 48, 32)
    r2 = Renderer(sc2, "shaders/")
    r2.Update(0.0); r2.Render()                     # dirty/preview frame
    r2.RenderSamples(2)
    assert r2.ctx.read_accum().tobytes() == acc.tobytes() and r2.GetSampleCount() == 3
    orc.close()


def test_preview_matches_oracle(oracle_mod):
    """preview.glsl: quarter-resolution, frame 1 seeds, full-resolution jitter scale, depth forced to 2 (Renderer.cpp:798)."""
    sc = scene_at("cornell_box_orig", 128, 128, 64, 64, 5)
    ctx = _ctx(sc); orc = oracle_mod.Oracle(sc)
    p = ctx.render_preview(32, 32)
    o = orc.render_preview(32, 32)
    assert p.shape == (32, 32, 4) and np.isfinite(p).all() and p[..., :3].max() > 0
    assert rel_mse(o, p) <= 1e-3 and np.isclose(p, o, rtol=1e-3, atol=1e-4).all(axis=-1).mean() > 0.95
    orc.close()
    # the preview never touches the accumulation buffer (Renderer.cpp:555-560)
    assert not ctx.read_accum().any()
    ctx.close()


@pytest.mark.parametrize("name", ["cornell_box_orig", "hyperion_rect_lights"])
def test_live_instance_edit_equals_fresh_context(name, oracle_mod):
    """N2 live update with data made by the reference itself (Scene::RebuildInstances on a scaled + translated instance,
    tests/golden/instance_edit.npz): ptb_update_instances on a running context == a context created from the edited scene, bitwise;
    node array byte-identical; image matches the oracle."""
    from conftest import edited_scene, ROOT
    sc, sc2 = edited_scene(name, 160, 90, 80, 45)
    ctx = _ctx(sc)
    ctx.render_samples(1, 2); before = ctx.read_accum()
    ctx.update_instances(sc2.transforms, sc2.materials, sc2.nodes[sc2.topLevelIndex:])
    assert ctx.read_nodes().tobytes() == np.ascontiguousarray(sc2.nodes, np.float32).tobytes()
    ctx.reset_accum(); ctx.render_samples(1, 4); after = ctx.read_accum()
    fresh = _ctx(sc2); fresh.render_samples(1, 4)
    assert after.tobytes() == fresh.read_accum().tobytes() and not np.array_equal(before, after)
    orc = oracle_mod.Oracle(sc2)
    o = orc.render(1, 4)
    assert rel_mse(np.nan_to_num(o) / 4, np.nan_to_num(after) / 4) <= 1e-3
    rays = orc.camera_rays(1)
    ctx.set_cull(False)
    g, h = ctx.trace_closest(rays), orc.trace_closest(rays)
    assert np.array_equal(g["primSlot"], h["primSlot"]) and np.array_equal(g["t"].view(np.uint32), h["t"].view(np.uint32))
    ctx.close(); fresh.close(); orc.close()


def test_update_envmap_equals_fresh_context(oracle_mod):
    """N2 env-map swap (Renderer.cpp:668-692): ptb_update_envmap with a different image of different size on a running context ==
    a context created with that environment, bitwise; image matches the oracle."""
    import copy
    sc = scene_at("ibl_spheres", 160, 90, 80, 45)
    img2 = np.ascontiguousarray((np.roll(sc.envImg, 97, axis=1) * np.float32(0.5) + np.float32(0.05))[::2, ::2], np.float32)
    lum = (np.float32(0.212671) * img2[..., 0] + np.float32(0.715160) * img2[..., 1] + np.float32(0.072169) * img2[..., 2]).astype(np.float32)
    cdf2 = np.cumsum(lum.ravel(), dtype=np.float32).reshape(lum.shape)       # EnvironmentMap::BuildCDF: flat fp32 running sum
    sc2 = copy.deepcopy(sc); sc2.envImg, sc2.envCdf, sc2.envTotalSum = img2, cdf2, float(cdf2[-1, -1])
    ctx = _ctx(sc)
    ctx.render_samples(1, 2); before = ctx.read_accum()
    ctx.update_envmap(img2, cdf2, sc2.envTotalSum)
    ctx.reset_accum(); ctx.render_samples(1, 4); after = ctx.read_accum()
    fresh = _ctx(sc2); fresh.render_samples(1, 4)
    assert after.tobytes() == fresh.read_accum().tobytes() and not np.array_equal(before, after)
    orc = oracle_mod.Oracle(sc2)
    o = orc.render(1, 4)
    assert rel_mse(np.nan_to_num(o) / 4, np.nan_to_num(after) / 4) <= 1e-3
    close = np.isclose(after[..., :3], o[..., :3], rtol=1e-3, atol=4e-4).all(axis=-1)
    assert close.mean() > 0.9
    ctx.close(); fresh.close(); orc.close()


def test_update_instances_moves_geometry(oracle_mod):
    import copy
    sc = scene_at("cornell_box_orig", 96, 96, 48, 48)
    ctx = _ctx(sc)
    rays = ctx.camera_rays(1)
    before = ctx.trace_closest(rays)
    # swap the materials of all instances to material 0 by rewriting the TLAS leaves, and keep transforms
    sc2 = copy.deepcopy(sc)
    tl = sc2.nodes[sc2.topLevelIndex:].copy()
    leaf = tl[:, 8] < 0
    tl[leaf, 7] = 0
    ctx.update_instances(sc2.transforms, sc2.materials, tl)
    after = ctx.trace_closest(rays)
    assert np.array_equal(before["primSlot"], after["primSlot"]) and (after["matID"][after["kind"] == 1] == 0).all()
    sc2.nodes[sc2.topLevelIndex:] = tl
    orc = oracle_mod.Oracle(sc2)
    o = orc.trace_closest(rays)
    assert np.array_equal(o["matID"], after["matID"]) and np.array_equal(o["t"].view(np.uint32), after["t"].view(np.uint32))
    assert ctx.read_nodes().tobytes() == sc2.nodes.tobytes()
    ctx.close(); orc.close()


# ---------------------------------------------------------------- full BASELINE size ---------------------------------
def test_full_size_hyperion_1080p(oracle_mod):
    """BASELINE configs[2] at its real size: primary-ray hits bit-exact for all 2 073 600 pixels, a 2-spp RNG-matched image against the
    oracle, and size-independent properties of the wavefront (strided shards sum to the whole; tile passes equal a whole-frame pass)."""
    sc = scene_at("hyperion_rect_lights", 1920, 1080)
    assert (sc.renderOptions.tileWidth, sc.renderOptions.tileHeight, sc.renderOptions.maxDepth) == (256, 144, 3)
    ctx = _ctx(sc); orc = oracle_mod.Oracle(sc); orc_c = oracle_mod.Oracle(sc, cull=True)
    rays = orc.camera_rays(1)
    assert np.array_equal(ctx.camera_rays(1).view(np.uint32), rays.view(np.uint32))
    g, o = ctx.trace_closest(rays), orc_c.trace_closest(rays)
    for f in ("kind", "instance", "matID", "primSlot", "triIDx", "lightIdx"):
        assert np.array_equal(g[f], o[f]), f
    assert np.array_equal(g["t"].view(np.uint32), o["t"].view(np.uint32))
    ctx.set_cull(False)
    g2, o2 = ctx.trace_closest(rays), orc.trace_closest(rays)
    assert g2.tobytes() == o2.tobytes()
    ctx.set_cull(True)
    ctx.render_samples(1, 2)
    a = ctx.read_accum()
    ref = orc.render(1, 2)
    assert rel_mse(ref / 2, a / 2) <= 1e-3
    assert np.isclose(a[..., :3], ref[..., :3], rtol=1e-3, atol=1e-4).all(axis=-1).mean() > 0.97
    # shards: pass 1 on one context + pass 2 on another == both passes on one
    s1 = _ctx(sc); s1.render_samples(1, 1, 2); s2 = _ctx(sc); s2.render_samples(2, 1, 2)
    np.testing.assert_allclose(s1.read_accum() + s2.read_accum(), a, rtol=1e-6, atol=1e-6)
    # one over-hanging tile (top-right: 128 x 72 visible pixels of a 256 x 144 tile) equals the same rectangle of a whole-frame pass
    t = _ctx(sc); t.render_tile(7, 7, 2 + 7)          # first tile of the schedule is (0,7); (7,7) is ordinal 7 of pass 1
    ta = t.read_accum()
    assert ta[1008:, 1792:].tobytes() == s1.read_accum()[1008:, 1792:].tobytes() and not ta[:1008].any()
    for c in (ctx, s1, s2, t): c.close()
    orc.close(); orc_c.close()


# ---------------------------------------------------------------- converged gates (G3 as the survey words it) ---------------------
def _signed_bias(g, o):
    """mean signed per-pixel difference of the rgb means, and the mean radiance"""
    d = (g[..., :3].astype(np.float64) - o[..., :3].astype(np.float64)).mean(axis=-1).ravel()
    return float(d.mean()), float(o[..., :3].mean())


@pytest.mark.parametrize("name,spp", [("hyperion_sphere_light", 256), ("hyperion_rect_lights", 128)])
def test_converged_image_is_unbiased_against_the_oracle(name, spp, oracle_mod):
    """480x270 at 256 / 128 spp.  hyperion_sphere_light is the scene where only a minority of the pixels stays RNG-matched
    (the reference's sphere light shadows its own grazing NEE rays within rounding of EPS, anyhit.glsl:56-61): "equal in expectation" is
    tested here — relMSE <= 1e-3 at equal spp AND the mean signed difference (a) within 3 sigma of the Monte-Carlo noise of the converged
    mean itself (sigma from two independent halves of the oracle's own passes) and (b) below 1e-4 of the mean radiance.
    hyperion_rect_lights (glass, clearcoat, near-mirror metal) bounds the bias of the 2-ulp shading arithmetic the same way.
    (With div.approx / rsqrt in the sphere-light sample geometry this test measured -0.5 % on hyperion_sphere_light; see SampleOneLight.)"""
    sc = scene_at(name, 480, 270, 256, 144)
    ctx = _ctx(sc); orc = oracle_mod.Oracle(sc)
    ctx.render_samples(1, spp)
    g = np.nan_to_num(ctx.read_accum() / spp)
    h = spp // 2
    o1 = np.nan_to_num(orc.render(1, h) / h); o2 = np.nan_to_num(orc.render(1 + h, h) / h)
    o = (o1 + o2) / 2
    r = rel_mse(o, g)
    mean, level = _signed_bias(g, o)
    half = (o1[..., :3].astype(np.float64) - o2[..., :3].astype(np.float64)).mean(axis=-1).ravel()
    sigma = float(half.std(ddof=1) / 2 / np.sqrt(half.size))          # standard error of the converged image's mean radiance
    print(f"{name}: relMSE {r:.3e}, signed bias {mean:+.3e} = {mean / level:+.2e} of the mean radiance {level:.4f}; MC sigma of that mean {sigma:.3e}; matched pixels "
          f"{np.isclose(g[..., :3], o[..., :3], rtol=1e-3, atol=1e-4).all(axis=-1).mean():.3f}")
    assert r <= 1e-3, f"relMSE {r}"
    assert abs(mean) <= 3.0 * sigma, f"signed bias {mean} exceeds 3 sigma ({sigma})"
    assert abs(mean) <= 1e-4 * level
    ctx.close(); orc.close()


# ---------------------------------------------------------------- the other BASELINE configs at their full sizes -------------------
@pytest.mark.parametrize("name,w,h,minfrac", [("volume_cube", 1920, 1080, 0.97), ("ibl_spheres", 1920, 1080, 0.95), ("hyperion_sphere_light", 1920, 1080, 0.85),
                                             ("instancing", 3840, 2160, 0.93)])
def test_full_size_other_baseline_configs(name, w, h, minfrac, oracle_mod):
    """BASELINE configs[1], [3], [4] (and the sphere-light variant of [2]) at the sizes they are quoted on: every primary ray bit-exact in
    both traversal variants (tile.glsl:41-68 + closest_hit.glsl), and a 1-spp RNG-matched image of the whole frame against the oracle.
    The 4K instancing frame is 8.3 M pixels x depth 8: 66 M path slots per default wave (auto wave split, uint32 slot indices)."""
    sc = scene_at(name, w, h)
    ctx = _ctx(sc); orc = oracle_mod.Oracle(sc); orc_c = oracle_mod.Oracle(sc, cull=True)
    rays = orc.camera_rays(1)
    assert np.array_equal(ctx.camera_rays(1).view(np.uint32), rays.view(np.uint32))
    g, o = ctx.trace_closest(rays), orc_c.trace_closest(rays)
    for f in ("kind", "instance", "matID", "primSlot", "triIDx", "lightIdx"):
        assert np.array_equal(g[f], o[f]), f
    assert np.array_equal(g["t"].view(np.uint32), o["t"].view(np.uint32))
    ctx.set_cull(False)
    assert ctx.trace_closest(rays).tobytes() == orc.trace_closest(rays).tobytes()
    ctx.set_cull(True)
    ctx.render_samples(1, 1)
    a = np.nan_to_num(ctx.read_accum()); ref = np.nan_to_num(orc.render(1, 1))
    close = np.isclose(a[..., :3], ref[..., :3], rtol=1e-3, atol=1e-4).all(axis=-1).mean()
    r = rel_mse(ref, a)
    print(f"{name} {w}x{h}: 1-spp relMSE {r:.3e}, RNG-matched pixels {close:.4f}")
    assert close > minfrac, close
    # ONE sample per pixel is not a converged image: a pixel whose path diverged differs by O(1), and the error of a k-spp mean of RNG-matched
    # paths falls as 1/k.  The relMSE <= 1e-3 gate is held at 8 spp by test_image_matches_oracle_rng_matched and at 128/256 spp by the
    # converged test above; here the same gate is scaled to one sample (8 x 1e-3).
    assert r <= 8e-3, r
    # size-independent property at the full size: two strided shards sum to a two-pass render
    ctx.render_samples(2, 1)
    s1 = _ctx(sc); s1.render_samples(1, 1, 2); s1.render_samples(2, 1, 2)
    np.testing.assert_allclose(np.nan_to_num(s1.read_accum()), np.nan_to_num(ctx.read_accum()), rtol=1e-6, atol=1e-6)
    for c in (ctx, s1): c.close()
    orc.close(); orc_c.close()
