"""G1 gate: the CUDA traversal, called through the C ABI, returns bit-identical hit kind / instance /
material / primitive slot / triangle id and t (tolerance 1e-5 relative, observed: bit-equal) versus the oracle's
reference-faithful host traversal of the same flattened BVH, on primary, random and bounce-like rays."""
import numpy as np
import pytest
from conftest import scene_at
from raysets import random_rays, boundary_rays, light_rays, any_hit_distances, assert_hits_equal, assert_hits_nearly_equal

pytestmark = pytest.mark.gpu
SCENES = ["cornell_box_orig", "cornell_box_sphere", "hyperion_rect_lights", "hyperion_sphere_light", "volume_cube", "teapot", "ibl_spheres", "instancing", "gltf_mix"]


def _ctx(sc):
    from glsl_pathtracer_b200 import capi
    return capi.Context(sc)


@pytest.mark.parametrize("name", SCENES)
def test_nodes_roundtrip_byte_exact(name, oracle_mod):
    sc = scene_at(name, 64, 64, 32, 32)
    ctx = _ctx(sc)
    assert ctx.read_nodes().tobytes() == np.ascontiguousarray(sc.nodes, np.float32).tobytes()
    assert ctx.stack_depth() <= 64
    ctx.close()


@pytest.mark.parametrize("name", SCENES)
def test_primary_rays_match_oracle(name, oracle_mod):
    sc = scene_at(name, 480, 270, 128, 72)
    ctx = _ctx(sc); orc = oracle_mod.Oracle(sc)
    rays_o = orc.camera_rays(1)
    rays_g = ctx.camera_rays(1)
    assert np.array_equal(rays_o.view(np.uint32), rays_g.view(np.uint32)), "pinhole primary rays must be bit-identical"
    orc_c = oracle_mod.Oracle(sc, cull=True)
    for depth in (0, 1):
        ctx.set_cull(False)                           # the reference's unculled traversal: bit-exact against the faithful host traversal
        faithful = orc.trace_closest(rays_o, depth)
        assert_hits_equal(ctx.trace_closest(rays_o, depth), faithful)
        ctx.set_cull(True)                            # default (t-culled) traversal: bit-exact against the culled host traversal ...
        got = ctx.trace_closest(rays_o, depth)
        assert_hits_equal(got, orc_c.trace_closest(rays_o, depth))
        assert_hits_nearly_equal(got, faithful)       # ... which equals the faithful one except on exact ties
    orc_c.close()
    ctx.close(); orc.close()


@pytest.mark.parametrize("name", SCENES)
def test_random_and_bounce_rays_match_oracle(name, oracle_mod):
    sc = scene_at(name, 64, 64, 32, 32)
    ctx = _ctx(sc); orc = oracle_mod.Oracle(sc)
    rays = random_rays(sc, 200_000, 7)
    h_o = orc.trace_closest(rays, 1)
    ctx.set_cull(False)
    assert_hits_equal(ctx.trace_closest(rays, 1), h_o)
    ctx.set_cull(True)
    orc_c = oracle_mod.Oracle(sc, cull=True)
    got = ctx.trace_closest(rays, 1)
    assert_hits_equal(got, orc_c.trace_closest(rays, 1))
    assert_hits_nearly_equal(got, h_o)
    ctx.set_cull(False)
    # bounce-like rays: start on the surfaces found above, cosine-ish random directions
    hit = h_o["kind"] == 1
    p = rays[hit, :3] + rays[hit, 3:] * h_o["t"][hit, None]
    rng = np.random.default_rng(11)
    d = rng.normal(size=p.shape).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
    b = np.concatenate([(p + d * np.float32(0.0003)).astype(np.float32), d], axis=1)
    assert_hits_equal(ctx.trace_closest(b, 1), orc.trace_closest(b, 1))
    ctx.set_cull(True)
    assert_hits_equal(ctx.trace_closest(b, 1), orc_c.trace_closest(b, 1))
    ctx.close(); orc.close(); orc_c.close()


@pytest.mark.parametrize("name", SCENES)
def test_culled_traversal_matches_oracle_culled(name, oracle_mod):
    """ptb_set_cull(1) and the oracle's culled variant implement the same rule (entry > t * 1.00001): bit-identical to each other,
    and identical to the faithful traversal except for exact-tie edge cases."""
    sc = scene_at(name, 320, 180, 80, 60)
    ctx = _ctx(sc); orc = oracle_mod.Oracle(sc)
    rays = np.concatenate([orc.camera_rays(1), random_rays(sc, 100_000, 3)])
    ref = orc.trace_closest(rays, 1)
    ctx.set_cull(True)
    got = ctx.trace_closest(rays, 1)
    assert_hits_nearly_equal(got, ref)
    orc_c = oracle_mod.Oracle(sc, cull=True)
    assert_hits_equal(got, orc_c.trace_closest(rays, 1))
    ctx.close(); orc.close(); orc_c.close()


@pytest.mark.parametrize("name", SCENES)
def test_any_hit_matches_oracle(name, oracle_mod):
    sc = scene_at(name, 64, 64, 32, 32)
    ctx = _ctx(sc); orc = oracle_mod.Oracle(sc)
    rays = random_rays(sc, 100_000, 5)
    rng = np.random.default_rng(9)
    ext = float(np.linalg.norm(np.array(sc.sceneBounds[1]) - np.array(sc.sceneBounds[0])))
    md = (rng.random(len(rays), dtype=np.float32) * ext).astype(np.float32)
    md[::3] = np.float32(1e6 - 0.0003)
    ctx.set_cull(False)
    a, b = ctx.trace_any(rays, md), orc.trace_any(rays, md)
    assert np.array_equal(a, b), f"{np.count_nonzero(a != b)} occlusion mismatches"
    ctx.set_cull(True)
    assert np.array_equal(ctx.trace_any(rays, md), b)
    ctx.close(); orc.close()


@pytest.mark.parametrize("name", SCENES)
def test_adversarial_and_light_rays_match_oracle(name, oracle_mod):
    """Origins exactly on node-box planes (in-plane axis-parallel and denormal directions: 0 * inf NaNs, 1/d = inf) and rays aimed at quad-light corners / edges
    (shared-plane light groups and their cell grid): closest hit bit-identical in both variants; any-hit — over the 4-wide hierarchy where it is admissible —
    the same boolean as the oracle's reference-order traversal."""
    sc = scene_at(name, 64, 64, 32, 32)
    ctx = _ctx(sc); orc = oracle_mod.Oracle(sc); orc_c = oracle_mod.Oracle(sc, cull=True)
    rays = np.concatenate([boundary_rays(sc, 100_000), light_rays(sc, 100_000)])
    for depth in (0, 1):
        ctx.set_cull(False); assert_hits_equal(ctx.trace_closest(rays, depth), orc.trace_closest(rays, depth))
        ctx.set_cull(True); assert_hits_equal(ctx.trace_closest(rays, depth), orc_c.trace_closest(rays, depth))
    md = any_hit_distances(sc, len(rays))
    want = orc.trace_any(rays, md)
    for cull in (False, True):
        ctx.set_cull(cull)
        got = ctx.trace_any(rays, md)
        assert np.array_equal(got, want), f"cull={cull}: {np.count_nonzero(got != want)} occlusion mismatches"
    ctx.close(); orc.close(); orc_c.close()


def test_empty_and_invalid_inputs():
    from glsl_pathtracer_b200 import capi
    sc = scene_at("cornell_box_orig", 32, 32, 16, 16)
    ctx = _ctx(sc)
    assert len(ctx.trace_closest(np.zeros((0, 6), np.float32))) == 0
    with pytest.raises(capi.PtbError):
        ctx.render_tile(99, 0, 2)
    with pytest.raises(capi.PtbError):
        ctx.render_samples(0, 1)
    # degenerate rays: zero direction / NaN origin must not hang or crash and must agree with "miss or something" deterministically
    rays = np.zeros((4, 6), np.float32); rays[1, 0] = np.nan; rays[2, 3:] = [0, 0, 1]; rays[3, 3:] = [1e-30, 1, 0]
    h1, h2 = ctx.trace_closest(rays), ctx.trace_closest(rays)
    assert h1.tobytes() == h2.tobytes()
    ctx.close()
