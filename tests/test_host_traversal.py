"""The product's hit-deciding SOURCE — csrc/ptb_device.cuh (two-level traversal state machine, analytic-light tests, camera rays) and
csrc/ptb_derive.cpp (derived 64-byte node / 48-byte triangle / instance / light-group layout) — compiled for the HOST
(tests/host_harness) and compared with the oracle bit for bit.  Every operation in that code is an IEEE round-to-nearest intrinsic that
ptxas cannot contract, so the host build executes the same arithmetic as the sm_100a build; this catches logic errors in traversal
changes where no GPU exists.  The G1 gate proper (tests/test_gpu_trace.py) runs the CUDA build through the C ABI on the B200."""
import numpy as np
import pytest
from conftest import scene_at
from raysets import random_rays, bounce_rays, boundary_rays, light_rays, any_hit_distances, assert_hits_equal, assert_hits_nearly_equal
from host_harness.binding import HostTrav

SCENES = ["cornell_box_orig", "cornell_box_sphere", "hyperion_rect_lights", "hyperion_sphere_light", "volume_cube", "teapot", "ibl_spheres", "instancing", "gltf_mix"]


@pytest.mark.parametrize("name", SCENES)
def test_host_compiled_traversal_equals_oracle(name, oracle_mod):
    sc = scene_at(name, 160, 90, 48, 32)          # 4 x 3 tiles, last column and top row over-hang
    ht = HostTrav(sc); orc = oracle_mod.Oracle(sc); orc_c = oracle_mod.Oracle(sc, cull=True)
    prim = orc.camera_rays(1)
    for sample in (1, 5):          # per-pixel arithmetic and the per-column / per-row tables give the oracle's rays bit for bit (over-hanging tiles included)
        want = orc.camera_rays(sample)
        for tables in (False, True):
            assert np.array_equal(ht.camera_rays(sample, tables).view(np.uint32), want.view(np.uint32)), f"camera rays, sample {sample}, tables {tables}"
    rays = np.concatenate([prim, random_rays(sc, 30_000, 7)])
    for depth in (0, 1):
        ht.set_cull(False)
        faithful = orc.trace_closest(rays, depth)
        assert_hits_equal(ht.trace_closest(rays, depth), faithful)
        ht.set_cull(True)
        got = ht.trace_closest(rays, depth)
        assert_hits_equal(got, orc_c.trace_closest(rays, depth))
        assert_hits_nearly_equal(got, faithful, max_frac=1e-4)
    b = bounce_rays(rays, faithful)
    ht.set_cull(False); assert_hits_equal(ht.trace_closest(b, 1), orc.trace_closest(b, 1))
    ht.set_cull(True); assert_hits_equal(ht.trace_closest(b, 1), orc_c.trace_closest(b, 1))
    # closest hit on adversarial rays (origins exactly on node-box planes, in-plane axis-parallel and denormal directions)
    adv = np.concatenate([boundary_rays(sc, 30_000), light_rays(sc, 30_000)])
    ht.set_cull(False); assert_hits_equal(ht.trace_closest(adv, 1), orc.trace_closest(adv, 1))
    ht.set_cull(True); assert_hits_equal(ht.trace_closest(adv, 1), orc_c.trace_closest(adv, 1))
    # any-hit: occlusion is the same boolean in both traversal variants AND over the 4-wide hierarchy (the production path of k_shadow)
    ar = np.concatenate([random_rays(sc, 30_000, 5), b[:20_000], adv])
    md = any_hit_distances(sc, len(ar))
    want = orc.trace_any(ar, md)
    assert ht.wide_nodes() > 0 or len(sc.nodes) < 4, "the wide hierarchy must be available for every fixture scene"
    for cull in (False, True):
        ht.set_cull(cull)
        for wide in (False, True):
            got = ht.trace_any(ar, md, wide=wide)
            assert np.array_equal(got, want), f"cull={cull} wide={wide}: {np.count_nonzero(got != want)} occlusion mismatches"
            if wide:
                assert ht.fallbacks < 0.6 * len(ar)      # the generic rays really went through the wide hierarchy
    assert ht.stack_depth() <= 96
    high, bound = ht.any_stack()
    assert 0 < high <= bound, f"any-hit stack reached {high} entries, the kernels reserve {bound}"
    ht.close(); orc.close(); orc_c.close()


@pytest.mark.parametrize("name", SCENES)
def test_tlas_rebuild_is_byte_identical_to_the_reference(name):
    """N3: the library's TLAS builder (instance world boxes, centre split, pre-order layout) applied to a scene's transforms reproduces the TLAS slice the
    reference's own Scene::createTLAS + Bvh::Build + BvhTranslator::ProcessTLAS put into the fixture, byte for byte (10 001 instances in `instancing`)."""
    from host_harness.binding import build_tlas
    sc = scene_at(name, 64, 64)
    nodes = np.ascontiguousarray(sc.nodes, np.float32).reshape(-1, 9)
    tlas, height = build_tlas(nodes, sc.topLevelIndex, sc.transforms)
    assert tlas.tobytes() == nodes[sc.topLevelIndex:].tobytes()
    assert height + 1 == sc.tlasHeight          # the fixture counts levels, Bvh::m_height counts edges


@pytest.mark.parametrize("name", ["cornell_box_orig", "hyperion_rect_lights"])
def test_tlas_rebuild_after_an_instance_edit_matches_the_reference(name):
    """... and for the instance edit the reference made itself (Scene::RebuildInstances on a moved + scaled instance, tests/golden/instance_edit.npz)."""
    from conftest import edited_scene
    from host_harness.binding import build_tlas
    sc, sc2 = edited_scene(name, 64, 64)
    nodes = np.ascontiguousarray(sc.nodes, np.float32).reshape(-1, 9)
    want = np.ascontiguousarray(sc2.nodes, np.float32).reshape(-1, 9)[sc.topLevelIndex:]
    leaves = want[want[:, 8] < 0]
    mats = np.zeros(len(leaves), np.int32); mats[(-leaves[:, 8] - 1).astype(int)] = leaves[:, 7].astype(np.int32)     # meshInstances[i].materialID after the edit
    tlas, _ = build_tlas(nodes, sc.topLevelIndex, sc2.transforms, mats)     # old node array (BLAS boxes, BLAS roots) + NEW transforms and material ids
    assert tlas.tobytes() == want.tobytes()


def test_env_cdf_guide_table_returns_the_texel_of_the_reference_search():
    """SampleEnvMap's BinarySearch (envmap.glsl:28-55) through the guide table of ptbd_build_env_guide == the two binary searches, bit for bit: on the
    fixture's CDF and on adversarial ones (long runs of equal values = black texels, a single bright texel, values on / next to every CDF entry,
    0, totalSum and beyond).  A non-monotone CDF gets no table."""
    from host_harness import binding as hb
    rng = np.random.default_rng(5)
    sc = scene_at("ibl_spheres", 64, 36)
    h, w = sc.envImg.shape[:2]
    cases = [(np.asarray(sc.envCdf, np.float32).ravel(), w, h, float(sc.envTotalSum))]
    for (cw, ch) in ((64, 32), (37, 19), (2048, 8)):
        lum = rng.random(cw * ch).astype(np.float32) ** 8
        lum[rng.random(cw * ch) < 0.6] = 0.0                         # runs of equal CDF values
        lum[rng.integers(cw * ch)] = 5000.0                          # one texel holds most of the mass
        cdf = np.zeros(cw * ch, np.float32); acc = np.float32(0)
        for i, v in enumerate(lum):                                  # the reference's float running sum (EnvironmentMap.cpp:52-58)
            acc = np.float32(acc + v); cdf[i] = acc
        cases.append((cdf, cw, ch, float(cdf[-1])))
    for cdf, cw, ch, total in cases:
        v = [rng.random(20000).astype(np.float32) * np.float32(total), cdf, np.nextafter(cdf, np.float32(-1)), np.nextafter(cdf, np.float32(1e30)),
             np.array([0.0, total, total * 1.5, -1.0, 1e-30], np.float32)]
        v = np.concatenate(v).astype(np.float32)
        have, fast, ref = hb.env_search(cdf, cw, ch, total, v)
        assert have
        assert fast.tobytes() == ref.tobytes()
    bad = cases[1][0].copy(); bad[100] = bad[99] - 1.0
    have, fast, ref = hb.env_search(bad, 64, 32, float(bad[-1]), np.array([1.0], np.float32))
    assert not have and fast.tobytes() == ref.tobytes()


@pytest.mark.parametrize("order", ["0", "1", "2", "3", "auto"])
def test_wide_hierarchy_slot_order_does_not_change_occlusion_and_bounds_the_stack(order, monkeypatch):
    """PTB_WIDE_ORDER=1 permutes the children of every 4-wide node by ascending stack need (deepest subtree last: nothing waits on the stack while it is
    traversed), which shrinks the worst-case any-hit stack (hyperion 28 -> 24 entries, instancing 42 -> 32).  Any-hit is order-free: same booleans as the
    oracle's reference-order traversal, and the stack never exceeds the bound the kernels reserve shared memory for."""
    from host_harness import binding as hb
    from oracle import binding as ob
    if order == "auto": monkeypatch.delenv("PTB_WIDE_ORDER", raising=False)      # the shipped rule: largest box first unless its bound exceeds 31 entries
    else: monkeypatch.setenv("PTB_WIDE_ORDER", order)
    bounds = {}
    for name in ("hyperion_rect_lights", "instancing", "ibl_spheres"):
        sc = scene_at(name, 96, 54)
        orc = ob.Oracle(sc); ht = hb.HostTrav(sc)
        prim = orc.camera_rays(1)
        rays = np.concatenate([prim, random_rays(sc, 40_000, 11), boundary_rays(sc, 20_000)])
        md = any_hit_distances(sc, len(rays))
        md[::3] = 1e6                                   # long rays: every box along the line is entered
        want = orc.trace_any(rays, md)
        ht.set_cull(True)
        assert np.array_equal(ht.trace_any(rays, md, wide=True), want)
        high, bound = ht.any_stack()
        assert 0 < high <= bound, f"{name}: any-hit stack reached {high} entries, the kernels reserve {bound}"
        bounds[name] = bound
        ht.close(); orc.close()
    if order == "1":
        assert bounds["hyperion_rect_lights"] <= 25 and bounds["instancing"] <= 32
    if order == "auto":
        assert bounds["hyperion_rect_lights"] <= 31 and bounds["instancing"] <= 32


@pytest.mark.parametrize("block_major,max_lps", [(False, 5), (True, 0), (True, 3), (True, 5)])
def test_slot_layout_of_a_wave_is_a_bijection_with_coherent_tiles(block_major, max_lps):
    """groupToPixel / slotOfSample (ptb_device.cuh) for pass counts 1..96 and a padded rectangle: every (pass, pixel) of the padded 8x4 blocks has exactly one slot,
    slotOfSample inverts the map, a 32-slot group is 2^lps passes of one pixel sub-block, and in block-major order the 32 * S slots of an 8x4 block are consecutive
    (that is what makes a 2048-slot sorting tile hold all passes of neighbouring pixels)."""
    from host_harness import binding as hb
    w, h = 21, 10                                   # padded to 24 x 12: 3 x 3 blocks with off-image pixels
    vw, vh = 24, 12
    for ns in (1, 2, 3, 4, 6, 8, 12, 16, 24, 32, 64, 96):
        m, lps = hb.slot_map(w, h, ns, block_major, max_lps)
        s, px, py, back = m[:, 0], m[:, 1], m[:, 2], m[:, 3]
        assert (s >= 0).all() and (s < ns).all() and (px >= 0).all() and (px < vw).all() and (py >= 0).all() and (py < vh).all()
        key = (s.astype(np.int64) * vh + py) * vw + px
        assert len(np.unique(key)) == len(key) == vw * vh * ns, f"{ns} passes: not a bijection"
        assert np.array_equal(back, np.arange(len(m))), f"{ns} passes: slotOfSample does not invert the map"
        if not (block_major and ns > 1):
            assert lps == 0
            continue
        assert ns % (1 << lps) == 0 and lps <= max_lps
        blk = (py // 4) * (vw // 8) + px // 8
        assert np.array_equal(blk, np.arange(len(m)) // (32 * ns)), "the slots of an 8x4 block must be consecutive"
        g = m.reshape(-1, 32, 4)
        assert ((g[:, :, 0].max(axis=1) - g[:, :, 0].min(axis=1)) == (1 << lps) - 1).all()          # 2^lps consecutive passes per group
        npix = np.array([len({(a, b) for a, b in zip(r[:, 1], r[:, 2])}) for r in g])
        assert (npix == 32 >> lps).all()                                                              # ... of 32 / 2^lps pixels
