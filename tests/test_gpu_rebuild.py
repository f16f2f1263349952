"""N3: TLAS rebuild after an instance edit, from the transforms alone, by the library — on the device (`k_tlas_build`) with the exact sequential
host builder as twin and fallback.  The rebuilt canonical node array must be byte-identical to what the reference's own Scene::createTLAS + Bvh::Build +
BvhTranslator::ProcessTLAS produce: the fixtures hold reference-made arrays for the scenes as loaded and for an instance edit (tests/golden/instance_edit.npz)."""
import copy
import numpy as np
import pytest
from conftest import scene_at, edited_scene

pytestmark = pytest.mark.gpu
SCENES = ["cornell_box_orig", "cornell_box_sphere", "hyperion_rect_lights", "hyperion_sphere_light", "volume_cube", "teapot", "ibl_spheres", "instancing", "gltf_mix"]


def _ctx(sc):
    from glsl_pathtracer_b200 import capi
    return capi.Context(sc)


@pytest.mark.parametrize("name", SCENES)
def test_device_rebuild_reproduces_the_reference_tlas(name):
    sc = scene_at(name, 64, 64, 32, 32)
    want = np.ascontiguousarray(sc.nodes, np.float32)
    ctx = _ctx(sc)
    # scramble the TLAS slice on the device first, so that a rebuild that writes nothing cannot pass
    junk = want.reshape(-1, 9)[sc.topLevelIndex:].copy(); junk[:, :6] += 1.0
    ctx.update_instances(sc.transforms, sc.materials, junk)
    assert ctx.read_nodes().tobytes() != want.tobytes()
    for on_host in (False, True):
        where = ctx.rebuild_instances(sc.transforms, sc.materials, on_host=on_host)
        assert where == (1 if on_host else 0), f"built at {where}"
        assert ctx.read_nodes().tobytes() == want.tobytes(), f"on_host={on_host}"
    ctx.close()


@pytest.mark.parametrize("name", ["cornell_box_orig", "hyperion_rect_lights"])
def test_device_rebuild_after_the_references_own_instance_edit(name, oracle_mod):
    """Transforms (and one material id) edited as the application edits them; the reference's RebuildInstances made the golden TLAS.  Device rebuild: node array
    byte-identical, render bitwise equal to a context created from the edited scene."""
    sc, sc2 = edited_scene(name, 160, 90, 80, 45)
    want = np.ascontiguousarray(sc2.nodes, np.float32).reshape(-1, 9)
    leaves = want[sc.topLevelIndex:][want[sc.topLevelIndex:, 8] < 0]
    mats = np.zeros(len(leaves), np.int32); mats[(-leaves[:, 8] - 1).astype(int)] = leaves[:, 7].astype(np.int32)
    ctx = _ctx(sc)
    ctx.render_samples(1, 1)
    assert ctx.rebuild_instances(sc2.transforms, sc2.materials, material_ids=mats) == 0
    assert ctx.read_nodes().tobytes() == want.tobytes()
    ctx.reset_accum(); ctx.render_samples(1, 3)
    fresh = _ctx(sc2); fresh.render_samples(1, 3)
    assert ctx.read_accum().tobytes() == fresh.read_accum().tobytes()
    ctx.close(); fresh.close()


def _jitter(sc, seed):
    rng = np.random.default_rng(seed)
    T = np.ascontiguousarray(sc.transforms, np.float32).reshape(-1, 4, 4).copy()
    n = len(T)
    ang = rng.uniform(0, 2 * np.pi, n).astype(np.float32); s = rng.uniform(0.5, 1.5, n).astype(np.float32)
    R = np.zeros((n, 4, 4), np.float32); R[:, 3, 3] = 1; R[:, 1, 1] = s
    R[:, 0, 0] = np.cos(ang) * s; R[:, 0, 2] = -np.sin(ang) * s; R[:, 2, 0] = np.sin(ang) * s; R[:, 2, 2] = np.cos(ang) * s
    T = np.einsum("nij,njk->nik", R, T).astype(np.float32)
    T[:, 3, :3] += rng.normal(0, 0.5, (n, 3)).astype(np.float32)
    return T.reshape(sc.transforms.shape)


def test_device_rebuild_equals_the_sequential_builder_on_10k_moved_instances():
    """10 001 instances re-posed (rotation, scale, translation): device == exact sequential host builder byte for byte; traversal after the rebuild matches a
    context created from the rebuilt arrays.  Prints the rebuild times (device path incl. readback + host-side derivation of the packed layouts)."""
    sc = scene_at("instancing", 96, 54, 48, 27)
    ctx = _ctx(sc); ref = _ctx(sc)
    T = _jitter(sc, 5)
    ctx.rebuild_instances(T, sc.materials); ref.rebuild_instances(T, sc.materials, on_host=True)      # first calls allocate the scratch buffers
    w_dev = ctx.rebuild_instances(T, sc.materials); dev = dict(ctx.last_rebuild)
    w_host = ref.rebuild_instances(T, sc.materials, on_host=True); host = dict(ref.last_rebuild)
    assert (w_dev, w_host) == (0, 1)
    a, b = ctx.read_nodes(), ref.read_nodes()
    assert a.tobytes() == b.tobytes()
    assert a.tobytes() != np.ascontiguousarray(sc.nodes, np.float32).tobytes()
    print(f"rebuild of {len(T.reshape(-1, 16))} instances: TLAS build {dev['build_ms']:.2f} ms on the device (cooperative grid, one CTA per SM) vs {host['build_ms']:.2f} ms sequential on the host; "
          f"whole call incl. derivation + upload of the packed layouts {dev['total_ms']:.1f} / {host['total_ms']:.1f} ms")
    sc2 = copy.deepcopy(sc); sc2.nodes = a.reshape(np.asarray(sc.nodes).shape); sc2.transforms = T
    fresh = _ctx(sc2)
    ctx.render_samples(1, 1); fresh.render_samples(1, 1)
    assert ctx.read_accum().tobytes() == fresh.read_accum().tobytes()
    for c in (ctx, ref, fresh): c.close()


def test_degenerate_input_falls_back_to_the_sequential_builder():
    """Two instances with the same centroid: the reference then splits by POSITION in its partially partitioned index array; the device builder detects the
    case and hands over to the sequential builder (where == 2) — same bytes as asking for the host build."""
    sc = scene_at("cornell_box_orig", 64, 64, 32, 32)
    T = np.ascontiguousarray(sc.transforms, np.float32).reshape(-1, 16).copy()
    # instances 0 and 1 use different meshes: give both a transform that collapses them onto one point (zero scale): equal centroids, zero extent
    for k in (0, 1):
        T[k] = 0; T[k, 15] = 1; T[k, 12:15] = [0.1, 0.2, 0.3]
    a = _ctx(sc); b = _ctx(sc)
    assert a.rebuild_instances(T, sc.materials) == 2
    assert b.rebuild_instances(T, sc.materials, on_host=True) == 1
    assert a.read_nodes().tobytes() == b.read_nodes().tobytes()
    a.close(); b.close()


def test_rebuild_rejects_bad_arguments():
    from glsl_pathtracer_b200 import capi
    sc = scene_at("cornell_box_orig", 32, 32, 16, 16)
    ctx = _ctx(sc)
    with pytest.raises(capi.PtbError):
        ctx.rebuild_instances(np.asarray(sc.transforms, np.float32).reshape(-1, 16)[:-1], sc.materials)
    with pytest.raises(capi.PtbError):
        ctx.rebuild_instances(sc.transforms, sc.materials, material_ids=np.full(len(np.asarray(sc.transforms).reshape(-1, 16)), 999, np.int32))
    ctx.close()


def test_rebuild_of_100k_instances():
    """Where a device-side rebuild starts to matter (the reference's CPU rebuild is 37 ms at 10^5 instances): a synthetic 100 000-instance scene
    (the instancing fixture's mesh on a jittered grid; its TLAS made by the exact sequential builder).  Device == sequential builder byte for byte; times printed."""
    from host_harness.binding import build_tlas
    sc = copy.deepcopy(scene_at("instancing", 64, 36, 32, 18))
    nodes = np.ascontiguousarray(sc.nodes, np.float32).reshape(-1, 9)
    top = sc.topLevelIndex
    leaf0 = nodes[top:][nodes[top:, 8] < 0][0]                       # one TLAS leaf: (BLAS root, material id)
    n = 100_000
    rng = np.random.default_rng(3)
    T = np.tile(np.eye(4, dtype=np.float32), (n, 1, 1))
    g = int(np.ceil(np.sqrt(n)))
    T[:, 3, 0] = (np.arange(n) % g) * 3.0 + rng.normal(0, 0.4, n); T[:, 3, 2] = (np.arange(n) // g) * 3.0 + rng.normal(0, 0.4, n); T[:, 3, 1] = rng.uniform(0, 2, n)
    T = T.astype(np.float32).reshape(n, 16)
    seed = np.zeros((2 * n, 9), np.float32)                          # a placeholder TLAS that only tells the builder each instance's BLAS root / material
    seed[:n, 6] = leaf0[6]; seed[:n, 7] = leaf0[7]; seed[:n, 8] = -(np.arange(n) + 1)
    blas = nodes[:top]
    tlas, _ = build_tlas(np.concatenate([blas, seed]), top, T)
    sc.nodes = np.concatenate([blas, tlas]).astype(np.float32); sc.transforms = T
    sc.instances = np.tile(np.asarray(sc.instances)[:1], (n, 1))
    ctx = _ctx(sc)
    assert ctx.read_nodes().tobytes() == sc.nodes.tobytes()
    T2 = T.copy(); T2[:, 12:15] += rng.normal(0, 0.7, (n, 3)).astype(np.float32)
    ctx.rebuild_instances(T2, sc.materials)
    assert ctx.rebuild_instances(T2, sc.materials) == 0; dev = dict(ctx.last_rebuild); a = ctx.read_nodes()
    assert ctx.rebuild_instances(T2, sc.materials, on_host=True) == 1; host = dict(ctx.last_rebuild)
    assert a.tobytes() == ctx.read_nodes().tobytes()
    want, _ = build_tlas(sc.nodes, top, T2)
    assert a.reshape(-1, 9)[top:].tobytes() == want.tobytes()
    print(f"rebuild of {n} instances: TLAS build {dev['build_ms']:.2f} ms on the device vs {host['build_ms']:.2f} ms sequential on the host; whole call {dev['total_ms']:.1f} / {host['total_ms']:.1f} ms")
    ctx.render_samples(1, 1)
    assert np.isfinite(np.nan_to_num(ctx.read_accum())).all()
    ctx.close()
