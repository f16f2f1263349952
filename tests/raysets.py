"""Ray sets and hit-record comparisons shared by the G1 tests on the GPU (test_gpu_trace.py) and the host-compiled traversal check
(test_host_traversal.py)."""
import numpy as np


def random_rays(sc, n, seed):
    rng = np.random.default_rng(seed)
    lo, hi = np.array(sc.sceneBounds[0], np.float32), np.array(sc.sceneBounds[1], np.float32)
    ext = hi - lo
    o = (lo - 0.25 * ext + rng.random((n, 3), dtype=np.float32) * 1.5 * ext).astype(np.float32)
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    # a share of exactly axis-parallel rays: 0*inf NaNs in the slab test
    k = n // 16
    d[:k] = 0; d[np.arange(k), rng.integers(0, 3, k)] = rng.choice([-1.0, 1.0], k)
    return np.concatenate([o, d.astype(np.float32)], axis=1)


def bounce_rays(rays, hits, seed=11):
    """bounce-like rays: start on the surfaces found by `hits`, random directions, offset by EPS along the direction"""
    hit = hits["kind"] == 1
    p = rays[hit, :3] + rays[hit, 3:] * hits["t"][hit, None]
    rng = np.random.default_rng(seed)
    d = rng.normal(size=p.shape).astype(np.float32); d /= np.linalg.norm(d, axis=1, keepdims=True)
    return np.concatenate([(p + d * np.float32(0.0003)).astype(np.float32), d], axis=1)


def any_hit_distances(sc, n, seed=9):
    rng = np.random.default_rng(seed)
    ext = float(np.linalg.norm(np.array(sc.sceneBounds[1]) - np.array(sc.sceneBounds[0])))
    md = (rng.random(n, dtype=np.float32) * ext).astype(np.float32)
    md[::3] = np.float32(1e6 - 0.0003)
    return md


def assert_hits_nearly_equal(a, b, max_frac=2e-5):
    """Culled traversal: identical except for exact-tie edge cases (two triangles sharing an edge, t equal to a few ulp)."""
    bad = np.nonzero((a["primSlot"] != b["primSlot"]) | (a["kind"] != b["kind"]) | (a["lightIdx"] != b["lightIdx"]))[0]
    assert bad.size <= max(1, int(max_frac * len(a))), f"{bad.size} mismatches"
    assert np.all(np.abs(a["t"] - b["t"]) <= 1e-5 * np.abs(b["t"]))
    good = np.ones(len(a), bool); good[bad] = False
    assert np.array_equal(a["t"][good].view(np.uint32), b["t"][good].view(np.uint32))


def assert_hits_equal(a, b):
    for f in ("kind", "instance", "matID", "primSlot", "triIDx", "lightIdx"):
        bad = np.nonzero(a[f] != b[f])[0]
        assert bad.size == 0, f"{f}: {bad.size} mismatches, first ray {bad[:5]}: {a[f][bad[:5]]} vs {b[f][bad[:5]]}"
    ta, tb = a["t"], b["t"]
    assert np.all(np.abs(ta - tb) <= 1e-5 * np.abs(tb)), "t beyond 1e-5 relative"
    assert np.array_equal(ta.view(np.uint32), tb.view(np.uint32)), "t not bit-identical"
    hit = a["kind"] == 1
    assert np.array_equal(a["bary"][hit].view(np.uint32), b["bary"][hit].view(np.uint32)), "barycentrics not bit-identical"


def boundary_rays(sc, n, seed=13):
    """Adversarial rays for the slab test: origins exactly ON planes of node boxes (world space), a third of them exactly axis-parallel inside such a
    plane (0 * inf = NaN in AABBIntersect), a third with one denormal-small direction component (1/d overflows to inf), the rest generic."""
    rng = np.random.default_rng(seed)
    nodes = np.ascontiguousarray(sc.nodes, np.float32).reshape(-1, 9)
    pick = nodes[rng.integers(0, len(nodes), n)]
    lo, hi = pick[:, 0:3], pick[:, 3:6]
    o = (lo + rng.random((n, 3), dtype=np.float32) * (hi - lo)).astype(np.float32)
    ax = rng.integers(0, 3, n); side = rng.integers(0, 2, n)
    o[np.arange(n), ax] = np.where(side == 0, lo[np.arange(n), ax], hi[np.arange(n), ax])       # exactly on a box plane
    d = rng.normal(size=(n, 3)).astype(np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    k = n // 3
    d[np.arange(k), ax[:k]] = np.float32(0.0) * rng.choice([-1.0, 1.0], k).astype(np.float32)   # in-plane, +-0 component
    d[np.arange(k, 2 * k), ax[k:2 * k]] = (np.float32(1e-41) * rng.choice([-1.0, 1.0], k)).astype(np.float32)   # denormal component
    return np.concatenate([o, d.astype(np.float32)], axis=1)


def light_rays(sc, n, seed=17):
    """Rays aimed at the analytic lights: targets spread over (and a little beyond) the lights' bounds, and targets exactly on quad corners and edges
    (a1 / a2 of RectIntersect at 0 or 1), from origins inside the scene bounds.  Exercises the light loops incl. the shared-plane groups and their cell grid."""
    L = np.ascontiguousarray(sc.lights, np.float32).reshape(-1, 15)
    if len(L) == 0:
        return np.zeros((0, 6), np.float32)
    rng = np.random.default_rng(seed)
    lo, hi = np.array(sc.sceneBounds[0], np.float32), np.array(sc.sceneBounds[1], np.float32)
    o = (lo + rng.random((n, 3), dtype=np.float32) * (hi - lo)).astype(np.float32)
    k = rng.integers(0, len(L), n)
    pos, u, v = L[k, 0:3], L[k, 6:9], L[k, 9:12]
    a = rng.random((n, 2), dtype=np.float32) * np.float32(1.6) - np.float32(0.3)          # inside and just outside the quad
    third = n // 3
    a[:third] = rng.choice(np.array([0.0, 0.5, 1.0], np.float32), (third, 2))            # corners, edge midpoints, centre
    tgt = pos + u * a[:, :1] + v * a[:, 1:]
    sph = L[k, 14] == 1
    tgt[sph] = pos[sph] + rng.normal(size=(int(sph.sum()), 3)).astype(np.float32) * L[k[sph], 12:13]
    d = (tgt - o).astype(np.float32)
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-20)
    return np.concatenate([o, d.astype(np.float32)], axis=1)
