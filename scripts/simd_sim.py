#!/usr/bin/env python
"""ANALYSIS TOOL: warp issue-slot model of the closest-hit traversal under different per-warp schedules (scripts/simd_sim/simd_sim.cpp)
on real ray streams captured from the oracle (orc_capture_rays), in the order the wavefront pipeline queues them (8x4 pixel blocks,
compacted).  CPU only.  Usage: python scripts/simd_sim.py [scene] [width height]"""
import ctypes as C, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import scene_at
from oracle import binding as ob


def build():
    src = os.path.join(ROOT, "scripts", "simd_sim", "simd_sim.cpp"); so = "/tmp/libsimd_sim.so"
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-fopenmp", "-fPIC", "-shared", src, "-o", so])
    L = C.CDLL(so)
    L.simd_sim.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return L


def inverse_rows(transforms):
    """rows of inverse(transform) per instance, fp32 adjugate like the shader (host: numpy float64 inverse rounded is NOT used)."""
    out = np.zeros_like(transforms)
    for k, m in enumerate(transforms.reshape(-1, 4, 4)):
        out.reshape(-1, 4, 4)[k] = np.linalg.inv(m.astype(np.float64)).astype(np.float32)
    return np.ascontiguousarray(out, np.float32)


def queue_order(valid):
    """Pixel order of the wavefront queue: 8x4 blocks row-major, lanes row-major inside a block; invalid entries compacted away."""
    h, w = valid.shape
    ys, xs = np.mgrid[0:h, 0:w]
    key = ((ys // 4) * ((w + 7) // 8) + (xs // 8)) * 32 + (ys % 4) * 8 + (xs % 8)
    order = np.argsort(key.ravel(), kind="stable")
    return order[valid.ravel()[order]]


def tile_sort(r, key, tile=2048):
    """Stable sort by key inside consecutive tiles of `tile` entries (what k_sort_tile_local does on the GPU)."""
    idx = np.arange(len(r))
    t = idx // tile
    order = np.lexsort((idx, key, t))
    return np.ascontiguousarray(r[order])


def reorder_keys(sc, r):
    lo, hi = np.array(sc.sceneBounds[0], np.float32), np.array(sc.sceneBounds[1], np.float32)
    d = r[:, 3:]
    octant = (d[:, 0] < 0).astype(np.int64) | ((d[:, 1] < 0).astype(np.int64) << 1) | ((d[:, 2] < 0).astype(np.int64) << 2)
    major = np.argmax(np.abs(d), axis=1) * 2 + (np.take_along_axis(d, np.argmax(np.abs(d), axis=1)[:, None], 1)[:, 0] < 0)
    cell = np.clip(((r[:, :3] - lo) / (hi - lo) * 8).astype(np.int64), 0, 7)
    def part(v):      # spread 3 bits
        return (v & 1) | ((v & 2) << 2) | ((v & 4) << 4)
    morton = part(cell[:, 0]) | (part(cell[:, 1]) << 1) | (part(cell[:, 2]) << 2)
    return {"direction octant": octant, "major axis (6)": major, "origin cell 8^3": morton, "octant | origin cell": octant * 512 + morton,
            "origin cell | octant": morton * 8 + octant}


def run(L, sc, arrays, r, policy, cull, warp=32, max_dist=None):
    nodes, vi, verts, invT = arrays
    r = np.ascontiguousarray(r, np.float32)
    t = np.zeros(len(r), np.float32); prim = np.zeros(len(r), np.int32); out = np.zeros(9)
    md = None if max_dist is None else np.ascontiguousarray(max_dist, np.float32)
    L.simd_sim(nodes.ctypes.data, len(nodes), sc.topLevelIndex, vi.ctypes.data, verts.ctypes.data, invT.ctypes.data, r.ctypes.data, len(r),
               policy, cull, warp, t.ctypes.data, prim.ctypes.data, out.ctypes.data, None if md is None else md.ctypes.data)
    return out


def reorder_study(L, sc, arrays, orc, depth=1):
    rays, valid = orc.capture_rays(1, depth)
    r = np.ascontiguousarray(rays.reshape(-1, 6)[queue_order(valid)], np.float32)
    b = run(L, sc, arrays, r, 0, 1)
    base = b[0] / len(r); basew = b[8] / len(r)
    print(f"  depth {depth}: queue order {base:.1f} slots/ray, {basew:.1f} L1 wavefronts/ray")
    for warp in (16, 8, 1):
        o = run(L, sc, arrays, r, 0, 1, warp)
        print(f"    (hypothetical {warp:2d}-lane warps: utilisation {o[1] / (warp * o[0]):.3f})")
    for name, key in reorder_keys(sc, r).items():
        for tile in (2048, 1 << 30):
            o = run(L, sc, arrays, tile_sort(r, key, tile), 0, 1)
            print(f"    sorted by {name:22s} in tiles of {'2048' if tile == 2048 else 'all '}: {o[0] / len(r):7.1f} slots/ray ({o[0] / len(r) / base * 100:5.1f} %), utilisation {o[1] / (32 * o[0]):.3f},"
                  f" L1 wavefronts/ray {o[8] / len(r):6.1f} ({o[8] / len(r) / basew * 100:5.1f} %)")


def shadow_study(L, sc, arrays, orc):
    """Light-NEE shadow rays (any-hit) of the first two shading iterations: today's queue (warp-compacted chunks appended in arrival order) vs a
    slot-ordered queue grouped inside 2048-slot tiles."""
    for depth in (0, 1):
        rays8, valid = orc.capture_shadow_rays(1, depth)
        order = queue_order(np.ones_like(valid))
        r_all = rays8.reshape(-1, 8)[order]; v_all = valid.ravel()[order]
        live = r_all[v_all]; n = len(live)
        sim = lambda q: run(L, sc, arrays, q[:, :6], 0, 1, 32, q[:, 6])
        chunks = [r_all[i:i + 32][v_all[i:i + 32]] for i in range(0, len(r_all), 32)]
        chunks = [c for c in chunks if len(c)]
        today = np.concatenate([chunks[i] for i in np.random.default_rng(0).permutation(len(chunks))])
        b = sim(today)[0] / n
        print(f"  shading depth {depth}: {n} shadow rays ({n / v_all.size * 100:.0f} % of the slots); today {b:.1f} slots/ray")
        keys = reorder_keys(sc, live); keys["light index"] = live[:, 7].astype(np.int64)
        starts = np.concatenate([[0], np.cumsum(v_all)])
        for kname in ("light index", "major axis (6)", "direction octant"):
            tot = 0.0
            for t0 in range(0, len(r_all), 2048):
                lt = r_all[t0:t0 + 2048][v_all[t0:t0 + 2048]]
                if len(lt):
                    kk = keys[kname][starts[t0]: starts[t0] + len(lt)]
                    tot += sim(lt[np.argsort(kk, kind="stable")])[0]
            print(f"    slot order, grouped by {kname:18s} in 2048-slot tiles: {tot / n:7.1f} ({tot / n / b * 100:5.1f} %)")
        q = tile_sort(today, today[:, 7].astype(np.int64), 2048)
        print(f"    arrival order grouped by light index (the rejected GPU experiment): {sim(q)[0] / n:7.1f} ({sim(q)[0] / n / b * 100:5.1f} %)")


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "hyperion_rect_lights"
    w, h = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (960, 540)
    sc = scene_at(name, w, h)
    orc = ob.Oracle(sc, cull=True)
    L = build()
    nodes = np.ascontiguousarray(sc.nodes, np.float32); vi = np.ascontiguousarray(sc.vertIndices, np.int32)
    verts = np.ascontiguousarray(sc.verticesUVX, np.float32); invT = inverse_rows(np.ascontiguousarray(sc.transforms, np.float32))
    print(f"{name} {w}x{h}: issue slots per ray (SASS-instruction weighted), lane utilisation, by path-loop depth and schedule")
    for depth in (0, 1, 2):
        rays, valid = orc.capture_rays(1, depth)
        if depth == 0: valid[:] = True
        idx = queue_order(valid)
        r = np.ascontiguousarray(rays.reshape(-1, 6)[idx], np.float32)
        ref = orc.trace_closest(r, 1)
        for cull in (1,):
            base = None
            for policy, pname in ((0, "while-while (shipped)"), (1, "postponed leaf"), (2, "if-if")):
                for warp in (32,):
                    t = np.zeros(len(r), np.float32); prim = np.zeros(len(r), np.int32); out = np.zeros(9)
                    L.simd_sim(nodes.ctypes.data, len(nodes), sc.topLevelIndex, vi.ctypes.data, verts.ctypes.data, invT.ctypes.data, r.ctypes.data, len(r),
                               policy, cull, warp, t.ctypes.data, prim.ctypes.data, out.ctypes.data, None)
                    tri = ref["kind"] == 1
                    # lights are not modelled: compare triangle hits only where the oracle's closest hit is a triangle
                    same = (prim[tri] == ref["primSlot"][tri]).mean() if tri.any() else 1.0
                    slots = out[0] / len(r)
                    base = base or slots
                    print(f"  depth {depth} {len(r):8d} rays  {pname:22s} slots/ray {slots:8.1f} ({slots / base * 100:5.1f} %)  utilisation {out[1] / (32 * out[0]):.3f}"
                          f"  inner {out[2] / len(r):7.1f} (util {out[3] / max(1, 32 * out[2]):.2f}, {out[7] / len(r):.1f} steps/ray)  tri {out[4] / len(r):6.1f} (util {out[5] / max(1, 32 * out[4]):.2f})"
                          f"  other {out[6] / len(r):5.1f}  L1 wavefronts/ray {out[8] / len(r):6.1f}  same prim as oracle {same:.5f}")
    print("ray reordering before the bounce trace (tile-local sort is ~0.05 ms per bounce on the GPU):")
    reorder_study(L, sc, (nodes, vi, verts, invT), orc, 1)
    print("light-NEE shadow rays (any-hit traversal):")
    shadow_study(L, sc, (nodes, vi, verts, invT), orc)
    orc.close()


if __name__ == "__main__":
    main()
