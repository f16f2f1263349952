#!/usr/bin/env python
"""ANALYSIS TOOL (CPU only): which ORDER of the path slots makes the tile-local grouping pay?  Feeds 32 sample passes of oracle-captured rays through the
warp issue-slot model of scripts/simd_sim (closest-hit and any-hit) under different slot layouts and sort keys.  This is the study behind the block-major
slot order, the octahedral direction classes of bounce 1 and the light-sorted NEE queue of DESIGN.md 3.3 — model and GPU agreed this time (bounce-1 trace:
model -28 %, GPU 6.14 -> 4.40 ms; first-bounce shadow rays: model -25 %, GPU 6.18 -> 4.87 ms).

  python scripts/simd_sim_order.py [scene] [study]      study: bounce1 (default) | shadow0 | primary | bounce2
"""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts")); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import simd_sim as ss
from conftest import scene_at
from oracle import binding as ob

NS = 32


def setup(name, w, h):
    sc = scene_at(name, w, h)
    orc = ob.Oracle(sc, cull=True)
    L = ss.build()
    arrays = (np.ascontiguousarray(sc.nodes, np.float32), np.ascontiguousarray(sc.vertIndices, np.int32), np.ascontiguousarray(sc.verticesUVX, np.float32),
              ss.inverse_rows(np.ascontiguousarray(sc.transforms, np.float32)))
    return sc, orc, L, arrays


def make_sim(L, sc, arrays):
    nodes, vi, verts, invT = arrays
    def sim(q):
        r = np.ascontiguousarray(q[:, :6], np.float32)
        t = np.zeros(len(r), np.float32); prim = np.zeros(len(r), np.int32); out = np.zeros(9)
        md = np.ascontiguousarray(q[:, 6], np.float32) if q.shape[1] > 6 else None
        L.simd_sim(nodes.ctypes.data, len(nodes), sc.topLevelIndex, vi.ctypes.data, verts.ctypes.data, invT.ctypes.data, r.ctypes.data, len(r),
                   0, 1, 32, t.ctypes.data, prim.ctypes.data, out.ctypes.data, None if md is None else md.ctypes.data)
        return out
    def sim_tiles(tiles):
        """every tile starts on a warp boundary (NaN rays = idle lanes pad the last warp)"""
        big = []
        for rt in tiles:
            big.append(rt)
            pad = (-len(rt)) % 32
            if pad: big.append(np.full((pad, rt.shape[1]), np.nan, np.float32))
        return sim(np.concatenate(big))
    return sim, sim_tiles


def dirclass6(d):
    a = np.abs(d); m = np.argmax(a, axis=1)
    return m * 2 + (np.take_along_axis(d, m[:, None], 1)[:, 0] < 0)


def oct_cell(d, g=8):
    """cell of a g x g octahedral map, rows walked boustrophedon (octCell in ptb_kernels.cu)"""
    a = np.abs(d).sum(axis=1, keepdims=True); p = d / a
    u = np.where(p[:, 2] >= 0, p[:, 0], (1 - np.abs(p[:, 1])) * np.sign(p[:, 0])); v = np.where(p[:, 2] >= 0, p[:, 1], (1 - np.abs(p[:, 0])) * np.sign(p[:, 1]))
    iu = np.clip(((u * 0.5 + 0.5) * g).astype(int), 0, g - 1); iv = np.clip(((v * 0.5 + 0.5) * g).astype(int), 0, g - 1)
    return iv * g + np.where(iv % 2 == 0, iu, g - 1 - iu)


def block_order(w, h):
    ys, xs = np.mgrid[0:h, 0:w]
    return np.argsort((((ys // 4) * (w // 8) + (xs // 8)) * 32 + (ys % 4) * 8 + (xs % 8)).ravel())


def region_tiles(R, V, bw, bh, keyf):
    tiles = []
    h, w = R.shape[1:3]
    for by in range(0, h, bh):
        for bx in range(0, w, bw):
            r = R[:, by:by + bh, bx:bx + bw].reshape(-1, R.shape[-1]); v = V[:, by:by + bh, bx:bx + bw].ravel()
            rt = r[v]
            if len(rt): tiles.append(rt[np.argsort(keyf(rt), kind="stable")])
    return tiles


def bounce_study(name, depth):
    w, h = 480, 272
    sc, orc, L, arrays = setup(name, w, h); sim, sim_tiles = make_sim(L, sc, arrays)
    R = []; V = []
    for s in range(1, NS + 1):
        rays, valid = orc.capture_rays(s, depth); R.append(rays.reshape(h, w, 6)); V.append(valid.reshape(h, w))
    R = np.stack(R); V = np.stack(V); n = int(V.sum())
    print(f"{name} {w}x{h}, {NS} passes, closest-hit rays of path-loop depth {depth}: {n} rays ({V.mean() * 100:.1f} % of the slots alive)")
    order = block_order(w, h)
    tiles = []
    for s in range(NS):       # round-2 layout: sample-major, 2048-slot tiles = strips of 64 8x4 blocks of one pass, 6 direction classes
        r = R[s].reshape(-1, 6)[order]; v = V[s].ravel()[order]
        for t in range(0, len(r), 2048):
            rt = r[t:t + 2048][v[t:t + 2048]]
            if len(rt): tiles.append(rt[np.argsort(dirclass6(rt[:, 3:]), kind="stable")])
    base = sim_tiles(tiles)
    print(f"  sample-major tiles, 6 direction classes (round 2): {base[0] / n:.1f} slots/ray, utilisation {base[1] / (32 * base[0]):.3f}")
    for (bw, bh) in ((8, 8), (8, 4), (16, 8)):
        for kname, kf in (("6 classes", lambda q: dirclass6(q[:, 3:6])), ("oct 8x8", lambda q: oct_cell(q[:, 3:6], 8)), ("oct 16x16", lambda q: oct_cell(q[:, 3:6], 16))):
            o = sim_tiles(region_tiles(R, V, bw, bh, kf))
            print(f"  block-major tile {bw}x{bh} pixels x {NS} passes = {bw * bh * NS} slots, sorted by {kname:10s}: {o[0] / n:.1f} ({o[0] / base[0] * 100:.1f} %), utilisation {o[1] / (32 * o[0]):.3f}")


def shadow_study(name, depth=0):
    w, h = 480, 272
    sc, orc, L, arrays = setup(name, w, h); sim, sim_tiles = make_sim(L, sc, arrays)
    R = []; V = []
    for s in range(1, NS + 1):
        rays8, valid = orc.capture_shadow_rays(s, depth); R.append(rays8.reshape(h, w, 8)); V.append(valid.reshape(h, w))
    R = np.stack(R); V = np.stack(V); n = int(V.sum())
    order = block_order(w, h)
    chunks = []
    for s in range(NS):
        r = R[s].reshape(-1, 8)[order]; v = V[s].ravel()[order]
        for t in range(0, len(r), 32):
            c = r[t:t + 32][v[t:t + 32]]
            if len(c): chunks.append(c)
    perm = np.random.default_rng(0).permutation(len(chunks))
    base = sim(np.concatenate([chunks[i] for i in perm]))
    print(f"{name}: light-NEE shadow rays of shading depth {depth}: {n}; compacted arrival-order chunks (round 2): {base[0] / n:.1f} slots/ray, utilisation {base[1] / (32 * base[0]):.3f}")
    for kname, kf in (("none", lambda q: np.zeros(len(q), int)), ("light", lambda q: q[:, 7].astype(int)), ("oct 8x8", lambda q: oct_cell(q[:, 3:6], 8)),
                      ("light, oct 8x8", lambda q: q[:, 7].astype(int) * 64 + oct_cell(q[:, 3:6], 8))):
        o = sim_tiles(region_tiles(R, V, 8, 8, kf))
        print(f"  block-major tile 8x8 pixels x {NS} passes, sorted by {kname:14s}: {o[0] / n:.1f} ({o[0] / base[0] * 100:.1f} %), utilisation {o[1] / (32 * o[0]):.3f}")


def primary_study(name):
    w, h = 960, 544
    sc, orc, L, arrays = setup(name, w, h); sim, _ = make_sim(L, sc, arrays)
    R = np.stack([orc.capture_rays(s, 0)[0].reshape(h, w, 6) for s in range(1, NS + 1)])[:, 100:420, 200:760]
    S, H, W_, _ = R.shape; n = S * H * W_; base = None
    for (pw, ph, ps) in ((8, 4, 1), (4, 4, 2), (4, 2, 4), (2, 2, 8), (2, 1, 16), (1, 1, 32)):
        a = R.reshape(S // ps, ps, H // ph, ph, W_ // pw, pw, 6).transpose(0, 2, 4, 1, 3, 5, 6).reshape(-1, 6)
        o = sim(a); base = base or o[0]
        print(f"  warp = {pw}x{ph} pixels x {ps} passes: {o[0] / n:.1f} slots/ray ({o[0] / base * 100:.1f} %), utilisation {o[1] / (32 * o[0]):.3f}, L1 wavefronts/ray {o[8] / n:.1f}")


if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "hyperion_rect_lights"
    study = sys.argv[2] if len(sys.argv) > 2 else "bounce1"
    {"bounce1": lambda: bounce_study(name, 1), "bounce2": lambda: bounce_study(name, 2), "shadow0": lambda: shadow_study(name, 0), "primary": lambda: primary_study(name)}[study]()
