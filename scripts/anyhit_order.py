#!/usr/bin/env python
"""ANALYSIS TOOL (CPU only): work of the production any-hit path (the product's traverseWideAny compiled for the host, tests/host_harness) on oracle-captured
light-NEE shadow rays, for the slot orders of the 4-wide hierarchy (PTB_WIDE_ORDER: 0 binary order, 1 ascending stack need, 2 largest box first, 3 smallest box
first, ...).  Work = bytes fetched per ray (nodes, triangles, instance rows).  Any-hit is order-free, so every order returns the same booleans (asserted).

  python scripts/anyhit_order.py [scene] [orders ...]
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import scene_at
from oracle import binding as ob
from host_harness import binding as hb


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "hyperion_rect_lights"
    orders = sys.argv[2:] or ["0", "1", "2", "3"]
    sc = scene_at(name, 480, 272)
    orc = ob.Oracle(sc, cull=True)
    sets = {}
    for depth in (0, 1):
        R = []
        for s in (1, 2, 3, 4):
            rays8, valid = orc.capture_shadow_rays(s, depth)
            R.append(rays8.reshape(-1, 8)[valid.ravel()])
        sets[depth] = np.concatenate(R)
    ref = None
    for o in orders:
        os.environ["PTB_WIDE_ORDER"] = o
        ht = hb.HostTrav(sc); ht.set_cull(True)
        line = f"order {o}: stack bound {ht.any_stack()[1]:2d}"
        for depth, r in sets.items():
            occ = ht.trace_any(r[:, :6], r[:, 6], wide=True)
            line += f" | shading depth {depth}: {ht.any_bytes() / len(r):7.1f} B/ray ({occ.mean() * 100:.1f} % occluded)"
            if ref is None: ref = {}
            if depth in ref: assert np.array_equal(ref[depth], occ)
            else: ref[depth] = occ
        print(line)
        ht.close()


if __name__ == "__main__":
    main()
