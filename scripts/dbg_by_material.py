"""Debug aid: CUDA vs oracle per first-hit material, for depth 0..D (which bounce introduces a difference)."""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import glsl_pathtracer_b200
from glsl_pathtracer_b200 import capi
from conftest import scene_at, rel_mse
from oracle import binding as ob
name, w, h, spp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
for depth in [int(a) for a in sys.argv[5:]] or [None]:
    sc = scene_at(name, w, h, 64, 36, depth)
    ctx = capi.Context(sc); orc = ob.Oracle(sc)
    ctx.render_samples(1, spp); g = ctx.read_accum(); o = orc.render(1, spp)
    rays = orc.camera_rays(1); hit = orc.trace_closest(rays, 0)
    mat = np.where(hit["kind"] == 1, hit["matID"], -hit["kind"] - 1).reshape(h, w)
    d = np.abs(g[..., :3] - o[..., :3]).max(-1); rel = d / (np.abs(o[..., :3]).max(-1) + 1e-4)
    print(f"depth {depth}: relMSE {rel_mse(o / spp, g / spp):.3g} frac rel>1e-3 {(rel > 1e-3).mean():.4f} alpha diff {np.abs(g[...,3]-o[...,3]).max():.3g}")
    for m in np.unique(mat):
        sel = mat == m
        print(f"   first-hit mat {m:3d}: px {sel.sum():6d} bad frac {(rel[sel] > 1e-3).mean():.4f} mean g {g[sel][:, :3].mean()/spp:.4f} o {o[sel][:, :3].mean()/spp:.4f}")
    sg, so = ctx.stats(), orc.stats()
    print("   segments", sg["pathSegments"], so["closestRays"], "shadow", sg.get("shadowRays"), so["anyRays"])
    os.makedirs('gpurun_out', exist_ok=True)
    np.save(f'gpurun_out/dbgm_{name}_{depth}_g.npy', g); np.save(f'gpurun_out/dbgm_{name}_{depth}_o.npy', o)
    ctx.close(); orc.close()
