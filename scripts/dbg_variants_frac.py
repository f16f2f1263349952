"""Debug aid: per feature variant, fraction of RNG-matched pixels CUDA vs oracle (native variant sizes)."""
import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import glsl_pathtracer_b200
from glsl_pathtracer_b200 import capi
from conftest import rel_mse
import feature_scenes as fs
from oracle import binding as ob
for name in fs.VARIANTS:
    for depth in (1, 2, None):
        sc = fs.build(name)
        if depth is not None: sc.renderOptions.maxDepth = depth
        ctx = capi.Context(sc); orc = ob.Oracle(sc)
        spp = 4
        ctx.render_samples(1, spp); g = np.nan_to_num(ctx.read_accum()); o = np.nan_to_num(orc.render(1, spp))
        d = np.abs(g[..., :3] - o[..., :3]).max(-1); rel = d / (np.abs(o[..., :3]).max(-1) + 1e-4)
        print(f"{name:42s} depth {str(depth):4s} relMSE {rel_mse(o / spp, g / spp):.3g} bad {(rel > 1e-3).mean():.4f}")
        ctx.close(); orc.close()
