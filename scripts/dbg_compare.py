import sys, os, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import glsl_pathtracer_b200
from glsl_pathtracer_b200 import capi
from conftest import scene_at, rel_mse
from oracle import binding as ob
name, w, h, spp = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
depth = int(sys.argv[5]) if len(sys.argv) > 5 else None
sc = scene_at(name, w, h, 64, 36, depth)
ctx = capi.Context(sc); orc = ob.Oracle(sc)
ctx.render_samples(1, spp); g = ctx.read_accum(); o = orc.render(1, spp)
os.makedirs('gpurun_out', exist_ok=True)
np.save(f'gpurun_out/dbg_{name}_g.npy', g); np.save(f'gpurun_out/dbg_{name}_o.npy', o)
d = np.abs(g[..., :3] - o[..., :3]).max(-1); rel = d / (np.abs(o[..., :3]).max(-1) + 1e-4)
print('relMSE', rel_mse(o / spp, g / spp), 'frac rel>1e-3', (rel > 1e-3).mean(), 'frac rel>1e-2', (rel > 1e-2).mean(), 'frac rel>0.5', (rel > 0.5).mean())
print(ctx.stats(), orc.stats())
