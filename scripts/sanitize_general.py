"""Small renders of every kernel specialisation (MODE 0 / 1 / 2, inline and deferred shadows, alpha, media, textures, instancing, preview,
tonemap, delta upload) for `compute-sanitizer --tool memcheck|racecheck python scripts/sanitize_general.py`."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
import glsl_pathtracer_b200  # noqa: F401
from glsl_pathtracer_b200 import capi
from conftest import scene_at
import feature_scenes as fs

cases = [(n, scene_at(n, 64, 36, 40, 20)) for n in ("cornell_box_orig", "hyperion_rect_lights", "ibl_spheres", "volume_cube", "instancing", "gltf_mix")]
cases += [("variant_" + v, fs.resized(fs.build(v), 64, 36, 40, 20)) for v in ("alpha_mask", "alpha_blend", "medium_scatter", "texture_maps_gl", "all_light_types")]
cases += [("variant_medium_deferred_transmittance", fs.resized(fs.media_no_blend(2), 64, 36, 40, 20))]      # k_shade<3> + k_transmit
for name, sc in cases:
    ctx = capi.Context(sc)
    ctx.render_samples(1, 2)
    ctx.render_tile(1, 1, 5)
    a = ctx.read_accum()
    ctx.render_preview(32, 18)
    ctx.read_output(0.5)
    ctx.update_instances(sc.transforms, sc.materials, sc.nodes[sc.topLevelIndex:])
    ctx.render_samples(3, 1)
    for s in (4, 5, 6):                       # coalesced-pass path: look-ahead wave + per-pass accumulate, frozen output
        ctx.render_pass(s, 3)
    ctx.snapshot_output(1.0 / 6); ctx.read_snapshot()
    rays = np.concatenate([np.tile(sc.camera.position, (256, 1)), np.random.default_rng(1).normal(size=(256, 3))], axis=1).astype(np.float32)
    rays[:16, 3:] = np.eye(3, dtype=np.float32)[np.arange(16) % 3]        # axis-parallel: the binary any-hit path next to the 4-wide one
    ctx.trace_closest(rays); ctx.trace_any(rays, 1e6)
    print(name, "ok", float(np.nan_to_num(a[..., :3]).mean()), ctx.stats()["kernelLaunches"], "launches")
    ctx.close()
# device-side TLAS rebuild (cooperative launch) on a scene with many instances and on a small one, then render
for name in ("instancing", "cornell_box_orig"):
    sc = scene_at(name, 48, 27, 24, 14)
    ctx = capi.Context(sc)
    T = np.ascontiguousarray(sc.transforms, np.float32).reshape(-1, 16).copy(); T[:, 12] += 0.25
    where = ctx.rebuild_instances(T, sc.materials)
    ctx.render_samples(1, 1)
    print("rebuild", name, "built at", where, float(np.nan_to_num(ctx.read_accum()[..., :3]).mean()))
    ctx.close()
