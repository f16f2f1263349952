#!/usr/bin/env python
"""Golden vectors from the EXECUTING REFERENCE SHADERS (oracle/glsl_ref: the reference's own tile.glsl / preview.glsl /
tonemap.glsl text compiled by g++, read in place from /root/reference).  Needs /root/reference; the outputs are committed so that
the pin also holds where the reference tree is absent (the GPU box).

Per case: 4 full-frame sample passes at 48x32 with 20x12 tiles (over-hanging last column / top row, so the tile-local RNG seeding
and the tile-offset uniforms are exercised), the 24x16 preview, and the tonemapped RGBA8 readback of the 4-pass sum.
Checked by tests/test_glsl_ref.py (oracle == golden, bit-exact) and tests/test_gpu_render.py (CUDA path vs golden)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import scene_at
import feature_scenes as fs
from oracle.glsl_ref import binding as gb

W, H, TW, TH, SPP = 48, 32, 20, 12, 4
SCENES = ("cornell_box_orig", "cornell_box_sphere", "hyperion_rect_lights", "hyperion_sphere_light", "volume_cube", "ibl_spheres",
          "teapot", "instancing", "gltf_mix")


def cases():
    for name in SCENES:
        yield name, scene_at(name, W, H, TW, TH)
    for name in fs.VARIANTS:
        yield "variant_" + name, fs.resized(fs.build(name), W, H, TW, TH)


def main():
    out = {}
    for name, sc in cases():
        g = gb.GlslRef(sc)
        accum = g.render(1, SPP)
        out[name + "/accum"] = accum
        out[name + "/preview"] = g.render_preview(W // 2, H // 2)
        out[name + "/rgba8"] = g.tonemap(accum, 1.0 / SPP, sc.renderOptions)
        print(name, "mean", float(np.nanmean(accum[..., :3])) / SPP)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "glslref_golden.npz"), **out)


if __name__ == "__main__":
    main()
