#!/usr/bin/env python
"""Render a scene blob with the CUDA path and write the tonemapped image (what Renderer::GetOutputBuffer + SaveFrame produce)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import glsl_pathtracer_b200  # noqa
from glsl_pathtracer_b200 import capi
from conftest import scene_at
from PIL import Image
name, w, h, spp, out = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
sc = scene_at(name, w, h)
ctx = capi.Context(sc)
ctx.render_samples(1, spp)
img = ctx.read_output(1.0 / spp)          # bottom row first
Image.fromarray(img[::-1, :, :3]).save(out, quality=90)
print(name, ctx.stats())
