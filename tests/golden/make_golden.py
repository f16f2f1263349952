#!/usr/bin/env python
"""Writes the golden oracle renders checked by tests/test_oracle.py::test_render_matches_committed_golden."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from conftest import scene_at
from oracle import binding as ob
for name, depth in (("cornell_box_orig", 3), ("hyperion_rect_lights", None), ("volume_cube", None)):
    o = ob.Oracle(scene_at(name, 48, 32, 24, 16, depth))
    np.save(os.path.join(ROOT, "tests", "golden", f"oracle_{name}_48x32_4spp.npy"), o.render(1, 4))
