#!/usr/bin/env python
"""Instance-edit fixtures for the live-update path: what the UNMODIFIED reference host code produces when an
instance is moved the way the application moves it (Main.cpp:495-510 -> Scene::RebuildInstances, Scene.cpp:200-214) — the new
transforms, materials and TLAS node slice that Renderer::Update re-uploads (Renderer.cpp:649-665).  Built by
oracle/_ref/instance_edit_dump (oracle/ref_host/instance_edit_dump.cpp); runs only where /root/reference exists.
-> tests/golden/instance_edit.npz, checked by tests/test_fixtures.py, tests/test_glsl_ref.py and tests/test_gpu_render.py."""
import os, subprocess, sys, tempfile
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("PTB_REFERENCE", "/root/reference")
# scene -> (instance, tx, ty, tz, scale, new material id or None)
EDITS = {"cornell_box_orig": (6, 0.05, 0.02, -0.03, 0.8, 2), "hyperion_rect_lights": (3, 4.0, 1.5, -3.0, 1.3, None)}


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle", "ref_host")], stdout=subprocess.DEVNULL)
    tool = os.path.join(ROOT, "oracle", "_ref", "instance_edit_dump")
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, (k, tx, ty, tz, sc, mat) in EDITS.items():
            f = os.path.join(tmp, name + ".bin")
            cmd = [tool, os.path.join(REF, "assets", name + ".scene"), f, str(k), str(tx), str(ty), str(tz), str(sc)] + ([str(mat)] if mat is not None else [])
            print([l for l in subprocess.check_output(cmd, text=True).splitlines() if l.startswith("INSTEDIT")][0])
            raw = open(f, "rb").read()
            ni, nm, nt, top = np.frombuffer(raw[:16], np.int32)
            o = 16
            out[name + "/transforms"] = np.frombuffer(raw, np.float32, ni * 16, o).reshape(ni, 16).copy(); o += ni * 64
            out[name + "/materials"] = np.frombuffer(raw, np.float32, nm * 32, o).reshape(nm, 32).copy(); o += nm * 128
            out[name + "/tlas"] = np.frombuffer(raw, np.float32, nt * 9, o).reshape(nt, 9).copy()
            out[name + "/top"] = np.int32(top)
            out[name + "/edit"] = np.array([k, tx, ty, tz, sc, -1 if mat is None else mat], np.float32)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "instance_edit.npz"), **out)


if __name__ == "__main__":
    main()
