"""The C++ drop-in: the reference's unmodified Scene/Loader/RadeonRays + its unmodified Renderer.h, with our Renderer_b200.cpp as the
implementation, driven by a headless main (oracle/_ref/ptb_headless, built where /root/reference exists).  Its accumulation buffer
must equal the Python host mirror's bit for bit: both feed the same arrays to the same C ABI."""
import os, subprocess
import numpy as np
import pytest
from conftest import ROOT, scene_at

pytestmark = pytest.mark.gpu
BIN = os.path.join(ROOT, "oracle", "_ref", "ptb_headless")
ASSETS = os.path.join(ROOT, "oracle", "_ref", "assets")


@pytest.mark.skipif(not (os.path.exists(BIN) and os.path.isdir(ASSETS)), reason="oracle/_ref/ptb_headless or assets not built (needs /root/reference at build time)")
@pytest.mark.parametrize("scene,w,h,spp,whole", [("cornell_box_orig", 128, 128, 3, False), ("hyperion_rect_lights", 256, 144, 2, True), ("volume_cube", 160, 90, 2, False),
                                                  # glTF + GLB through the reference's GLTFLoader inside the C++ process (inputs written by tests/golden/gen_gltf.py)
                                                  ("gltf_mix/gltf_mix", 160, 90, 2, True)])
def test_cpp_renderer_equals_python_mirror(scene, w, h, spp, whole, tmp_path):
    from glsl_pathtracer_b200 import capi
    if not os.path.exists(os.path.join(ASSETS, scene + ".scene")):
        pytest.skip("generated inputs not present (tests/golden/gen_gltf.py)")
    acc_file, png = str(tmp_path / "a.f32"), str(tmp_path / "o.png")
    cmd = [BIN, "-s", os.path.join(ASSETS, scene + ".scene"), "-o", png, "--spp", str(spp), "--res", str(w), str(h), "--accum", acc_file]
    if whole:
        cmd.append("--whole-frame")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert f"rendered {spp} spp" in out.stdout and os.path.getsize(png) > 1000
    a = np.fromfile(acc_file, np.float32).reshape(h, w, 4)
    sc = scene_at(os.path.basename(scene), w, h)     # same .scene through the committed blob (tile size / depth from the file)
    ctx = capi.Context(sc)
    ctx.render_samples(1, spp)
    b = ctx.read_accum()
    assert a.tobytes() == b.tobytes()
    ctx.close()
