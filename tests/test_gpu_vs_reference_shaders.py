"""G3 against the EXECUTING reference: the CUDA path (through the C ABI) vs the accumulation buffers, previews and RGBA8
readbacks the reference's own shader text produced (tests/golden/glslref_golden.npz; generator make_glslref_golden.py runs
oracle/glsl_ref where /root/reference exists).  8 scenes + 15 feature variants at 48x32, 20x12 tiles, 4 passes, RNG-matched:
relMSE <= 1e-3 and most pixels equal to float rounding.  Small images make single divergent paths visible, so the per-case
pixel fraction is looser than in test_gpu_render.py; the aggregate over all cases is tight."""
import os
import numpy as np
import pytest
from conftest import rel_mse
from test_glsl_ref import CASES, case_scene, W, H, SPP, GOLDEN

pytestmark = pytest.mark.gpu
_FRAC = {}
# hyperion_sphere_light: the REFERENCE algorithm lets the sphere light occlude its own NEE shadow ray whenever SphereIntersect's t
# lands below dist - EPS (anyhit.glsl:56-61 vs sampling.glsl:203-205); for grazing samples that is decided by rounding noise of the
# inputs (libm vs CUDA sin/cos), see tests/test_gpu_render.py.  Those pixels are equal in expectation, not per sample, so at
# 1536 pixels x 4 spp the relMSE of this one case is noise-limited; its converged gate (1e-3) is the 240x136 x 8 spp case of
# test_gpu_render.py against the oracle, which tests/test_glsl_ref.py shows bit-identical to the reference shaders.
RELMSE = {"hyperion_sphere_light": 1e-2}
MINFRAC = {"hyperion_sphere_light": 0.5}


@pytest.fixture(scope="module")
def golden():
    return np.load(GOLDEN)


@pytest.mark.parametrize("cull", [0, 1])
@pytest.mark.parametrize("case", CASES)
def test_cuda_matches_reference_shader_output(case, cull, golden):
    from glsl_pathtracer_b200 import capi
    sc = case_scene(case)
    ctx = capi.Context(sc)
    ctx.set_cull(bool(cull))
    ctx.render_samples(1, SPP)
    g = ctx.read_accum(); ref = golden[case + "/accum"]
    assert np.isfinite(g[..., :3]).all() == np.isfinite(ref[..., :3]).all()
    g = np.nan_to_num(g) / SPP; ref = np.nan_to_num(ref) / SPP
    r = rel_mse(ref, g)
    assert r <= RELMSE.get(case, 1e-3), f"relMSE {r}"
    close = np.isclose(g[..., :3], ref[..., :3], rtol=1e-3, atol=1e-4).all(axis=-1)
    _FRAC[(case, cull)] = close.mean()
    assert close.mean() >= MINFRAC.get(case, 0.85), f"only {close.mean():.4f} of pixels match"
    np.testing.assert_allclose(g[..., 3], ref[..., 3], atol=1e-6)
    if cull == 0:
        # tonemap kernel vs the reference's tonemap.glsl + RGBA8 readback on IDENTICAL input: load the reference's own sum
        ctx.write_accum(np.nan_to_num(golden[case + "/accum"]))
        out = ctx.read_output(1.0 / SPP).astype(np.int32); ref8 = golden[case + "/rgba8"].astype(np.int32)
        fin = np.isfinite(golden[case + "/accum"]).all(axis=-1)
        assert np.abs(out - ref8)[fin].max() <= 1, "tonemapped bytes differ by more than 1 LSB"
        assert (out == ref8)[fin].mean() > 0.99
    ctx.close()


def test_cuda_preview_matches_reference_preview_shader(golden):
    from glsl_pathtracer_b200 import capi
    for case in ("cornell_box_orig", "hyperion_rect_lights", "ibl_spheres", "volume_cube"):
        sc = case_scene(case)
        ctx = capi.Context(sc)
        p = ctx.render_preview(W // 2, H // 2); ref = golden[case + "/preview"]
        assert rel_mse(np.nan_to_num(ref), np.nan_to_num(p)) <= 1e-3
        close = np.isclose(p[..., :3], ref[..., :3], rtol=1e-3, atol=1e-4).all(axis=-1)
        assert close.mean() >= 0.9, (case, close.mean())
        ctx.close()


def test_aggregate_pixel_agreement():
    """Over all cases, >= 97 % of pixels are RNG-matched to float rounding (the rest are paths that a 1-ulp difference in the
    shading math sends the other way at a Russian-roulette / lobe-selection / refraction decision)."""
    if not _FRAC:
        pytest.skip("runs after the per-case tests")
    rest = [v for (c, _), v in _FRAC.items() if c not in MINFRAC]
    assert np.mean(rest) >= 0.97, _FRAC
