"""Several GPUs and the coalesced drop-in path, on hardware, through the C ABI.

* `ptb_render_pass` (what the drop-in's Render() calls at the first tile of a pass): bitwise equal to `ptb_render_samples`, also when
  the speculative wave has to be thrown away because a uniform changed between two passes.
* `ptb_mgpu_*` (N contexts + NCCL in one process): 1 device degenerates to a plain context; 2 devices — skipped on a 1-GPU box —
  give the single-GPU image up to fp32 add order, also on a SECOND readback after more passes (the reduce goes into a scratch sum).
* one process per GPU under torchrun with torch.distributed/NCCL (bench.py's arrangement): same check, twice.
"""
import os, subprocess, sys, textwrap
import numpy as np
import pytest
from conftest import ROOT, scene_at

pytestmark = pytest.mark.gpu


def n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def test_render_pass_equals_render_samples_bitwise():
    from glsl_pathtracer_b200 import capi
    sc = scene_at("cornell_box_orig", 160, 96, 64, 48, 3)
    a = capi.Context(sc, samples_per_wave=4); b = capi.Context(sc, samples_per_wave=4)
    for s in range(1, 8):                    # waves of 4: passes 1-4 from the first wave, 5-7 from the second (look-ahead capped at 3)
        a.render_pass(s, max_lookahead=8 - s)
        b.reset_accum(); b.render_samples(1, s)
        assert a.read_accum().tobytes() == b.read_accum().tobytes(), s
    assert a.stats()["samplesRendered"] == 7
    c = capi.Context(sc); c.render_samples(1, 7)
    assert a.stats()["pathSegments"] == c.stats()["pathSegments"]      # every pass was traced exactly once (2 waves: 4 + 3 passes)
    c.close()
    a.close(); b.close()


def test_render_pass_discards_its_wave_when_a_uniform_changes():
    import copy
    from glsl_pathtracer_b200 import capi
    sc = scene_at("cornell_box_orig", 128, 96, 64, 48, 3)
    a = capi.Context(sc, samples_per_wave=4); b = capi.Context(sc, samples_per_wave=4)
    a.render_pass(1); b.render_samples(1, 1)
    cam = copy.deepcopy(sc.camera); cam.position = tuple(np.asarray(cam.position, np.float32) + np.float32([0.01, 0, 0]))
    a.set_camera(cam); b.set_camera(cam)    # the application would also mark the scene dirty; the library must not rely on that
    a.render_pass(2); b.render_samples(2, 1)
    assert a.read_accum().tobytes() == b.read_accum().tobytes()
    o = capi.make_options(sc, samples_per_wave=4); o.maxDepth = 2
    a.set_options(o); b.set_options(o)
    a.render_pass(3); b.render_samples(3, 1)
    assert a.read_accum().tobytes() == b.read_accum().tobytes()
    a.render_pass(5); b.render_samples(5, 1)          # out of order: not the next pass of the resident wave
    assert a.read_accum().tobytes() == b.read_accum().tobytes()
    a.close(); b.close()


def test_snapshot_is_frozen_and_read_lazily():
    from glsl_pathtracer_b200 import capi
    sc = scene_at("cornell_box_orig", 96, 64, 48, 32, 3)
    c = capi.Context(sc)
    assert not c.read_snapshot().any()                     # nothing completed yet: the cleared texture
    c.render_samples(1, 2)
    c.snapshot_output(0.5)
    want = c.read_output(0.5)
    c.render_samples(3, 2)                                 # rendering goes on; the frozen image must not move
    assert np.array_equal(c.read_snapshot(), want) and not np.array_equal(c.read_output(0.25), want)
    pinned = c.read_output(0.25, pinned=True)
    assert np.array_equal(pinned, c.read_output(0.25))
    c.close()


def test_renderer_mirror_coalesced_equals_tile_walk(oracle_mod):
    from glsl_pathtracer_b200.renderer import Renderer
    imgs = []
    for coalesce in (True, False):
        sc = scene_at("cornell_box_orig", 100, 72, 48, 32)          # 3 x 3 tiles, over-hanging last column and top row
        sc.renderOptions.maxSpp = 4
        r = Renderer(sc, "shaders/", coalesce=coalesce)
        while r.GetSampleCount() < 4:
            r.Update(0.016); r.Render(); r.Present()
        imgs.append((r.ctx.read_accum(), r.GetOutputBuffer()[0], r.ctx.stats()["kernelLaunches"]))
    assert imgs[0][0].tobytes() == imgs[1][0].tobytes()             # same seeds, same frame numbers, same sum
    assert np.array_equal(imgs[0][1], imgs[1][1])                   # the completed (tonemapped) image: pass 3 with uniform 1/3
    assert imgs[0][2] < imgs[1][2] / 3                              # 1 wave (+2 accumulates) instead of 27 tile waves


def test_mgpu_one_device_is_a_plain_context():
    from glsl_pathtracer_b200 import capi
    sc = scene_at("cornell_box_orig", 128, 96, 64, 48, 3)
    m = capi.Mgpu(sc, devices=[0]); c = capi.Context(sc)
    m.render_samples(1, 5); c.render_samples(1, 5)
    assert m.read_accum().tobytes() == c.read_accum().tobytes()
    assert np.array_equal(m.read_output(0.2), c.read_output(0.2))
    for s in (6, 7):
        m.render_pass(s); c.render_samples(s, 1)
    m.snapshot_output(1 / 7)
    assert np.array_equal(m.read_snapshot(), c.read_output(1 / 7))
    assert m.stats()["samplesRendered"] == 7
    m.close(); c.close()


@pytest.mark.skipif(n_gpus() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("scene", ["cornell_box_orig", "hyperion_rect_lights"])
def test_mgpu_two_devices_equal_one_gpu_and_readback_is_repeatable(scene):
    from glsl_pathtracer_b200 import capi
    sc = scene_at(scene, 256, 144, 128, 72)
    m = capi.Mgpu(sc, devices=[0, 1]); c = capi.Context(sc, device=0)
    m.render_samples(1, 9); c.render_samples(1, 9)                 # odd split: GPU 0 gets 5 passes, GPU 1 gets 4
    a1 = m.read_accum(); b1 = c.read_accum()
    np.testing.assert_allclose(a1, b1, rtol=2e-5, atol=1e-5)
    assert np.array_equal(m.read_accum(), a1)                      # a second readback without new passes: identical (scratch, not in place)
    m.render_samples(10, 7); c.render_samples(10, 7)
    np.testing.assert_allclose(m.read_accum(), c.read_accum(), rtol=2e-5, atol=1e-5)
    d = np.abs(m.read_output(1 / 16).astype(int) - c.read_output(1 / 16).astype(int))
    assert d.max() <= 1
    # each context still holds only its own passes
    own = capi.Context(sc, device=0); own.render_samples(1, 5, 2); own.render_samples(10, 4, 2)
    assert m.contexts[0].read_accum().tobytes() == own.read_accum().tobytes()
    # the drop-in's pass-by-pass path over both GPUs
    m.reset_accum(); c.reset_accum()
    for s in range(1, 7):
        m.render_pass(s, 3)
    c.render_samples(1, 6)
    m.snapshot_output(1 / 6)
    np.testing.assert_allclose(m.read_accum(), c.read_accum(), rtol=2e-5, atol=1e-5)
    assert np.abs(m.read_snapshot().astype(int) - c.read_output(1 / 6).astype(int)).max() <= 1
    m.close(); c.close(); own.close()


WORKER = textwrap.dedent("""
    import os, sys
    sys.path.insert(0, {root!r}); sys.path.insert(0, os.path.join({root!r}, 'tests'))
    import numpy as np, torch, torch.distributed as dist
    from conftest import scene_at
    from glsl_pathtracer_b200 import capi, multigpu
    rank, world, local = int(os.environ['RANK']), int(os.environ['WORLD_SIZE']), int(os.environ['LOCAL_RANK'])
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    sc = scene_at('hyperion_rect_lights', 480, 270, 256, 144)
    ctx = capi.Context(sc, device=local)
    stream = torch.cuda.Stream(device=local); ctx.set_stream(stream.cuda_stream)
    t = multigpu.DeviceAccumView(ctx).tensor(torch.device('cuda', local))
    drv = multigpu.ShardedRenderer(ctx, rank, world)
    outs = []
    for batch in (16, 7):                                  # passes 1..16, then 17..23: two progressive readbacks
        drv.render(batch)
        with torch.cuda.stream(stream):
            s = drv.reduced(t)
        stream.synchronize()
        if rank == 0:
            img = ctx.read_output(1.0 / drv.samples_total(), dev_accum=s.data_ptr())
            outs.append((s.cpu().numpy().copy(), img.copy()))
    dist.barrier()
    if rank == 0:
        one = capi.Context(sc, device=local)
        for (acc, img), n in zip(outs, (16, 23)):
            one.reset_accum(); one.render_samples(1, n)
            np.testing.assert_allclose(acc, one.read_accum(), rtol=2e-5, atol=1e-5)
            assert np.abs(img.astype(int) - one.read_output(1.0 / n).astype(int)).max() <= 1
        print('NCCL_OK', world)
    dist.barrier(); dist.destroy_process_group()
""")


@pytest.mark.skipif(n_gpus() < 2, reason="needs 2 GPUs")
def test_torchrun_two_ranks_nccl_image_equals_one_gpu(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=ROOT))
    port = 29600 + (os.getpid() % 1000)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), str(script)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "NCCL_OK 2" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
