"""Host-side mirror of GLSLPT::Renderer (state machine of reference Renderer.cpp:281-300,546-590,641-812) exercised on the CPU
with a recording stand-in for the C-ABI context, plus the sample-sharding arithmetic of the multi-GPU driver."""
import numpy as np
import pytest
from conftest import scene_at
from glsl_pathtracer_b200 import capi, multigpu
from glsl_pathtracer_b200 import renderer as R


class FakeCtx:
    """Records the calls the Renderer makes instead of touching a GPU (test double for capi.Context)."""

    def __init__(self, scene, device=0, samples_per_wave=0, features=None):
        self.scene, self.calls = scene, []
        self.opts = capi.PtbOptions()
        self.opts.renderW, self.opts.renderH = scene.renderOptions.renderResolution
        self.opts.features = 6

    def __getattr__(self, name):
        def rec(*a, **k):
            self.calls.append((name, a))
            if name in ("read_output", "read_snapshot"):
                return np.zeros((self.opts.renderH, self.opts.renderW, 4), np.uint8)
            if name == "render_preview":
                return np.zeros((a[1], a[0], 4), np.float32)
            return None
        return rec

    def close(self):
        pass


@pytest.fixture
def fake(monkeypatch):
    monkeypatch.setattr(capi, "Context", FakeCtx)
    monkeypatch.setattr(capi, "make_options", lambda scene, features=None, samples_per_wave=0: capi.PtbOptions())


def test_initial_state_matches_initfbos(fake):
    sc = scene_at("hyperion_rect_lights", 1920, 1080, 256, 144)
    r = R.Renderer(sc, "shaders/")
    assert r.numTiles == (8, 8)                       # ceil(1920/256), ceil(1080/144): over-hanging last column and top row
    assert r.tile == [-1, 7] and r.sampleCounter == 1 and r.frameCounter == 1 and r.GetSampleCount() == 1
    assert abs(r.invNumTiles[0] - 256 / 1920) < 1e-7 and r.pixelRatio == 0.25


def test_tile_schedule_and_frame_numbers(fake):
    """First Update is the dirty one (preview); afterwards tiles run x-fastest from the TOP row and frameNum increments per tile:
    frameNum = 2 + (s-1)*T + j  (SURVEY §8(b) state machine)."""
    sc = scene_at("cornell_box_orig", 100, 72, 48, 32)          # 3 x 3 tiles
    r = R.Renderer(sc, "", coalesce=False)                      # one wavefront per Render() tile, as the reference draws
    seen = []
    for _ in range(1 + 2 * 9 + 1):
        r.Update(0.0)
        r.Render()
    tiles = [c for c in r.ctx.calls if c[0] == "render_tile"]
    previews = [c for c in r.ctx.calls if c[0] == "render_preview"]
    assert len(previews) == 1 and previews[0][1] == (25, 18)    # windowSize * 0.25
    assert len(tiles) == 19
    for k, (_, (tx, ty, frame)) in enumerate(tiles):
        s, j = k // 9 + 1, k % 9
        assert (tx, ty) == (j % 3, 2 - j // 3) and frame == 2 + (s - 1) * 9 + j
    assert r.GetSampleCount() == 3                             # two passes complete, third in progress
    outs = [c for c in r.ctx.calls if c[0] == "snapshot_output"]
    assert [round(1 / o[1][0]) for o in outs] == [1, 2]        # tonemap uniform invSampleCounter = 1/sampleCounter of the finished pass
    assert not [c for c in r.ctx.calls if c[0] in ("read_output", "read_snapshot")]     # no host copy until GetOutputBuffer
    r.GetOutputBuffer()
    assert r.ctx.calls[-1][0] == "read_snapshot"


def test_coalesced_passes_keep_the_tile_walk_and_render_once_per_pass(fake):
    """Default mode: the whole pass is rendered at its first tile (render_pass(sample, remaining passes)); the other tiles' Render() calls
    draw nothing, counters / tile walk / tonemap uniforms are those of the reference's loop; a dirty scene restarts at pass 1."""
    sc = scene_at("cornell_box_orig", 100, 72, 48, 32)          # 3 x 3 tiles
    sc.renderOptions.maxSpp = 5
    r = R.Renderer(sc, "")
    walk = []
    for _ in range(1 + 2 * 9 + 4):
        r.Update(0.0); walk.append((tuple(r.tile), r.frameCounter)); r.Render()
    assert [c[0] for c in r.ctx.calls if c[0] in ("render_tile", "render_pass", "render_preview")] == ["render_preview", "render_pass", "render_pass", "render_pass"]
    assert [c[1] for c in r.ctx.calls if c[0] == "render_pass"] == [(1, 4), (2, 3), (3, 2)]      # (sample, passes left before maxSpp stops the loop)
    assert walk[1] == ((0, 2), 2) and walk[9] == ((2, 0), 10) and walk[10] == ((0, 2), 11)
    assert [round(1 / c[1][0]) for c in r.ctx.calls if c[0] == "snapshot_output"] == [1, 2]
    sc.dirty = True
    r.Update(0.0); r.Render(); r.Update(0.0); r.Render()
    assert [c[1] for c in r.ctx.calls if c[0] == "render_pass"][-1] == (1, 4) and r.sampleCounter == 1


def test_maxspp_renders_maxspp_minus_one_passes(fake):
    """Render/Update return early once sampleCounter >= maxSpp."""
    sc = scene_at("cornell_box_orig", 96, 64, 48, 32)
    sc.renderOptions.maxSpp = 4
    r = R.Renderer(sc, "", coalesce=False)
    for _ in range(200):
        r.Update(0.0); r.Render()
    assert r.GetSampleCount() == 4 and r.GetProgress() == 100.0
    assert len([c for c in r.ctx.calls if c[0] == "render_tile"]) == 3 * 4
    sc.renderOptions.maxSpp = -1
    assert r.GetProgress() == 0.0


def test_dirty_scene_resets_and_uploads_instances(fake):
    sc = scene_at("cornell_box_orig", 96, 64, 48, 32)
    r = R.Renderer(sc, "")
    for _ in range(6):
        r.Update(0.0); r.Render()
    assert r.sampleCounter == 2
    sc.dirty = True; sc.instancesModified = True                # Scene::RebuildInstances (Scene.cpp:200-214)
    r.Update(0.0)
    assert r.sampleCounter == 1 and r.frameCounter == 1 and r.tile == [-1, 1]
    names = [c[0] for c in r.ctx.calls]
    assert "update_instances" in names and names.count("reset_accum") == 2
    r.Render()
    assert not sc.dirty and not sc.instancesModified


def test_null_scene_prints_and_returns(capsys):
    r = R.Renderer(None, "")
    assert "No Scene Found" in capsys.readouterr().out and r.ctx is None


def test_shard_passes_partition_every_pass_exactly_once():
    for first, total, world in [(1, 32, 8), (5, 7, 4), (1, 3, 8), (9, 0, 2), (1, 1000, 3)]:
        got = sorted(p for r in range(world) for p in multigpu.passes_of(first, total, r, world))
        assert got == list(range(first, first + total))
        counts = [multigpu.shard_passes(first, total, r, world)[1] for r in range(world)]
        assert max(counts) - min(counts) <= 1
    with pytest.raises(ValueError):
        multigpu.shard_passes(0, 4, 0, 2)
