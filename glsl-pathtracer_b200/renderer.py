"""`Renderer`: host-side mirror of `GLSLPT::Renderer` (reference src/core/Renderer.h:104-185) over the C ABI.

Same public methods, same argument meaning, same state machine (sampleCounter/frameCounter/tile schedule,
reference Renderer.cpp:281-300, 546-590, 641-812); the GL body (3 draws per tile) is replaced by calls into
libptb200.so.  The C++ twin of this class lives in host/Renderer.{h,cpp}.
"""
from __future__ import annotations
import math
import numpy as np
from . import capi


class Renderer:
    def __init__(self, scene, shadersDirectory: str = "", device: int = 0, samples_per_wave: int = 0, coalesce: bool = True):
        """Renderer::Renderer (Renderer.cpp:41-87).  `shadersDirectory` is accepted for signature parity and ignored:
        the OPT_* defines select precompiled sm_100a kernels instead of GLSL sources."""
        self.scene = scene
        self.shadersDirectory = shadersDirectory
        self.ctx = None
        if scene is None:
            print("No Scene Found")                      # Renderer.cpp:72-76
            return
        # (scene->initialized / ProcessScene: scene blobs are always processed, Renderer.cpp:78-79)
        self.ctx = capi.Context(scene, device=device, samples_per_wave=samples_per_wave)      # InitGPUDataBuffers + InitShaders
        self.coalesce = coalesce                         # a whole sample pass per first-tile Render() (False: one wavefront per tile)
        self.pixelRatio = 0.25                           # Renderer.cpp:83
        self._init_fbos()

    # Renderer::InitFBOs (Renderer.cpp:281-379): counters, tile grid, buffers
    def _init_fbos(self):
        ro = self.scene.renderOptions
        self.sampleCounter = 1
        self.currentBuffer = 0
        self.frameCounter = 1
        self.renderSize = tuple(ro.renderResolution)
        self.windowSize = tuple(ro.windowResolution)
        self.tileWidth, self.tileHeight = ro.tileWidth, ro.tileHeight
        self.invNumTiles = (np.float32(self.tileWidth) / np.float32(self.renderSize[0]), np.float32(self.tileHeight) / np.float32(self.renderSize[1]))
        self.numTiles = (int(math.ceil(np.float32(self.renderSize[0]) / np.float32(self.tileWidth))),
                         int(math.ceil(np.float32(self.renderSize[1]) / np.float32(self.tileHeight))))
        self.tile = [-1, self.numTiles[1] - 1]
        self._have_completed = False      # tileOutputTexture[1-currentBuffer] lives on the device (ctx.snapshot_output)
        self._pass_coalesced = False
        self._preview = None
        self._invSampleCounter = 1.0
        self._maxDepthUniform = ro.maxDepth

    def __del__(self):
        if getattr(self, "ctx", None) is not None:
            self.ctx.close()

    def ResizeRenderer(self):
        """Renderer.cpp:251-279"""
        ro = self.scene.renderOptions
        self.ctx.set_options(capi.make_options(self.scene, samples_per_wave=self.ctx.opts.samplesPerWave))
        self.ctx.resize(ro.renderResolution[0], ro.renderResolution[1], ro.tileWidth, ro.tileHeight)
        self._init_fbos()

    def ReloadShaders(self):
        """Renderer.cpp:381-390: re-derive the OPT_* feature set from the current options/scene."""
        self.ctx.set_options(capi.make_options(self.scene, samples_per_wave=self.ctx.opts.samplesPerWave))

    def _maxspp_reached(self):
        ro = self.scene.renderOptions
        return (not self.scene.dirty) and ro.maxSpp != -1 and self.sampleCounter >= ro.maxSpp

    def Render(self):
        """Renderer.cpp:546-590: preview while dirty, else one tile of one sample pass."""
        if self._maxspp_reached():
            return
        sc = self.scene
        if sc.dirty:
            w, h = int(self.windowSize[0] * self.pixelRatio), int(self.windowSize[1] * self.pixelRatio)
            self._preview = self.ctx.render_preview(w, h)
            sc.instancesModified = False
            sc.dirty = False
            sc.envMapModified = False
        else:
            # Only the buffer of a completed pass is observable (Renderer.cpp:628-633): the whole pass is rendered as one wavefront at
            # its first tile (same frameNum per tile, same tile-local seeds), the other tiles of the pass draw nothing.
            T = self.numTiles[0] * self.numTiles[1]
            if self.tile[0] == 0 and self.tile[1] == self.numTiles[1] - 1:
                self._pass_coalesced = self.coalesce and self.frameCounter == 2 + (self.sampleCounter - 1) * T
                if self._pass_coalesced:
                    maxSpp = sc.renderOptions.maxSpp
                    self.ctx.render_pass(self.sampleCounter, 0 if maxSpp == -1 else maxSpp - self.sampleCounter)
            if not self._pass_coalesced:
                self.ctx.render_tile(self.tile[0], self.tile[1], self.frameCounter)

    def Present(self):
        """Renderer.cpp:592-611 draws to the window; headless build: no-op."""
        return

    def GetProgress(self) -> float:
        maxSpp = self.scene.renderOptions.maxSpp
        return 0.0 if maxSpp <= 0 else self.sampleCounter * 100.0 / maxSpp

    def GetSampleCount(self) -> int:
        return self.sampleCounter

    def GetOutputBuffer(self):
        """Renderer.cpp:619-634: tonemapped RGBA8 of the last COMPLETED pass, bottom row first.  Returns (data, w, h)."""
        w, h = self.renderSize
        return self.ctx.read_snapshot(), w, h            # zeros before the first completed pass (the cleared texture)

    def Update(self, secondsElapsed: float):
        """Renderer.cpp:641-812"""
        if self._maxspp_reached():
            return
        sc = self.scene
        ro = sc.renderOptions
        if sc.instancesModified:                           # :649-665
            self.ctx.update_instances(sc.transforms, sc.materials, sc.nodes[sc.topLevelIndex:])
        if sc.envMapModified and sc.envImg is not None:    # :668-692
            self.ctx.update_envmap(sc.envImg, sc.envCdf, sc.envTotalSum)
        if sc.dirty:                                       # :733-744
            self.tile = [-1, self.numTiles[1] - 1]
            self.sampleCounter = 1
            self.frameCounter = 1
            self.ctx.reset_accum()
        else:                                              # :745-762
            self.frameCounter += 1
            self.tile[0] += 1
            if self.tile[0] >= self.numTiles[0]:
                self.tile[0] = 0
                self.tile[1] -= 1
                if self.tile[1] < 0:
                    self.tile[0] = 0
                    self.tile[1] = self.numTiles[1] - 1
                    # all tiles of a pass done: the buffer written with the uniform of that pass becomes the displayed one
                    self.ctx.snapshot_output(self._invSampleCounter)      # frozen on the device; copied to the host by GetOutputBuffer
                    self.sampleCounter += 1
                    self.currentBuffer = 1 - self.currentBuffer
        # uniforms (:766-811)
        self.ctx.set_camera(sc.camera)
        o = capi.make_options(sc, features=self.ctx.opts.features, samples_per_wave=self.ctx.opts.samplesPerWave)
        self.ctx.set_options(o)
        self._invSampleCounter = float(np.float32(1.0) / np.float32(self.sampleCounter))     # :806

    # ---- B200 fast path (no reference counterpart): whole-frame passes with the same per-pixel seeds ------------
    def RenderSamples(self, n: int):
        """Equivalent to n * numTiles.x * numTiles.y Update()+Render() pairs on a non-dirty scene."""
        first = self.sampleCounter
        self.ctx.render_samples(first, n, 1)
        T = self.numTiles[0] * self.numTiles[1]
        self.frameCounter += n * T
        self.sampleCounter += n
        self.ctx.snapshot_output(float(np.float32(1.0) / np.float32(self.sampleCounter - 1)))
