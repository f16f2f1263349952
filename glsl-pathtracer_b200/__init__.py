"""ptb200: B200-native CUDA wavefront path tracer behind GLSL-PathTracer's `Renderer` surface.

Only the hot path (reference `tile.glsl` render loop) lives here: `csrc/` holds the sm_100a kernels and the C-ABI
(`include/ptb200.h`), `renderer.py` mirrors `GLSLPT::Renderer`, `scene_io.py` holds the scene data contract.
"""
from . import scene_io  # noqa: F401
