import os, sys
import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
import glsl_pathtracer_b200  # noqa: F401  (import shim for the hyphenated package directory)
from glsl_pathtracer_b200 import scene_io


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_sessionstart(session):
    """The suites load the in-tree libptb200.so; if a fresh checkout has not been built yet (__graft_entry__.build()), build it once
    here (nvcc cross-compiles sm_100a without a GPU).  Building is all this does — there is still no CPU path behind the library."""
    import subprocess
    lib = os.path.join(ROOT, "glsl-pathtracer_b200", "libptb200.so")
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(lib) and os.path.exists(nvcc):
        subprocess.run(["make", "-C", os.path.join(ROOT, "glsl-pathtracer_b200", "csrc")], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=False)


_SCENES = {}


def load_scene_cached(name):
    if name not in _SCENES:
        _SCENES[name] = scene_io.load_scene(name)
    return _SCENES[name]


def scene_at(name, w, h, tw=None, th=None, depth=None):
    """Fresh copy of a fixture scene with the render size / tile size / depth overridden."""
    import copy
    sc = copy.deepcopy(load_scene_cached(name))
    ro = sc.renderOptions
    ro.renderResolution = (w, h); ro.windowResolution = (w, h)
    if tw: ro.tileWidth = tw
    if th: ro.tileHeight = th
    if depth is not None: ro.maxDepth = depth
    return sc


@pytest.fixture(scope="session")
def oracle_mod():
    from oracle import binding
    binding.lib()
    return binding


def rel_mse(a, b, eps=1e-2):
    """relMSE of b against reference a over rgb: mean((a-b)^2 / (a^2 + eps))."""
    a = np.asarray(a, np.float64)[..., :3]; b = np.asarray(b, np.float64)[..., :3]
    return float(np.mean((a - b) ** 2 / (a * a + eps)))


def edited_scene(name, w, h, tw=None, th=None):
    """(scene, edited scene): the instance edit of tests/golden/instance_edit.npz — transforms / materials / TLAS slice as the
    reference's Scene::RebuildInstances() produced them (tests/golden/make_instance_edit_fixture.py)."""
    import copy
    fx = np.load(os.path.join(ROOT, "tests", "golden", "instance_edit.npz"))
    sc = scene_at(name, w, h, tw, th)
    sc2 = copy.deepcopy(sc)
    assert int(fx[name + "/top"]) == sc.topLevelIndex
    sc2.nodes = sc.nodes.copy(); sc2.nodes[sc.topLevelIndex:] = fx[name + "/tlas"]
    sc2.transforms = fx[name + "/transforms"].reshape(sc.transforms.shape).copy()
    sc2.materials = fx[name + "/materials"].copy()
    return sc, sc2
