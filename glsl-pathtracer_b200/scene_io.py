"""Scene blobs: the arrays `Renderer::InitGPUDataBuffers` uploads (reference Renderer.cpp:135-249), as dumped
by the unmodified reference host code (the `scene_dump` tool of the test infrastructure) into a `.ptscene` file.

Host-side data contract only - no compute happens here.
"""
from __future__ import annotations
import lzma, os, struct
from dataclasses import dataclass, field
import numpy as np

# Mirrors `struct Scalars` of the scene_dump tool (camera: Camera.h:46-54, options: Renderer.h:37-100).
_SCALARS = np.dtype([
    ("topLevelIndex", "<i4"), ("numNodes", "<i4"), ("numIndices", "<i4"), ("numVertices", "<i4"), ("numMaterials", "<i4"),
    ("numInstances", "<i4"), ("numLights", "<i4"), ("numTextures", "<i4"),
    ("texW", "<i4"), ("texH", "<i4"), ("envW", "<i4"), ("envH", "<i4"), ("envTotalSum", "<f4"),
    ("camPosition", "<f4", 3), ("camUp", "<f4", 3), ("camRight", "<f4", 3), ("camForward", "<f4", 3),
    ("camFov", "<f4"), ("camFocalDist", "<f4"), ("camAperture", "<f4"),
    ("renderW", "<i4"), ("renderH", "<i4"), ("windowW", "<i4"), ("windowH", "<i4"), ("tileW", "<i4"), ("tileH", "<i4"),
    ("maxDepth", "<i4"), ("maxSpp", "<i4"), ("RRDepth", "<i4"), ("denoiserFrameCnt", "<i4"),
    ("uniformLightCol", "<f4", 3), ("backgroundCol", "<f4", 3), ("envMapIntensity", "<f4"), ("envMapRot", "<f4"),
    ("roughnessMollificationAmt", "<f4"),
    ("enableRR", "<i4"), ("enableDenoiser", "<i4"), ("enableTonemap", "<i4"), ("enableAces", "<i4"), ("simpleAcesFit", "<i4"),
    ("openglNormalMap", "<i4"), ("enableEnvMap", "<i4"), ("enableUniformLight", "<i4"), ("hideEmitters", "<i4"),
    ("enableBackground", "<i4"), ("transparentBackground", "<i4"), ("independentRenderSize", "<i4"),
    ("enableRoughnessMollification", "<i4"), ("enableVolumeMIS", "<i4"),
    ("sceneBoundsMin", "<f4", 3), ("sceneBoundsMax", "<f4", 3), ("tlasHeight", "<i4"), ("maxBlasHeight", "<i4"),
])


def fnv1a64(data: bytes) -> int:
    """FNV-1a 64 over raw bytes (the hash SURVEY.md §8(c) pins the flattened BVH with)."""
    h = 1469598103934665603
    # vectorised: process with python ints in chunks would be slow for MBs; use numpy loop-free trick is not possible
    # for a sequential hash, so fall back to a tight loop over a memoryview (few MB -> ~1 s).
    for b in memoryview(data):
        h = ((h ^ b) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


@dataclass
class RenderOptions:
    """Mirror of `GLSLPT::RenderOptions` (reference Renderer.h:37-100), same field names and defaults."""
    renderResolution: tuple = (1280, 720)
    windowResolution: tuple = (1280, 720)
    uniformLightCol: tuple = (0.3, 0.3, 0.3)
    backgroundCol: tuple = (1.0, 1.0, 1.0)
    tileWidth: int = 100
    tileHeight: int = 100
    maxDepth: int = 2
    maxSpp: int = -1
    RRDepth: int = 2
    texArrayWidth: int = 2048
    texArrayHeight: int = 2048
    denoiserFrameCnt: int = 20
    enableRR: bool = True
    enableDenoiser: bool = False
    enableTonemap: bool = True
    enableAces: bool = False
    simpleAcesFit: bool = False
    openglNormalMap: bool = True
    enableEnvMap: bool = False
    enableUniformLight: bool = False
    hideEmitters: bool = False
    enableBackground: bool = False
    transparentBackground: bool = False
    independentRenderSize: bool = False
    enableRoughnessMollification: bool = False
    enableVolumeMIS: bool = False
    envMapIntensity: float = 1.0
    envMapRot: float = 0.0
    roughnessMollificationAmt: float = 0.0


@dataclass
class Camera:
    """The camera members the renderer reads (reference Camera.h:46-54, Renderer.cpp:769-775)."""
    position: np.ndarray
    up: np.ndarray
    right: np.ndarray
    forward: np.ndarray
    fov: float          # radians
    focalDist: float
    aperture: float


@dataclass
class Scene:
    """The data members of `GLSLPT::Scene` the renderer consumes (reference Scene.h:87-125)."""
    nodes: np.ndarray            # (numNodes, 9) f32  bvhTranslator.nodes
    topLevelIndex: int
    vertIndices: np.ndarray      # (n, 3) i32
    verticesUVX: np.ndarray      # (n, 4) f32
    normalsUVY: np.ndarray       # (n, 4) f32
    materials: np.ndarray        # (n, 32) f32
    transforms: np.ndarray       # (n, 16) f32
    lights: np.ndarray           # (n, 15) f32
    textures: np.ndarray         # (layers, H, W, 4) u8
    envImg: np.ndarray | None    # (H, W, 3) f32
    envCdf: np.ndarray | None    # (H, W) f32
    envTotalSum: float
    camera: Camera
    renderOptions: RenderOptions
    instances: np.ndarray        # (n, 2) i32 (meshID, materialID)
    sceneBounds: tuple = ((0, 0, 0), (0, 0, 0))
    tlasHeight: int = 0
    maxBlasHeight: int = 0
    dirty: bool = True
    instancesModified: bool = False
    envMapModified: bool = False
    initialized: bool = True
    name: str = ""
    extra: dict = field(default_factory=dict)


def _read_bytes(path: str) -> bytes:
    if path.endswith(".xz"):
        with lzma.open(path, "rb") as f:
            return f.read()
    with open(path, "rb") as f:
        return f.read()


def load_ptscene(path: str) -> Scene:
    raw = _read_bytes(path)
    if raw[:8] != b"PTBSCN01":
        raise ValueError(f"{path}: not a ptscene blob")
    (nsec,) = struct.unpack_from("<I", raw, 8)
    secs = {}
    for i in range(nsec):
        name, nbytes, off = struct.unpack_from("<16sQQ", raw, 12 + 32 * i)
        secs[name.rstrip(b"\0").decode()] = raw[off:off + nbytes]
    sc = np.frombuffer(secs["scalars"], dtype=_SCALARS, count=1)[0]

    def arr(name, dt, cols):
        b = secs.get(name, b"")
        a = np.frombuffer(b, dtype=dt).copy()
        return a.reshape(-1, cols) if cols else a

    ro = RenderOptions(
        renderResolution=(int(sc["renderW"]), int(sc["renderH"])), windowResolution=(int(sc["windowW"]), int(sc["windowH"])),
        uniformLightCol=tuple(float(x) for x in sc["uniformLightCol"]), backgroundCol=tuple(float(x) for x in sc["backgroundCol"]),
        tileWidth=int(sc["tileW"]), tileHeight=int(sc["tileH"]), maxDepth=int(sc["maxDepth"]), maxSpp=int(sc["maxSpp"]),
        RRDepth=int(sc["RRDepth"]), texArrayWidth=int(sc["texW"]), texArrayHeight=int(sc["texH"]),
        denoiserFrameCnt=int(sc["denoiserFrameCnt"]), enableRR=bool(sc["enableRR"]), enableDenoiser=bool(sc["enableDenoiser"]),
        enableTonemap=bool(sc["enableTonemap"]), enableAces=bool(sc["enableAces"]), simpleAcesFit=bool(sc["simpleAcesFit"]),
        openglNormalMap=bool(sc["openglNormalMap"]), enableEnvMap=bool(sc["enableEnvMap"]),
        enableUniformLight=bool(sc["enableUniformLight"]), hideEmitters=bool(sc["hideEmitters"]),
        enableBackground=bool(sc["enableBackground"]), transparentBackground=bool(sc["transparentBackground"]),
        independentRenderSize=bool(sc["independentRenderSize"]),
        enableRoughnessMollification=bool(sc["enableRoughnessMollification"]), enableVolumeMIS=bool(sc["enableVolumeMIS"]),
        envMapIntensity=float(sc["envMapIntensity"]), envMapRot=float(sc["envMapRot"]),
        roughnessMollificationAmt=float(sc["roughnessMollificationAmt"]))
    cam = Camera(position=sc["camPosition"].astype(np.float32).copy(), up=sc["camUp"].astype(np.float32).copy(),
                 right=sc["camRight"].astype(np.float32).copy(), forward=sc["camForward"].astype(np.float32).copy(),
                 fov=float(sc["camFov"]), focalDist=float(sc["camFocalDist"]), aperture=float(sc["camAperture"]))
    ntex, tw, th = int(sc["numTextures"]), int(sc["texW"]), int(sc["texH"])
    tex = np.frombuffer(secs.get("textures", b""), dtype=np.uint8).copy()
    tex = tex.reshape(ntex, th, tw, 4) if ntex else np.zeros((0, 1, 1, 4), np.uint8)
    ew, eh = int(sc["envW"]), int(sc["envH"])
    envImg = arr("envImg", "<f4", 0).reshape(eh, ew, 3) if ew else None
    envCdf = arr("envCdf", "<f4", 0).reshape(eh, ew) if ew else None
    return Scene(
        nodes=arr("nodes", "<f4", 9), topLevelIndex=int(sc["topLevelIndex"]), vertIndices=arr("vertIndices", "<i4", 3),
        verticesUVX=arr("verticesUVX", "<f4", 4), normalsUVY=arr("normalsUVY", "<f4", 4), materials=arr("materials", "<f4", 32),
        transforms=arr("transforms", "<f4", 16), lights=arr("lights", "<f4", 15), textures=tex, envImg=envImg, envCdf=envCdf,
        envTotalSum=float(sc["envTotalSum"]), camera=cam, renderOptions=ro, instances=arr("instances", "<i4", 2),
        sceneBounds=(tuple(float(x) for x in sc["sceneBoundsMin"]), tuple(float(x) for x in sc["sceneBoundsMax"])),
        tlasHeight=int(sc["tlasHeight"]), maxBlasHeight=int(sc["maxBlasHeight"]),
        name=os.path.basename(path).split(".")[0])


_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SCENE_DIRS = [d for d in os.environ.get("PTB_SCENE_DIRS", "").split(os.pathsep) if d] + [os.path.join(_ROOT, "tests", "golden", "scenes")]


def find_scene(name: str) -> str:
    """Locate `<name>.ptscene[.xz]` in $PTB_SCENE_DIRS or tests/golden/scenes (committed fixtures)."""
    for d in SCENE_DIRS:
        for ext in (".ptscene", ".ptscene.xz"):
            p = os.path.join(d, name + ext)
            if os.path.exists(p):
                return p
    raise FileNotFoundError(f"scene blob '{name}' not found in {SCENE_DIRS}")


def load_scene(name: str) -> Scene:
    return load_ptscene(find_scene(name))


# ---- feature defines: Renderer::InitShaders (reference Renderer.cpp:401-459) ------------------------------------
OPT_ENVMAP, OPT_LIGHTS, OPT_RR, OPT_UNIFORM_LIGHT, OPT_OPENGL_NORMALMAP, OPT_HIDE_EMITTERS, OPT_BACKGROUND, \
    OPT_TRANSPARENT_BACKGROUND, OPT_ALPHA_TEST, OPT_ROUGHNESS_MOLLIFICATION, OPT_MEDIUM, OPT_VOL_MIS = (1 << i for i in range(12))


def derive_features(scene: Scene) -> int:
    ro = scene.renderOptions
    m = 0
    if ro.enableEnvMap and scene.envImg is not None:
        m |= OPT_ENVMAP
    if len(scene.lights):
        m |= OPT_LIGHTS
    if ro.enableRR:
        m |= OPT_RR
    if ro.enableUniformLight:
        m |= OPT_UNIFORM_LIGHT
    if ro.openglNormalMap:
        m |= OPT_OPENGL_NORMALMAP
    if ro.hideEmitters:
        m |= OPT_HIDE_EMITTERS
    if ro.enableBackground:
        m |= OPT_BACKGROUND
    if ro.transparentBackground:
        m |= OPT_TRANSPARENT_BACKGROUND
    if np.any(scene.materials[:, 29].astype(np.int32) != 0):     # alphaMode != Opaque
        m |= OPT_ALPHA_TEST
    if ro.enableRoughnessMollification:
        m |= OPT_ROUGHNESS_MOLLIFICATION
    if np.any(scene.materials[:, 18].astype(np.int32) != 0):     # mediumType != None
        m |= OPT_MEDIUM
    if ro.enableVolumeMIS:
        m |= OPT_VOL_MIS
    return m
