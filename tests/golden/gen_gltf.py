#!/usr/bin/env python
"""glTF input fixture (north star: "the same Scene, loaders and .scene/GLTF inputs").

Generates a .gltf (+ external .bin and PNG images) and the same model as a .glb, referenced from one .scene file through the
reference's `gltf { file ... }` blocks (Loader.cpp:454-511), and has the UNMODIFIED reference host code load it
(GLTFLoader.cpp via tinygltf, Texture resize to the array size, RadeonRays BVH, BvhTranslator) -> tests/golden/scenes/gltf_mix.ptscene.xz.
What the model exercises in GLTFLoader.cpp: uint8 / uint16 / uint32 index buffers (:166-199), interleaved vertex buffer with
byteStride (:84-86), multi-primitive mesh -> one BLAS per primitive (:44-46), node hierarchy with TRS and `matrix` nodes (:297-356),
metallic-roughness + normal + emissive textures, baseColor alpha with MASK and BLEND (:246-271), KHR_materials_transmission (:273-279),
sqrt(roughnessFactor) (:259), and the second load's texture-index offset incl. the normal-map `-1 + sceneTexIdx` quirk (:264).
Runs only where /root/reference exists."""
import base64, json, lzma, math, os, struct, subprocess, sys, tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gen_synthetic import write_hdr, sky

FLOAT, UBYTE, USHORT, UINT = 5126, 5121, 5123, 5125


def sphere(nu, nv, r):
    P, N, T, I = [], [], [], []
    for j in range(nv + 1):
        t = math.pi * j / nv
        for i in range(nu + 1):
            p = 2 * math.pi * i / nu
            n = (math.sin(t) * math.cos(p), math.cos(t), math.sin(t) * math.sin(p))
            P.append([r * x for x in n]); N.append(n); T.append((2.0 * i / nu, 1.0 * j / nv))
    idx = lambda i, j: j * (nu + 1) + i
    for j in range(nv):
        for i in range(nu):
            a, b, c, d = idx(i, j), idx(i + 1, j), idx(i + 1, j + 1), idx(i, j + 1)
            if j > 0: I += [a, b, c]
            if j < nv - 1: I += [a, c, d]
    return np.array(P, np.float32), np.array(N, np.float32), np.array(T, np.float32), np.array(I)


def box(h):
    P, N, T, I = [], [], [], []
    for ax in range(3):
        for s in (-1, 1):
            n = [0, 0, 0]; n[ax] = s
            u = [0, 0, 0]; u[(ax + 1) % 3] = 1
            v = np.cross(n, u)
            b = len(P)
            for (a, c) in ((-1, -1), (1, -1), (1, 1), (-1, 1)):
                P.append((np.array(n) + a * np.array(u) + c * v) * h); N.append(n); T.append(((a + 1) / 2, (c + 1) / 2))
            I += [b, b + 1, b + 2, b, b + 2, b + 3]
    return np.array(P, np.float32), np.array(N, np.float32), np.array(T, np.float32), np.array(I)


def plane(half, uvs):
    P = np.array([(-half, 0, -half), (half, 0, -half), (half, 0, half), (-half, 0, half)], np.float32)
    N = np.array([(0, 1, 0)] * 4, np.float32)
    T = np.array([(0, 0), (uvs, 0), (uvs, uvs), (0, uvs)], np.float32)
    return P, N, T, np.array([0, 2, 1, 0, 3, 2])


def images(tmp):
    from PIL import Image
    n = 64
    y, x = np.mgrid[0:n, 0:n]
    c = (((x // 8) + (y // 8)) % 2).astype(np.uint8)
    alpha = np.where(((x // 4) + (y // 4)) % 3 == 0, 30, 255).astype(np.uint8)
    Image.fromarray(np.stack([70 + 170 * c, 200 - 120 * c, 90 + 40 * c, alpha], -1).astype(np.uint8), "RGBA").save(os.path.join(tmp, "base.png"))
    rough = (40 + 180 * ((x // 16) % 2)).astype(np.uint8); metal = (255 * ((y // 16) % 2)).astype(np.uint8)
    Image.fromarray(np.stack([np.zeros_like(rough), rough, metal], -1).astype(np.uint8), "RGB").save(os.path.join(tmp, "mr.png"))
    nx = np.sin(x / n * 8 * math.pi) * 0.35; ny = np.cos(y / n * 6 * math.pi) * 0.35
    nz = np.sqrt(np.clip(1 - nx * nx - ny * ny, 0, 1))
    Image.fromarray(((np.stack([nx, ny, nz], -1) * 0.5 + 0.5) * 255).astype(np.uint8), "RGB").save(os.path.join(tmp, "normal.png"))
    r = np.hypot(x - n / 2, y - n / 2)
    e = np.clip(1.2 - r / 20, 0, 1)
    Image.fromarray((np.stack([e, e * 0.6, e * 0.2], -1) * 255).astype(np.uint8), "RGB").save(os.path.join(tmp, "emissive.png"))


def build_model(tmp):
    blob = bytearray()
    views, accessors = [], []

    def view(data, stride=None):
        while len(blob) % 4: blob.append(0)
        v = {"buffer": 0, "byteOffset": len(blob), "byteLength": len(data)}
        if stride: v["byteStride"] = stride
        blob.extend(data); views.append(v)
        return len(views) - 1

    def accessor(view_i, ctype, count, typ, offset=0, minmax=None):
        a = {"bufferView": view_i, "byteOffset": offset, "componentType": ctype, "count": count, "type": typ}
        if minmax is not None:
            a["min"], a["max"] = [float(v) for v in minmax[0]], [float(v) for v in minmax[1]]
        accessors.append(a)
        return len(accessors) - 1

    def prim(P, N, T, I, index_type, material, interleave=False):
        if interleave:      # one vertex buffer view, stride 32: position | normal | uv
            inter = np.concatenate([P, N, T], axis=1).astype(np.float32)
            v = view(inter.tobytes(), stride=32)
            ap = accessor(v, FLOAT, len(P), "VEC3", 0, (P.min(0), P.max(0))); an = accessor(v, FLOAT, len(P), "VEC3", 12); at = accessor(v, FLOAT, len(P), "VEC2", 24)
        else:
            ap = accessor(view(P.tobytes()), FLOAT, len(P), "VEC3", 0, (P.min(0), P.max(0)))
            an = accessor(view(N.tobytes()), FLOAT, len(P), "VEC3"); at = accessor(view(T.tobytes()), FLOAT, len(P), "VEC2")
        dt = {UBYTE: np.uint8, USHORT: np.uint16, UINT: np.uint32}[index_type]
        ai = accessor(view(I.astype(dt).tobytes()), index_type, len(I), "SCALAR")
        return {"attributes": {"POSITION": ap, "NORMAL": an, "TEXCOORD_0": at}, "indices": ai, "material": material, "mode": 4}

    sp = sphere(24, 16, 1.0); bx = box(0.7); pl = plane(6.0, 6.0)
    half = len(sp[3]) // 2 // 3 * 3
    meshes = [
        {"name": "ball", "primitives": [prim(*sp, USHORT, 0, interleave=True)]},
        # two primitives with different materials: upper / lower half of the sphere (GLTFLoader.cpp:44-46 -> two BLAS)
        {"name": "split_ball", "primitives": [prim(sp[0], sp[1], sp[2], sp[3][:half], UINT, 1), prim(sp[0], sp[1], sp[2], sp[3][half:], USHORT, 3)]},
        {"name": "crate", "primitives": [prim(*bx, UBYTE, 2)]},
        {"name": "ground", "primitives": [prim(*pl, UBYTE, 5)]},
        {"name": "veil", "primitives": [prim(*plane(1.2, 2.0), UBYTE, 4)]},
    ]
    materials = [
        {"name": "textured", "pbrMetallicRoughness": {"baseColorFactor": [0.9, 0.9, 0.9, 1.0], "baseColorTexture": {"index": 0},
                                                       "metallicRoughnessTexture": {"index": 1}, "roughnessFactor": 0.5, "metallicFactor": 1.0},
         "normalTexture": {"index": 2}},
        {"name": "glass", "pbrMetallicRoughness": {"baseColorFactor": [1.0, 1.0, 1.0, 1.0], "roughnessFactor": 0.0, "metallicFactor": 0.0},
         "extensions": {"KHR_materials_transmission": {"transmissionFactor": 1.0}}},
        {"name": "cutout", "pbrMetallicRoughness": {"baseColorFactor": [1.0, 1.0, 1.0, 1.0], "baseColorTexture": {"index": 0}, "roughnessFactor": 0.64,
                                                     "metallicFactor": 0.0}, "alphaMode": "MASK", "alphaCutoff": 0.5, "doubleSided": True},
        {"name": "glow", "pbrMetallicRoughness": {"baseColorFactor": [0.2, 0.2, 0.2, 1.0], "roughnessFactor": 0.81, "metallicFactor": 0.0},
         "emissiveFactor": [1.0, 1.0, 1.0], "emissiveTexture": {"index": 3}},
        {"name": "veil", "pbrMetallicRoughness": {"baseColorFactor": [0.3, 0.5, 0.9, 0.45], "roughnessFactor": 0.25, "metallicFactor": 0.0}, "alphaMode": "BLEND"},
        {"name": "ground", "pbrMetallicRoughness": {"baseColorFactor": [0.8, 0.8, 0.8, 1.0], "baseColorTexture": {"index": 0}, "roughnessFactor": 0.36,
                                                     "metallicFactor": 0.0}},
    ]
    q = lambda ax, ang: [ax[0] * math.sin(ang / 2), ax[1] * math.sin(ang / 2), ax[2] * math.sin(ang / 2), math.cos(ang / 2)]
    c30, s30 = math.cos(math.radians(30)), math.sin(math.radians(30))
    nodes = [
        {"name": "root", "children": [1, 2, 3], "scale": [1.0, 1.0, 1.0], "translation": [0.0, 0.0, 0.0]},
        {"name": "group", "children": [4, 5], "rotation": q((0, 1, 0), 0.6), "translation": [0.0, 1.0, 0.0], "scale": [1.0, 1.1, 1.0]},
        {"name": "ground", "mesh": 3},
        # column-major matrix node: rotation about y by 30 deg, non-uniform scale, translation (5 cm above the ground: no coplanar faces)
        {"name": "crate", "mesh": 2, "matrix": [c30 * 1.2, 0, -s30 * 1.2, 0, 0, 0.9, 0, 0, s30, 0, c30, 0, 2.2, 0.68, 1.8, 1]},
        {"name": "ball", "mesh": 0, "translation": [-1.4, 0.0, -0.8]},
        {"name": "pair", "children": [6, 7], "translation": [1.3, 0.1, -1.6], "scale": [0.8, 0.8, 0.8]},
        {"name": "split_ball", "mesh": 1},
        {"name": "veil", "mesh": 4, "translation": [0.0, 1.6, 0.0], "rotation": q((1, 0, 0), 0.4)},
    ]
    model = {
        "asset": {"version": "2.0", "generator": "tests/golden/gen_gltf.py"},
        "extensionsUsed": ["KHR_materials_transmission"],
        "scene": 0, "scenes": [{"nodes": [0]}], "nodes": nodes, "meshes": meshes, "materials": materials,
        "textures": [{"source": i, "sampler": 0} for i in range(4)],
        "samplers": [{"magFilter": 9729, "minFilter": 9729, "wrapS": 10497, "wrapT": 10497}],
        "images": [{"uri": "base.png"}, {"uri": "mr.png"}, {"uri": "normal.png"}, {"uri": "emissive.png"}],
        "accessors": accessors, "bufferViews": views,
    }
    return model, bytes(blob)


def write_gltf(tmp, model, blob):
    m = dict(model); m["buffers"] = [{"uri": "mix.bin", "byteLength": len(blob)}]
    with open(os.path.join(tmp, "mix.bin"), "wb") as f: f.write(blob)
    with open(os.path.join(tmp, "mix.gltf"), "w") as f: json.dump(m, f)


def write_glb(tmp, model, blob):
    """Binary container: images embedded as bufferViews, one BIN chunk."""
    m = json.loads(json.dumps(model))
    blob = bytearray(blob)
    for img in m["images"]:
        data = open(os.path.join(tmp, img.pop("uri")), "rb").read()
        while len(blob) % 4: blob.append(0)
        m["bufferViews"].append({"buffer": 0, "byteOffset": len(blob), "byteLength": len(data)})
        blob.extend(data)
        img["bufferView"] = len(m["bufferViews"]) - 1; img["mimeType"] = "image/png"
    while len(blob) % 4: blob.append(0)
    m["buffers"] = [{"byteLength": len(blob)}]
    js = json.dumps(m).encode()
    js += b" " * (-len(js) % 4)
    with open(os.path.join(tmp, "mix.glb"), "wb") as f:
        f.write(struct.pack("<III", 0x46546C67, 2, 12 + 8 + len(js) + 8 + len(blob)))
        f.write(struct.pack("<II", len(js), 0x4E4F534A)); f.write(js)
        f.write(struct.pack("<II", len(blob), 0x004E4942)); f.write(bytes(blob))


SCENE = """renderer
{
	resolution 640 360
	maxdepth 4
	tilewidth 160
	tileheight 90
	envmapfile sky.hdr
	envmapintensity 1.5
	texarraywidth 128
	texarrayheight 128
}
camera
{
	position 6.5 3.2 5.0
	lookat 0.3 0.9 0.0
	fov 45
}
light
{
	type quad
	position -1.5 5.0 -1.5
	v1 1.5 5.0 -1.5
	v2 -1.5 5.0 1.5
	emission 12 11 10
}
gltf
{
	file mix.gltf
}
# the second copy stands 7 cm above the first one's ground plane: coplanar overlapping floors would make every floor hit an
# exact-tie z-fight between two instances, decided by the last ulp of the bounce ray (no implementation can be RNG-matched there)
gltf
{
	file mix.glb
	position -4.5 0.07 1.0
	scale 0.6 0.6 0.6
	rotation 0.0 0.3826834 0.0 0.9238795
}
"""


def main():
    dump = os.path.join(ROOT, "oracle", "_ref", "scene_dump")
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle", "ref_host")], stdout=subprocess.DEVNULL)
    out_dir = os.path.join(ROOT, "tests", "golden", "scenes")
    with tempfile.TemporaryDirectory() as tmp:
        write_hdr(os.path.join(tmp, "sky.hdr"), sky(256, 128))
        images(tmp)
        model, blob = build_model(tmp)
        write_gltf(tmp, model, blob); write_glb(tmp, model, blob)
        with open(os.path.join(tmp, "gltf_mix.scene"), "w") as f: f.write(SCENE)
        raw = os.path.join(tmp, "gltf_mix.ptscene")
        out = subprocess.check_output([dump, os.path.join(tmp, "gltf_mix.scene"), raw], text=True)
        print([l for l in out.splitlines() if l.startswith("PTSCENE")][0])
        # direct load of the .gltf / .glb (Main.cpp:129-134 dispatch): the binary container must give the same arrays as the ASCII one
        a = subprocess.check_output([dump, os.path.join(tmp, "mix.gltf"), os.path.join(tmp, "a.ptscene")], text=True)
        b = subprocess.check_output([dump, os.path.join(tmp, "mix.glb"), os.path.join(tmp, "b.ptscene")], text=True)
        la = [l for l in a.splitlines() if l.startswith("PTSCENE")][0].split(" ", 2)[2]; lb = [l for l in b.splitlines() if l.startswith("PTSCENE")][0].split(" ", 2)[2]
        assert la == lb, (la, lb)
        print("gltf == glb:", la)
        with open(raw, "rb") as f, lzma.open(os.path.join(out_dir, "gltf_mix.ptscene.xz"), "wb", preset=6) as g:
            g.write(f.read())
        # the generated inputs themselves, for the C++ drop-in binary on the GPU box (oracle/_ref is git-ignored but travels)
        import shutil
        dst = os.path.join(ROOT, "oracle", "_ref", "assets", "gltf_mix")
        shutil.rmtree(dst, ignore_errors=True); os.makedirs(dst)
        for fn in ("gltf_mix.scene", "mix.gltf", "mix.bin", "mix.glb", "sky.hdr", "base.png", "mr.png", "normal.png", "emissive.png"):
            shutil.copy(os.path.join(tmp, fn), dst)


if __name__ == "__main__":
    main()
