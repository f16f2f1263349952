#!/usr/bin/env python
"""Synthetic stand-ins for the BASELINE configs whose assets are missing from the reference checkout (SURVEY asset note):
  ibl_spheres  (C2: teapot.scene stand-in)  two glossy spheres + checker-textured floor lit by a generated lat-long HDR
  instancing   (C5)                         N instanced spheres (one BLAS) in one TLAS + floor + the same HDR, depth 8
Inputs (OBJ, HDR, PNG, .scene) are generated here with a fixed seed, then loaded and processed by the UNMODIFIED reference host
code (oracle/_ref/scene_dump: Loader.cpp, stbi_loadf, EnvironmentMap::BuildCDF, RadeonRays BVH build, BvhTranslator) so the
committed blobs are reference-built.  Runs only where /root/reference exists."""
import lzma, math, os, struct, subprocess, sys, tempfile
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def write_hdr(path, img):
    """Radiance RGBE, flat (non-RLE) scanlines, top row first: readable by stb_image's stbi_loadf (EnvironmentMap.cpp:65)."""
    h, w, _ = img.shape
    m = img.max(axis=2)
    e = np.where(m > 1e-32, np.floor(np.log2(np.maximum(m, 1e-38))) + 1, 0)
    scale = np.where(m > 1e-32, 256.0 / np.exp2(e), 0)
    rgbe = np.zeros((h, w, 4), np.uint8)
    rgbe[..., :3] = np.clip(img * scale[..., None], 0, 255).astype(np.uint8)
    rgbe[..., 3] = np.where(m > 1e-32, e + 128, 0).astype(np.uint8)
    with open(path, "wb") as f:
        f.write(b"#?RADIANCE\nFORMAT=32-bit_rle_rgbe\n\n")
        f.write(f"-Y {h} +X {w}\n".encode())
        f.write(rgbe.tobytes())


def sky(w, h):
    v = (np.arange(h) + 0.5) / h; u = (np.arange(w) + 0.5) / w
    theta = v[:, None] * math.pi; phi = u[None, :] * 2 * math.pi
    d = np.stack([-np.sin(theta) * np.cos(phi), np.cos(theta) * np.ones_like(phi), -np.sin(theta) * np.sin(phi)], -1)
    up = np.clip(d[..., 1], 0, 1)
    img = np.zeros((h, w, 3)); img[...] = (0.25, 0.35, 0.6)
    img = img * (0.3 + 0.7 * up[..., None]) + np.array([0.5, 0.4, 0.3]) * (1 - up[..., None]) ** 4 * (d[..., 1:2] > 0)
    img[d[..., 1] <= 0] = (0.12, 0.10, 0.09)
    sun = np.array([0.45, 0.65, -0.6]); sun /= np.linalg.norm(sun)
    c = (d * sun).sum(-1)
    img += np.array([60.0, 52.0, 40.0]) * np.clip((c - 0.995) / 0.005, 0, 1)[..., None] ** 2
    return img.astype(np.float32)


def write_sphere_obj(path, nu, nv, r=1.0):
    vs, ns, ts, fs = [], [], [], []
    for j in range(nv + 1):
        t = math.pi * j / nv
        for i in range(nu + 1):
            p = 2 * math.pi * i / nu
            n = (math.sin(t) * math.cos(p), math.cos(t), math.sin(t) * math.sin(p))
            vs.append(tuple(r * x for x in n)); ns.append(n); ts.append((i / nu, 1 - j / nv))
    idx = lambda i, j: j * (nu + 1) + i + 1
    for j in range(nv):
        for i in range(nu):
            a, b, c, d = idx(i, j), idx(i + 1, j), idx(i + 1, j + 1), idx(i, j + 1)
            if j > 0: fs.append((a, b, c))
            if j < nv - 1: fs.append((a, c, d))
    with open(path, "w") as f:
        for v in vs: f.write("v %.6f %.6f %.6f\n" % v)
        for t in ts: f.write("vt %.6f %.6f\n" % t)
        for n in ns: f.write("vn %.6f %.6f %.6f\n" % n)
        for a, b, c in fs: f.write(f"f {a}/{a}/{a} {b}/{b}/{b} {c}/{c}/{c}\n")
    return len(fs)


def write_floor_obj(path, half, uvscale):
    with open(path, "w") as f:
        for x, z in ((-half, -half), (half, -half), (half, half), (-half, half)):
            f.write(f"v {x} 0 {z}\n")
        for u, v in ((0, 0), (uvscale, 0), (uvscale, uvscale), (0, uvscale)):
            f.write(f"vt {u} {v}\n")
        f.write("vn 0 1 0\nf 1/1/1 3/3/1 2/2/1\nf 1/1/1 4/4/1 3/3/1\n")


def write_checker_png(path, n=64, cells=8):
    from PIL import Image
    y, x = np.mgrid[0:n, 0:n]
    c = (((x // (n // cells)) + (y // (n // cells))) % 2).astype(np.uint8)
    img = np.stack([60 + 180 * c, 60 + 170 * c, 70 + 150 * c, np.full_like(c, 255)], -1).astype(np.uint8)
    Image.fromarray(img, "RGBA").save(path)


MATERIALS = """
material white
{
	color 0 0.29 0.88
	roughness 0.0
}
material checker
{
	albedotexture checker.png
	roughness 0.5
}
material gold
{
	color 1.0 0.71 0.29
	metallic 1.0
	roughness 0.15
}
material glass
{
	color 1.0 1.0 1.0
	spectrans 1.0
	roughness 0.0
	ior 1.45
}
material coat
{
	color 0.8 0.05 0.05
	roughness 0.3
	clearcoat 1.0
	clearcoatgloss 0.9
}
material rough
{
	color 0.7 0.7 0.7
	roughness 0.9
}
material sheen
{
	color 0.2 0.5 0.3
	roughness 0.7
	sheen 1.0
	sheentint 0.5
	subsurface 0.4
}
material aniso
{
	color 0.9 0.9 0.9
	metallic 1.0
	roughness 0.3
	anisotropic 0.8
}
material mirror
{
	color 0.95 0.95 0.95
	metallic 1.0
	roughness 0.001
}
"""


def gen(tmp, n_inst):
    write_hdr(os.path.join(tmp, "sky.hdr"), sky(512, 256))
    write_sphere_obj(os.path.join(tmp, "sphere.obj"), 32, 20)
    write_floor_obj(os.path.join(tmp, "floor.obj"), 60.0, 30.0)
    write_checker_png(os.path.join(tmp, "checker.png"))
    with open(os.path.join(tmp, "ibl_spheres.scene"), "w") as f:
        f.write("renderer\n{\n\tresolution 1280 720\n\tmaxdepth 2\n\ttilewidth 320\n\ttileheight 180\n\tenvmapfile sky.hdr\n\tenvmapintensity 5.0\n\ttexarraywidth 256\n\ttexarrayheight 256\n}\n")
        f.write("camera\n{\n\tposition 9 5 0\n\tlookat 0 1.2 0\n\tfov 60\n}\n" + MATERIALS)
        f.write("mesh\n{\n\tfile sphere.obj\n\tmaterial white\n\tposition 0 1.5 -2.0\n\tscale 1.5 1.5 1.5\n}\n")
        f.write("mesh\n{\n\tfile sphere.obj\n\tmaterial gold\n\tposition 0 1.0 2.2\n}\n")
        f.write("mesh\n{\n\tfile floor.obj\n\tmaterial checker\n}\n")
    rng = np.random.default_rng(42)
    side = int(round(math.sqrt(n_inst)))
    mats = ["white", "gold", "glass", "coat", "rough", "sheen", "aniso", "mirror"]
    with open(os.path.join(tmp, "instancing.scene"), "w") as f:
        f.write("renderer\n{\n\tresolution 1280 720\n\tmaxdepth 8\n\ttilewidth 320\n\ttileheight 180\n\tenvmapfile sky.hdr\n\tenvmapintensity 2.0\n\ttexarraywidth 256\n\ttexarrayheight 256\n}\n")
        f.write(f"camera\n{{\n\tposition {side * 0.9:.3f} {side * 0.35:.3f} {side * 0.9:.3f}\n\tlookat 0 0 0\n\tfov 50\n}}\n" + MATERIALS)
        k = 0
        for iz in range(side):
            for ix in range(side):
                x = (ix - side / 2 + 0.5) * 1.6 + rng.uniform(-0.25, 0.25); z = (iz - side / 2 + 0.5) * 1.6 + rng.uniform(-0.25, 0.25)
                s = rng.uniform(0.35, 0.65); y = s + rng.uniform(0, 0.6)
                q = rng.normal(size=4); q /= np.linalg.norm(q)
                f.write(f"mesh\n{{\n\tfile sphere.obj\n\tmaterial {mats[k % 8]}\n\tposition {x:.4f} {y:.4f} {z:.4f}\n\tscale {s:.4f} {s * rng.uniform(0.7, 1.0):.4f} {s:.4f}\n"
                        f"\trotation {q[0]:.5f} {q[1]:.5f} {q[2]:.5f} {q[3]:.5f}\n}}\n")
                k += 1
        f.write(f"mesh\n{{\n\tfile floor.obj\n\tmaterial checker\n\tscale {side / 30:.4f} 1 {side / 30:.4f}\n}}\n")


def main():
    n_inst = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    dump = os.path.join(ROOT, "oracle", "_ref", "scene_dump")
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle", "ref_host")], stdout=subprocess.DEVNULL)
    out_dir = os.path.join(ROOT, "tests", "golden", "scenes")
    with tempfile.TemporaryDirectory() as tmp:
        gen(tmp, n_inst)
        for name in ("ibl_spheres", "instancing"):
            raw = os.path.join(tmp, name + ".ptscene")
            out = subprocess.check_output([dump, os.path.join(tmp, name + ".scene"), raw], text=True)
            print([l for l in out.splitlines() if l.startswith("PTSCENE")][0])
            with open(raw, "rb") as f, lzma.open(os.path.join(out_dir, name + ".ptscene.xz"), "wb", preset=6) as g:
                g.write(f.read())


if __name__ == "__main__":
    main()
