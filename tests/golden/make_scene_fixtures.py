#!/usr/bin/env python
"""Regenerates tests/golden/scenes/*.ptscene.xz from the reference's own assets with the UNMODIFIED reference
host code (oracle/ref_host -> oracle/_ref/scene_dump).  Runs only where /root/reference exists (the build
container); the committed .xz blobs are what travels to the GPU box.

The printed FNV-1a-64 of the flattened BVH must equal the values pinned in SURVEY.md §8(c) (checked in
tests/test_fixtures.py)."""
import lzma, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = os.environ.get("PTB_REFERENCE", "/root/reference")
SCENES = ["cornell_box_orig", "cornell_box_sphere", "hyperion_rect_lights", "hyperion_sphere_light", "volume_cube", "teapot"]


def main():
    subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle", "ref_host"), f"REF={REF}"])
    dump = os.path.join(ROOT, "oracle", "_ref", "scene_dump")
    raw_dir = os.path.join(ROOT, "oracle", "_ref", "scenes"); os.makedirs(raw_dir, exist_ok=True)
    out_dir = os.path.join(ROOT, "tests", "golden", "scenes"); os.makedirs(out_dir, exist_ok=True)
    for s in SCENES:
        raw = os.path.join(raw_dir, s + ".ptscene")
        out = subprocess.check_output([dump, os.path.join(REF, "assets", s + ".scene"), raw], text=True)
        print([l for l in out.splitlines() if l.startswith("PTSCENE")][0])
        with open(raw, "rb") as f, lzma.open(os.path.join(out_dir, s + ".ptscene.xz"), "wb", preset=6) as g:
            g.write(f.read())


if __name__ == "__main__":
    sys.exit(main())
