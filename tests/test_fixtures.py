"""The scene fixtures are the output of the UNMODIFIED reference host code (oracle/ref_host); their flattened-BVH hashes must
equal the values pinned in SURVEY.md §8(c), and the array element sizes must match the reference structs."""
import os
import numpy as np
import pytest
from conftest import load_scene_cached
from glsl_pathtracer_b200 import scene_io

PINNED = {  # FNV-1a-64 over the raw bytes of bvhTranslator.nodes (g++ -O2, no FMA contraction)
    "cornell_box_orig": (41, 27, 0xF0664A6098E22B9E),
    "cornell_box_sphere": (14781, 14767, 0xBC1D6B3755866FC8),
    "hyperion_rect_lights": (100393, 100371, 0x92FC573B9AD0AC4E),
    "hyperion_sphere_light": (100393, 100371, 0x004B190F9B53FFDB),
    "volume_cube": (16, 12, 0x1193AFF733BF53F0),
    "teapot": (3, 1, 0x6FB4E90F01A5D9BA),
    # synthetic stand-ins (tests/golden/gen_synthetic.py, seed 42) built by the same unmodified reference host code; regression pins
    "ibl_spheres": (1190, 1184, 0x07664C04E651EC2E),
    "instancing": (21186, 1184, 0x4C5BB29432A17014),
    # glTF input (tests/golden/gen_gltf.py): a .gltf and the same model as .glb loaded by the reference's GLTFLoader.cpp through
    # `gltf {}` blocks of one .scene file
    "gltf_mix": (2732, 2708, 0x80FF272081464E8A),
}


@pytest.mark.parametrize("name", sorted(PINNED))
def test_flattened_bvh_hash_matches_survey(name):
    sc = load_scene_cached(name)
    n, top, h = PINNED[name]
    assert sc.nodes.shape == (n, 9) and sc.topLevelIndex == top
    assert scene_io.fnv1a64(sc.nodes.tobytes()) == h


def test_hyperion_counts_match_survey():
    sc = load_scene_cached("hyperion_rect_lights")
    assert len(sc.vertIndices) == 107534 and len(sc.verticesUVX) == 322602 and len(sc.transforms) == 11 and len(sc.lights) == 17
    assert sc.renderOptions.maxDepth == 3 and (sc.renderOptions.tileWidth, sc.renderOptions.tileHeight) == (256, 144)
    sp = load_scene_cached("hyperion_sphere_light")
    assert len(sp.lights) == 1 and sp.lights[0, 14] == 1.0 and sp.lights[0, 12] == 6.0
    assert np.array_equal(sc.verticesUVX, sp.verticesUVX)


def test_struct_sizes_match_reference():
    sc = load_scene_cached("cornell_box_orig")
    assert sc.nodes.dtype == np.float32 and sc.nodes.strides[0] == 36          # BvhTranslator::Node
    assert sc.materials.strides[0] == 128 and sc.lights.strides[0] == 60 and sc.transforms.strides[0] == 64 and sc.vertIndices.strides[0] == 12


def test_feature_derivation_matches_renderer_cpp():
    f = scene_io.derive_features
    S = scene_io
    assert f(load_scene_cached("cornell_box_orig")) == S.OPT_LIGHTS | S.OPT_RR | S.OPT_OPENGL_NORMALMAP
    assert f(load_scene_cached("volume_cube")) == S.OPT_LIGHTS | S.OPT_RR | S.OPT_OPENGL_NORMALMAP | S.OPT_ALPHA_TEST | S.OPT_MEDIUM | S.OPT_VOL_MIS
    # teapot.scene names an HDR that is missing from the checkout: enableEnvMap is set by the loader but scene->envMap is null
    t = load_scene_cached("teapot")
    assert t.renderOptions.enableEnvMap and t.envImg is None and not (f(t) & S.OPT_ENVMAP) and not (f(t) & S.OPT_LIGHTS)


def test_synthetic_env_map_cdf_matches_buildcdf():
    """EnvironmentMap::BuildCDF (EnvironmentMap.cpp:39-61): flat fp32 running sum of luminance; totalSum = last element."""
    sc = load_scene_cached("ibl_spheres")
    assert sc.envImg.shape == (256, 512, 3) and sc.envCdf.shape == (256, 512)
    w = (np.float32(0.212671) * sc.envImg[..., 0] + np.float32(0.715160) * sc.envImg[..., 1] + np.float32(0.072169) * sc.envImg[..., 2]).astype(np.float32).ravel()
    cdf = np.empty_like(w); acc = np.float32(0)
    for i in range(0, len(w), 1):      # sequential fp32 accumulation, as the reference loop
        acc = np.float32(acc + w[i]); cdf[i] = acc
        if i == 4096: break
    assert np.array_equal(cdf[:4097], sc.envCdf.ravel()[:4097])
    assert sc.envTotalSum == sc.envCdf.ravel()[-1] and sc.tlasHeight == 3
    inst = load_scene_cached("instancing")
    assert len(inst.transforms) == 10001 and inst.tlasHeight == 15 and inst.renderOptions.maxDepth == 8


def test_vert_indices_follow_scene_cpp_packing():
    """Scene.cpp:235-246: vertIndices = (tri*3+0, tri*3+1, tri*3+2) + mesh vertex base."""
    sc = load_scene_cached("hyperion_rect_lights")
    vi = sc.vertIndices
    assert np.array_equal(vi[:, 1], vi[:, 0] + 1) and np.array_equal(vi[:, 2], vi[:, 0] + 2) and (vi[:, 0] % 3 == 0).all()


def test_gltf_fixture_shows_the_loader_semantics():
    """GLTFLoader.cpp: one mesh instance per primitive (:44-46, 6 per load), sqrt(roughnessFactor) (:259), MASK/BLEND alpha modes,
    KHR_materials_transmission, texture ids offset by the textures already in the scene on the second load, including the
    reference's `normalTexture.index + sceneTexIdx` for materials WITHOUT a normal map (:264: -1 + 4 = texture 3)."""
    sc = load_scene_cached("gltf_mix")
    m = sc.materials
    assert len(sc.transforms) == 12 and len(m) == 13 and sc.textures.shape == (8, 128, 128, 4) and len(sc.lights) == 1
    first, second = m[1:7], m[7:13]                       # material 0 is the loader's default material
    assert np.allclose(first[0, 9], np.sqrt(0.5)) and first[1, 16] == 1.0 and first[2, 29] == 2 and first[4, 29] == 1 and first[4, 28] == np.float32(0.45)
    assert list(first[0, 24:27]) == [0, 1, 2] and list(second[0, 24:27]) == [4, 5, 6]
    assert list(first[:, 26]) == [2, -1, -1, -1, -1, -1] and list(second[:, 26]) == [6, 3, 3, 3, 3, 3]
    f = scene_io.derive_features(sc)
    assert f & scene_io.OPT_ALPHA_TEST and f & scene_io.OPT_ENVMAP and f & scene_io.OPT_LIGHTS


@pytest.mark.parametrize("name", ["cornell_box_orig", "hyperion_rect_lights"])
def test_instance_edit_fixture_is_a_tlas_only_change(name):
    """Scene::RebuildInstances (Scene.cpp:200-214) rebuilds only the TLAS slice: same node count, BLAS part untouched, one transform
    row changed; the material id of the edited instance travels in its TLAS leaf (bvh_translator.cpp:70-78)."""
    from conftest import edited_scene
    sc, sc2 = edited_scene(name, 64, 64)
    top = sc.topLevelIndex
    assert sc2.nodes.shape == sc.nodes.shape and np.array_equal(sc.nodes[:top], sc2.nodes[:top]) and not np.array_equal(sc.nodes[top:], sc2.nodes[top:])
    assert (sc.transforms.reshape(-1, 16) != sc2.transforms.reshape(-1, 16)).any(axis=1).sum() == 1
    fx = np.load(os.path.join(os.path.dirname(__file__), "golden", "instance_edit.npz"))
    k, mat = int(fx[name + "/edit"][0]), int(fx[name + "/edit"][5])
    leaves = sc2.nodes[top:][sc2.nodes[top:, 8] < 0]
    assert sorted((-leaves[:, 8] - 1).astype(int).tolist()) == list(range(len(sc.transforms.reshape(-1, 16))))
    if mat >= 0:
        assert int(leaves[(-leaves[:, 8] - 1).astype(int) == k][0, 7]) == mat
