"""The CPU oracle against every fixture that pins it (tests/golden, made by tools/make_goldens.py
from cv2 4.13.0 and the reference datasets). No GPU involved."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN


def _bits_equal(a, b):
    return np.array_equal(np.ascontiguousarray(a).view(np.uint32), np.ascontiguousarray(b).view(np.uint32))


def test_project_matches_cv2_gemm_bitwise(oracle, golden):
    """VoxelCarving.cpp:19 second product == cv2.gemm(3x4, 4x1) on 52 800 vectors, incl. the near-tie
    rows that separate ((s0+s1)+s2)+s3 from every other f64 association."""
    g = golden("gemm_kat.npz")
    bad = sum(not _bits_equal(oracle.project(P, w), r) for P, w, r in zip(g["P"], g["w"], g["proj"]))
    assert bad == 0


def test_near_tie_vectors_discriminate_association(golden):
    """the fixture is strong enough: the order SURVEY §8c-5 states fails on it, the sequential one does not"""
    g = golden("gemm_kat.npz")
    P, w, r = g["P"].astype(np.float64), g["w"].astype(np.float64), g["proj"]
    s = P * w[:, None, :]
    seq = (((s[..., 0] + s[..., 1]) + s[..., 2]) + s[..., 3]).astype(np.float32)
    alt = (s[..., 0] + ((s[..., 1] + s[..., 2]) + s[..., 3])).astype(np.float32)
    assert np.array_equal(seq.view(np.uint32), r.view(np.uint32))
    assert (alt.view(np.uint32) != r.view(np.uint32)).sum() > 100


def test_intrinsics_times_pose_matches_cv2_gemm(oracle, golden):
    g = golden("gemm_kat.npz")
    bad = sum(not _bits_equal(oracle.gemm3x3_3x4(K, M), KM) for K, M, KM in zip(g["K"], g["M"], g["KM"]))
    assert bad == 0


def test_cached_P_is_K32_times_M(oracle, golden):
    for ds in ("box", "human"):
        v = golden(f"{ds}_views.npz")
        for M, P in zip(v["M"], v["P"]):
            assert _bits_equal(oracle.gemm3x3_3x4(v["K32"], M), P)


@pytest.mark.parametrize("val,exp", [
    (0.5, 1), (1.5, 2), (2.5, 3), (-0.5, -1), (-1.5, -2), (0.49999997, 0), (-0.49999997, 0), (0.0, 0),
    (639.5, 640), (639.49994, 639), (8388609.0, 8388609), (2147483520.0, 2147483520),
    (float("nan"), -2 ** 31), (float("inf"), -2 ** 31), (float("-inf"), -2 ** 31), (3e9, -2 ** 31), (-3e9, -2 ** 31),
])
def test_round_half_away_and_overflow(oracle, val, exp):
    """(int)std::round(float) incl. the x86 INT_MIN result for NaN/inf/overflow (VoxelCarving.cpp:44)"""
    assert oracle.round_to_int(val) == exp


@pytest.mark.parametrize("ds", ["box", "human"])
def test_carve_matches_literal_cv2_run(oracle, golden, ds):
    v, L = golden(f"{ds}_views.npz"), golden(f"{ds}_literal.npz")
    X, Y, Z, s = int(L["X"]), int(L["Y"]), int(L["Z"]), L["s"]
    for nthreads in (1, 3):
        occ, seen = oracle.carve(X, Y, Z, s, v["P"], int(v["W"]), int(v["H"]), mask_bits=v["mask_bits"], nthreads=nthreads)
        assert np.array_equal(oracle.unpack(occ, X), L["occ"])
        assert np.array_equal(oracle.unpack(seen, X), L["seen"])
    # z-slabs of the oracle tile the full result
    parts = [oracle.carve(X, Y, Z, s, v["P"], int(v["W"]), int(v["H"]), mask_bits=v["mask_bits"], z0=a, z1=b)
             for a, b in ((0, 5), (5, 6), (6, Z))]
    assert np.array_equal(np.concatenate([p[0] for p in parts]), occ)
    assert np.array_equal(np.concatenate([p[1] for p in parts]), seen)


@pytest.mark.parametrize("ds", ["box", "human"])
def test_bgr_masks_equal_bit_masks(oracle, golden, ds):
    """mask test `pixel == (0,0,0)` (VoxelCarving.cpp:50): near-black pixels must NOT carve"""
    from ar_voxel_project_b200.synth import unpack_bits
    v, L = golden(f"{ds}_views.npz"), golden(f"{ds}_literal.npz")
    X, Y, Z, s, W, H = int(L["X"]), int(L["Y"]), int(L["Z"]), L["s"], int(v["W"]), int(v["H"])
    bg = unpack_bits(v["mask_bits"], W)
    rng = np.random.default_rng(3)
    bgr = rng.integers(1, 6, size=(*bg.shape, 3), dtype=np.uint8)  # "almost black" foreground
    bgr[..., rng.integers(0, 3)] = 0
    bgr[bg] = 0
    occ, seen = oracle.carve(X, Y, Z, s, v["P"], W, H, mask_bgr=bgr)
    assert np.array_equal(oracle.unpack(occ, X), L["occ"]) and np.array_equal(oracle.unpack(seen, X), L["seen"])


@pytest.mark.parametrize("ds", ["box", "human"])
def test_colour_matches_literal_cv2_run(oracle, golden, ds):
    from ar_voxel_project_b200.api import ViewSet
    vs = ViewSet.from_npz(os.path.join(GOLDEN, f"{ds}_views.npz"))
    L = golden(f"{ds}_literal.npz")
    X, Y, Z, s = int(L["X"]), int(L["Y"]), int(L["Z"]), L["s"]
    occ, _ = oracle.carve(X, Y, Z, s, vs.P, vs.W, vs.H, mask_bits=vs.mask_bits)
    sf = L["surf"].astype(np.int64)
    fl = sf[:, 0] + X * (sf[:, 1] + Y * sf[:, 2])
    order = np.argsort(fl)
    for mode, key in ((1, "closest"), (2, "avg")):
        idx, rgbn = oracle.color(X, Y, Z, s, vs.P, vs.M, vs.W, vs.H, vs.images_bgr, occ, mode)
        assert np.array_equal(idx, fl[order].astype(np.uint64))
        assert np.array_equal(rgbn[:, :3], L[key][order])
        assert np.array_equal(rgbn[:, 3], np.minimum(L["nobs"][order], 255))


def test_fast_carve_is_flood_of_carved_set_from_origin(oracle, golden):
    """closed form of VoxelCarving.cpp:100-164: carved2 = 6-connected component of the
    method-1 carved set containing (0,0,0); seen2 = carved2 plus its popped, uncarved neighbours."""
    from scipy import ndimage
    for ds in ("box", "human"):
        v, L = golden(f"{ds}_views.npz"), golden(f"{ds}_literal.npz")
        X, Y, Z, s = int(L["X"]), int(L["Y"]), int(L["Z"]), L["s"]
        occ2, seen2 = oracle.fast_carve(X, Y, Z, s, v["P"], int(v["W"]), int(v["H"]), mask_bits=v["mask_bits"])
        carved1 = ~L["occ"]
        lab, _ = ndimage.label(carved1)  # default structure = 6-connectivity
        comp = (lab == lab[0, 0, 0]) & carved1 if carved1[0, 0, 0] else np.zeros_like(carved1)
        assert np.array_equal(~oracle.unpack(occ2, X), comp)
        grown = ndimage.binary_dilation(comp) if comp.any() else comp
        exp_seen = grown.copy()
        exp_seen[0, 0, 0] = True
        assert np.array_equal(oracle.unpack(seen2, X), exp_seen)


def test_mc_tables_and_classify_small(oracle):
    tc = oracle.tri_counts()
    assert tc.sum() == 820 and np.bincount(tc, minlength=6).tolist() == [2, 16, 50, 80, 76, 32]
    # testMarchingCubes() model (MarchingCubes.cpp:38-74): 4x4x4 solid minus the four corner columns
    from ar_voxel_project_b200.synth import pack_bits
    occ = np.ones((4, 4, 4), bool)
    for x in (0, 3):
        for y in (0, 3):
            occ[:, y, x] = False
    hist, na, nt = oracle.mc_classify(4, 4, 4, pack_bits(occ))
    assert hist.sum() == 125 and hist[0] == 0 + (occ[:-1, :-1, :-1] & occ[1:, :-1, :-1] & occ[:-1, 1:, :-1] & occ[1:, 1:, :-1]
                                                  & occ[:-1, :-1, 1:] & occ[1:, :-1, 1:] & occ[:-1, 1:, 1:] & occ[1:, 1:, 1:]).sum()
    # brute force per cell in python
    def get(x, y, z):
        return 0 <= x < 4 and 0 <= y < 4 and 0 <= z < 4 and occ[z, y, x]
    h2 = np.zeros(256, np.uint64)
    for x in range(-1, 4):
        for y in range(-1, 4):
            for z in range(-1, 4):
                c = [get(x + 1, y, z), get(x, y, z), get(x, y + 1, z), get(x + 1, y + 1, z),
                     get(x + 1, y, z + 1), get(x, y, z + 1), get(x, y + 1, z + 1), get(x + 1, y + 1, z + 1)]
                h2[sum((not b) << i for i, b in enumerate(c))] += 1
    assert np.array_equal(hist, h2)
    assert nt == int((h2 * tc).sum()) and na == int(h2[1:255].sum())


def test_soft_golden_box_mesh(oracle, golden):
    """Data/box_dataset/generated_models/1.off (16 352 faces) pins the pipeline only softly: poses come
    from a different OpenCV build.  carve -> closure(=3x3x3 dilation) -> MC triangle count
    must land within 0.5 %."""
    from scipy import ndimage
    soft = json.load(open(os.path.join(GOLDEN, "soft_box_1off.json")))
    v = golden("box_views.npz")
    occ, seen = oracle.carve(100, 100, 50, np.float32(0.0028), v["P"], int(v["W"]), int(v["H"]), mask_bits=v["mask_bits"], nthreads=0)
    o = oracle.unpack(occ, 100)
    closed = ndimage.binary_dilation(o, structure=np.ones((3, 3, 3), bool))  # Postprocessing3d.cpp:20-58
    from ar_voxel_project_b200.synth import pack_bits
    _, _, ntris = oracle.mc_classify(100, 100, 50, pack_bits(closed))
    assert abs(ntris - soft["faces"]) / soft["faces"] < 0.005, ntris


def test_full_pipeline_against_soft_golden_mesh(oracle, golden, tmp_path):
    """carve -> handleUnseen -> applyClosure(3) -> marchingCubes -> WriteMesh on box_dataset at 100x100x50 (the command behind
    generated_models/1.off). Soft: the golden's poses come from another OpenCV build. File layout, number formatting and the
    first lines are identical; >= 95 % of the golden's vertex multiset is reproduced; all faces carry MODEL_COLOR."""
    v, g = golden("box_views.npz"), golden("soft_box_1off_vertices.npz")
    s = np.float32(0.0028)
    occ, seen = oracle.carve(100, 100, 50, s, v["P"], int(v["W"]), int(v["H"]), mask_bits=v["mask_bits"], nthreads=0)
    m = oracle.closure(100, 100, 50, oracle.dense_model(100, 100, 50, occ, seen), 3)
    verts, rgb = oracle.marching_cubes(100, 100, 50, m, 0.5)
    out = str(tmp_path / "mesh.off")
    oracle.write_off(out, verts, rgb, np.float32(1.0) * s)
    lines = open(out).read().splitlines()
    assert lines[0] == "OFF" and lines[1].split()[2] == "0" and int(lines[1].split()[0]) == 3 * int(lines[1].split()[1])
    assert lines[2:12] == [str(x) for x in g["first_lines"][2:12]]          # same first triangles, same %g formatting
    assert {tuple(c) for c in rgb.tolist()} == {tuple(c) for c in g["face_colors"].tolist()} == {(50, 168, 141)}
    ours, cnt = np.unique(verts.reshape(-1, 3).astype(np.int16), axis=0, return_counts=True)
    gold = {tuple(k): int(c) for k, c in zip(g["uniq"].tolist(), g["counts"].tolist())}
    common = sum(min(c, gold.get(tuple(k), 0)) for k, c in zip(ours.tolist(), cnt.tolist()))
    assert common / g["counts"].sum() >= 0.95, common / g["counts"].sum()


def test_closure_rejects_even_kernel_and_dilates(oracle):
    rgba = np.zeros((5 * 4 * 3, 4), np.float32)
    rgba[1 + 5 * (1 + 4 * 1)] = (10, 20, 30, 1)
    rgba[2 + 5 * (1 + 4 * 1)] = (50, 60, 70, 1)
    with pytest.raises(ValueError):
        oracle.closure(5, 4, 3, rgba, 2)
    out = oracle.closure(5, 4, 3, rgba, 3)
    assert (out[:, 3] > 0).sum() == 4 * 3 * 3            # x in 0..3, all y in 0..2 (y = 3 is 2 away), all z
    assert np.array_equal(out[0 + 5 * (0 + 4 * 0)], (10, 20, 30, 1))          # only the first voxel in reach
    assert np.array_equal(out[1 + 5 * (0 + 4 * 0)], (30, 40, 50, 1))          # mean of both


def test_undistort_matches_cv2_kat(oracle, golden):
    """cv::undistort (VoxelCarving.cpp:36): the oracle restatement against cv2 4.13.0 outputs, byte for byte — dataset images
    and masks, a strongly distorted random image, the 8-coefficient rational model with a skewed K, a tall 33-px-wide image."""
    import cv2  # PNG decode of the fixture only
    z = golden("undistort_kat.npz")
    for n in z["names"]:
        src, dst = cv2.imdecode(z[f"{n}_src"], 1), cv2.imdecode(z[f"{n}_dst"], 1)
        assert np.array_equal(oracle.undistort(src, z[f"{n}_K"], z[f"{n}_dist"]), dst), n
    with pytest.raises(ValueError):
        oracle.undistort(src, z[f"{n}_K"], [0.1, 0.2, 0.3])
