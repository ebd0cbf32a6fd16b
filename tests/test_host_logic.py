"""Host-side logic that needs no GPU: the synthetic workload generator, the Model mirror, the .off writer, view caches."""
import os

import numpy as np
import pytest

from conftest import GOLDEN


def test_pack_unpack_roundtrip_and_layout():
    from ar_voxel_project_b200.synth import pack_bits, unpack_bits
    rng = np.random.default_rng(0)
    for W in (1, 31, 32, 33, 100, 640):
        b = rng.random((3, 5, W)) < 0.5
        w = pack_bits(b)
        assert w.shape == (3, 5, (W + 31) // 32) and w.dtype == np.uint32
        assert np.array_equal(unpack_bits(w, W), b)
        x = W - 1
        assert ((w[1, 2, x >> 5] >> np.uint32(x & 31)) & 1) == b[1, 2, x]      # bit x&31 of word x>>5
        if W % 32:
            assert (w[..., -1] >> np.uint32(W % 32)).max() == 0                 # padding bits are 0


def test_synthetic_P_follows_the_cv_gemm_rule(oracle, golden):
    """P = K32 * M must be evaluated like cv::gemm does (f32, left to right, no FMA): same bits as the oracle's pinned rule"""
    from ar_voxel_project_b200.synth import cameras, gemm_k_m_f32
    K32, M, P = cameras(9, 640, 480)
    for m, p in zip(M, P):
        assert np.array_equal(oracle.gemm3x3_3x4(K32, m).view(np.uint32), p.view(np.uint32))
    g = golden("gemm_kat.npz")
    for K, m, km in zip(g["K"][:500], g["M"][:500], g["KM"][:500]):
        assert np.array_equal(gemm_k_m_f32(K, m).view(np.uint32), km.view(np.uint32))


def test_workload_is_deterministic_and_sane():
    from ar_voxel_project_b200.synth import Workload, unpack_bits
    a, b = Workload(64, 6, 160, 120, seed=3), Workload(64, 6, 160, 120, seed=3)
    assert np.array_equal(a.mask_bits, b.mask_bits) and np.array_equal(a.P, b.P)
    assert not np.array_equal(a.mask_bits, Workload(64, 6, 160, 120, seed=4).mask_bits)
    fg = 1.0 - unpack_bits(a.mask_bits, 160).mean()
    assert 0.01 < fg < 0.3                                   # the object is in view and small
    assert a.s == np.float32(0.28) / np.float32(64)
    bgr = a.mask_bgr()
    assert bgr.shape == (6, 120, 160, 3) and np.array_equal((bgr == 0).all(-1), unpack_bits(a.mask_bits, 160))
    img = a.images_bgr()
    assert img.shape == (6, 120, 160, 3) and img.dtype == np.uint8 and img.std() > 50


def test_model_mirror_follows_model_h():
    from ar_voxel_project_b200 import Model
    m = Model(4, 3, 2, 0.5)
    assert (m.getX(), m.getY(), m.getZ(), m.getSize()) == (4, 3, 2, np.float32(0.5))
    assert m.flatten(1, 2, 1) == 1 + 4 * (2 + 3 * 1)                                    # Model.h:104-106
    assert np.array_equal(m.toWord(1, 2, 1), np.array([1.0, 0.5, -0.5, 1.0], np.float32))  # (y*s, x*s, -z*s, 1) Model.h:134-136
    assert tuple(m.get(0, 0, 0)) == (50, 168, 141, 1) and tuple(m.get(-1, 0, 0)) == (0, 0, 0, 0) and tuple(m.get(0, 3, 0)) == (0, 0, 0, 0)
    assert not m.isInner(1, 1, 0)                                                        # z-1 is outside the grid -> alpha 0
    big = Model(3, 3, 3, 1.0)
    assert big.isInner(1, 1, 1)
    big.set(1, 1, 0, np.zeros(4, np.float32))
    assert not big.isInner(1, 1, 1)
    big.see(0, 0, 0)
    big.visit((2, 2, 2))
    assert big.visited((0, 0, 0)) and big.visited((2, 2, 2)) and not big.visited((1, 0, 0))
    big.handleUnseen()                                                                   # Model.cpp:36-47: unseen -> (204,0,0,1), stays solid
    assert tuple(big.get(1, 0, 0)) == (204, 0, 0, 1) and tuple(big.get(0, 0, 0)) == (50, 168, 141, 1) and tuple(big.get(1, 1, 0)) == (204, 0, 0, 1)
    for bad in ((0, 1, 1, 0.1), (1, 1, 1, 0.0), (1, 1, 1, -1.0)):
        with pytest.raises(ValueError):
            Model(*bad)                                                                  # main.cpp:232-246


def test_off_writer_equals_oracle_writer(oracle, tmp_path):
    """SimpleMesh::WriteMesh formatting (MarchingCubes.h:59-87): %g numbers, f32 scale-then-translate, unshared vertices"""
    from ar_voxel_project_b200.mesh import write_off
    rng = np.random.default_rng(1)
    verts = rng.integers(-1, 120, (57, 3, 3)).astype(np.float32)
    verts[3] += 0.5
    rgb = rng.integers(0, 256, (57, 3)).astype(np.uint32)
    for scale, t in ((np.float32(0.0028), (0.0, 0.0, 0.0)), (np.float32(1.5) * np.float32(0.0028), (0.5, -0.25, 2.0)), (np.float32(1e-7), (0.0, 1e6, 0.0))):
        a, b = str(tmp_path / "a.off"), str(tmp_path / "b.off")
        write_off(a, verts, rgb, scale, t)
        oracle.write_off(b, verts, rgb, scale, t)
        assert open(a).read() == open(b).read()
    lines = open(a).read().splitlines()
    assert lines[0] == "OFF" and lines[1] == "171 57 0" and lines[2 + 171].startswith("3 0 1 2 ")


def test_view_cache_loads_and_matches_calibration():
    from ar_voxel_project_b200.api import ViewSet
    vs = ViewSet.from_npz(os.path.join(GOLDEN, "box_views.npz"))
    assert (vs.V, vs.W, vs.H) == (8, 640, 480) and vs.images_bgr.shape == (8, 480, 640, 3)
    z = np.load(os.path.join(GOLDEN, "box_views.npz"))
    assert abs(float(z["K32"][0, 0]) - 496.50601) < 1e-3 and z["mask_bits"].shape == (8, 480, 20)   # cameracalibration.yml, 640/32 words
    with pytest.raises(ValueError):
        ViewSet(vs.P, vs.M, 640, 480, mask_bits=vs.mask_bits[:3])                        # main.cpp:228-231 count mismatch


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` runs on the host alone (no GPU): one JSON line with the contract's keys, the CPU restatement
    as the thing measured, zero transfer bytes; under torchrun every rank but 0 exits 0 without printing."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--config", "C3", "--steps", "1", "--warmup", "0",
           "--cpu-seconds", "0.2"]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "voxel_view_projections_per_s" and d["unit"] == "voxel-views/s"
    assert d["higher_is_better"] is True and d["vs_baseline"] is None and d["gpu_launches"] == 0
    assert d["config"]["workload"].startswith("synthetic 512x512x512 x 36 views")
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2")
    r = subprocess.run(cmd + ["--gpus", "2"], capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_output_buffers_are_never_copied_behind_the_callers_back():
    """engine._host_out_ptr: an output buffer must be usable as it stands (4-byte integers, C-contiguous, writeable); anything
    else raises instead of letting the library write into a temporary copy"""
    from ar_voxel_project_b200.engine import _host_out_ptr
    ok = np.zeros((4, 3, 2), np.uint32)
    assert _host_out_ptr(ok, ok.nbytes, "t") == ok.ctypes.data
    assert _host_out_ptr(ok.view(np.int32), ok.nbytes, "t") == ok.ctypes.data
    for bad in (np.zeros((4, 3, 2), np.int64), np.zeros((4, 3, 2), np.float32), np.zeros((4, 3, 4), np.uint32)[:, :, ::2],
                np.zeros((2, 3, 2), np.uint32)):
        with pytest.raises(ValueError):
            _host_out_ptr(bad, ok.nbytes, "t")
    ro = np.zeros((4, 3, 2), np.uint32)
    ro.flags.writeable = False
    with pytest.raises(ValueError):
        _host_out_ptr(ro, ok.nbytes, "t")
    with pytest.raises(TypeError):
        _host_out_ptr([0] * 24, ok.nbytes, "t")


def test_patch_pass_byte_tricks_match_the_per_brick_definition():
    """vc_patch4_planes (vc_kernels.cuh) decides for four bricks at once which words the blind fill left wrong: byte k of a 32-bit
    flag word is skipped iff brick k is carved AND seen (the blind pattern is right) or listed (the work items own it).  The two
    SWAR expressions are restated here and checked against the per-byte definition for every pair of adjacent flag bytes and a
    random sample of whole words (carries between bytes are what could go wrong)."""
    CARVED, SEEN, DECIDED, LISTED = 1, 2, 4, 8
    rng = np.random.default_rng(3)
    words = [a | (b << 8) | (c << 16) | (d << 24) for a in range(16) for b in range(16) for c in (0, 5, 15) for d in (0, 3, 15)]
    words += [int(x) for x in rng.integers(0, 16, size=(4000, 4), dtype=np.uint32) @ np.array([1, 1 << 8, 1 << 16, 1 << 24], dtype=np.uint64)]
    words += [w | 0xf0f0f0f0 for w in words[:500]]   # high nibbles are unused today: they must not leak into the low bits
    for f4 in words:
        f4 &= 0xffffffff
        cs = f4 & (f4 >> 1) & (CARVED * 0x01010101)
        skip = cs | ((f4 // LISTED) & 0x01010101)
        for k in range(4):
            b = (f4 >> (8 * k)) & 0xff
            want = ((b & (CARVED | SEEN)) == (CARVED | SEEN)) or bool(b & LISTED)
            assert ((skip >> (8 * k)) & 1) == int(want), (hex(f4), k)
            assert (skip >> (8 * k)) & 0xfe == 0
    # the sentinel for layers beyond the slab is skipped whole
    sf = DECIDED | CARVED | SEEN
    f4 = (sf & 0xff) * 0x01010101
    assert (f4 & (f4 >> 1) & 0x01010101) == 0x01010101
