"""The header-only C++ host layer (include/voxcarve_host.hpp): compiles with plain g++ (no OpenCV/Eigen), and on a
GPU runs the main.cpp:248-303 sequence on a stand-in Model and matches the oracle."""
import os
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT


def _build(tmp_path, lib_built, name="host_roundtrip"):
    exe = str(tmp_path / name)
    libdir = os.path.dirname(lib_built)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-Wall", "-pthread", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", name + ".cpp"), "-o", exe,
                           "-L", libdir, "-lvoxcarve", f"-Wl,-rpath,{libdir}"])
    return exe


def test_host_layer_compiles_without_opencv(tmp_path, lib_built):
    assert os.path.exists(_build(tmp_path, lib_built))
    assert os.path.exists(_build(tmp_path, lib_built, "multi_gpu"))
    # the OpenCV shim must at least be syntactically inert where OpenCV is absent
    subprocess.check_call(["g++", "-std=c++17", "-fsyntax-only", "-I", os.path.join(ROOT, "include"), "-x", "c++",
                           os.path.join(ROOT, "include", "voxcarve_shim.hpp")])


@pytest.mark.parametrize("replace_post", [False, True])
def test_opencv_shim_compiles_against_the_reference_interface(replace_post):
    """include/voxcarve_shim.hpp is the file a maintainer drops into the reference; OpenCV and Eigen are absent here, so it is
    type-checked against interface-only stand-ins of the headers it includes (tests/cpp/shim_stubs: the reference's Model,
    Benchmark, estimatePoseFromImage, marchingCubes / applyClosure prototypes; cv::Mat, Eigen vectors) from a caller written
    like main.cpp:260-303."""
    stubs = os.path.join(ROOT, "tests", "cpp", "shim_stubs")
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-I", stubs, "-I", os.path.join(ROOT, "include")]
    if replace_post:
        cmd.append("-DVOXCARVE_SHIM_REPLACE_POSTPROCESSING")
    subprocess.check_call(cmd + [os.path.join(ROOT, "tests", "cpp", "shim_user.cpp")])


@pytest.mark.gpu
def test_cpp_pipeline_matches_oracle(tmp_path, lib_built, oracle):
    from ar_voxel_project_b200.api import ViewSet
    vs = ViewSet.from_npz(os.path.join(GOLDEN, "box_views.npz"))
    X, Y, Z, s = 37, 30, 21, np.float32(0.008)
    case = tmp_path / "case.bin"
    with open(case, "wb") as f:
        np.array([X, Y, Z, vs.V, vs.W, vs.H], np.int32).tofile(f)
        np.array([s], np.float32).tofile(f)
        vs.P.astype(np.float32).tofile(f)
        vs.M.astype(np.float32).tofile(f)
        vs.mask_bits.astype(np.uint32).tofile(f)
        vs.images_bgr.astype(np.uint8).tofile(f)
    out = tmp_path / "out.bin"
    off = tmp_path / "mesh.off"
    inter = tmp_path / "intermediate"
    inter.mkdir()
    subprocess.check_call([_build(tmp_path, lib_built), str(case), str(out), str(off), str(inter)], stdout=subprocess.DEVNULL)
    raw = open(out, "rb").read()
    n = X * Y * Z
    vox = np.frombuffer(raw, np.float32, n * 4).reshape(n, 4)
    seen = np.frombuffer(raw, np.uint8, n, n * 16).astype(bool)
    hist = np.frombuffer(raw, np.uint64, 256, n * 17)
    ntris = int(np.frombuffer(raw, np.uint64, 1, n * 17 + 2048)[0])
    nv = int(np.frombuffer(raw, np.uint64, 1, n * 17 + 2056)[0])
    per_view = np.frombuffer(raw, np.uint64, nv, n * 17 + 2064)
    ro, rs = oracle.carve(X, Y, Z, s, vs.P, vs.W, vs.H, mask_bits=vs.mask_bits)
    occ = oracle.unpack(ro, X).reshape(-1)
    assert np.array_equal(vox[:, 3] != 0, occ) and np.array_equal(seen, oracle.unpack(rs, X).reshape(-1))
    idx, rgbn = oracle.color(X, Y, Z, s, vs.P, vs.M, vs.W, vs.H, vs.images_bgr, ro, 2)
    exp = np.tile(np.array([50, 168, 141], np.float32), (n, 1))
    exp[~occ] = 0
    m = rgbn[:, 3] > 0
    exp[idx[m].astype(np.int64)] = rgbn[m, :3]
    exp[~oracle.unpack(rs, X).reshape(-1)] = (204, 0, 0)  # handleUnseen (Model.cpp:36-47)
    assert np.array_equal(vox[:, :3], exp)
    rh, _, rnt = oracle.mc_classify(X, Y, Z, ro)
    assert np.array_equal(hist, rh) and ntris == rnt
    assert nv == vs.V and per_view[-1] == rnt  # -intermediateMesh: the last per-view summary is the final one
    # -intermediateMesh (VoxelCarving.cpp:65-68): after view i, marchingCubes(&model, 1.0f, (i*(X+2)*size, 0, 0), 0.5f,
    # "out/intermediate/image_<i>_mesh.off") of the Model carved by views 0..i - text-identical to the oracle's
    acc = np.full_like(ro, 0xffffffff)
    for i in range(vs.V):
        oi, _ = oracle.carve(X, Y, Z, s, vs.P[i:i + 1], vs.W, vs.H, mask_bits=vs.mask_bits[i:i + 1])
        acc &= oi
        mv, mc = oracle.marching_cubes(X, Y, Z, oracle.dense_model(X, Y, Z, acc), 0.5)
        assert len(mv) == per_view[i]
        ref_i = tmp_path / f"ref_{i}.off"
        oracle.write_off(str(ref_i), mv, mc, np.float32(1.0) * s, (np.float32(i * (X + 2)) * s, 0.0, 0.0))
        assert open(inter / f"image_{i}_mesh.off").read() == open(ref_i).read(), i
    # applyClosure(&model, 3) + marchingCubes(&model, 1.5, (0.5,-0.25,2), 0.5, file): the .off text equals the oracle's
    dense = exp.copy()
    full = np.concatenate([dense, (vox[:, 3:4] != 0).astype(np.float32)], axis=1)
    closed = oracle.closure(X, Y, Z, full, 3)
    rv, rc = oracle.marching_cubes(X, Y, Z, closed, 0.5)
    ref = tmp_path / "ref.off"
    oracle.write_off(str(ref), rv, rc, np.float32(1.5) * s, (0.5, -0.25, 2.0))
    assert open(off).read() == open(ref).read()


@pytest.mark.gpu
def test_cpp_host_drives_several_engines(tmp_path, lib_built, oracle):
    """a C++ host (no Python, no torch in the process) runs more than one engine: one thread + one NCCL rank per
    visible GPU through vc_comm_init / vc_exchange_halos / vc_comm_allreduce_u64 / vc_gather when there are >= 2 GPUs, several
    engines on device 0 with vc_exchange_halos_peer otherwise; slab results, summed histograms, concatenated colour records, the
    gathered grid and vc::carveOnDevices' Model all equal the single engine's - which equals the oracle's."""
    import torch
    from ar_voxel_project_b200.synth import Workload
    X, Y, Z = 90, 50, 70
    w = Workload(90, 7, 256, 192, seed=9, dims=(X, Y, Z))
    case = tmp_path / "case.bin"
    with open(case, "wb") as f:
        np.array([X, Y, Z, w.V, w.W, w.H], np.int32).tofile(f)
        np.array([w.s], np.float32).tofile(f)
        w.P.astype(np.float32).tofile(f)
        w.M.astype(np.float32).tofile(f)
        w.mask_bits.astype(np.uint32).tofile(f)
        w.images_bgr().astype(np.uint8).tofile(f)
    n_dev = torch.cuda.device_count()
    n_slabs = min(n_dev, 4) if n_dev >= 2 else 3
    out = tmp_path / "out.bin"
    r = subprocess.run([_build(tmp_path, lib_built, "multi_gpu"), str(case), str(n_dev), str(n_slabs), str(out)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    raw = open(out, "rb").read()
    n = Z * Y * ((X + 31) // 32)
    occ = np.frombuffer(raw, np.uint32, n).reshape(Z, Y, -1)
    seen = np.frombuffer(raw, np.uint32, n, n * 4).reshape(Z, Y, -1)
    hist = np.frombuffer(raw, np.uint64, 256, n * 8)
    ro, rs = oracle.carve(X, Y, Z, w.s, w.P, w.W, w.H, mask_bits=w.mask_bits)
    assert np.array_equal(occ, ro) and np.array_equal(seen, rs)
    assert np.array_equal(hist, oracle.mc_classify(X, Y, Z, ro)[0])
