#!/usr/bin/env python3
"""bench.py — voxel-view projections/s of the carve hot path on N B200s (z-slab sharded), next to the
reference's CPU carve (the oracle restatement) timed on the same box.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C4|C3|C5] [--impl ours|reference]

Step = one full carve (Model ctor state -> occupied + seen volumes complete) of the synthetic workload
BASELINE.json quotes the metric on (default C4: 1024^3 x 72 views, 1920x1080 silhouettes).
`value`  = X*Y*Z*V (nominal voxel-view projections of the job) / device time, masks resident in HBM.
`e2e`    = same metric through the C ABI with HOST buffers: per step H2D of P/M + the 8UC3 masks from pinned
           memory, reset, carve, D2H of both bit volumes into pinned memory; wall clock, max over ranks.
`roofline` = projections executed by the dominant kernel (vc_carve_bricks: the corner projections of its 8x8x8 sub-brick
           classification + the per-voxel projections left after three levels of classification and the early exits)
           * 23 FLOP / that kernel's time against the measured FFMA peak of this GPU (CUDA-core bound);
           `roofline_hbm` = grid-write bound of the whole carve against the copy bandwidth of MEASURED_PEAKS.json.
`cpu_baseline` / --impl reference = oracle/ (C restatement of VoxelCarving.cpp:60-72) on a bounded z-slab sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

# stdout carries exactly ONE line, the JSON result: library chatter written to file descriptor 1 (NCCL prints its version banner
# there) is sent to stderr instead, and emit() writes the result to the real stdout
_RESULT_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(obj):
    _RESULT_OUT.write(json.dumps(obj) + "\n")
    _RESULT_OUT.flush()


import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "voxel_view_projections_per_s"
UNIT = "voxel-views/s"
F_ALG = 23.0  # FLOP per voxel-view as the reference evaluates it: 12 mul + 9 add + 2 div


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="C4", choices=["C3", "C4", "C5"])
    ap.add_argument("--cpu-seconds", type=float, default=8.0, help="target seconds of CPU work per reference step")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-color", action="store_true", help="skip the colour-reconstruction timing (uploads V x H x W x 3 image bytes)")
    ap.add_argument("--uniform-slabs", action="store_true", help="equal-height z-slabs instead of the planner's balanced ones")
    return ap.parse_args()


def config_dict(w, n_gpus):
    return {"workload": w.name, "grid": [w.X, w.Y, w.Z], "views": w.V, "image": [w.W, w.H],
            "voxel_size": float(w.s), "partition": f"z-slabs x{n_gpus}", "arithmetic": "exact (bit-identical to reference)"}


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])), mx.append(float(f[1]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = sorted(sm)[len(sm) // 2:] if sm else []
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------- CPU reference
def cpu_reference(w, seconds, steps, warmup):
    """oracle.carve on a contiguous z-slab of the same workload, all host threads. -> (vv/s, cores, sample text, ms/step)"""
    from oracle import oracle as O
    nthr = O.max_threads()
    z_mid = w.Z // 2

    def run(n):
        z0 = max(0, min(w.Z - n, z_mid - n // 2))
        t = time.perf_counter()
        O.carve(w.X, w.Y, w.Z, w.s, w.P, w.W, w.H, mask_bits=w.mask_bits, z0=z0, z1=z0 + n, nthreads=nthr)
        return time.perf_counter() - t, z0

    n = min(w.Z, nthr)
    t1, _ = run(n)  # calibration (also warms the page cache / threads)
    n = int(max(n, min(w.Z, round(n * seconds / max(t1, 1e-3)))))
    n = max(nthr, (n // nthr) * nthr) if n >= nthr else n
    n = min(n, w.Z)
    for _ in range(max(0, warmup - 1)):
        run(n)
    ts = [run(n) for _ in range(steps)]
    t = float(np.mean([a for a, _ in ts]))
    vv = w.X * w.Y * n * w.V
    sample = (f"{n} contiguous z-planes [{ts[0][1]},{ts[0][1] + n}) of {w.Z} ({vv / 1e9:.2f} G of {w.X * w.Y * w.Z * w.V / 1e9:.1f} G voxel-views), "
              f"{nthr} pthreads, gcc -O2 -ffp-contract=off, bit-packed masks; CPU cost per voxel-view is position-independent")
    return vv / t, nthr, sample, t * 1e3


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from ar_voxel_project_b200.synth import Workload, CONFIGS
    w = Workload(**CONFIGS[args.config])
    v, cores, sample, ms = cpu_reference(w, args.cpu_seconds, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic", "config": config_dict(w, args.gpus),
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
            "note": "reference C++ cannot be built here (needs OpenCV C++ + Eigen); this is oracle/, its pinned C restatement"}
    emit(line)
    return 0


# --------------------------------------------------------------------------------- ours
def run_ours(args):
    import torch
    import torch.distributed as dist
    import ar_voxel_project_b200 as A
    from ar_voxel_project_b200.engine import measure_peaks
    from ar_voxel_project_b200.synth import Workload, CONFIGS

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the engine has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    # a non-default stream: its handle is what the engine launches on, so torch events bracket our kernels
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    def allgather(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world == 1:
            return [float(x)]
        out = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(out, t)
        return [float(o.item()) for o in out]

    w = Workload(**CONFIGS[args.config])
    X, Y, Z, V = w.X, w.Y, w.Z, w.V
    Wx = (X + 31) // 32
    nominal_total = X * Y * Z * V

    # whole-grid device buffers (torch = plumbing): slabs are carved in place, gathered in place
    occ_full = torch.empty((Z, Y, Wx), dtype=torch.int32, device=dev)
    seen_full = torch.empty_like(occ_full)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    # every rank holds all views; z-slab boundaries come from the engine's planner (super-brick classification of the
    # whole grid, identical on every rank) so that the per-voxel work, not the plane count, is balanced
    eng = A.VoxelEngine(X, Y, Z, w.s, device=local)
    eng.set_stream(torch.cuda.current_stream().cuda_stream)
    eng.set_views(w.P, w.W, w.H, w.M)
    eng.set_masks_bits(w.mask_bits)
    bounds = eng.plan_slabs(world) if (world > 1 and not args.uniform_slabs) else [(r * Z) // world for r in range(world)] + [Z]
    z0, z1 = bounds[rank], bounds[rank + 1]
    if world > 1:
        eng.set_slab(z0, z1)
    eng.bind_volumes(occ_full.data_ptr(), seen_full.data_ptr())

    def step():
        eng.reset()
        eng.carve(A._lib.VC_EXACT)

    clocks = ClockSampler(local)  # spans warm-up, the timed steps and an identical untimed tail (nvidia-smi period: 100 ms)
    t_clk = time.perf_counter()
    for _ in range(args.warmup):
        flush.fill_(1)
        step()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    kernel_ms = []
    barrier()
    for a, b in ev:
        flush.fill_(1)  # L2 flush between timed iterations (untimed)
        a.record()
        step()
        b.record()
        kernel_ms.append(None)
    barrier()
    n_tail = 0
    while time.perf_counter() - t_clk < 2.0:  # keep the same load up until the sampler has ~20 samples
        flush.fill_(1)
        step()
        n_tail += 1
    torch.cuda.synchronize()
    clk = clocks.stop()
    clk["sampled_over"] = f"{args.warmup} warm-up + {args.steps} timed + {n_tail} identical untimed steps"
    step_ms = [a.elapsed_time(b) for a, b in ev]
    ms_per_step = allmax(float(np.mean(step_ms)))
    launches = 3 * args.steps  # vc_brick_classify_kernel<1>, <0>, vc_carve_bricks (its first blocks do the fill pass, which absorbs the reset), per step

    # carve-kernel-only time (events inside vc_carve) and executed voxel-views (separate, untimed counting pass)
    kt, ct = [], []
    for _ in range(5):
        flush.fill_(1)
        step()
        st = eng.stats()
        kt.append(st["last_carve_ms"]), ct.append(st["last_classify_ms"])
    kernel_ms = float(np.mean(kt))
    classify_ms = float(np.mean(ct))
    eng.reset()
    eng.carve(A._lib.VC_EXACT, count_executed=True)
    st = eng.stats()
    executed_total = allsum(float(st["executed_voxel_views"]))
    corner_total = allsum(float(st["brick_corner_views"]))
    sub_corner_total = allsum(float(st["subbrick_corner_views"]))
    filter_rows, filter_slow, filter_bad = (allsum(float(st[k])) for k in ("filter_rows", "filter_slow_rows", "filter_mismatches"))
    bricks_listed = allsum(float(st["bricks_listed"]))
    bricks_total = allsum(float(st["bricks_total"]))
    # the same job without brick classification (VC_EXACT_FLAT), for reference
    flat = []
    for _ in range(2):
        flush.fill_(1)
        eng.reset()
        eng.carve(A._lib.VC_EXACT_FLAT)
        flat.append(eng.stats()["last_carve_ms"])
    flat_ms = allmax(float(flat[-1]))
    n_occ, n_seen = eng.count_occupied()
    occupied_total = allsum(float(n_occ))
    kernel_ms_max = allmax(kernel_ms)
    fine_ms_max = allmax(kernel_ms - classify_ms)
    classify_ms_max = allmax(classify_ms)
    per_rank = {"carve_ms": allgather(kernel_ms), "classify_ms": allgather(classify_ms),
                "per_voxel_projections": allgather(float(st["executed_voxel_views"]) - float(st["brick_corner_views"])),
                "bricks_listed": allgather(float(st["bricks_listed"]))}

    # on-demand assembly of the bit-packed grid (NCCL all-gather over NVLink), timed separately
    allgather_ms = None
    if world > 1:
        from ar_voxel_project_b200.dist import all_gather_slabs
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for it in range(3):
            barrier()
            g0.record()
            all_gather_slabs(occ_full, Z, world, bounds=bounds)
            all_gather_slabs(seen_full, Z, world, bounds=bounds)
            g1.record()
            torch.cuda.synchronize()
        allgather_ms = allmax(g0.elapsed_time(g1))
        eng.set_gathered(True)
    mc_ms = None
    if world == 1 or allgather_ms is not None:
        m0, m1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        eng.mc_classify()
        m0.record()
        eng.mc_classify()
        m1.record()
        torch.cuda.synchronize()
        mc_ms = allmax(m0.elapsed_time(m1))
        hist, na, nt = eng.download_mc()
        mc_tris = allsum(float(nt))
    else:
        mc_tris = None
    # per-surface-voxel colouring of the carved grid (ColorReconstruction.cpp:22-70) on the same device-resident volume;
    # the 8UC3 images (V x H x W x 3, hash-coloured) are uploaded outside the timed region
    color_ms = None
    if (world == 1 or allgather_ms is not None) and not args.no_color:
        eng.set_images(w.images_bgr())
        color_ms = {}
        for name, mode in (("average", A._lib.VC_COLOR_AVG), ("closest", A._lib.VC_COLOR_CLOSEST)):
            eng.color(mode)
            torch.cuda.synchronize()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            eng.color(mode)
            c1.record()
            torch.cuda.synchronize()
            color_ms[name] = allmax(c0.elapsed_time(c1))
        color_ms["surface_voxels"] = allsum(float(eng.surface_count()))

    # end to end through the C ABI with host buffers: (a) the cached bit-packed silhouettes (SURVEY §7-1 cache format,
    # what the CPU arm consumes too) and (b) the 8UC3 undistorted masks as the reference holds them in memory
    e2e = None
    e2e_bgr = None
    if not args.no_e2e:
        out_occ = torch.empty((z1 - z0, Y, Wx), dtype=torch.int32).pin_memory()
        out_seen = torch.empty_like(out_occ).pin_memory()
        d2h = 2 * out_occ.numel() * 4
        bits_pinned = torch.from_numpy(w.mask_bits.view(np.int32)).pin_memory()
        bgr = torch.from_numpy(w.mask_bgr()).pin_memory()

        def e2e_run(use_bits):
            def one():
                eng.set_views(w.P, w.W, w.H, w.M)
                if use_bits:
                    eng.set_masks_bits_async(bits_pinned)
                else:
                    eng.set_masks_bgr(bgr, sync=False)
                eng.reset()
                eng.carve_download(out_occ, out_seen)   # carve + D2H of both volumes, z-chunks overlapped
            for _ in range(2):
                one()
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                one()
            barrier()
            return allmax((time.perf_counter() - t0) / args.steps)

        s_bits = e2e_run(True)
        s_bgr = e2e_run(False)
        # kernels per e2e step: [pack_bgr] + 3 SAT passes + 4 z-chunks x (2 classify + fill + carve_bricks)
        e2e = {"value": nominal_total / s_bits, "unit": UNIT, "h2d_bytes_per_step": int(bits_pinned.numel() * 4 + w.P.nbytes + w.M.nbytes),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": s_bits * 1e3,
               "input": "cached bit-packed undistorted silhouettes (VC_MASK_BITS) + P/M, pinned host memory",
               "call": "vc_set_views + vc_set_masks + vc_reset + vc_carve_download", "gpu_launches": 15 * args.steps}
        e2e_bgr = {"value": nominal_total / s_bgr, "unit": UNIT, "h2d_bytes_per_step": int(bgr.numel() + w.P.nbytes + w.M.nbytes),
                   "d2h_bytes_per_step": int(d2h), "ms_per_step": s_bgr * 1e3,
                   "input": "8UC3 undistorted masks (VC_MASK_BGR8), packed on device", "gpu_launches": 16 * args.steps}
        eng.set_masks_bits(w.mask_bits)

    out = None
    if rank == 0:
        ffma, dfma = measure_peaks(local)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        hbm_src = "MEASURED_PEAKS.json hbm_gbs" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            key = f"{args.config}_x{world}"
            if key in tr:
                traffic = tr[key]["vc_carve_bricks_dram_bytes"]
        except Exception:
            pass
        kt_s = kernel_ms_max * 1e-3
        fine_s = fine_ms_max * 1e-3
        # projections of one vc_carve_bricks launch (rank average): per-voxel ones + the corners of its sub-brick classification
        exec_rank = (executed_total - corner_total + sub_corner_total) / world
        ach = exec_rank * F_ALG / fine_s / 1e12
        alg_bytes = 2.0 * (Z / world) * Y * Wx * 4 + w.mask_bits.nbytes
        out = {
            "metric": METRIC, "value": nominal_total / (ms_per_step * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(config_dict(w, world), l2="flushed between timed iterations (256 MiB fill)", slab_bounds=bounds),
            "executed_voxel_views": executed_total, "executed_fraction": executed_total / nominal_total,
            "executed_value": executed_total / (ms_per_step * 1e-3),
            "occupied_voxels": occupied_total, "carve_kernel_ms": kernel_ms_max,
            "kernels_ms": {"vc_brick_classify_kernel<1>+<0>": classify_ms_max, "vc_carve_bricks (incl. the fill pass)": fine_ms_max,
                           "flat_vc_carve_rows_same_job": flat_ms},
            "bricks": {"total": bricks_total, "needing_per_voxel_work": bricks_listed, "corner_projections": corner_total,
                       "of_which_sub_brick_level": sub_corner_total},
            "filter": {"evaluations_x32": filter_rows, "exact_reevaluations_x32": filter_slow, "mismatches_vs_exact": filter_bad,
                       "note": "per-voxel f32 filter with rigorous radius; every decision cross-checked against the exact path in the counting run"},
            "roofline": {"bound": "fp32", "achieved": ach, "peak": ffma, "unit": "TFLOP/s", "frac": ach / ffma if ffma else None,
                         "traffic": traffic, "algorithmic_bytes": exec_rank / 8.0 + 2.0 * bricks_listed / world * 2048 / 8,
                         "kernel": "vc_carve_bricks",
                         "how": f"projections of one vc_carve_bricks launch ({exec_rank:.4g}: per-voxel + sub-brick corners) x {F_ALG:.0f} FLOP / its time "
                                f"({fine_ms_max:.3f} ms, CUDA events inside vc_carve; the kernel's first blocks also do the fill pass); peak = FFMA microbenchmark on this GPU "
                                f"(vc_measure_peaks); DFMA peak {dfma:.1f} TFLOP/s. The kernel is issue-bound, see profiles/"},
            "roofline_hbm": {"bound": "hbm", "achieved": alg_bytes / kt_s / 1e9, "peak": hbm_peak, "unit": "GB/s",
                             "frac": alg_bytes / kt_s / 1e9 / hbm_peak, "traffic": None, "write_only_floor_ms": 2.0 * (Z / world) * Y * Wx * 4 / 3.68e12 * 1e3,
                             "how": f"grid-write bound: occupied+seen slab written once + masks read once = {alg_bytes / 1e6:.1f} MB / kernel time; peak = {hbm_src}; "
                                    "write_only_floor_ms = the two volumes at the 3.68 TB/s a cudaMemset reaches on this GPU (tools/memset_bench.py)"},
            "allgather_ms": allgather_ms, "mc_classify_ms": mc_ms, "mc_triangles": mc_tris, "color_ms": color_ms, "per_rank": per_rank,
            "e2e": e2e, "e2e_bgr8": e2e_bgr, "gpu_launches": launches, "clocks": clk,
        }
        if world == 1 and not args.no_cpu_baseline:
            v, cores, sample, _ = cpu_reference(w, args.cpu_seconds, 1, 1)
            out["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}
    eng.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        emit(out)
    return 0


if __name__ == "__main__":
    a = parse()
    sys.exit(run_reference(a) if a.impl == "reference" else run_ours(a))
