#!/usr/bin/env python
"""bench.py -- RANSAC hypotheses/s on BASELINE config C2 (fit_plane + fit_sphere + fit_cylinder,
1M-point synthetic cloud, 10k hypotheses each) on N x B200, beside the CPU reference path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One step = the three fits (30 000 hypotheses per GPU; probability = 1.0 so that every hypothesis
is evaluated, ransac.h:601-606).  For N > 1 (torchrun, one rank per GPU) the hypothesis batch
grows with N (10k per primitive per GPU: weak scaling); rank r scores its shard of the SAME
global sample table and one NCCL all-gather of the inlier counts per fit lets every rank replay
the identical ordered scan.

Printed JSON (one line, rank 0):
  value     hypotheses/s, cloud resident in HBM (m3d_ransac_fit_cloud), CUDA-event timed
  e2e       the same through the host-buffer C-ABI entry point (m3d_ransac_fit): pinned host
            cloud -> H2D, fit, inlier indices -> D2H, all inside the timed region
  roofline  the scoring kernel against the HBM roofline (compulsory bytes 24*N + 64*H per launch,
            SURVEY.md §8d) -- and `roofline_alu`, the roofline that actually binds it
  cpu_baseline  the reference's own ransac.h (compiled into oracle/_ref with -O3 -fopenmp) on the box's
            host cores; the oracle's OpenMP restatement if that library is absent
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_POINTS = 1_000_000
H_PER_PRIM = 10_000
THR = 0.01
SEED = 20240917
KINDS = (0, 1, 2)  # plane, sphere, cylinder
FLOPS_PER_UNIT = {0: 8, 1: 10, 2: 24}      # SURVEY.md §8(d), fp64 ops of the reference as written
FAST_FFMA_PER_UNIT = {0: 3, 1: 4, 2: 8}    # fp32 FMA-pipe ops of the scoring kernel's inner loop


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured", d
    return 6650.0, "fallback", {}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons while the timed region runs"""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def cpu_path():
    """The CPU implementation both CPU legs time.  Preferred: the REFERENCE'S OWN sources (ransac.h's
    OpenMP loop, mutex-guarded sampler and per-hypothesis SelectByIndex included) compiled into
    oracle/_ref/libm3d_ref_omp.so with the reference's flags (-O3 -fopenmp; Eigen/Open3D stood in by
    oracle/shim/) -> kind "reference".  Fallback when that library did not travel: the oracle's OpenMP
    restatement of the same loop -> kind "port".  Returns (kind, cores, fit(kind, xyz, nrm, h, seed), what)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import refc
    if refc.available(omp=True):
        cores = refc.use_all_cores()

        def fit(kind, xyz, nrm, h, seed):
            refc.ransac_fit(kind, xyz, nrm if kind == 2 else None, THR, h, 1.0, seed, omp=True)
        return "reference", cores, fit, ("the reference's own ransac.h compiled with -O3 -fopenmp (oracle/_ref; "
                                         "Eigen/Open3D stand-ins), RANSAC<>::FitModel incl. RefineModel")
    import orc
    orc.build()
    cores = orc.use_all_cores()

    def fit(kind, xyz, nrm, h, seed):
        orc.ransac_fit(kind, xyz, nrm if kind == 2 else None, thr=THR, max_it=h, prob=1.0, seed=seed, omp=True,
                       faithful=True)
    return "port", cores, fit, ("oracle OpenMP restatement of ransac.h:571-614 incl. the per-hypothesis O(N) "
                                "SelectByIndex pass")


def calibrate(fit, xyz, cores, seconds_per_step):
    """hypotheses per primitive such that one plane+sphere+cylinder step is about `seconds_per_step`
    (two-point: a call also has O(N) fixed costs -- cloud copy, RefineModel -- that do not scale with H)"""
    ts = []
    for h in (cores, 5 * cores):
        t0 = time.perf_counter()
        fit(0, xyz, None, h, 1)
        ts.append(time.perf_counter() - t0)
    per_hyp = max((ts[1] - ts[0]) / (4 * cores), 1e-6)
    # plane is the cheapest of the three primitives: sphere + cylinder cost about 4x a plane hypothesis
    h = (seconds_per_step - 3 * ts[0]) / (5.0 * per_hyp)
    return int(min(H_PER_PRIM, max(cores, round(h / cores) * cores)))


def pin_to_gpu_numa_node(local_rank):
    """Opt-in (M3D_BENCH_PIN=1): restrict this rank to the CPUs of its GPU's NUMA node.  Every fit ends with an
    exchange that waits for the slowest rank, so per-rank host latency jitter costs all ranks; torchrun does not
    pin its workers.  Off by default until it has been measured on the 8-GPU box."""
    try:
        import torch
        bdf = torch.cuda.get_device_properties(local_rank).pci_bus_id if hasattr(
            torch.cuda.get_device_properties(local_rank), "pci_bus_id") else None
        if bdf is None:
            import pynvml
            pynvml.nvmlInit()
            bdf = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(local_rank)).busId
            bdf = bdf.decode() if isinstance(bdf, bytes) else bdf
        bdf = bdf.lower()
        if len(bdf.split(":")[0]) == 8:      # nvml prints a 32-bit domain, sysfs a 16-bit one
            bdf = bdf[4:]
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read())
        if node < 0:
            return
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if len(cpus) >= 2:
            os.sched_setaffinity(0, cpus)
    except Exception:
        pass


def reference_arm(args, rank, world):
    """The reference's own CPU implementation of the path on the host cores (see cpu_path): each step =
    a bounded sample of the C2 workload (H_cpu hypotheses per primitive on the full 1M-point cloud)."""
    if rank != 0:
        return
    from misc3d_b200 import synth
    kind_name, cores, fit, what = cpu_path()
    xyz, nrm = synth.make_c2(N_POINTS, SEED)
    h_cpu = calibrate(fit, xyz, cores, 4.5)

    def step(seed):
        for kind in KINDS:
            fit(kind, xyz, nrm, h_cpu, seed)

    for w in range(args.warmup):
        step(100 + w)
    t0 = time.perf_counter()
    for s in range(args.steps):
        step(200 + s)
    dt = time.perf_counter() - t0
    value = 3 * h_cpu * args.steps / dt
    sample = (f"{h_cpu} hypotheses per primitive per step (of {H_PER_PRIM}) on the full {N_POINTS}-point cloud; "
              f"{what}; {cores} threads")
    print(json.dumps({
        "impl": "reference", "metric": "ransac_hypotheses_per_sec", "value": value, "unit": "hypotheses/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2: fit_plane+fit_sphere+fit_cylinder, 1M-point cloud, 10k hypotheses each "
                               "(bounded sample per step)", "n_points": N_POINTS, "threshold": THR,
                   "probability": 1.0, "hypotheses_per_primitive_per_step": h_cpu},
        "cpu_baseline": {"value": value, "unit": "hypotheses/s", "cores": cores, "kind": kind_name, "sample": sample},
        "e2e": {"value": value, "unit": "hypotheses/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "point_hypotheses_per_sec": value * N_POINTS,
    }))


def cpu_baseline(xyz, nrm, budget_s=12.0):
    kind_name, cores, fit, what = cpu_path()
    h_cpu = calibrate(fit, xyz, cores, budget_s)
    t0 = time.perf_counter()
    for kind in KINDS:
        fit(kind, xyz, nrm, h_cpu, 3)
    value = 3 * h_cpu / (time.perf_counter() - t0)
    out = {"value": value, "unit": "hypotheses/s", "cores": cores, "kind": kind_name,
           "sample": f"{h_cpu} of {H_PER_PRIM} hypotheses per primitive on the full {N_POINTS}-point cloud; {what}"}
    if kind_name == "reference":   # the oracle port beside it, for continuity with earlier runs
        import orc
        orc.build()
        orc.use_all_cores()
        t0 = time.perf_counter()
        for kind in KINDS:
            orc.ransac_fit(kind, xyz, nrm if kind == 2 else None, thr=THR, max_it=h_cpu, prob=1.0, seed=3, omp=True,
                           faithful=True)
        out["port_value"] = 3 * h_cpu / (time.perf_counter() - t0)
        out["port_note"] = "oracle OpenMP restatement of the same loop (what earlier runs reported)"
    return out


def kernel_profile():
    """ncu figures of the dominant kernel kept under profiles/ (written by tools/ncu_summary.py from an ncu --set full
    capture of THIS kernel build: the file records the sha of the kernel source, so a stale profile is visible)."""
    import hashlib
    p = os.path.join(ROOT, "profiles", "score_kernel_ncu.json")
    src = os.path.join(ROOT, "misc3d_b200", "csrc", "score_cell.cuh")
    if not os.path.exists(p):
        return None
    try:
        d = json.load(open(p))
        d["stale"] = d.get("kernel_source_sha16") != hashlib.sha256(open(src, "rb").read()).hexdigest()[:16]
        return d
    except Exception:
        return None


def c5_leg(ctx, capi, synth, dist, dev, rank, world, reps=3):
    """BASELINE config C5 -- fit_plane, 4M points x 100k hypotheses, STRONG scaling: the 100k hypotheses are sharded
    over the ranks (one 64-byte best record per rank is exchanged).  Returns ms per fit (max over ranks)."""
    import torch
    n, H = 4_000_000, 100_000
    xyz = synth.make_c1(n=n, seed=SEED)
    cloud = ctx.upload(xyz)
    res = None
    for _ in range(2):
        res = ctx.ransac_fit_cloud(0, cloud, THR, H, 1.0, seed=7, want_inliers=False)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    score = []
    for r in range(reps):
        res = ctx.ransac_fit_cloud(0, cloud, THR, H, 1.0, seed=7, want_inliers=False)
        score.append(res[3]["score_ms"])
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = 1e3 * float(t.item())
    cloud.free()
    st = res[3]
    return {"workload": "C5: fit_plane, 4M points x 100k hypotheses, probability 1.0; hypotheses sharded over the ranks "
                        "(strong scaling), cloud resident", "ms_per_fit": ms, "hypotheses_per_sec": H / (ms * 1e-3),
            "point_hypotheses_per_sec": float(n) * H / (ms * 1e-3), "score_ms_rank0": float(np.mean(score)),
            "best_index": int(st["best_index"]), "best_count": int(st["best_count"]), "n_gpus": world}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the pageable / pybind / C5 legs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
        return
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    from misc3d_b200 import capi, synth

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    if world > 1 and os.environ.get("M3D_BENCH_PIN") == "1":
        pin_to_gpu_numa_node(local_rank)   # opt-in experiment (not measured yet): see the function
    xyz, nrm = synth.make_c2(N_POINTS, SEED)
    stream = torch.cuda.current_stream()
    ctx = capi.Context(local_rank, stream=stream.cuda_stream)
    if world > 1:
        ids = [capi.nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(ids, src=0)
        ctx.init_nccl(ids[0], rank, world)
    H = H_PER_PRIM * world  # weak scaling: 10k hypotheses per primitive per GPU

    # pinned host copies (e2e leg) and the HBM-resident clouds (value leg)
    h_xyz = torch.from_numpy(xyz).pin_memory()
    h_nrm = torch.from_numpy(nrm).pin_memory()
    np_xyz, np_nrm = h_xyz.numpy(), h_nrm.numpy()
    cloud = ctx.upload(np_xyz, np_nrm)
    inl_buf = torch.empty(N_POINTS, dtype=torch.int64).pin_memory().numpy().view(np.uint64)  # pinned result buffer
    flush = torch.empty(512 * 1024 * 1024, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # value leg, N > 1: every rank gets the model, the inlier COUNT and the stats of every fit; the inlier index
    # list (2.7 MB per fit) is copied to the host on rank 0 only -- the job has one result, and eight identical
    # device->host copies at the same instant share PCIe uplinks (tools/numa_probe.py: GPUs 0-3 of the 8-GPU box
    # drop to 16 GB/s each).  Every rank still builds the list on its GPU (RefineModel needs it).
    # M3D_BENCH_INLIERS=all restores the copy on every rank.
    inl_everywhere = os.environ.get("M3D_BENCH_INLIERS") == "all"
    want_inl = rank == 0 or inl_everywhere

    def step_resident(seed):
        out = []
        for kind in KINDS:
            rc, model, inl, st = ctx.ransac_fit_cloud(kind, cloud, THR, H, 1.0, seed=seed + kind,
                                                      inl_buf=inl_buf if want_inl else None, want_inliers=want_inl)
            out.append((rc, len(inl) if inl is not None else 0, st))
        return out

    def step_e2e(seed, pts, nrms, buf):
        d2h = 0
        for kind in KINDS:
            rc, model, inl, st = ctx.ransac_fit(kind, pts, nrms if kind == 2 else None, THR, H, 1.0, seed=seed + kind,
                                                inl_buf=buf if want_inl else None, want_inliers=want_inl)
            d2h += (inl.nbytes if inl is not None else 0) + 64 * world
        return d2h

    def timed_e2e(step, n_warm=2):
        for w in range(n_warm):
            step(3000 + w)
        barrier()
        t0 = time.perf_counter()
        out = None
        for s in range(args.steps):
            out = step(4000 + 3 * s)
        barrier()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return 3.0 * H * args.steps / float(t.item()), out

    # ---------------------------------------------------------------- value: resident cloud
    for w in range(args.warmup):
        step_resident(1000 + w)
    # one nvidia-smi poller for the job (rank 0's GPU): a poller per rank makes N processes query the driver
    # every 100 ms while N ranks are launching kernels
    clocks = ClockSampler(local_rank) if rank == 0 else None
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    score_ms = {k: [] for k in KINDS}
    fit_ms = {k: [] for k in KINDS}
    refine_ms = {k: [] for k in KINDS}
    draw_ms = {k: [] for k in KINDS}
    n_inl_k = {k: 0 for k in KINDS}
    resolves = 0
    launches0 = ctx.launches
    barrier()
    for s in range(args.steps):
        flush.fill_(s & 0xFF)  # L2 flush between timed iterations (not timed)
        barrier()
        ev[s][0].record(stream)
        res = step_resident(2000 + 3 * s)
        ev[s][1].record(stream)
        for kind, (rc, n_inl, st) in zip(KINDS, res):
            score_ms[kind].append(st["score_ms"])
            fit_ms[kind].append(st["device_ms"])
            refine_ms[kind].append(st["refine_ms"])
            draw_ms[kind].append(st["draw_ms"])
            n_inl_k[kind] = n_inl
            resolves += st["exact_resolves"]
    barrier()
    launches = ctx.launches - launches0
    clk = clocks.stop() if clocks else None
    total_ms = float(sum(a.elapsed_time(b) for a, b in ev))
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    value = 3.0 * H * args.steps / (total_ms * 1e-3)

    # ---------------------------------------------------------------- e2e: host buffers through the C-ABI
    e2e_value, d2h = timed_e2e(lambda seed: step_e2e(seed, np_xyz, np_nrm, inl_buf))
    # three cloud uploads + the normals of the cylinder's sample points (2 per hypothesis; the caller's normal array
    # itself stays on the host) + the cylinder's host-drawn sample table (plane / sphere tables are drawn on the device)
    # (N > 1: every rank uploads 1/N of the cloud and the slices are all-gathered over NVLink; bytes are rank 0's)
    sharded_upload = world > 1 and os.environ.get("M3D_SHARD_UPLOAD") != "0"
    h2d = 3 * xyz.nbytes // (world if sharded_upload else 1) + 2 * 24 * H + 2 * 4 * H
    extras = {}
    if not args.no_extras:
        # the same call with what an Open3D / numpy caller actually holds: pageable arrays, pageable result buffer
        pg_xyz, pg_nrm = xyz.copy(), nrm.copy()
        pg_buf = np.empty(N_POINTS, dtype=np.uint64)
        v, _ = timed_e2e(lambda seed: step_e2e(seed, pg_xyz, pg_nrm, pg_buf), n_warm=1)
        extras["e2e_pageable"] = {"value": v, "unit": "hypotheses/s", "api": "m3d_ransac_fit (host buffers, pageable numpy arrays)"}
        # the same arrays page-locked in place by the library on first use (M3D_FLAG_REGISTER_HOST): what a caller that
        # fits the same cloud repeatedly can ask for
        def step_reg(seed):
            for kind in KINDS:
                ctx.ransac_fit(kind, pg_xyz, pg_nrm if kind == 2 else None, THR, H, 1.0, seed=seed + kind,
                               inl_buf=pg_buf if want_inl else None, want_inliers=want_inl, flags=capi.FLAG_REGISTER_HOST)
            return 0
        v, _ = timed_e2e(step_reg, n_warm=1)
        ctx.host_unregister_all()
        extras["e2e_pageable_registered"] = {"value": v, "unit": "hypotheses/s",
                                             "api": "m3d_ransac_fit (pageable numpy arrays, M3D_FLAG_REGISTER_HOST: "
                                                    "page-locked in place on first use)"}
        if world == 1:
            try:   # the reference-facing python call: misc3d.common.fit_* (pybind11 shim over the C++ facade)
                sys.path.insert(0, os.path.join(ROOT, "python"))
                import misc3d

                class _PC:   # duck-typed open3d.geometry.PointCloud
                    def __init__(self, p, n):
                        self.points, self.normals = p, n
                pc = _PC(pg_xyz, pg_nrm)
                fits = (misc3d.common.fit_plane, misc3d.common.fit_sphere, misc3d.common.fit_cylinder)

                def step_py(seed):
                    for kind, f in zip(KINDS, fits):
                        f(pc, THR, H, 1.0, seed=seed + kind)
                    return 0
                v, _ = timed_e2e(step_py, n_warm=1)
                extras["e2e_pybind"] = {"value": v, "unit": "hypotheses/s",
                                        "api": "misc3d.common.fit_plane/fit_sphere/fit_cylinder (pybind11 shim; returns the "
                                               "inlier indices as a python list like the reference: O(n_inl) boxing included)"}
            except Exception as e:  # the shim is optional on the bench box
                extras["e2e_pybind"] = {"value": None, "error": str(e)[:200]}
        if world == 1:
            # throughput of a caller that keeps three clouds in flight: three host threads, one context (own stream and
            # scratch) each, the plain synchronous m3d_ransac_fit call in both -- the upload of one thread's cloud
            # overlaps the scoring of the other's.  Same work per step as `e2e` (which stays the one-call-at-a-time number).
            import threading
            n_ctx = max(2, int(os.environ.get("M3D_BENCH_CTXS", "3")))
            ctxs = [capi.Context(local_rank) for _ in range(n_ctx)]
            bufs = [torch.empty(N_POINTS, dtype=torch.int64).pin_memory().numpy().view(np.uint64) for _ in range(n_ctx)]

            def run(i, seeds):
                for seed in seeds:
                    for kind in KINDS:
                        ctxs[i].ransac_fit(kind, np_xyz, np_nrm if kind == 2 else None, THR, H, 1.0,
                                           seed=seed + kind, inl_buf=bufs[i])

            def both(seed_lists):
                th = [threading.Thread(target=run, args=(i, seed_lists[i])) for i in range(n_ctx)]
                for t_ in th:
                    t_.start()
                for t_ in th:
                    t_.join()
            both([[3000 + i] for i in range(n_ctx)])
            steps_c = max(args.steps, n_ctx)
            seeds = [[4000 + 3 * s_ for s_ in range(steps_c) if s_ % n_ctx == i] for i in range(n_ctx)]
            barrier()
            t0 = time.perf_counter()
            both(seeds)
            barrier()
            dt = time.perf_counter() - t0
            extras["e2e_concurrent"] = {"value": 3.0 * H * steps_c / dt, "unit": "hypotheses/s", "contexts": n_ctx,
                                        "api": f"m3d_ransac_fit (host buffers, pinned) called from {n_ctx} host threads, "
                                               "one context each"}
            for c_ in ctxs:
                c_.close()
        extras["c5"] = c5_leg(ctx, capi, synth, dist, dev, rank, world)

    # work counters of the kernel (statistics build, one sharded launch per primitive, outside the timed region;
    # every rank takes part in the fit's exchange, rank 0 reports its shard)
    stats = {}
    for k in KINDS:
        ctx.ransac_fit_cloud(k, cloud, THR, H, 1.0, seed=2000 + k, flags=capi.FLAG_STATS, want_inliers=False)
        stats[k] = ctx.score_stats()

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------------------------------------------------------- roofline of the dominant kernel
    hbm_peak, which, pk = peaks()
    ffma_peak = ctx.probe_fp32_ffma()  # FFMA lane-ops/s measured on this device
    dfma_peak = ctx.probe_fp64_dfma()  # DFMA lane-ops/s (resolve / refine kernels)
    h_local = H // world  # hypotheses one launch scores on this rank
    per_kind = {}
    tot_score = sum(np.mean(score_ms[k]) for k in KINDS)
    useful_ops = 0.0
    for k, name in zip(KINDS, ("plane", "sphere", "cylinder")):
        ms = float(np.mean(score_ms[k]))
        units = float(N_POINTS) * h_local
        pairs = 32.0 * stats[k]["cell_pairs"]
        slots = 32.0 * (32 * stats[k]["passes_1"] + 64 * stats[k]["passes_2"])
        # fp32 FMA-pipe lane-ops the kernel executed for point tests: per evaluated pair FAST_FFMA_PER_UNIT FFMA + 1 FADD
        ops = (FAST_FFMA_PER_UNIT[k] + 1) * slots
        useful_ops += ops
        rbytes = 48.0 * N_POINTS + 8.0 * n_inl_k[k]
        rms = float(np.mean(refine_ms[k]))
        per_kind[name] = {"score_kernel_ms": ms, "fit_ms": float(np.mean(fit_ms[k])),
                          "refine_ms": rms, "draw_ms": float(np.mean(draw_ms[k])),
                          "refine_GBps": rbytes / (rms * 1e-3) / 1e9 if rms > 0 else None,
                          "refine_frac_of_hbm_peak": rbytes / (rms * 1e-3) / 1e9 / hbm_peak if rms > 0 else None,
                          "point_hypotheses_per_sec": units / (ms * 1e-3),
                          "pairs_evaluated_frac": pairs / units, "tile_survivor_frac": stats[k]["tile_survivors"] / max(stats[k]["tile_tests"], 1),
                          "lane_efficiency": pairs / max(slots, 1.0), "guard_band_rescans": stats[k]["rescans"],
                          "fp32_pipe_frac": ops / (ms * 1e-3) / ffma_peak,
                          "dense_equiv_ffma_frac": FAST_FFMA_PER_UNIT[k] * units / (ms * 1e-3) / ffma_peak}
    # ALGORITHMIC bytes per launch: 24 B per point-hypothesis unit (one Vector3d streamed per
    # evaluation, what the reference's loop moves) x N*H units.  `traffic` is what ncu measured at the DRAM
    # (one read of the cloud): achieved / peak is >> 1 because every point fetched is re-used on chip and ~90 % of
    # the pairs are decided per cell / tile -- the contract figure is not a utilisation; `binding` is.
    launch_ms = tot_score / 3.0
    algo_bytes = 24.0 * N_POINTS * h_local
    ach = algo_bytes / (launch_ms * 1e-3) / 1e9
    compulsory = 24.0 * N_POINTS + 64.0 * h_local
    prof = kernel_profile()
    traffic = prof.get("dram_bytes_per_launch") if prof else None
    binding = {"resource": "warp-instruction issue slots (the kernel is instruction-issue / latency bound; HBM idle)",
               "frac": prof.get("issue_slots_busy_frac") if prof else None,
               "frac_source": ("ncu smsp__issue_active of this kernel build, profiles/score_kernel_ncu.json"
                               + (" [STALE: kernel source changed since the capture]" if prof and prof.get("stale") else ""))
               if prof else "no ncu capture in profiles/",
               "ncu": {k: prof.get(k) for k in ("issue_slots_busy_frac", "fma_pipe_frac", "smem_wavefront_frac", "warps_active_frac",
                                                "warp_instructions_per_launch", "duration_ms", "capture")} if prof else None,
               # measured in THIS run from the kernel's own counters: fp32 lane-operations issued for point tests
               # (evaluated lane slots x (FFMA + FADD per pair)) over the FFMA rate measured by m3d_probe_fp32_ffma
               "fp32_pipe_frac_live": useful_ops / (tot_score * 1e-3) / ffma_peak,
               "pairs_evaluated_frac_live": {n: per_kind[n]["pairs_evaluated_frac"] for n in per_kind},
               "lane_efficiency_live": {n: per_kind[n]["lane_efficiency"] for n in per_kind}}
    roofline = {"bound": "hbm", "kernel": "score_cell_kernel<KIND,512,1024> (average of the plane, sphere and cylinder launches)",
                "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak, "traffic": traffic,
                "peak_source": f"{which} (MEASURED_PEAKS.json hbm_gbs)" if which == "measured" else "fallback",
                "algorithmic_bytes_per_launch": algo_bytes, "launch_ms": launch_ms,
                "note": "24 B x N x H streaming-equivalent bytes (SURVEY.md 8d); frac >> 1 is NOT a utilisation: every point "
                        "fetched from HBM once per launch is re-used on chip by the whole hypothesis batch and ~90 % of the "
                        "pairs are decided by one bounding-sphere test per cell / tile; see `binding`",
                "binding": binding,
                "compulsory_bytes_per_launch": compulsory,
                "compulsory_GBps": compulsory / (launch_ms * 1e-3) / 1e9,
                "compulsory_frac": compulsory / (launch_ms * 1e-3) / 1e9 / hbm_peak}
    ffma_ops = sum(FAST_FFMA_PER_UNIT[k] * float(N_POINTS) * h_local for k in KINDS)
    roofline_alu = {"bound": "fp32-fma-pipe (dense-equivalent)", "achieved": ffma_ops / (tot_score * 1e-3) / 1e12,
                    "peak": ffma_peak / 1e12, "unit": "TFFMA/s",
                    "frac": ffma_ops / (tot_score * 1e-3) / ffma_peak, "fp64_dfma_peak_T": dfma_peak / 1e12,
                    "note": "FFMA lane-ops a dense evaluation of all N x H pairs would need (3/4/8 per pair) over the "
                            "FFMA rate measured by m3d_probe_fp32_ffma on this device: a speed-up measure against the dense "
                            "kernel's binding roofline, not a utilisation (the kernel evaluates ~7-11 % of the pairs)"}
    rb = sum(48.0 * N_POINTS + 8.0 * n_inl_k[k] for k in KINDS)
    rt = sum(float(np.mean(refine_ms[k])) for k in KINDS)
    roofline_refine = {"bound": "hbm", "kernels": "refine_count / refine_scan / refine_write / refine_final (RefineModel, ransac.h:534-549)",
                       "achieved": rb / (rt * 1e-3) / 1e9 if rt > 0 else None, "peak": hbm_peak, "unit": "GB/s",
                       "frac": rb / (rt * 1e-3) / 1e9 / hbm_peak if rt > 0 else None,
                       "algorithmic_bytes_per_fit": "2 x 24 N read + 8 n_inl written",
                       "note": "timed with CUDA events around the four passes of one fit (the 24 MB cloud is L2-resident after the "
                               "scoring kernel; launch gaps of the four short kernels are inside the interval)"}

    cpu = None if args.no_cpu else cpu_baseline(xyz, nrm)
    line = {
        "metric": "ransac_hypotheses_per_sec", "value": value, "unit": "hypotheses/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+f64",
        "data": "synthetic",
        "config": {"workload": "C2: fit_plane+fit_sphere+fit_cylinder on one 1M-point synthetic cloud, "
                               f"{H_PER_PRIM} hypotheses per primitive per GPU, probability 1.0 (no early exit)",
                   "n_points": N_POINTS, "hypotheses_per_primitive": H, "threshold": THR,
                   "l2": "512 MiB flush write between timed steps", "sharding": f"hypotheses over {world} rank(s)",
                   "seeds": "a different sample-table seed every step and primitive (no table re-use)",
                   "results_to_host": ("model, inlier count and stats on every rank; inlier index list on "
                                       + ("every rank" if inl_everywhere or world == 1 else "rank 0"))},
        "point_hypotheses_per_sec": value * N_POINTS,
        "e2e": {"value": e2e_value, "unit": "hypotheses/s", "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h),
                "api": "m3d_ransac_fit (host buffers, pinned)" + (
                    "; each rank uploads 1/N of the replicated cloud, NVLink all-gather of the rest" if sharded_upload else "")},
        "gpu_launches": int(launches),
        "clocks": clk,
        "roofline": roofline, "roofline_alu": roofline_alu, "roofline_refine": roofline_refine, "per_primitive": per_kind,
        "exact_resolves_per_step": resolves / args.steps,
        "cpu_baseline": cpu,
    }
    line.update(extras)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
