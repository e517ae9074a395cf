#!/usr/bin/env python
"""Headline benchmark: Vicon CSV loader throughput (GB of CSV per second, output bit-exact)
on synthetic 10-minute trials of the dynamic_trial.csv layout (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One step = one pass of the hot path over one trial per GPU: ms_scan + header parse +
ms_parse (all data rows -> channel-major float64 in HBM) + Segmenter (40 transitions) +
gather of the 32 phase windows of the EMG device.  `value` times it with the CSV bytes already
resident in HBM; `e2e` times the public batch call `ViconLoader.load_many` from pinned HOST
bytes to HOST arrays (H2D of every CSV byte and D2H of every parsed double inside the timed
region, overlapped across consecutive trials on three streams).
Multi-GPU: one process per GPU, one distinct trial per rank per step, no collective on the
data path (weak scaling); NCCL is used only for the barrier and the max-over-ranks time.

`--impl reference` times the reference's own CPU algorithm (oracle/vicon_oracle.py, the
plain-Python port of load_vicon_file) on all host cores, on a bounded sample of the same
layout.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "vicon_csv_loader_throughput"
UNIT = "GB/s"
WORKLOAD = "T10: synthetic 10-min Vicon trial, 2 force plates + 16 EMG @2 kHz, 40 markers @100 Hz, load + segment"


# ---- CPU baseline (the reference's algorithm, Python port, all cores) ------------------------------
def _cpu_worker(args):
    path, reps = args
    from oracle.vicon_oracle import load_vicon_file_oracle

    n = 0
    for _ in range(reps):
        res = load_vicon_file_oracle(path)
        n += res.num_frames
    return n


def cpu_baseline(seconds_of_trial=6.0, target_wall=12.0, reps=None):
    """Loads a `seconds_of_trial` T10-layout sample once per core per rep with the Python port of
    the reference loader; returns (GB/s, cores, sample description, wall seconds)."""
    import multiprocessing as mp

    from tools.synth_vicon import synth_vicon

    cores = os.cpu_count() or 1
    blob = synth_vicon(seed=1234, seconds=seconds_of_trial, n_emg=16, n_markers=40)
    base = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
    path = os.path.join(base, f"ms_b200_cpu_sample_{os.getpid()}.csv")
    blob.tofile(path)
    try:
        ctx = mp.get_context("fork")
        with ctx.Pool(cores) as pool:
            pool.map(_cpu_worker, [(path, 1)] * cores)  # warm-up: imports, page cache
            if reps is None:
                t = time.perf_counter()
                pool.map(_cpu_worker, [(path, 1)] * cores)
                one = time.perf_counter() - t
                reps = max(1, int(target_wall / max(one, 1e-3)))
            t = time.perf_counter()
            pool.map(_cpu_worker, [(path, reps)] * cores)
            wall = time.perf_counter() - t
    finally:
        os.unlink(path)
    total = blob.nbytes * reps * cores
    sample = (f"{cores} processes x {reps} loads of a {seconds_of_trial:g} s T10-layout trial "
              f"({blob.nbytes / 1e6:.1f} MB), python port of load_vicon_file")
    return total / wall / 1e9, cores, sample, wall


# ---- clocks ---------------------------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.samples = []
        self._stop = threading.Event()
        self._thread = None

    def _run(self):
        # NVML in-process (a sample every 10 ms: the timed region of a short run is tens of milliseconds);
        # nvidia-smi, one process per sample, only when the binding is missing
        try:
            import pynvml

            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu_index)
            smax = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self._stop.is_set():
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                try:
                    why = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:  # noqa: BLE001 - older bindings
                    why = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                flags = ["Active" if why & b else "Not Active" for b in (0x8, 0x40, 0x20, 0x4)]
                self.samples.append(["", str(sm), str(smax), "", ""] + flags)
                self._stop.wait(0.01)
            return
        except Exception:  # noqa: BLE001
            pass
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-i", str(self.gpu_index)],
                    capture_output=True, text=True, timeout=5,
                ).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:  # noqa: BLE001
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._thread.start()
        return self

    def __exit__(self, *exc):
        self._stop.set()
        self._thread.join(timeout=6)

    def summary(self):
        sm, reasons, smax = [], set(), None
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            try:
                sm.append(float(s[1]))
                smax = float(s[2])
                for name, val in zip(names, s[5:9]):
                    if val.lower().startswith("active"):
                        reasons.add(name)
            except Exception:  # noqa: BLE001
                continue
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


# ---- NMF-MU extension -------------------------------------------------------------------------------------
def nmf_envelopes(seed=1, n=200, m=16, k_true=4):
    import numpy as np

    rng = np.random.default_rng(seed)
    t = np.linspace(0, 1, n)[:, None]
    basis = np.abs(np.sin(np.pi * (rng.uniform(0.5, 3, (1, k_true)) * t + rng.uniform(0, 1, (1, k_true))))) ** 2
    X = basis @ rng.uniform(0, 1, (k_true, m)) + 0.02 * rng.uniform(0, 1, (n, m))
    return X / X.max(axis=0)


def bench_nmf(dev, iters=2000):
    """160 problems (k=1..8 x 20 restarts) x `iters` MU iterations in one launch; sklearn on one
    core for a bounded subset beside it."""
    import warnings

    import numpy as np
    import torch

    from muscle_synergies_b200 import analysis

    X = nmf_envelopes()
    ranks = [k for k in range(1, 9) for _ in range(20)]
    seeds = [r for _ in range(1, 9) for r in range(20)]
    analysis.nmf_mu_batched(X, ranks, seeds, max_iter=50, tol=0.0)  # warm-up
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # time the launch itself: host init / upload excluded by timing a second, longer run around events
    t = time.perf_counter()
    res = analysis.nmf_mu_batched(X, ranks, seeds, max_iter=iters, tol=0.0)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t
    t = time.perf_counter()
    analysis.nmf_mu_batched(X, ranks, seeds, max_iter=1, tol=0.0)
    torch.cuda.synchronize()
    overhead = time.perf_counter() - t
    kernel_s = max(wall - overhead, 1e-6)
    total_iters = int(res.n_iter.sum())
    out = {"problems": len(ranks), "shape": [200, 16], "iterations_per_problem": iters,
           "iterations_per_s": total_iters / kernel_s, "kernel_s": kernel_s, "host_overhead_s": overhead,
           "regime": "shared-memory resident (no HBM traffic between iterations): bound by SM issue, not HBM"}
    # long-signal variant: X and W stream from HBM every iteration (the 200 x 16 case never touches HBM).
    # Timed with CUDA events around the C entry point on device-resident factors (the host-side upload of a
    # 2 M-row initialisation would drown the kernels in a wall-clock measurement); two iteration counts, the
    # difference is the cost of the extra iterations alone.
    try:
        import ctypes

        from muscle_synergies_b200 import _native as nat

        n_long, m, k, P = 2_000_000, 16, 8, 4
        lib = nat.lib()
        Xl = torch.rand((n_long, m), device=dev, dtype=torch.float32)
        W0 = torch.rand((P, n_long, k), device=dev, dtype=torch.float32)
        H0 = torch.rand((P, k, m), device=dev, dtype=torch.float32)
        work = torch.empty(int(lib.ms_nmf_stream_workspace_bytes(m, P)), dtype=torch.uint8, device=dev)
        d_iter = torch.empty(P, dtype=torch.int32, device=dev)
        d_err = torch.empty(P, dtype=torch.float32, device=dev)
        d_vaf = torch.empty((P, m + 1), dtype=torch.float32, device=dev)
        ranks_c = (ctypes.c_int32 * P)(*([k] * P))
        stream = torch.cuda.current_stream(dev)
        sptr = ctypes.c_void_p(stream.cuda_stream)

        def run(iters):
            W, H = W0.clone(), H0.clone()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            nat.check(lib.ms_nmf_mu_stream(Xl.data_ptr(), n_long, m, ranks_c, None, P, W.data_ptr(), H.data_ptr(), iters,
                                           ctypes.c_float(0.0), 10, work.data_ptr(), d_iter.data_ptr(), d_err.data_ptr(),
                                           d_vaf.data_ptr(), sptr), "ms_nmf_mu_stream")
            e1.record(stream)
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) * 1e-3

        run(3)
        its = 40
        t_small = min(run(5) for _ in range(2))
        t_big = min(run(5 + its) for _ in range(2))
        per_iter = max(t_big - t_small, 1e-6) / its
        bytes_iter = P * (4 * n_long * m + 8 * n_long * k)
        out["long_signal"] = {"shape": [n_long, m], "rank": k, "problems": P, "iterations_per_s": P / per_iter,
                              "ms_per_iteration_all_problems": per_iter * 1e3,
                              "algorithmic_gbs": bytes_iter / per_iter / 1e9,
                              "note": "4nm + 8nk bytes per iteration and problem; X re-read per problem; CUDA events"}
        del Xl, W0, H0
    except Exception as exc:  # noqa: BLE001
        out["long_signal"] = {"error": f"{type(exc).__name__}: {exc}"}
    try:
        from sklearn.decomposition import NMF

        t = time.perf_counter()
        done = 0
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for k, s in list(zip(ranks, seeds))[::8]:
                m = NMF(n_components=k, solver="mu", init="random", random_state=s, max_iter=200, tol=0.0)
                m.fit_transform(X)
                done += m.n_iter_
        out["sklearn_iterations_per_s_1core"] = done / (time.perf_counter() - t)
    except Exception as exc:  # noqa: BLE001
        out["sklearn_iterations_per_s_1core"] = None
        out["sklearn_error"] = str(exc)
    return out


def bench_pipeline(loader, d_bytes, n, layout, trials=3):
    """BASELINE configs[4] per trial: device-resident CSV -> load -> segment -> 8 gait cycles ->
    envelopes -> NMF sweep k=1..8 x 20 restarts x 200 iterations (1280 problems, one launch);
    wall clock, results (factors, errors, VAF of every restart) on the host."""
    import torch

    from muscle_synergies_b200.pipeline import trial_synergies

    def one():
        data = loader.load_device(d_bytes, n=n, name=layout)
        return trial_synergies(data, 1, 8, n_restarts=20, random_state=0, max_iter=200, tol=0.0)

    one()
    torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(trials):
        res = one()
    torch.cuda.synchronize()
    per_trial = (time.perf_counter() - t) / trials
    return {"workload": "per trial: load + segment + 8 cycles x (envelope, time-normalise 200) + NMF k=1..8 x 20 restarts x 200 it",
            "cycles_per_s": len(res.cycles) / per_trial, "ms_per_trial": per_trial * 1e3,
            "nmf_problems_per_trial": int(len(res.restarts)), "timing": "wall clock; factors, errors and VAF of every restart on the host (DataFrames are built on access)"}


def bench_files(loader, blob, layout, copies=6):
    """SURVEY.md section 8d, third number: files -> host arrays through `ViconLoader.load_files` (reader thread ->
    pinned ring -> H2D / parse / D2H pipeline).  `copies` files of the trial in shared memory (or the temp
    directory), page cache warm, wall clock."""
    import shutil
    import tempfile

    import torch

    base = "/dev/shm" if os.path.isdir("/dev/shm") and shutil.disk_usage("/dev/shm").free > (copies + 1) * blob.nbytes else None
    tmp = tempfile.mkdtemp(prefix="ms_b200_bench_", dir=base)
    try:
        paths = []
        for i in range(copies):
            path = os.path.join(tmp, f"trial{i}.csv")
            with open(path, "wb") as f:
                f.write(memoryview(blob))
            paths.append(path)

        def run():
            last = None
            for _name, data in loader.load_files(paths, to_host=True):
                if isinstance(data, Exception):
                    raise data
                last = data
            for blk in last.blocks:
                blk.host()
            torch.cuda.synchronize()

        run()
        t = time.perf_counter()
        run()
        wall = time.perf_counter() - t
        return {"value": copies * blob.nbytes / wall / 1e9, "unit": UNIT, "files": copies, "ms_per_file": wall / copies * 1e3,
                "where": "tmpfs" if base else "temp directory", "api": "ViconLoader.load_files -> host arrays, wall clock"}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


# ---- reference arm ----------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    # each step: every core loads the sample once
    vals = []
    cores = os.cpu_count() or 1
    for _ in range(args.warmup):
        cpu_baseline(seconds_of_trial=3.0, reps=1)
    t_total = 0.0
    nbytes = 0.0
    sample = ""
    for _ in range(max(1, args.steps)):
        gbs, cores, sample, wall = cpu_baseline(seconds_of_trial=3.0, reps=1)
        vals.append(gbs)
        t_total += wall
        nbytes += gbs * wall
    value = nbytes / t_total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "reference CPU algorithm (python port) on a bounded sample per step"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---- our arm ------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device(f"cuda:{local_rank}")
    # staging buffers and the threads that fill them next to the GPU (restored before the CPU baseline)
    from muscle_synergies_b200.sharding import bind_to_gpu_numa

    all_cpus = os.sched_getaffinity(0)
    numa_cpus = bind_to_gpu_numa(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    import __graft_entry__ as entry

    if rank == 0:
        entry.build()
    if dist is not None:
        dist.barrier()
    import muscle_synergies_b200 as ms
    from muscle_synergies_b200 import _native
    from muscle_synergies_b200.segment import Segmenter
    from tools.synth_vicon import synth_layout

    layout = args.layout
    blob = synth_layout(layout, seed=1000 + rank)
    n = int(blob.nbytes)
    loader = ms.ViconLoader(dev)
    pinned_in = torch.empty(loader.padded_size(n), dtype=torch.uint8, pin_memory=True)
    pinned_in.numpy()[:n] = blob
    d_bytes = torch.empty(loader.padded_size(n), dtype=torch.uint8, device=dev)
    d_bytes[:n].copy_(pinned_in[:n])
    torch.cuda.synchronize()

    def step_resident():
        # one submission: parse, transition search and the gather of the 32 EMG phase windows are queued
        # back to back; the parse status, the transitions and the window bounds come back in one wait
        data = loader.load_device(d_bytes, n=n, name=layout, defer_check=True)
        seg = Segmenter(data, cut_phases_of=(data.emg,))
        cuts = seg.phase_cuts(data.emg)
        return data, cuts

    # shapes for the algorithmic byte count: B_alg = B_csv + 8 * N_kept
    data, cuts = step_resident()
    n_kept = sum(int(d.tensor.numel()) for d in list(data.forcepl) + [data.emg] + list(data.traj))
    b_alg = n + 8 * n_kept

    def run_e2e(steps):
        """`steps` trials through the public pipelined API: pinned host CSV -> host arrays."""
        last = None
        for d in loader.load_many([pinned_in[:n]] * steps, names=[layout] * steps, to_host=True, host_slots=2):
            last = d
        for blk in last.blocks:
            blk.host()
        return last

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms_total = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms_total], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_total = float(t.item())
        barrier()
        return ms_total

    for _ in range(max(3, args.warmup)):
        step_resident()
    launches0 = _native.launch_count()
    with ClockSampler(local_rank) as clocks:
        ms_total = timed(step_resident, args.steps)
    launches = _native.launch_count() - launches0
    value = world * n * args.steps / (ms_total * 1e-3) / 1e9

    # ---- kernel-only timings on the launching stream (roofline of the dominant kernel)
    import ctypes

    from muscle_synergies_b200.vicon_data import loader as loader_mod

    src = loader_mod._Source(d_bytes, n, None)
    summary, ws = loader._scan(src)
    plan = loader_mod._plan(src, summary, layout)
    sections = (_native.Section * _native.MS_MAX_SECTIONS)()
    blocks = []
    k = 0
    for lay, (r0, r1) in zip(plan.layouts, plan.data_rows):
        blk = torch.empty((lay.n_keep, r1 - r0), dtype=torch.float64, device=dev)
        blocks.append(blk)
        s = sections[k]
        s.row_begin, s.row_end, s.num_cols, s.n_keep, s.d_out, s.stride = r0, r1, lay.num_cols, lay.n_keep, blk.data_ptr(), r1 - r0
        k += 1
    d_status = torch.empty(1, dtype=torch.int64, device=dev)
    d_summary = torch.empty(ctypes.sizeof(_native.ScanSummary), dtype=torch.uint8, device=dev)
    stream = torch.cuda.current_stream()
    sptr = ctypes.c_void_p(stream.cuda_stream)
    lib = _native.lib()

    def time_kernel(call, reps=10):
        for _ in range(3):
            call()
        torch.cuda.synchronize()
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
        for a, b in evs:
            a.record(stream)
            call()
            b.record(stream)
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in evs) / reps

    t_parse = time_kernel(lambda: lib.ms_parse(d_bytes.data_ptr(), n, ws.data_ptr(), sections, k, d_status.data_ptr(), sptr))
    t_scan = time_kernel(lambda: lib.ms_scan(d_bytes.data_ptr(), n, ws.data_ptr(), ws.numel(), d_summary.data_ptr(), sptr))
    assert int(d_status.item()) == -1, "parse reported an error on the benchmark input"

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    achieved = b_alg / (t_parse * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "parse_traffic.json")
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")

    # ---- end to end from host memory: the public batch API, H2D / compute / D2H overlapped
    run_e2e(3)
    e2e_steps = max(4, min(args.steps, 8))
    barrier()
    t_e2e = time.perf_counter()
    run_e2e(e2e_steps)
    torch.cuda.synchronize()
    ms_e2e = (time.perf_counter() - t_e2e) * 1e3
    if dist is not None:
        t = torch.tensor([ms_e2e], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_e2e = float(t.item())
    e2e_value = world * n * e2e_steps / (ms_e2e * 1e-3) / 1e9
    d2h = int(8 * n_kept)

    # ---- NMF-MU extension: rank sweep k=1..8 x 20 restarts on 200 x 16 envelopes (configs[3])
    nmf = bench_nmf(dev) if rank == 0 else None
    pipeline = None
    if rank == 0:
        try:
            pipeline = bench_pipeline(loader, d_bytes, n, layout)
        except Exception as exc:  # noqa: BLE001
            pipeline = {"error": f"{type(exc).__name__}: {exc}"}

    from_files = None
    if rank == 0 and world == 1:
        try:
            from_files = bench_files(loader, blob, layout)
        except Exception as exc:  # noqa: BLE001
            from_files = {"error": f"{type(exc).__name__}: {exc}"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        os.sched_setaffinity(0, all_cpus)
        gbs, cores, sample, _wall = cpu_baseline()
        cpu = {"value": gbs, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": WORKLOAD if layout == "T10" else layout, "csv_bytes_per_gpu": n, "kept_doubles_per_gpu": n_kept,
                "l2": "input (CSV) and output are each larger than the 126 MB L2; no explicit flush",
                "parallelism": f"{world} ranks, one trial per rank per step, no data-path collective",
                "host_cpus_bound": len(numa_cpus) if numa_cpus else None,
            },
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / e2e_steps, "steps": e2e_steps,
                    "api": "ViconLoader.load_many(pinned host CSV) -> host arrays, 3-stream pipeline, wall clock"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "ms_parse_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": b_alg, "kernel_ms": t_parse},
            "kernels_ms": {"ms_parse": t_parse, "ms_scan+resolve": t_scan},
            "nmf": nmf,
            "pipeline": pipeline,
            "from_files": from_files,
            "cpu_baseline": cpu,
            "clocks": clocks.summary(),
        }
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--layout", default="T10", help="synthetic layout (tools/synth_vicon.py LAYOUTS)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
