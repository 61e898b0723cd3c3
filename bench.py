#!/usr/bin/env python
"""Benchmark of the hot path on B200:  python bench.py --gpus N --steps K --warmup W [--impl reference]

Workload (BASELINE.json configs[1]): StandardAtmosphere.csv profile (66 layers), H2O+CO2+O3 absorption and
layer-integrated TUD over 700-1500 cm^-1 at 0.001 cm^-1 (800 000 grid points), N_angle=30.  One "step" = one TUD
spectrum of one atmosphere: Voigt line sum for all 66 layers (optical depth) + layer integration.  Each step
uses a different perturbed atmosphere (config 3's generator), so nothing can be cached between steps.  Lines are
synthetic (seed 1: 30 000 H2O + 80 000 CO2 + 100 000 O3) because the reference ships no line data.

Metric: Voigt line-gridpoint evals/s, where an eval is one (line, grid point) pair inside the reference's wing
window (misc/hapi.py:11131-11134), counted exactly on the device outside the timed region.  The whole step
(including the TUD integration) is inside the timed region; `tud_spectra_per_s` reports the same time as
spectra/s.  N>1: one process per GPU (torchrun), each rank integrates its own atmospheres (weak scaling) and the
ranks all-gather the (tau, Lu, Ld) spectra every step -- the path's one exchange.

`--impl reference` times the CPU oracle port (oracle/port.py: the reference's per-line numpy algorithm; the
reference itself is pure Python and cannot travel to the GPU box) on all host cores over a bounded sample of the
same workload.
"""
import argparse
import contextlib
import io
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

XMIN, XMAX, DV = 700.0, 1500.0, 0.001
N_ANGLE = 30
METRIC = "voigt_line_gridpoint_evals_per_s"
UNIT = "evals/s"
WORKLOAD = ("config2: StandardAtmosphere 66 layers, H2O+CO2+O3 (210k synthetic lines, seed 1), TUD over "
            "700-1500 cm^-1 @0.001 (800000 pts), N_angle=30")


# --------------------------------------------------------------------------------------------- CPU arm
def _cpu_task(args):
    """One CPU work item: Voigt sum of a slice of one molecule's lines for one layer, by the oracle port."""
    mol, lo, hi, T, p = args
    from oracle import port
    from radtxfr_b200 import synthetic as syn
    g = _cpu_task.__dict__
    if "tabs" not in g:
        g["tabs"] = syn.config2_tables()
        g["X"] = port.make_spectral_axis(XMIN, XMAX, DV)
    t = g["tabs"][mol]
    cols = {k: v[lo:hi] for k, v in t.items()}
    _, xs, ev, evw = port.absorption_coefficient_voigt(cols, T=T, p=p, OmegaGrid=g["X"], return_evals=True)
    return ev, evw, float(xs.sum())


def cpu_sample_tasks(lines_per_layer, layers=(0, 22, 44, 65), chunk=250, seed=0):
    from radtxfr_b200 import synthetic as syn
    atm = syn.standard_atmosphere()
    sizes = {1: 30000, 2: 80000, 3: 100000}
    total = sum(sizes.values())
    tasks = []
    rng = np.random.default_rng(seed)
    for L in layers:
        for mol, n in sizes.items():
            want = max(chunk, int(round(lines_per_layer * n / total)))
            nchunks = max(1, want // chunk)
            starts = rng.choice(max(1, n // chunk), size=min(nchunks, n // chunk), replace=False) * chunk
            for s in starts:
                tasks.append((mol, int(s), int(min(n, s + chunk)), float(atm["Ts"][L]), float(atm["Ps"][L] / 101325.0)))
    return tasks


def run_cpu_sample(lines_per_layer, cores, seed=0):
    """Returns (evals, seconds, n_tasks): the oracle port over a stratified sample, fanned over `cores`
    processes (the reference's own idiom: multiprocessing.Pool, Generate_LWIR_TUD.py:137-149)."""
    import multiprocessing as mp
    tasks = cpu_sample_tasks(lines_per_layer, seed=seed)
    ctx = mp.get_context("spawn")
    with ctx.Pool(cores) as pool:
        pool.map(_cpu_task, tasks[:cores])          # warm the workers (imports, table generation)
        t0 = time.perf_counter()
        res = pool.map(_cpu_task, tasks, chunksize=1)
        dt = time.perf_counter() - t0
    return sum(r[0] for r in res), dt, len(tasks)


def cpu_baseline(cores=None, budget_lines_per_core=2000):
    cores = cores or os.cpu_count() or 1
    lines = min(210000, budget_lines_per_core * cores)
    ev, dt, nt = run_cpu_sample(lines, cores)
    return {"value": ev / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "oracle/port.py Voigt sum, layers 0/22/44/65 x ~%d lines x full 800000-pt grid = %d evals "
                      "in %.1f s over %d processes" % (lines, ev, dt, cores)}


def reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    lines = min(210000, 2000 * cores)
    vals = []
    for i in range(args.warmup + args.steps):
        ev, dt, _ = run_cpu_sample(lines, cores, seed=i)
        if i >= args.warmup:
            vals.append((ev, dt))
    ev = sum(v[0] for v in vals)
    dt = sum(v[1] for v in vals)
    val = ev / dt
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "sample": "layers 0/22/44/65 x ~%d lines per step" % lines},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d evals in %.1f s over %d steps" % (ev, dt, args.steps)},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append((time.perf_counter(), ln.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        for t, ln in self.rows:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            if t0 - 0.05 <= t <= t1 + 0.15:
                try:
                    sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no sample inside the timed region"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


def banded_leg(args, rank, world, dev):
    """Config 5 (BASELINE.json configs[4]): 500-6000 cm^-1, --banded-points grid points x 60 layers, --banded-lines
    synthetic lines of eight gases, sharded by wavenumber band over the ranks with ONE all_gather
    (distributed.compute_TUD_banded) -- the strong-scaling partitioning of the path.  Reports the device-timed rate
    (max over ranks, results left in HBM), the end-to-end time with every rank copying the full spectra to numpy,
    rank 0's single-GPU time for the same problem (the efficiency denominator), and whether the gathered result equals
    the single-GPU result bit for bit on every rank; at N > 1 the atmosphere-sharded batch is checked the same way."""
    import torch
    import torch.distributed as dist
    from radtxfr_b200 import distributed as D, radiative_transfer as rt, synthetic as syn
    if world == 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        dist.init_process_group("nccl", rank=0, world_size=1, device_id=dev)
        own_group = True
    else:
        own_group = False
    atm = syn.standard_atmosphere()
    n_layers = 60
    tabs = syn.config5_tables(args.banded_lines)
    opts = dict(rt.options)
    t0 = time.perf_counter()
    db = rt.set_line_database(tabs, device=dev, opts=opts)
    ingest_s = time.perf_counter() - t0
    nx = args.banded_points
    dv = (6000.0 - 500.0) / nx
    kw = dict(Zs=atm["Zs"][:n_layers], Ts=atm["Ts"][:n_layers], Ps=atm["Ps"][:n_layers], PLs=atm["PLs"][:n_layers],
              MFs_VAL=atm["MFs_VAL"][:n_layers], MFs_ID=atm["MFs_ID"], DVOUT=dv, N_angle=N_ANGLE)
    reps = 3

    def timed(fn):
        for _ in range(4):                               # warm-up (workspaces, pinned staging, band-balance feedback)
            fn()
        best_dev, best_wall = 1e30, 1e30
        for _ in range(reps):
            dist.barrier(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            tw = time.perf_counter(); tw_epoch = time.time(); e0.record()
            out = fn()
            e1.record(); torch.cuda.synchronize()
            wall = time.perf_counter() - tw
            t = torch.tensor([e0.elapsed_time(e1), 1e3 * wall], dtype=torch.float64, device=dev)
            per_rank = [torch.zeros(3, dtype=torch.float64, device=dev) for _ in range(world)]
            dist.all_gather(per_rank, torch.tensor([e0.elapsed_time(e1), tw_epoch, 1e3 * wall], dtype=torch.float64, device=dev))
            dist.all_reduce(t, op=dist.ReduceOp.MAX)      # max over ranks
            if t[0].item() < best_dev:
                timed.detail = {"per_rank_ms": [float(x[0]) for x in per_rank],
                                "start_skew_ms": [1e3 * float(x[1] - per_rank[0][1]) for x in per_rank]}
            best_dev, best_wall = min(best_dev, t[0].item()), min(best_wall, t[1].item())
        return out, best_dev, best_wall

    (X, tau_d, lu_d, ld_d), ms_dev, _ = timed(lambda: D.compute_TUD_banded(500.0, 6000.0, dict(opts), to_host=False, **kw))
    dev_detail = getattr(timed, "detail", None)
    # how well the bands are balanced: per-rank compute times (before the gather) of the last call, from the feedback
    # records of compute_TUD_banded
    fbs = [v for k, v in D._band_feedback.items() if k[2] == X.size]
    bands = fbs[-1]["events"][2] if fbs else D.plan_bands(X.size, world)
    own_ms = 0.0
    if fbs:
        e0, e1, _ = fbs[-1]["events"]
        e1.synchronize()
        own_ms = e0.elapsed_time(e1)
    allms = [torch.zeros(1, dtype=torch.float64, device=dev) for _ in range(world)]
    dist.all_gather(allms, torch.tensor([own_ms], dtype=torch.float64, device=dev))
    rank_ms = [float(t.item()) for t in allms]
    (_, tau_h, lu_h, ld_h), _, ms_e2e = timed(lambda: D.compute_TUD_banded(500.0, 6000.0, dict(opts), **kw))
    # exact evals of the whole problem (outside any timed region)
    p, w = rt.layer_weights(db, kw["Ts"], kw["Ps"], kw["PLs"], kw["MFs_VAL"], kw["MFs_ID"])
    grid = torch.from_numpy(X).to(dev)
    ev, evw = db.count_evals(kw["Ts"], p, w, grid)
    del grid
    # the same problem on ONE GPU (every rank does it, so every rank checks its gathered copy): device time of the
    # Voigt sum + layer integration, then bit-for-bit comparison
    o1 = dict(opts); o1["voigt"] = {"tile_k": 8}
    ident, ms_one = True, None
    if world > 1:
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        rt.compute_TUD(500.0, 6000.0, opts=dict(o1), **kw)           # warm-up: workspaces of the full width
        e0.record()
        X1, t1, u1, d1 = rt.compute_TUD(500.0, 6000.0, opts=dict(o1), **kw)
        e1.record(); torch.cuda.synchronize()
        ms_one = e0.elapsed_time(e1)
        ident = bool(np.array_equal(t1, tau_h) and np.array_equal(u1, lu_h) and np.array_equal(d1, ld_h)
                     and np.array_equal(tau_d.cpu().numpy(), tau_h))
        del X1, t1, u1, d1
        # atmosphere-sharded batch (config 3's partitioning) against the unsharded batch on this rank
        Ts, MF = syn.perturbed_atmospheres(5, 2 * world + 1)
        kb = dict(Zs=atm["Zs"], Ps=atm["Ps"], PLs=atm["PLs"], MFs_ID=atm["MFs_ID"], DVOUT=dv * 8, N_angle=N_ANGLE)
        Xb, tb, ub, db_ = D.compute_TUD_batch_sharded(900.0, 1100.0, Ts, MF, dict(opts), **kb)
        Xs, ts, us, ds = rt.compute_TUD_batch(900.0, 1100.0, Ts, MF, opts=dict(opts), **kb)
        ident = ident and bool(torch.equal(tb, ts) and torch.equal(ub, us) and torch.equal(db_, ds))
    flag = torch.tensor([int(ident)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if own_group:
        dist.destroy_process_group()
    out = {"bit_identical": bool(flag.item() == 1) if world > 1 else None,
           "workload": "config5: 500-6000 cm^-1, %d points x %d layers, %d synthetic lines of 8 gases, band-sharded over %d "
                       "GPU(s), one all_gather of (tau, Lu, Ld)" % (nx, n_layers, args.banded_lines, world),
           "scaling": "strong", "evals": ev, "weideman_fraction": evw / max(1, ev), "ms": ms_dev, "evals_per_s": ev / (ms_dev * 1e-3),
           "band_points": [b - a for a, b in bands], "band_compute_ms": rank_ms,
           "timed_call_detail": dev_detail,
           "balance_history": [{"band_points": h[0], "compute_ms": [round(x, 2) for x in h[1]]} for h in fbs[-1].get("history", [])] if fbs else None,
           "host_prep_ms_rank0": fbs[-1].get("host_prep_ms") if fbs else None,
           "gather_ms_rank0": (fbs[-1]["tail_events"][0].elapsed_time(fbs[-1]["tail_events"][1]) if fbs and "tail_events" in fbs[-1] else None),
           "assemble_ms_rank0": (fbs[-1]["tail_events"][1].elapsed_time(fbs[-1]["tail_events"][2]) if fbs and "tail_events" in fbs[-1] else None),
           "e2e_ms": ms_e2e, "e2e_evals_per_s": ev / (ms_e2e * 1e-3),
           "e2e_note": "every rank copies the full (tau, Lu, Ld) = %d MB to fresh numpy arrays" % (3 * 8 * nx // 1000000),
           "timing": "best of %d, CUDA events on the launching stream / wall clock, max over ranks" % reps,
           "ingest_s": ingest_s}
    if ms_one is not None:
        out["single_gpu_ms"] = ms_one
        out["single_gpu_evals_per_s"] = ev / (ms_one * 1e-3)
        out["efficiency_vs_single_gpu"] = ms_one / (world * ms_dev)
        out["single_gpu_note"] = ("radiative_transfer.compute_TUD of the same problem on rank 0's GPU, numpy out (includes "
                                  "its 264 MB device-to-host copy, so the efficiency is slightly flattering; compare "
                                  "evals_per_s across N in SCALE for the strict figure)")
    return out


# --------------------------------------------------------------------------------------------- GPU arm
def gpu_arm(args):
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline()        # spawned workers; runs before any kernel of ours is in flight

    from radtxfr_b200 import engine, radiative_transfer as rt, synthetic as syn
    atm = syn.standard_atmosphere()
    tabs = syn.config2_tables()
    t_ing = time.perf_counter()
    db = engine.LineDB([tabs[m] for m in sorted(tabs)], dev)
    ingest_s = time.perf_counter() - t_ing
    X = rt.make_spectral_axis(XMIN, XMAX, DV)
    nX, nL = X.size, atm["Ts"].size
    grid = torch.from_numpy(X).to(dev)
    nsteps = args.warmup + args.steps
    Ts_all, MF_all = syn.perturbed_atmospheres(1000 + rank, nsteps)
    Ts_all[0], MF_all[0] = atm["Ts"], atm["MFs_VAL"]          # first (warm-up) step is the unperturbed profile
    states = []
    for i in range(nsteps):
        p, w = rt.layer_weights(db, Ts_all[i], atm["Ps"], atm["PLs"], MF_all[i], atm["MFs_ID"])
        states.append((torch.from_numpy(Ts_all[i].copy()).to(dev), torch.from_numpy(np.ascontiguousarray(p)).to(dev),
                       torch.from_numpy(w).to(dev)))
    od = torch.empty((nL, nX), dtype=torch.float64, device=dev)
    gathered = torch.empty((world, 3, nX), dtype=torch.float64, device=dev) if world > 1 else None
    h0 = engine.default_handle(dev)

    # The one collective of the path: every rank's three spectra to every rank.  It runs on a side
    # stream so that step k's gather overlaps step k+1's Voigt kernel; the timed region ends after the last gather.
    side = torch.cuda.Stream(device=dev) if world > 1 else None

    # Results go to two persistent (3, nX) buffers used in turn -- (tau, Lu, Ld) of a step are one contiguous block, the
    # gather reads it in place (no stack, no allocation inside the timed loop), and the buffer of step k is free again
    # for step k+2 once the gather of step k has fired its event.
    res = [torch.empty((3, nX), dtype=torch.float64, device=dev) for _ in range(2)]
    res_free = [None, None]

    def tud_into(k, T):
        b = res[k % 2]
        if res_free[k % 2] is not None:
            torch.cuda.current_stream(dev).wait_event(res_free[k % 2])
        return engine.tud(od, T[None, :], grid, atm["Zs"], [500.0], [1.0], N_ANGLE, out=(b[0], b[1], b[2]))

    def gather(k):
        done = torch.cuda.Event()
        done.record()
        with torch.cuda.stream(side):
            side.wait_event(done)
            dist.all_gather_into_tensor(gathered, res[k % 2])
            res_free[k % 2] = torch.cuda.Event()
            res_free[k % 2].record(side)

    def step(i):
        T, p, w = states[i]
        db.voigt_sum(T, p, w, grid, out=od, flags=args.flags)
        tau, lu, ld = tud_into(i, T)
        if world > 1:
            gather(i)
        return tau, lu, ld

    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    # exact eval counts of the timed steps (outside the timed region)
    E = Ew = 0
    for i in range(args.warmup, nsteps):
        T, p, w = states[i]
        e, ew = db.count_evals(T, p, w, grid)
        E += e; Ew += ew
    launches0 = db.handle.launches() + h0.launches()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 1)]
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ev[0].record()
    for k in range(args.steps):
        i = args.warmup + k
        T, p, w = states[i]
        db.voigt_sum(T, p, w, grid, out=od, flags=args.flags)
        ev[2 * k + 1].record()
        tau, lu, ld = tud_into(i, T)
        ev[2 * k + 2].record()
        if world > 1:
            gather(i)
    if world > 1:
        torch.cuda.current_stream(dev).wait_stream(side)      # the last gather is inside the timed region
    ev_end = torch.cuda.Event(enable_timing=True)
    ev_end.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t1 = time.perf_counter()
    ms_total = ev[0].elapsed_time(ev_end)
    v_steps = [ev[2 * k].elapsed_time(ev[2 * k + 1]) for k in range(args.steps)]
    t_steps = [ev[2 * k + 1].elapsed_time(ev[2 * k + 2]) for k in range(args.steps)]
    ms_voigt, ms_tud = sum(v_steps), sum(t_steps)

    def stats(v):
        # every step's device time as well: an outlier (one slow launch, a hiccup of the box) is then visible as such
        return {"min": float(np.min(v)), "median": float(np.median(v)), "max": float(np.max(v)),
                "slowest_step": int(np.argmax(v)), "all": [round(float(x), 3) for x in v]}
    launches = db.handle.launches() + h0.launches() - launches0
    tot = torch.tensor([ms_total, float(E), float(Ew), float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = tot.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tot.clone(); dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        ms_total = mx[0].item(); E_all, Ew_all, launches_all = sm[1].item(), sm[2].item(), sm[3].item()
    else:
        E_all, Ew_all, launches_all = float(E), float(Ew), float(launches)
    clk = clocks.stop(t0, t1) if rank == 0 else None

    # ---- end-to-end through the reference-facing API: numpy in, numpy out, copies inside the timed region.
    # N = 1: radiative_transfer.compute_TUD.  N > 1: distributed.compute_TUD_batch_sharded -- every step hands N
    # atmospheres (host arrays) to the ranks, each rank integrates its own, the spectra are all-gathered over NCCL and
    # rank 0 copies the gathered result of all N atmospheres to host memory.
    e2e = None
    if not args.no_e2e:
        from radtxfr_b200 import distributed as D
        opts = dict(rt.options)
        opts["LINE_DB"] = db
        kw = dict(Zs=atm["Zs"], Ps=atm["Ps"], PLs=atm["PLs"], MFs_ID=atm["MFs_ID"], DVOUT=DV, N_angle=N_ANGLE)
        if world == 1:
            def e2e_step(i):
                return rt.compute_TUD(XMIN, XMAX, opts=dict(opts), Ts=Ts_all[i], MFs_VAL=MF_all[i], **kw)
            h2d = 8 * (3 * nL + nL * db.n_iso)  # T, p, T (TUD) and column amounts; the spectral axis is cached on the device
            d2h = 8 * 3 * nX
            api = "radtxfr_b200.radiative_transfer.compute_TUD (numpy in, numpy out; line table resident)"
        else:
            # the same N atmospheres on every rank's host (rank r's draw for step i is atmosphere r of the step)
            allT = [torch.empty((nsteps, nL), dtype=torch.float64, device=dev) for _ in range(world)]
            allM = [torch.empty((nsteps,) + MF_all.shape[1:], dtype=torch.float64, device=dev) for _ in range(world)]
            dist.all_gather(allT, torch.from_numpy(Ts_all).to(dev))
            dist.all_gather(allM, torch.from_numpy(MF_all).to(dev))
            T_steps = torch.stack(allT, dim=1).cpu().numpy()          # [nsteps, world, nL]
            M_steps = torch.stack(allM, dim=1).cpu().numpy()

            def e2e_step(i):
                Xo, tau_d, lu_d, ld_d = D.compute_TUD_batch_sharded(XMIN, XMAX, T_steps[i], M_steps[i], dict(opts), **kw)
                # every rank holds all N spectra on its GPU after the gather and copies the atmosphere it owns to its host
                return rt._to_host(tau_d[rank], lu_d[rank], ld_d[rank])
            h2d = 8 * (3 * nL + nL * db.n_iso) * world
            d2h = 8 * 3 * nX * world
            api = ("radtxfr_b200.distributed.compute_TUD_batch_sharded (N atmospheres as numpy in on every rank; NCCL "
                   "all_gather of (tau, Lu, Ld) so that every GPU holds all N spectra; every rank copies the spectra of the "
                   "atmosphere it owns to numpy); byte counts are summed over the ranks")
        e2e_step(0)   # warm
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0e = time.perf_counter()
        for k in range(args.steps):
            e2e_step(args.warmup + k)
        torch.cuda.synchronize()
        dte = time.perf_counter() - t0e
        if world > 1:
            t = torch.tensor([dte], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dte = t.item()
        e2e = {"value": E_all / dte, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": 1e3 * dte / args.steps, "api": api}

    # ---- multi-GPU parity and the band-sharded strong-scaling leg (config 5), outside the timed region above
    banded = None
    if not args.no_banded:
        banded = banded_leg(args, rank, world, dev)

    if rank == 0:
        secs = ms_total * 1e-3
        value = E_all / secs
        flops = 16.0 * (E - Ew) + 200.0 * Ew            # algorithmic flops, SURVEY.md 8d (rank 0's steps)
        peak = db.handle.fp64_peak()
        achieved = flops / (ms_voigt * 1e-3)
        prof = os.path.join(ROOT, "profiles", "roofline_traffic.json")
        traffic, ncu = None, {}
        if os.path.isfile(prof):
            try:
                ncu = json.load(open(prof))
                traffic = ncu.get("voigt_tile_kernel_dram_bytes_per_launch")
            except Exception:
                traffic = None
        hbm_peak, hbm_src = 6650.0, "fallback of B200_PROFILING.md"
        try:
            hbm_peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
            hbm_src = "MEASURED_PEAKS.json"
        except Exception:
            pass
        voigt_bytes = 8.0 * nL * nX                      # algorithmic HBM bytes per launch: one double per (point, layer)
        tud_bytes = 8.0 * nL * nX + 24.0 * nX            # 552 B per grid point at 66 layers
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "atmospheres_per_step_per_gpu": 1,
                       "l2": "each step writes and re-reads a 422 MB optical-depth buffer (> 126 MB L2); a "
                             "different atmosphere every step",
                       "parallelism": "1 process per GPU, atmospheres sharded, one all_gather of (tau,Lu,Ld) per step"
                       if world > 1 else "single GPU"},
            "tud_spectra_per_s": world * args.steps / secs,
            "evals_per_step": E / args.steps, "weideman_fraction": Ew / max(1, E),
            "ms_voigt_per_step": ms_voigt / args.steps, "ms_tud_per_step": ms_tud / args.steps,
            "per_step_ms": {"voigt": stats(v_steps), "tud": stats(t_steps)},
            "ingest_s": ingest_s,
            "roofline": {"bound": "fp64", "kernel": "voigt_tile_kernel (K = 8 and K = 4 instantiations, picked per state on the device)", "achieved": achieved / 1e12,
                         "peak": peak / 1e12, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "measured in this run: dependent-free DFMA loop (rtx_measure_fp64_peak); "
                                        "MEASURED_PEAKS.json has no FP64 figure",
                         "flops_model": "16 flop per asymptote-branch eval, 200 per Weideman eval (SURVEY.md 8d)",
                         "note": "achieved counts the reference's ALGORITHMIC flops (every (line, grid point) pair of its "
                                 "windows); the kernel evaluates wing contributions at 32 Chebyshev nodes per 256-point slice "
                                 "(far-field blocks, at tile level for wings >= 4000 points; deviation < 1e-13) and so issues fewer operations than it is credited "
                                 "with -- frac above 1 means faster than a perfect FP64 pipe could run the reference's "
                                 "arithmetic.  The pipe-level view of the same kernel is in `issued`.",
                         "issued": {"fp64_pipe_pct_of_peak": ncu.get("voigt_tile_kernel_fp64_pipe_pct"),
                                    "issue_slots_active_pct": ncu.get("voigt_tile_kernel_issue_active_pct"),
                                    "source": "profiles/roofline_traffic.json (ncu --set full of this command, one launch)"},
                         "hbm_view": {"voigt_tile_kernel": {"algorithmic_bytes": voigt_bytes,
                                                            "achieved_GBps": voigt_bytes / (ms_voigt / args.steps) / 1e6,
                                                            "frac_of_hbm_peak": voigt_bytes / (ms_voigt / args.steps) / 1e6 / hbm_peak},
                                      "tud_kernel": {"algorithmic_bytes": tud_bytes,
                                                     "achieved_GBps": tud_bytes / (ms_tud / args.steps) / 1e6,
                                                     "frac_of_hbm_peak": tud_bytes / (ms_tud / args.steps) / 1e6 / hbm_peak,
                                                     "note": "exp-bound: 2113 FP64 exp per grid point (SURVEY.md 8d)"},
                                      "hbm_peak_GBps": hbm_peak, "hbm_peak_source": hbm_src,
                                      "note": "both kernels are FP64-pipe bound; their HBM traffic is <1% / ~2% of the "
                                              "time the copy rate would need"}},
            "clocks": clk, "gpu_launches": int(launches_all),
        }
        # layer-integration kernel against the same measured FP64 rate: SURVEY.md 8d counts 5.9e4 flop per grid point
        # (2113 exp x 25 + recurrences) at 66 layers and 30 angles; the kernel issues ~10 FP64 operations per exp
        tud_flops = 5.9e4 * nX
        line["roofline_tud"] = {"bound": "fp64", "kernel": "tud_kernel", "achieved": tud_flops / (ms_tud / args.steps * 1e-3) / 1e12,
                                "peak": peak / 1e12, "unit": "TFLOP/s", "frac": tud_flops / (ms_tud / args.steps * 1e-3) / peak,
                                "flops_model": "5.9e4 algorithmic flop per grid point: 2113 FP64 exp x 25 + 3 x 2046 "
                                               "recurrence + 200 (SURVEY.md 8d)",
                                "issued": {"fp64_pipe_pct_of_peak": ncu.get("tud_kernel_fp64_pipe_pct"),
                                           "note": "the kernel's exp costs 10 FP64-pipe operations, not the 25 of the model, "
                                                   "so the pipe is the honest denominator: see profiles/"}}
        if banded:
            line["multi_gpu_bit_identical"] = banded.pop("bit_identical")
            line["banded"] = banded
        if e2e:
            line["e2e"] = e2e
        if cpu:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--flags", type=int, default=0, help="RTX_VOIGT_* flags for rtx_voigt_sum")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-banded", action="store_true", help="skip the config-5 band-sharded leg and the multi-GPU parity check")
    ap.add_argument("--banded-points", type=int, default=11_000_000)
    ap.add_argument("--banded-lines", type=int, default=5_000_000)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        reference_arm(args)
    else:
        if args.warmup < 3:
            print("note: fewer than 3 warm-up steps requested", file=sys.stderr)
        gpu_arm(args)


if __name__ == "__main__":
    # The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its version banner to stdout
    # when NCCL_DEBUG is set), so file descriptor 1 is pointed at stderr for the whole run and the JSON line is
    # written to the real stdout at the end.
    sys.stdout.flush()
    _real_stdout = os.dup(1)
    os.dup2(2, 1)
    _buf = io.StringIO()
    with contextlib.redirect_stdout(_buf):
        main()
    sys.stdout.flush()
    os.write(_real_stdout, _buf.getvalue().encode())     # fd 1 stays on stderr: libraries also print at exit
    os.close(_real_stdout)
