#!/usr/bin/env python
"""bench.py - converged NMPC solves/sec of the batched solver (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--config cfg2|cfg3|cfg4|cfg5] [--scaling weak|strong]
    python bench.py --impl reference ...      CPU arm: the compiled oracle port (oracle/cxx, C++17 + OpenMP,
                                              all host cores) of the reference's IPOPT path on the SAME
                                              workload; CasADi/IPOPT is not installable offline

Configurations (BASELINE.json `configs`; cfg2 is the headline the metric is quoted on):
  cfg2  dynamic-bicycle NLP, N = 30, Ts = 0.05, 16384 problems per GPU (4096 reference-sampled Shanghai
        start states x 4 seeded perturbations), cold start.  A "step" = one mpcr_solve_batch call.
  cfg3  the same with the lane-bound row on t1_triple (16384 problems per GPU).
  cfg4  closed loop: 16384 parallel receding-horizon rollouts (blended plant, N = 15, warm-started) on
        t3 / t4; a "step" = --inner control steps of every rollout (project, refit, solve, plant).
  cfg5  GA fitness sweep: population 100 x the 5 tracks = 500 rollouts sharded over the ranks; a
        "step" = --inner control steps of every rollout + the all-gather of the fitness vector.
All configs use the options of the drop-in MPC class (control/MPC.py): restoration phase ON.  The
headline line also carries the figure without it (`no_restoration`).
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

TS = 0.05
BATCH = 16384
METRIC = "converged_nmpc_solves_per_sec"
UNIT = "solves/s"
TRACKS5 = ["shanghai_intl_circuit", "t1_triple", "t2_triple", "t3", "t4"]

CONFIGS = {
    "cfg2": dict(kind="solve", fixture="cfg2_shanghai_4096.npz", N=30, model="dynamic", lane=False,
                 workload="cfg2-shanghai-dynamic-N30-Ts0.05: 4096 reference-sampled start states "
                          "(seed 1234) x seeded perturbations, cold start"),
    "cfg3": dict(kind="solve", fixture="cfg3_t1_triple_512.npz", N=30, model="dynamic", lane=True,
                 workload="cfg3-t1_triple-dynamic-N30-Ts0.05-lane-rows: 512 reference-sampled start states "
                          "(seed 1234) x seeded perturbations, cold start"),
    "cfg4": dict(kind="rollout", N=15, model="dynamic", plant="blended", tracks=["t3", "t4"],
                 workload="cfg4-closed-loop-blended-plant-N15-Ts0.05: parallel warm-started receding-horizon "
                          "rollouts on t3 / t4 (start states seed 1234, v0 = 5 m/s)"),
    "cfg5": dict(kind="ga", N=15, model="dynamic", plant="blended", tracks=TRACKS5, population=100,
                 workload="cfg5-GA-fitness-sweep-N15-Ts0.05: population 100 x 5 tracks closed-loop rollouts "
                          "(GA/mpcGA.py parameter sets, seed 1234), fitness all-gathered"),
}


def flop_per_iter(N):
    return 4400 * N                      # algorithmic FLOP of one IPM iteration


def config_block(name, cfg, batch_per_gpu, scaling, inner=None):
    """The keys both arms print under "config" (identical in the two arms)."""
    if cfg["kind"] == "ga":
        scaling = "strong"               # population x tracks is a fixed job, sharded over the ranks
    c = {"workload": cfg["workload"], "name": name, "batch_per_gpu": batch_per_gpu, "N": cfg["N"], "Ts": TS,
         "model": cfg["model"], "lane_rows": bool(cfg.get("lane", False)), "restoration": True,
         "scaling": scaling}
    if inner is not None:
        c["control_steps_per_step"] = inner
    return c


# ---------------------------------------------------------------------------------------
def clocks_sampler(stop, samples, gpu_index, n_gpus=1):
    """Samples SM clocks, throttle reasons, power and temperature every 0.25 s during the timed region.
    Rank 0 samples every GPU of the job (ids gpu_index .. gpu_index + n_gpus - 1) in one nvidia-smi
    call, so that a slow rank can be told from its clocks (`clocks.per_gpu`)."""
    q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap,power.draw,temperature.gpu")
    ids = ",".join(str(gpu_index + i) for i in range(max(1, n_gpus)))
    while not stop.is_set():
        try:
            out = subprocess.run(["nvidia-smi", f"--id={ids}", f"--query-gpu={q}",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                 timeout=5).stdout.strip()
            rows = [[x.strip() for x in line.split(",")] for line in out.splitlines() if line.strip()]
            if rows:
                samples.append(rows)
        except Exception:
            pass
        stop.wait(0.25)


def _num(v):
    try:
        return float(v)
    except (TypeError, ValueError):
        return None


def summarize_clocks(samples):
    """`sm_mhz` / `sm_max_mhz` / `reasons` describe the first GPU (rank 0's) as before - reasons are
    collected over ALL sampled GPUs -, `per_gpu` the medians of every GPU of the job."""
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    n = min(len(rows) for rows in samples)
    per_gpu, reasons = [], []
    for g in range(n):
        col = lambda i: [v for v in (_num(rows[g][i]) for rows in samples if len(rows[g]) > i) if v is not None]
        sm, mx, pw, tc = col(0), col(1), col(6), col(7)
        rs = [nm for i, nm in enumerate(names)
              if any(len(rows[g]) > 2 + i and rows[g][2 + i].lower().startswith("active") for rows in samples)]
        reasons += [r for r in rs if r not in reasons]
        per_gpu.append({"sm_mhz": float(np.median(sm)) if sm else None, "sm_min_mhz": min(sm) if sm else None,
                        "sm_max_mhz": max(mx) if mx else None, "power_w": float(np.median(pw)) if pw else None,
                        "temp_c": float(np.median(tc)) if tc else None, "reasons": rs})
    out = {"sm_mhz": per_gpu[0]["sm_mhz"], "sm_max_mhz": per_gpu[0]["sm_max_mhz"], "reasons": reasons,
           "samples": len(samples)}
    if n > 1:
        out["per_gpu"] = per_gpu
    else:
        out["power_w"], out["temp_c"] = per_gpu[0]["power_w"], per_gpu[0]["temp_c"]
    return out


def pin_to_gpu_numa(local):
    """Run this rank on the host cores next to its GPU (PCIe root complex's local_cpulist)."""
    try:
        import torch
        bus = torch.cuda.get_device_properties(local).pci_bus_id
        dom = getattr(torch.cuda.get_device_properties(local), "pci_domain_id", 0)
        path = f"/sys/bus/pci/devices/{dom:04x}:{bus:02x}:00.0/local_cpulist"
        txt = open(path).read().strip()
        cpus = set()
        for part in txt.split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return f"{len(cpus)} cores of {txt}"
    except Exception as exc:
        return f"not pinned ({type(exc).__name__})"
    return "not pinned"


# ---------------------------------------------------------------------------------------
# CPU arm: the compiled oracle (oracle/cxx) - bench.py's cpu_baseline leg and --impl reference
# ---------------------------------------------------------------------------------------
def host_cores():
    """Host threads this process may use (torchrun exports OMP_NUM_THREADS=1: the CPU arm asks for
    its threads explicitly)."""
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cxx_solve(batch, cfg, threads=0):
    from oracle import cxx
    threads = threads or host_cores()
    t0 = time.perf_counter()
    r = cxx.solve_batch(batch["x0"], batch["u_prev"], batch["s0"], batch["cx"], batch["cy"], N=cfg["N"], Ts=TS,
                        model=cfg["model"], lane=cfg.get("lane", False),
                        max_err=batch["max_err"] if cfg.get("lane") else None, restoration=True,
                        threads=threads)
    dt = time.perf_counter() - t0
    return int((r["status"] >= 0).sum()), dt


def _numpy_one(args):
    from oracle.nlp import RacingNLP
    from oracle import ipm
    x0, up, s0, cx, cy, N = args
    r = ipm.solve(RacingNLP(x0, up, s0, cx, cy, N=N, Ts=TS))
    return r.status


def numpy_oracle_rate(batch, n, cores, N):
    """Context figure: the numpy oracle (oracle/ipm.py) in a process pool.  "spawn": this process has
    live CUDA and OpenMP runtimes, neither survives a fork."""
    from multiprocessing import get_context
    env = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS")}
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    jobs = [(batch["x0"][b], batch["u_prev"][b], batch["s0"][b], batch["cx"][b], batch["cy"][b], N) for b in range(n)]
    with get_context("spawn").Pool(cores) as pool:
        pool.map(_numpy_one, jobs[:cores])
        t0 = time.perf_counter()
        st = pool.map(_numpy_one, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    for k, v in env.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    return sum(1 for s in st if s >= 0) / dt


class CpuClosedLoop:
    """The reference's per-step pipeline on the host for R vehicles: projection + degree-4 refit
    (oracle/splines.py = the reference's SciPy code path), the solve (oracle/cxx, all cores), the
    plant step (oracle/plant.py)."""

    def __init__(self, track, cfg):
        from oracle import splines as osp
        self.cl = osp.Centerline(track, lanes=False)
        self.cfg = cfg

    def run(self, x0, s0, params, K):
        from oracle import cxx, plant
        R = x0.shape[0]
        x, prog = x0.copy(), s0.copy()
        u_prev = np.zeros((R, 2))
        u_warm = None
        conv = 0
        L = self.cl.length
        for _ in range(K):
            cx, cy = np.empty((R, 5)), np.empty((R, 5))
            for r in range(R):
                lo, hi = (prog[r] - 2) % L, (prog[r] + 2) % L
                b = (lo, hi) if lo < hi else (hi, lo)
                if b[1] - b[0] <= 5:
                    prog[r], _ = self.cl.projection_local(x[r, 0], x[r, 1], bounds=b)
                else:
                    prog[r], _ = self.cl.projection_global(x[r, 0], x[r, 1], seed=0)
                cx[r] = self.cl.x_as_coeffs(prog[r] - 5, 45, 4)
                cy[r] = self.cl.y_as_coeffs(prog[r] - 5, 45, 4)
            res = cxx.solve_batch(x, u_prev, prog, cx, cy, N=self.cfg["N"], Ts=TS, model=self.cfg["model"],
                                  params=params, u_warm=u_warm, restoration=True, threads=host_cores())
            conv += int((res["status"] >= 0).sum())
            u0 = res["U"][:, :, 0]
            for r in range(R):
                x[r] = plant.blended_step(x[r], u0[r], TS)
            u_prev = u0.copy()
            u_warm = res["U"]
        return conv


def rollout_starts(cfg, R, rank=0, world=1):
    """Start states of config 4 / 5: per track, s0 ~ U[0, L) on the centerline, tangent yaw, v0 = 5."""
    from mpc_racing_b200.sharding import shard_range
    tracks = cfg["tracks"]
    out = []
    if cfg["kind"] == "ga":
        from mpc_racing_b200.GA import mpcGA
        P = cfg["population"]
        pop = mpcGA.initialize_population(P, mpcGA.num_parameters, np.random.RandomState(1234))
        prm = mpcGA.decode(pop)
        pairs = [(t, i) for t in range(len(tracks)) for i in range(P)]      # (track, individual)
        lo, hi = shard_range(len(pairs), rank, world)
        mine = pairs[lo:hi]
        for t in sorted(set(p[0] for p in mine)):
            ind = [i for (tt, i) in mine if tt == t]
            out.append((tracks[t], len(ind), prm[ind], 1234 + t))
        return out, len(pairs)
    per = R // len(tracks)
    for j, t in enumerate(tracks):
        out.append((t, per if j < len(tracks) - 1 else R - per * (len(tracks) - 1), None, 1234 + 17 * rank + j))
    return out, R


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    cores = host_cores()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    B = args.batch if args.scaling == "weak" else max(1, args.batch // max(1, world))
    conv = 0
    if cfg["kind"] == "solve":
        from mpc_racing_b200.workload import make_batch
        batch = make_batch(B, fixture=cfg["fixture"])
        for _ in range(args.warmup):
            cxx_solve({k: v[:512] for k, v in batch.items()}, cfg)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            c, _ = cxx_solve(batch, cfg)
            conv += c
        dt = time.perf_counter() - t0
        sample = (f"every step = the full rank-0 batch of the GPU arm ({B} problems), oracle/cxx "
                  f"(C++17, AD derivatives, block-tridiagonal LDL^T, restoration on), OpenMP over problems")
        conf = config_block(args.config, cfg, B, args.scaling)
    else:
        # closed loop on the host: a bounded sample of the GPU arm's rollouts (the per-vehicle SciPy
        # projection / refit of the reference is Python), same control steps per step
        from mpc_racing_b200.splines.ParameterizedCenterline import ParameterizedCenterline  # noqa: F401 (tables)
        R_total = args.batch if cfg["kind"] == "rollout" else None
        groups, total = rollout_starts(cfg, R_total or 0, 0, 1)
        per_track = 32
        loops = []
        from oracle import splines as osp
        for track, n, prm, seed in groups:
            n = min(n, per_track)
            cl = CpuClosedLoop(track, cfg)
            rng = np.random.default_rng(seed)
            s0 = rng.uniform(0, cl.cl.length, n)
            x0 = np.stack([cl.cl.Gx(s0), cl.cl.Gy(s0), np.arctan2(cl.cl.dGy(s0), cl.cl.dGx(s0)),
                           np.full(n, 5.0), np.zeros(n), np.zeros(n)], axis=1)
            loops.append((cl, x0, s0, None if prm is None else prm[:n]))
        for _ in range(args.warmup):
            for cl, x0, s0, prm in loops[:1]:
                cl.run(x0[:4], s0[:4], None if prm is None else prm[:4], 1)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            for cl, x0, s0, prm in loops:
                conv += cl.run(x0, s0, prm, args.inner)
        dt = time.perf_counter() - t0
        sample = (f"{sum(l[1].shape[0] for l in loops)} of the GPU arm's {total} rollouts x {args.inner} control "
                  f"steps per step: SciPy projection + refit (oracle/splines.py), oracle/cxx solve (OpenMP), "
                  f"numpy plant step")
        conf = config_block(args.config, cfg, args.batch if cfg["kind"] == "rollout" else total, args.scaling,
                            args.inner)
    value = conv / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps),
        "higher_is_better": True, "scaling": conf["scaling"], "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": conf,
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "port": "oracle/cxx",
                         "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------
# GPU arm
# ---------------------------------------------------------------------------------------
def rank_stats(vals, world, dist, dev):
    """min / median / max over ranks of a per-rank vector (gathered once, after the timed region)."""
    import torch
    t = torch.tensor(vals, dtype=torch.float64, device=dev)
    if world > 1:
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        m = torch.stack(parts).cpu().numpy()
    else:
        m = t.cpu().numpy()[None]
    return m


def run_gpu(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    pinned = pin_to_gpu_numa(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = CONFIGS[args.config]
    if cfg["kind"] == "solve":
        run_gpu_solve(args, cfg, torch, dist, world, rank, local, dev, pinned)
    else:
        run_gpu_rollout(args, cfg, torch, dist, world, rank, local, dev, pinned)
    if world > 1:
        dist.destroy_process_group()


def _barrier(torch, dist, world, dev):
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)


def run_gpu_solve(args, cfg, torch, dist, world, rank, local, dev, pinned):
    from mpc_racing_b200.solver import BatchedMPC
    from mpc_racing_b200.workload import make_batch
    from mpc_racing_b200.sharding import shard_range
    N = cfg["N"]
    # weak: B problems per GPU; strong: B problems in all.  Rank r owns a contiguous index range of the
    # global batch (independent problems, no collective on the solve path).
    total = args.batch * world if args.scaling == "weak" else args.batch
    full = make_batch(total, fixture=cfg["fixture"])
    lo, hi = shard_range(total, rank, world)
    B = hi - lo
    keys = ("x0", "u_prev", "s0", "cx", "cy") + (("max_err",) if cfg["lane"] else ())
    host = {k: np.ascontiguousarray(full[k][lo:hi]) for k in keys}
    mk = lambda on_device, resto: BatchedMPC(N=N, Ts=TS, model=cfg["model"], device=local, on_device=on_device,
                                             lane_constraint=cfg["lane"], restoration=resto)
    dsolver, hsolver, plain = mk(True, True), mk(False, True), mk(True, False)
    din = {k: torch.from_numpy(host[k]).to(dev) for k in keys}
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    def solve(s, out=None):
        return s.solve(din["x0"], din["u_prev"], din["s0"], din["cx"], din["cy"],
                       max_err=din.get("max_err"), out=out)

    def as_out(res):
        return dict(U=res.U, X=res.X, S=res.S, eC=res.eC, eL=res.eL, lam_g=res.lam_g, obj=res.obj,
                    status=res.status, iters=res.iters)

    res = solve(dsolver)
    out = as_out(res)
    for _ in range(max(0, args.warmup - 1)):
        res = solve(dsolver, out)
    stop = threading.Event()
    samples = []
    th = threading.Thread(target=clocks_sampler, args=(stop, samples, local, world), daemon=True)
    if rank == 0:
        th.start()
    # ---- timed region: K steps, inputs resident in HBM; L2 flushed between steps (untimed)
    _barrier(torch, dist, world, dev)
    step_ms, kern_ms = [], []
    for _ in range(args.steps):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        res = solve(dsolver, out)
        e1.record()
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        kern_ms.append(dsolver.last_kernel_ms())
    _barrier(torch, dist, world, dev)
    serial_ms = float(sum(step_ms))
    status = res.status.cpu().numpy()
    iters = res.iters.cpu().numpy()
    conv_local = int((status >= 0).sum())
    launch = dsolver.launch_info()
    # ---- timed region of the headline: the same K steps through the solve pipeline (--pipeline-depth, default 3)
    #      (mpc_racing_b200/pipeline.py): step i+1 is enqueued on a second handle / stream while step i
    #      runs, so the tail of one persistent launch (its last long problems) is filled by the next
    #      batch.  Inputs resident in HBM; the L2 flush precedes every step on that step's stream.
    from mpc_racing_b200.pipeline import SolvePipeline
    mkpipe = lambda host_side: SolvePipeline(depth=max(1, args.pipeline_depth), device=local, host=host_side, N=N, Ts=TS, model=cfg["model"],
                                             lane_constraint=cfg["lane"], restoration=True)
    pipe = mkpipe(False)
    psub = lambda: pipe.submit(din["x0"], din["u_prev"], din["s0"], din["cx"], din["cy"], max_err=din.get("max_err"))
    for _ in range(max(2, args.pipeline_depth)):
        psub()
    pipe.drain()
    _barrier(torch, dist, world, dev)
    main = torch.cuda.current_stream(dev)
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record(main)
    for st in pipe.streams:
        st.wait_stream(main)
    last = -1
    for _ in range(args.steps):
        with torch.cuda.stream(pipe.next_stream()):
            flush.zero_()
        last = psub()
    pipe.join(main)
    e1.record(main)
    e1.synchronize()
    total_ms = e0.elapsed_time(e1)
    _barrier(torch, dist, world, dev)
    pres_ = pipe.result(last)
    pipeline_identical = bool(torch.equal(pres_.status, res.status) and torch.equal(pres_.U, res.U)
                              and torch.equal(pres_.obj, res.obj))
    pipe.close()
    # ---- the same without the restoration phase (round-1 headline configuration), for context
    pres = solve(plain)
    pout = as_out(pres)
    pk = []
    for _ in range(3):
        flush.zero_()
        solve(plain, pout)
        pk.append(plain.last_kernel_ms())
    plain_conv = int((pres.status >= 0).sum().item())
    plain_ms = float(np.mean(pk[1:]))
    # ---- e2e: the same step through the C ABI with HOST buffers (pinned), H2D + D2H inside
    pin = {k: torch.from_numpy(host[k]).pin_memory().numpy() for k in keys}
    hout = hsolver.alloc_host_outputs(B, want_lam_g=True, pinned=True)

    def hsolve(p, o):
        return hsolver.solve(p["x0"], p["u_prev"], p["s0"], p["cx"], p["cy"], max_err=p.get("max_err"), out=o)

    hsolve(pin, hout)
    # the same pipeline through the C ABI with host arrays (mpcr_solve_batch_async): the copies of
    # one step overlap the kernels of the other; every step's inputs go H2D and its results D2H
    hpipe = mkpipe(True)
    hsub = lambda: hpipe.submit(pin["x0"], pin["u_prev"], pin["s0"], pin["cx"], pin["cy"], max_err=pin.get("max_err"))
    for _ in range(max(2, args.pipeline_depth)):
        hsub()
    hpipe.drain()
    _barrier(torch, dist, world, dev)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        hlast = hsub()
    hpipe.drain()
    e2e_s = time.perf_counter() - t0
    _barrier(torch, dist, world, dev)
    hres = hpipe.result(hlast)
    e2e_conv = int((hres.status >= 0).sum())
    e2e_identical = bool(np.array_equal(hres.U, hout["U"]) and np.array_equal(hres.status, hout["status"]))
    hpipe.close()
    # ---- p50 single-solve latency (BASELINE.json metric, second half): B = 1 through the C ABI
    #      with host buffers, H2D + kernel + D2H inside each timed call (rank 0 only)
    lat = None
    if rank == 0:
        o1 = hsolver.alloc_host_outputs(1, want_lam_g=True, pinned=True)
        ts = []
        for i in range(60):
            j = i % min(B, 50)
            one = {k: pin[k][j:j + 1] for k in pin}
            t1 = time.perf_counter()
            hsolve(one, o1)
            ts.append(time.perf_counter() - t1)
        ts = np.array(ts[10:]) * 1e3
        lat = {"p50_ms": float(np.percentile(ts, 50)), "p90_ms": float(np.percentile(ts, 90)),
               "batch": 1, "calls": int(ts.size), "includes": "H2D + solve kernel + D2H, host pointers"}
    # ---- the production use (agent.py:138-205): ONE vehicle in closed loop, N = 15, warm-started from
    #      the previous plan; a tick = prepare (projection + refit) + solve + plant step, host-synchronised
    tick = None
    if rank == 0 and not cfg["lane"]:
        try:
            from mpc_racing_b200.rollout import ClosedLoopRollouts, start_states
            from mpc_racing_b200.splines.ParameterizedCenterline import ParameterizedCenterline
            cl = ParameterizedCenterline("t4", lanes=False, error=False, device=local)
            ro = ClosedLoopRollouts(cl, N=15, Ts=TS, plant="blended", device=local, restoration=True, use_graph=False)
            x1, s1 = start_states(cl, 1, seed=4321, v0=10.0)
            S1 = ro.begin(x1, s1)
            tt = []
            for i in range(120):
                t1 = time.perf_counter()
                with torch.cuda.stream(ro.stream):
                    ro._tick(S1, S1.k, ro.stream)
                    S1.k += 1
                ro.stream.synchronize()
                tt.append(time.perf_counter() - t1)
            tt = np.array(tt[20:]) * 1e3
            tick = {"p50_ms": float(np.percentile(tt, 50)), "p90_ms": float(np.percentile(tt, 90)), "ticks": int(tt.size),
                    "what": "one vehicle on t4, N = 15, blended plant, warm-started: prepare + solve + apply, "
                            "synchronised every tick (the reference agent's per-tick MPC time, agent.py:170-185)"}
            ro.close()
        except Exception as exc:           # the closed-loop probe is context, never a reason to lose the line
            tick = {"unavailable": repr(exc)[:200]}
    h2d = sum(pin[k].nbytes for k in pin)
    d2h = sum(v.nbytes for v in hout.values() if v is not None)
    stop.set()
    # ---- reduce over ranks: max time, summed work (the only collectives: after the timed regions)
    t = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    c = torch.tensor([conv_local, e2e_conv, int(iters.sum()), B, plain_conv], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    per_rank = rank_stats([float(np.mean(kern_ms)), total_ms / args.steps, e2e_s * 1e3 / args.steps, plain_ms,
                           float(B), float(iters.sum()), float(iters.max()),
                           float(np.sort(iters)[-max(1, B // 1000):].mean()), float(np.mean(step_ms))], world, dist, dev)
    total_ms_max, e2e_ms_max = t.tolist()
    conv_all, e2e_conv_all, _, b_all, plain_conv_all = c.tolist()
    if rank != 0:
        return
    value = conv_all * args.steps / (total_ms_max * 1e-3)
    e2e_value = e2e_conv_all * args.steps / (e2e_ms_max * 1e-3)
    # roofline of the dominant kernel: algorithmic fp64 FLOP / measured launch time (rank 0)
    kms = float(np.mean(kern_ms))
    flop = float(iters.sum()) * flop_per_iter(N)
    achieved = flop / (kms * 1e-3) / 1e12
    peak = dsolver.fp64_peak_tflops()
    bytes_alg = B * (84 + 736) * 8.0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "solve_kernel_traffic.json")))["bytes_per_launch"]
    except Exception:
        pass
    cpu = None
    if not args.no_cpu_baseline:
        cores = host_cores()
        conv, dt = cxx_solve(host, cfg)
        cpu = {"value": conv / dt, "unit": UNIT, "cores": cores, "kind": "port", "port": "oracle/cxx",
               "sample": f"rank 0's whole batch ({B} problems, same options), oracle/cxx (C++17 + OpenMP, "
                         f"independent of the kernel source), {dt:.1f} s"}
        try:
            n = max(4 * cores, 32)
            cpu["numpy_port_value"] = numpy_oracle_rate(host, n, cores, N) if not cfg["lane"] else None
            cpu["numpy_port_sample"] = f"first {n} problems, oracle/ipm.py, {cores} processes (context only)"
        except Exception as exc:
            cpu["numpy_port_value"] = None
            cpu["numpy_port_sample"] = repr(exc)[:120]
    # ipm_iterations: the work of a rank's slice - a rank whose kernel takes longer because its problems
    # need more iterations is not a straggler
    # longest_solve_iterations / top_0.1pct_mean_iterations: the tail of a rank's slice - the launch cannot end
    # before its longest problem does (one warp, ~0.1 ms per iteration), however the rest is balanced
    names = ["kernel_ms", "step_ms", "e2e_step_ms", "kernel_ms_no_restoration", "problems", "ipm_iterations",
             "longest_solve_iterations", "top_0.1pct_mean_iterations", "serial_step_ms"]
    ranks = {n: {"min": float(per_rank[:, i].min()), "median": float(np.median(per_rank[:, i])),
                 "max": float(per_rank[:, i].max()), "per_rank": [float(v) for v in per_rank[:, i]]}
             for i, n in enumerate(names)}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms_max / max(1, args.steps),
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": config_block(args.config, cfg, args.batch if args.scaling == "weak" else B, args.scaling),
        "value_is": "restoration phase ON (the drop-in MPC class's options); the K steps go through a "
                    f"{max(1, args.pipeline_depth)}-deep pipeline (one handle and stream per slot, mpc_racing_b200/pipeline.py): the tail of one launch overlaps "
                    "the head of the next.  `serial` = one launch at a time, each step timed on its own",
        "serial": {"value": conv_all / (float(per_rank[:, 8].max()) * 1e-3), "unit": UNIT,
                   "ms_per_step": float(per_rank[:, 8].max()),
                   "note": "max over ranks of the mean serial step; kernel_ms and the roofline are measured on these launches"},
        "detail": {"l2": "flushed before every timed step (256 MB write on the step's stream, inside the timed region)",
                   "pipeline_results_bit_identical_to_serial": pipeline_identical and e2e_identical,
                   "pipeline_depth": max(1, args.pipeline_depth),
                   "converged_frac": conv_all / b_all, "mean_iters": float(iters.mean()), "launch": launch,
                   "host_affinity": pinned},
        "no_restoration": {"value": plain_conv_all / (float(per_rank[:, 3].max()) * 1e-3), "unit": UNIT,
                           "kernel_ms": float(per_rank[:, 3].max()), "converged_frac": plain_conv_all / b_all,
                           "note": "kernel-side, the build without the restoration phase (round-1 headline configuration)"},
        "clocks": summarize_clocks(samples),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": args.steps * max(1, launch["kernels_per_solve"]),
        "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                     "frac": achieved / peak if peak else None, "traffic": traffic,
                     "peak_source": "measured here: DFMA-chain microbenchmark (mpcr_fp64_peak_tflops); "
                                    "MEASURED_PEAKS.json has no fp64 entry",
                     "kernel": "mpcr::solve_kernel", "kernel_ms": kms, "flop_per_launch": flop,
                     "kernel_ms_is": "mean duration of the solve launches of the serial steps (CUDA events around each launch, "
                                     "rank 0); in the pipelined region launches overlap, so a launch has no duration of its own",
                     "pipelined": {"achieved": flop / (total_ms / max(1, args.steps) * 1e-3) / 1e12,
                                   "frac": (flop / (total_ms / max(1, args.steps) * 1e-3) / 1e12) / peak if peak else None,
                                   "note": "the same FLOP over rank 0's time per step of the pipelined timed region"},
                     "hbm": {"achieved_gbs": bytes_alg / (kms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                             "frac": bytes_alg / (kms * 1e-3) / 1e9 / hbm_peak,
                             "algorithmic_bytes_per_launch": bytes_alg}},
        "ranks": ranks,
        "cpu_baseline": cpu,
        "single_solve_latency": lat,
        "closed_loop_tick_latency": tick,
    }
    print(json.dumps(line), flush=True)


def run_gpu_rollout(args, cfg, torch, dist, world, rank, local, dev, pinned):
    """cfg4 / cfg5: closed-loop control steps; every control step contains exactly one NMPC solve."""
    from mpc_racing_b200.rollout import ClosedLoopRollouts, start_states
    from mpc_racing_b200.splines.ParameterizedCenterline import ParameterizedCenterline
    from mpc_racing_b200.sharding import shard_range
    R = args.batch
    if cfg["kind"] == "rollout" and args.scaling == "strong":
        lo, hi = shard_range(args.batch, rank, world)
        R = hi - lo
    groups, total = rollout_starts(cfg, R, rank, world)
    split = max(1, args.engines_per_track)
    engines = []
    for track, n, prm, seed in groups:
        cl = ParameterizedCenterline(track, lanes=False, error=False, device=local)
        x0, s0 = start_states(cl, n, seed=seed, v0=5.0)
        # --engines-per-track: the track's fleet as independent sub-fleets (same vehicles), each engine on its
        # own stream, so that the tail of one engine's solve launch is filled by the others
        for q in range(split):
            lo, hi = shard_range(n, q, split)
            if hi <= lo:
                continue
            ro = ClosedLoopRollouts(cl, N=cfg["N"], Ts=TS, nlp_model=cfg["model"], plant=cfg["plant"], device=local,
                                    restoration=True)
            engines.append(dict(ro=ro, x0=x0[lo:hi], s0=s0[lo:hi], prm=None if prm is None else prm[lo:hi],
                                n=hi - lo, track=track))
    K = args.inner
    fit_local = None

    def one_step(state_in_hbm):
        """K control steps of every rollout of this rank; the tracks' engines run concurrently, each on
        its own stream.  Returns (converged solves, fitness vector)."""
        conv = 0
        fit = []
        sess = []
        for e in engines:
            s = e["sess"] if state_in_hbm else e["ro"].begin(e["x0"], e["s0"], params=e["prm"])   # H2D of the start states
            e["c0"] = e["ro"].converged_so_far(s)
            e["ro"].advance(s, K, sync=False)
            sess.append(s)
        for e, s in zip(engines, sess):
            if state_in_hbm:
                e["ro"].stream.synchronize()
                conv += e["ro"].converged_so_far(s) - e["c0"]
                fit.append(s.travelled)
            else:
                r = e["ro"].fetch(s)                                          # D2H of the result
                conv += r["converged_solves"] - e["c0"]
                fit.append(r["travelled"])
        return conv, fit

    for e in engines:
        e["sess"] = e["ro"].begin(e["x0"], e["s0"], params=e["prm"])
    for _ in range(args.warmup):
        one_step(True)
    stop = threading.Event()
    samples = []
    th = threading.Thread(target=clocks_sampler, args=(stop, samples, local, world), daemon=True)
    if rank == 0:
        th.start()
    _barrier(torch, dist, world, dev)
    conv_local = 0
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        c, fit = one_step(True)
        conv_local += int(c)
        if cfg["kind"] == "ga":
            # the only exchange of the fitness sweep: the fitness vector of all (track, individual) pairs
            f = torch.cat(fit)
            pad = torch.zeros((total + world - 1) // world, dtype=torch.float64, device=dev)
            pad[:f.numel()] = f
            if world > 1:
                parts = [torch.empty_like(pad) for _ in range(world)]
                dist.all_gather(parts, pad)
                fit_local = torch.cat(parts)
            else:
                fit_local = pad
    e1.record()
    e1.synchronize()
    total_ms = e0.elapsed_time(e1)
    _barrier(torch, dist, world, dev)
    solve_ms = float(np.mean([e["ro"].solver.last_kernel_ms() for e in engines]))
    # kernels of this library per control step: prepare + apply + the solve's 1 (or 3 with the dispatch order)
    launches = sum(2 + max(1, e["ro"].solver.launch_info()["kernels_per_solve"]) for e in engines) * K * args.steps
    # ---- e2e: the public API with host arrays in and out every step
    _barrier(torch, dist, world, dev)
    t0 = time.perf_counter()
    e2e_conv = 0
    for _ in range(args.steps):
        c, _ = one_step(False)
        e2e_conv += int(c)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    _barrier(torch, dist, world, dev)
    stop.set()
    n_local = sum(e["n"] for e in engines)
    t = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    c = torch.tensor([conv_local, e2e_conv, n_local], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    per_rank = rank_stats([total_ms / args.steps, solve_ms, float(n_local)], world, dist, dev)
    if rank != 0:
        return
    total_ms_max, e2e_ms_max = t.tolist()
    conv_all, e2e_all, n_all = c.tolist()
    value = conv_all / (total_ms_max * 1e-3)
    h2d = sum(e["x0"].nbytes + e["s0"].nbytes + (0 if e["prm"] is None else e["prm"].nbytes) for e in engines)
    d2h = sum(e["ro"].fetch_bytes(e["n"]) for e in engines)
    names = ["step_ms", "last_solve_kernel_ms", "rollouts"]
    ranks = {n: {"min": float(per_rank[:, i].min()), "median": float(np.median(per_rank[:, i])),
                 "max": float(per_rank[:, i].max()), "per_rank": [float(v) for v in per_rank[:, i]]}
             for i, n in enumerate(names)}
    N = cfg["N"]
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms_max / max(1, args.steps),
        "higher_is_better": True, "scaling": args.scaling if cfg["kind"] == "rollout" else "strong",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_block(args.config, cfg, args.batch if cfg["kind"] == "rollout" else total, args.scaling, K),
        "value_is": "converged solves inside closed-loop control steps (one solve per vehicle and control step); "
                    "restoration phase ON",
        "detail": {"rollouts_all_ranks": n_all, "control_steps_per_sec": n_all * K * args.steps / (total_ms_max * 1e-3),
                   "converged_frac": conv_all / (n_all * K * args.steps),
                   "ms_per_control_step": total_ms_max / args.steps / K,
                   "tracks": [e["track"] for e in engines], "host_affinity": pinned,
                   "fitness_vector": None if fit_local is None else int(fit_local.numel())},
        "clocks": summarize_clocks(samples),
        "e2e": {"value": e2e_all / (e2e_ms_max * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "roofline": {"bound": "fp64", "achieved": None, "peak": None, "unit": "TFLOP/s", "frac": None,
                     "traffic": None, "kernel": "mpcr::solve_kernel",
                     "kernel_ms": solve_ms,
                     "note": "kernel_ms = the last UNGRAPHED tick's solve (graph replays carry no timing events); "
                             "iteration counts are not read back inside the loop; see the cfg2 line for the roofline"},
        "ranks": ranks,
        "cpu_baseline": None,
    }
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--inner", type=int, default=20, help="control steps per bench step (cfg4 / cfg5)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--pipeline-depth", type=int, default=3,
                    help="cfg2 / cfg3: handles (streams) of the solve pipeline; 1 = one launch at a time")
    ap.add_argument("--engines-per-track", type=int, default=1,
                    help="cfg4: independent sub-fleets per track, each on its own stream")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_gpu(a)
