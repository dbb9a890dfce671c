#!/usr/bin/env python
"""bench.py - converged NMPC solves/sec of the batched solver (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (one rank per GPU under torchrun)
    python bench.py --impl reference ...                     (CPU arm: the oracle port of the
                                                              reference's IPOPT path, all host cores)

A "step" is one pass of the hot path (mpcr_solve_batch) over one batch of B = 16384 synthetic
problems per GPU: config 2 of SURVEY 8(d) widened to the metric's batch size - dynamic-bicycle
NLP (what control/MPC.py solves), default horizon N = 30, Ts = 0.05, Shanghai start states.
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

N_HORIZON = 30
TS = 0.05
BATCH = 16384
FLOP_PER_ITER = 4400 * N_HORIZON        # algorithmic FLOP of one IPM iteration
METRIC = "converged_nmpc_solves_per_sec"
UNIT = "solves/s"
WORKLOAD = ("cfg2-shanghai-dynamic-N30-Ts0.05-B16384/GPU: 4096 reference-sampled start states "
            "(seed 1234) x 4 seeded perturbations, cold start")


# ---------------------------------------------------------------------------------------
def clocks_sampler(stop, samples, gpu_index):
    q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")
    while not stop.is_set():
        try:
            out = subprocess.run(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={q}",
                                  "--format=csv,noheader,nounits"], capture_output=True, text=True,
                                 timeout=5).stdout.strip()
            if out:
                samples.append([x.strip() for x in out.split(",")])
        except Exception:
            pass
        stop.wait(0.2)


def summarize_clocks(samples):
    if not samples:
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
    sm = [float(s[0]) for s in samples if s[0].replace('.', '').isdigit()]
    mx = [float(s[1]) for s in samples if s[1].replace('.', '').isdigit()]
    names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
    reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in samples)]
    return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
            "reasons": reasons, "samples": len(samples)}


def _pool_init():
    try:
        from threadpoolctl import threadpool_limits
        threadpool_limits(1)
    except Exception:
        pass


def oracle_solve_one(args):
    from oracle.nlp import RacingNLP
    from oracle import ipm
    x0, up, s0, cx, cy = args
    nlp = RacingNLP(x0, up, s0, cx, cy, N=N_HORIZON, Ts=TS)
    r = ipm.solve(nlp)
    return r.status, r.iters


def cpu_pool_rate(batch, n_problems, cores):
    """Oracle (numpy port of the reference's IPOPT path) in a process pool: converged/s."""
    from multiprocessing import get_context
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["MKL_NUM_THREADS"] = "1"
    jobs = [(batch["x0"][b], batch["u_prev"][b], batch["s0"][b], batch["cx"][b], batch["cy"][b])
            for b in range(n_problems)]
    ctx = get_context("fork")
    with ctx.Pool(cores, initializer=_pool_init) as pool:
        pool.map(oracle_solve_one, jobs[:cores])            # warm the workers
        t0 = time.perf_counter()
        res = pool.map(oracle_solve_one, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    conv = sum(1 for st, _ in res if st >= 0)
    return conv / dt, conv, dt


def _cxx_chunk(args):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from emu_util import emu_solve
    x0, up, s0, cx, cy = args
    r = emu_solve(x0, up, s0, cx, cy, N=N_HORIZON, Ts=TS, lanes=True)
    return int((r["status"] >= 0).sum())


def cpu_cxx_rate(batch, n_problems, cores):
    """Context only: the lane-solver SOURCE (mpcr_ipm_lanes.cuh: same IPM + Riccati as the kernels)
    compiled for the host by tests/emu with g++ -O2, one process per core.  Not the oracle and not
    the reference - an optimised C++ CPU port of the same algorithm."""
    from multiprocessing import get_context
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import emu_util
    emu_util.build_emu()                 # once, in the parent (the workers only load it)
    per = max(1, n_problems // cores)
    jobs = [tuple(batch[k][i * per:(i + 1) * per] for k in ("x0", "u_prev", "s0", "cx", "cy"))
            for i in range(cores)]
    ctx = get_context("fork")
    with ctx.Pool(cores, initializer=_pool_init) as pool:
        pool.map(_cxx_chunk, [tuple(a[:1] for a in j) for j in jobs])      # build + warm
        t0 = time.perf_counter()
        conv = sum(pool.map(_cxx_chunk, jobs, chunksize=1))
        dt = time.perf_counter() - t0
    return conv / dt, conv, dt


# ---------------------------------------------------------------------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from mpc_racing_b200.workload import make_batch
    cores = os.cpu_count() or 1
    per_step = max(4 * cores, 32)
    batch = make_batch(per_step * (args.steps + args.warmup))
    from multiprocessing import get_context
    os.environ["OMP_NUM_THREADS"] = "1"
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    jobs = [(batch["x0"][b], batch["u_prev"][b], batch["s0"][b], batch["cx"][b], batch["cy"][b])
            for b in range(batch["x0"].shape[0])]
    ctx = get_context("fork")
    conv = 0
    with ctx.Pool(cores, initializer=_pool_init) as pool:
        for w in range(args.warmup):
            pool.map(oracle_solve_one, jobs[w * per_step:(w + 1) * per_step], chunksize=1)
        t0 = time.perf_counter()
        for k in range(args.steps):
            lo = (args.warmup + k) * per_step
            res = pool.map(oracle_solve_one, jobs[lo:lo + per_step], chunksize=1)
            conv += sum(1 for st, _ in res if st >= 0)
        dt = time.perf_counter() - t0
    value = conv / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": {"workload": WORKLOAD, "sample_per_step": per_step},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{per_step} problems/step of the same workload, oracle/ipm.py "
                                   "(numpy restatement of the reference's IPOPT path; CasADi/IPOPT "
                                   "is not installable offline), one process per core"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------
def run_gpu(args):
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from mpc_racing_b200.solver import BatchedMPC
    from mpc_racing_b200.workload import make_batch

    B = args.batch
    # independent problems: rank r owns batch indices [r*B, (r+1)*B) of the global batch
    from mpc_racing_b200.sharding import shard_range
    full = make_batch(B * world)
    sl = slice(*shard_range(B * world, rank, world))
    host = {k: np.ascontiguousarray(v[sl]) for k, v in full.items()}
    dsolver = BatchedMPC(N=N_HORIZON, Ts=TS, model="dynamic", device=local, on_device=True)
    hsolver = BatchedMPC(N=N_HORIZON, Ts=TS, model="dynamic", device=local, on_device=False)
    din = {k: torch.from_numpy(host[k]).to(dev) for k in ("x0", "u_prev", "s0", "cx", "cy")}
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)   # > 126 MB L2

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    out = None
    res = None
    for _ in range(args.warmup):
        res = dsolver.solve(din["x0"], din["u_prev"], din["s0"], din["cx"], din["cy"], out=out)
        out = dict(U=res.U, X=res.X, S=res.S, eC=res.eC, eL=res.eL, lam_g=res.lam_g, obj=res.obj,
                   status=res.status, iters=res.iters)
    if out is None:
        res = dsolver.solve(din["x0"], din["u_prev"], din["s0"], din["cx"], din["cy"])
        out = dict(U=res.U, X=res.X, S=res.S, eC=res.eC, eL=res.eL, lam_g=res.lam_g, obj=res.obj,
                   status=res.status, iters=res.iters)
    stop = threading.Event()
    samples = []
    th = threading.Thread(target=clocks_sampler, args=(stop, samples, local), daemon=True)
    if rank == 0:
        th.start()
    # ---- timed region: K steps, inputs resident in HBM; L2 flushed between steps (untimed)
    barrier()
    step_ms = []
    kern_ms = []
    for _ in range(args.steps):
        flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        res = dsolver.solve(din["x0"], din["u_prev"], din["s0"], din["cx"], din["cy"], out=out)
        e1.record()
        e1.synchronize()
        step_ms.append(e0.elapsed_time(e1))
        kern_ms.append(dsolver.last_kernel_ms())
    barrier()
    total_ms = float(sum(step_ms))
    status = res.status.cpu().numpy()
    iters = res.iters.cpu().numpy()
    conv_local = int((status >= 0).sum())
    # ---- e2e: the same step through the C ABI with HOST buffers (pinned), H2D+D2H inside
    pin = {k: torch.from_numpy(host[k]).pin_memory().numpy() for k in ("x0", "u_prev", "s0", "cx", "cy")}
    hout = hsolver.alloc_host_outputs(B, want_lam_g=True, pinned=True)
    hsolver.solve(pin["x0"], pin["u_prev"], pin["s0"], pin["cx"], pin["cy"], out=hout)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        hres = hsolver.solve(pin["x0"], pin["u_prev"], pin["s0"], pin["cx"], pin["cy"], out=hout)
    torch.cuda.synchronize(dev)
    e2e_s = time.perf_counter() - t0
    barrier()
    e2e_conv = int((hres.status >= 0).sum())
    # ---- p50 single-solve latency (BASELINE.json metric, second half): B = 1 through the C ABI
    #      with host buffers, H2D + kernel + D2H inside each timed call (rank 0 only)
    lat = None
    if rank == 0:
        one = {k: pin[k][:1] for k in pin}
        o1 = hsolver.alloc_host_outputs(1, want_lam_g=True, pinned=True)
        ts = []
        for i in range(60):
            j = i % min(B, 50)
            one = {k: pin[k][j:j + 1] for k in pin}
            t1 = time.perf_counter()
            hsolver.solve(one["x0"], one["u_prev"], one["s0"], one["cx"], one["cy"], out=o1)
            ts.append(time.perf_counter() - t1)
        ts = np.array(ts[10:]) * 1e3
        lat = {"p50_ms": float(np.percentile(ts, 50)), "p90_ms": float(np.percentile(ts, 90)),
               "batch": 1, "calls": int(ts.size), "includes": "H2D + solve kernel + D2H, host pointers"}
    h2d = sum(pin[k].nbytes for k in pin)
    d2h = sum(v.nbytes for v in hout.values() if v is not None)
    stop.set()
    # ---- reduce over ranks: max time, summed work (the only collective: a final gather)
    t = torch.tensor([total_ms, e2e_s * 1e3], dtype=torch.float64, device=dev)
    c = torch.tensor([conv_local, e2e_conv, int(iters.sum())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c, op=dist.ReduceOp.SUM)
    total_ms_max, e2e_ms_max = t.tolist()
    conv_all, e2e_conv_all, _ = c.tolist()
    if rank == 0:
        value = conv_all * args.steps / (total_ms_max * 1e-3)
        e2e_value = e2e_conv_all * args.steps / (e2e_ms_max * 1e-3)
        # roofline of the dominant (only) kernel: algorithmic fp64 FLOP / measured launch time
        kms = float(np.mean(kern_ms))
        flop = float(iters.sum()) * FLOP_PER_ITER
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
            cores = os.cpu_count() or 1
            nprob = max(8 * cores, 64)          # ~10-20 s of CPU work on the box's host cores
            rate, conv, dt = cpu_pool_rate(host, nprob, cores)
            cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"first {nprob} problems of rank 0's batch, oracle/ipm.py (numpy port of "
                             f"the reference's IPOPT path), {cores} processes, {dt:.1f} s"}
        cxx = None
        if not args.no_cpu_baseline:
            try:
                cores = os.cpu_count() or 1
                rate, conv, dt = cpu_cxx_rate(host, min(B, 512 * cores), cores)
                cxx = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port-cxx",
                       "sample": f"first {min(B, 512 * cores)} problems, lane-solver source compiled for the host "
                                 f"(g++ -O2, tests/emu), {cores} processes, {dt:.1f} s; context only"}
            except Exception as exc:      # the C++ host build is test infrastructure: optional here
                cxx = {"unavailable": repr(exc)[:200]}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms_max / max(1, args.steps),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "batch_per_gpu": B, "N": N_HORIZON, "Ts": TS,
                       "model": "dynamic", "l2": "flushed between timed steps (256 MB write, untimed)",
                       "converged_frac": conv_all / (B * world),
                       "mean_iters": float(iters.mean()), "launch": dsolver.launch_info()},
            "clocks": summarize_clocks(samples),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": args.steps * max(1, dsolver.launch_info()["kernels_per_solve"]),
            "roofline": {"bound": "fp64", "achieved": achieved, "peak": peak, "unit": "TFLOP/s",
                         "frac": achieved / peak if peak else None, "traffic": traffic,
                         "peak_source": "measured here: DFMA-chain microbenchmark (mpcr_fp64_peak_tflops)",
                         "kernel": "mpcr::solve_kernel", "kernel_ms": kms,
                         "flop_per_launch": flop,
                         "hbm": {"achieved_gbs": bytes_alg / (kms * 1e-3) / 1e9, "peak_gbs": hbm_peak,
                                 "frac": bytes_alg / (kms * 1e-3) / 1e9 / hbm_peak,
                                 "algorithmic_bytes_per_launch": bytes_alg}},
            "cpu_baseline": cpu,
            "single_solve_latency": lat,
            "cpu_baseline_cxx": cxx,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    a = ap.parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_gpu(a)
