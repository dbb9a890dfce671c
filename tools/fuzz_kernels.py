"""Prints the per-trial figures of tests/test_gpu_fuzz.py (warp kernel vs lane kernel on random
configurations).  Usage: python tools/fuzz_kernels.py [n_trials] [seed]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_gpu_fuzz import compare, fuzz_trials          # noqa: E402
from mpc_racing_b200.solver import BatchedMPC           # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
seed = int(sys.argv[2]) if len(sys.argv) > 2 else 2024
bad = 0
for trial, name, B, kw, args, opt in fuzz_trials(n, seed):
    a = BatchedMPC(solve_kernel="warp", **kw).solve(*args, **opt)
    b = BatchedMPC(solve_kernel="lanes", **kw).solve(*args, **opt)
    m = compare(a, b)
    flag = "" if (m["finite"] and m["class_eq"] >= 0.97 and m["worst_same"] <= 1e-7 and m["obj_close"] >= 0.95) else "  <== CHECK"
    bad += bool(flag)
    print(f"{trial:3d} {name} B={B:3d} {kw} prm={opt['params'] is not None} warm={opt['u_warm'] is not None}: {m}{flag}", flush=True)
print("flagged:", bad)
