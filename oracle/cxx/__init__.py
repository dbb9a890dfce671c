"""Oracle (test infrastructure): ctypes loader of the compiled CPU oracle (racing_oracle.cpp).

Independent of the CUDA kernels' source: AD derivatives of the reference's expressions
(control/MPC.py:57-149, :186-283), oracle/ipm.py's algorithm, block-tridiagonal Bunch-Kaufman
LDL^T.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle_cxx.so")
SRC = os.path.join(HERE, "racing_oracle.cpp")


class Options(C.Structure):
    _fields_ = [(n, C.c_int32) for n in ("N", "model", "wrap", "lane", "integrator", "max_iter",
                                         "restoration", "linsol", "threads", "no_refine")] + \
               [(n, C.c_double) for n in ("Ts", "tol", "acceptable_tol", "polish_tol", "mu_init",
                                          "lambda_s", "alpha_L", "min_steer", "max_steer", "min_throttle",
                                          "max_steer_delta", "min_steer_delta", "max_throttle_delta",
                                          "min_throttle_delta", "q_v_max", "v_max", "min_s_delta",
                                          "alpha_c", "d_max", "q_v_y", "n_exp", "beta_delta")]


def build(force=False):
    """g++ -O3 -funroll-loops -fopenmp, in-tree (the .so is git-ignored but travels to the GPU box)."""
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        subprocess.run(["g++", "-O3", "-funroll-loops", "-std=c++17", "-fopenmp", "-fPIC", "-shared", "-o", LIB, SRC],
                       check=True)
    return LIB


_lib = None


def load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_options_default.argtypes = [C.POINTER(Options)]
        _lib.orc_solve_batch.restype = C.c_int
        _lib.orc_model_eval.restype = C.c_int
    return _lib


MODELS = {"dynamic": 0, "kinematic": 1, "blended": 2}


def lam_g_stride(N, lane=False):
    return 7 + N * (7 + (1 if lane else 0)) + 6 * N + 2


def solve_batch(x0, u_prev, s0, cx, cy, N=30, Ts=0.05, model="dynamic", wrap=True, lane=False,
                max_err=None, params=None, u_warm=None, integrator="euler", restoration=False,
                tol=1e-4, acceptable_tol=1e-2, max_iter=500, polish_tol=0.0, linsol="blocktri",
                threads=0, fixed=None, refine=True):
    """Solve B problems on the host cores; arrays as mpcr_solve_batch (include/mpcr.h)."""
    lib = load()
    o = Options()
    lib.orc_options_default(C.byref(o))
    o.N, o.Ts = int(N), float(Ts)
    o.model = MODELS[model] if isinstance(model, str) else int(model)
    o.wrap, o.lane = int(bool(wrap)), int(bool(lane))
    o.integrator = {"euler": 0, "rk4": 1}.get(integrator, integrator)
    o.restoration = int(bool(restoration))
    o.tol, o.acceptable_tol, o.max_iter, o.polish_tol = float(tol), float(acceptable_tol), int(max_iter), float(polish_tol)
    o.linsol = {"blocktri": 0, "dense": 1}[linsol]
    o.threads = int(threads)
    o.no_refine = int(not refine)
    for k, v in (fixed or {}).items():
        setattr(o, k, float(v))
    f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
    x0, u_prev, s0, cx, cy, max_err, params, u_warm = map(f, (x0, u_prev, s0, cx, cy, max_err, params, u_warm))
    B = x0.shape[0]
    assert x0.shape == (B, 6) and u_prev.shape == (B, 2) and s0.shape == (B,) and cx.shape == (B, 5)
    assert not lane or max_err is not None
    ls = lam_g_stride(N, lane)
    out = dict(U=np.zeros((B, 2, N)), X=np.zeros((B, 6, N + 1)), S=np.zeros((B, N + 1)),
               lam_g=np.zeros((B, ls)), obj=np.zeros(B), status=np.zeros(B, np.int32),
               iters=np.zeros(B, np.int32), polish_iters=np.zeros(B, np.int32))
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    rc = lib.orc_solve_batch(C.byref(o), B, p(x0), p(u_prev), p(s0), p(cx), p(cy), p(max_err), p(params),
                             p(u_warm), p(out["U"]), p(out["X"]), p(out["S"]), p(out["lam_g"]), p(out["obj"]),
                             p(out["status"]), p(out["iters"]), p(out["polish_iters"]))
    if rc != 0:
        raise RuntimeError(f"orc_solve_batch rc={rc}")
    return out


def model_eval(model, x, u, Ts, lam=None, integrator="euler"):
    lib = load()
    x = np.ascontiguousarray(x, dtype=np.float64)
    u = np.ascontiguousarray(u, dtype=np.float64)
    f, J, H = np.zeros(6), np.zeros((6, 8)), np.zeros((8, 8))
    lam_ = None if lam is None else np.ascontiguousarray(lam, dtype=np.float64)
    p = lambda a: None if a is None else a.ctypes.data_as(C.c_void_p)
    lib.orc_model_eval(MODELS[model] if isinstance(model, str) else int(model),
                       {"euler": 0, "rk4": 1}.get(integrator, integrator), p(x), p(u), C.c_double(Ts), p(lam_),
                       p(f), p(J), p(H))
    return f, J, (H if lam is not None else None)


def bk_solve(A, b):
    """Unit-test hook: dense Bunch-Kaufman LDL^T solve + inertia of a symmetric matrix."""
    lib = load()
    A = np.ascontiguousarray(A, dtype=np.float64)
    b = np.ascontiguousarray(b, dtype=np.float64)
    x = np.zeros_like(b)
    inertia = np.zeros(3, np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    lib.orc_bk_solve(A.shape[0], p(A), p(b), p(x), p(inertia))
    return x, tuple(int(v) for v in inertia)
