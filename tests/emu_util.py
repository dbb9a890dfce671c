"""TEST INFRASTRUCTURE: build + call the SIMT-emulated solver (tests/emu)."""
import ctypes as C
import os
import subprocess

import numpy as np

from mpc_racing_b200._capi import MpcrConfig

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "emu")
EMU_LIB = os.path.join(EMU_DIR, "libmpcr_emu.so")
_lib = None


def build_emu(force=False):
    srcs = [os.path.join(EMU_DIR, "emu_main.cpp"), os.path.join(EMU_DIR, "simt_emu.h")]
    csrc = os.path.join(HERE, "..", "mpc_racing_b200", "csrc")
    srcs += [os.path.join(csrc, f) for f in os.listdir(csrc) if f.endswith((".cuh", ".h"))]
    if not force and os.path.exists(EMU_LIB):
        t = os.path.getmtime(EMU_LIB)
        if all(os.path.getmtime(s) <= t for s in srcs):
            return EMU_LIB
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", EMU_LIB,
                           os.path.join(EMU_DIR, "emu_main.cpp")])
    return EMU_LIB


def emu_lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build_emu())
        _lib.emu_config_default.argtypes = [C.POINTER(MpcrConfig)]
        _lib.emu_lam_g_stride.argtypes = [C.POINTER(MpcrConfig)]
        _lib.emu_solve_batch.argtypes = [C.POINTER(MpcrConfig), C.c_int] + [C.c_void_p] * 18
    return _lib


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def emu_solve(x0, u_prev, s0, cx, cy, max_err=None, params=None, u_warm=None, trace=False, **cfg_kw):
    lib = emu_lib()
    cfg = MpcrConfig()
    lib.emu_config_default(C.byref(cfg))
    for k, v in cfg_kw.items():
        assert hasattr(cfg, k), k
        setattr(cfg, k, v)
    f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
    x0, u_prev, s0, cx, cy, max_err, params, u_warm = map(f, (x0, u_prev, s0, cx, cy, max_err, params, u_warm))
    B, N = x0.shape[0], cfg.N
    ls = lib.emu_lam_g_stride(C.byref(cfg))
    out = dict(U=np.zeros((B, 2, N)), X=np.zeros((B, 6, N + 1)), S=np.zeros((B, N + 1)),
               eC=np.zeros((B, N)), eL=np.zeros((B, N)), lam_g=np.zeros((B, ls)), obj=np.zeros(B),
               status=np.zeros(B, dtype=np.int32), iters=np.zeros(B, dtype=np.int32))
    tr = np.zeros((B, cfg.max_iter + 1, 8)) if trace else None
    fn = lib.emu_solve_batch
    rc = fn(C.byref(cfg), B, _ptr(x0), _ptr(u_prev), _ptr(s0), _ptr(cx), _ptr(cy),
                             _ptr(max_err), _ptr(params), _ptr(u_warm), _ptr(out['U']), _ptr(out['X']),
                             _ptr(out['S']), _ptr(out['eC']), _ptr(out['eL']), _ptr(out['lam_g']),
                             _ptr(out['obj']), _ptr(out['status']), _ptr(out['iters']), _ptr(tr))
    assert rc == 0
    if trace:
        out['trace'] = tr
    return out
