"""Memory-safety of the solver source: the warp-level solver is compiled for the host (SIMT emulator)
with AddressSanitizer + UBSan and run over the shape variants (horizons 2/15/16/32,
models, lane rows, no wrap rows, warm start, RK4, polish, restoration).  The same source is what
nvcc compiles for the device (compute-sanitizer is not available on the GPU pool)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

DRIVER = r'''
import sys; sys.path.insert(0, %(tests)r); sys.path.insert(0, %(root)r)
import ctypes as C, numpy as np
import emu_util
from mpc_racing_b200._capi import MpcrConfig
l = C.CDLL(%(lib)r)
emu_util._lib = l
l.emu_config_default.argtypes = [C.POINTER(MpcrConfig)]
l.emu_lam_g_stride.argtypes = [C.POINTER(MpcrConfig)]
l.emu_solve_batch.argtypes = [C.POINTER(MpcrConfig), C.c_int] + [C.c_void_p] * 18
from emu_util import emu_solve
from mpc_racing_b200.workload import load_fixture
d = load_fixture(); d3 = load_fixture("cfg3_t1_triple_512.npz")
idx = [7]
a = lambda k: d[k][idx]
# the solver runs through the fiber emulator (slow under ASan): one problem per variant
for kw in (dict(N=32), dict(N=2), dict(N=15, model=2), dict(N=5, wrap_constraint=0)):
    r = emu_solve(a('x0'), a('u_prev'), a('s0'), a('cx'), a('cy'), trace=True, **kw)
    assert (r['status'] == 0).all(), (kw, r['status'])
r = emu_solve(a('x0'), np.full((1, 2), np.nan), a('s0'), a('cx'), a('cy'), N=10, u_warm=np.zeros((1, 2, 10)))
r = emu_solve(a('x0'), a('u_prev'), a('s0'), a('cx'), a('cy'), N=8, integrator=1)
r = emu_solve(a('x0'), a('u_prev'), a('s0'), a('cx'), a('cy'), N=8, polish_tol=1e-9)
# lane row + both restoration modes (two-stage re-solve, and the build with the phase for everybody)
for mode in (1, 2):
    r = emu_solve(d3['x0'][[478]], d3['u_prev'][[478]], d3['s0'][[478]], d3['cx'][[478]], d3['cy'][[478]],
                  max_err=d3['max_err'][[478]], lane_constraint=1, restoration=mode)
    assert r['status'][0] == 0
print("SANITIZED RUN OK")
'''


def test_solver_sources_under_asan_ubsan(tmp_path):
    asan = subprocess.run(["gcc", "-print-file-name=libasan.so"], capture_output=True, text=True).stdout.strip()
    ubsan = subprocess.run(["gcc", "-print-file-name=libubsan.so"], capture_output=True, text=True).stdout.strip()
    if not (os.path.isabs(asan) and os.path.exists(asan) and os.path.exists(ubsan)):
        pytest.skip("libasan / libubsan not installed")
    lib = str(tmp_path / "libmpcr_emu_asan.so")
    subprocess.check_call(["g++", "-O0", "-std=c++17", "-shared", "-fPIC", "-fsanitize=address,undefined",
                           "-fno-omit-frame-pointer", "-o", lib, os.path.join(HERE, "emu", "emu_main.cpp")])
    env = dict(os.environ, LD_PRELOAD=asan + ":" + ubsan, ASAN_OPTIONS="detect_leaks=0:abort_on_error=1",
               UBSAN_OPTIONS="halt_on_error=1:print_stacktrace=1")
    out = subprocess.run([sys.executable, "-c", DRIVER % dict(tests=HERE, root=ROOT, lib=lib)], env=env,
                         capture_output=True, text=True, timeout=900)
    assert out.returncode == 0 and "SANITIZED RUN OK" in out.stdout, (out.stdout[-1500:], out.stderr[-3000:])
    assert "runtime error" not in out.stderr and "AddressSanitizer" not in out.stderr.replace(
        "ASan doesn't fully support makecontext", ""), out.stderr[-3000:]
