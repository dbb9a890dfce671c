"""Batched NMPC solve: the ``solve(states[B])`` entry point over the C ABI.

``BatchedMPC`` owns one ``mpcr_handle`` (one CUDA device).  ``solve`` accepts
either numpy arrays (host pointers; the library stages H2D/D2H itself - this is
the end-to-end path) or torch CUDA tensors (device pointers; nothing leaves the
GPU).  PyTorch is only plumbing here: device memory and streams.
"""
import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _capi
from .control.ControllerParameters import FixedControllerParameters, RuntimeControllerParameters

MODEL_IDS = {"dynamic": 0, "kinematic": 1, "blended": 2}
STATUS_NAMES = {0: "Solve_Succeeded", 1: "Solved_To_Acceptable_Level",
                -1: "Maximum_Iterations_Exceeded", -2: "Restoration_Failed",
                -3: "Error_In_Step_Computation", -4: "Invalid_Number_Detected"}


@dataclass
class BatchResult:
    """Outputs of one batched solve, arrays on the side the inputs came from."""
    U: object        # [B,2,N]
    X: object        # [B,6,N+1]
    S: object        # [B,N+1]
    eC: object       # [B,N]
    eL: object       # [B,N]
    lam_g: object    # [B,stride] or None
    obj: object      # [B]
    status: object   # [B] int32
    iters: object    # [B] int32
    trace: object = None

    @property
    def converged(self):
        return self.status >= 0


def _is_torch(a):
    return type(a).__module__.startswith("torch")


class BatchedMPC:
    """B independent progress-maximising NLPs (control/MPC.py) per call."""

    def __init__(self, N: Optional[int] = None, Ts: Optional[float] = None, model="dynamic",
                 fixed_params: Optional[FixedControllerParameters] = None,
                 runtime_params: Optional[RuntimeControllerParameters] = None,
                 wrap_constraint=True, lane_constraint=False, device=0, on_device=False,
                 tol=1e-4, acceptable_tol=1e-2, max_iter=None, honor_instance_d_max=False,
                 solve_kernel=0, integrator="euler", restoration=False, schedule="auto", polish_tol=0.0):
        self.lib = _capi.load_library()
        fp = fixed_params or FixedControllerParameters()
        rp = runtime_params or RuntimeControllerParameters()
        cfg = _capi.default_config(self.lib)
        cfg.N = int(fp.N if N is None else N)
        cfg.Ts = float(FixedControllerParameters.Ts if Ts is None else Ts)   # control/MPC.py:57
        cfg.model = MODEL_IDS[model] if isinstance(model, str) else int(model)
        cfg.wrap_constraint = int(bool(wrap_constraint))
        cfg.lane_constraint = int(bool(lane_constraint))
        cfg.device = int(device)
        cfg.pointers_on_device = int(bool(on_device))
        cfg.max_iter = int(FixedControllerParameters.max_iter if max_iter is None else max_iter)
        for name in ("lambda_s", "alpha_L", "min_steer", "max_steer", "min_throttle",
                     "max_steer_delta", "min_steer_delta", "max_throttle_delta",
                     "min_throttle_delta", "q_v_max", "v_max"):
            setattr(cfg, name, float(getattr(fp, name)))
        cfg.alpha_c = float(rp.alpha_c)
        cfg.q_v_y = float(rp.q_v_y)
        cfg.n_exp = float(rp.n)
        cfg.beta_delta = float(rp.beta_delta)
        # control/MPC.py:50 reads d_max from the CLASS, not the instance
        cfg.d_max = float(rp.d_max if honor_instance_d_max else RuntimeControllerParameters.d_max)
        cfg.tol = float(tol)
        cfg.acceptable_tol = float(acceptable_tol)
        # layout: 0 = auto, 1 = one problem per warp (the only layout)
        cfg.solve_kernel = int({"auto": 0, "warp": 1}.get(solve_kernel, solve_kernel))
        # "euler" = the reference's discretisation; "rk4" = extension (same kernel, 32-entry stage Jacobian)
        cfg.integrator = int({"euler": 0, "rk4": 1}.get(integrator, integrator))
        # feasibility restoration phase (off = status -2 instead).  True picks the launch mode: two-stage
        # (plain solver + re-solve of the few problems that need the phase) unless the lane rows are on,
        # where ~5 % of the problems need it and the build with the phase for everybody is cheaper;
        # "two-stage" / "always" (or 1 / 2) force a mode
        mode = {False: 0, True: (2 if lane_constraint else 1), "two-stage": 1, "always": 2}.get(restoration, restoration)
        cfg.restoration = int(mode)
        # dispatch order of the batch: "auto" = expected-longest-first for batches above one wave
        cfg.schedule = int({"auto": 0, "index": 1, "curvature": 2}.get(schedule, schedule))
        # > 0: keep iterating to this tolerance after the first termination (status / iters report the first)
        cfg.polish_tol = float(polish_tol)
        self.cfg = cfg
        self.N = cfg.N
        self.on_device = bool(on_device)
        self._h = C.c_void_p()
        rc = self.lib.mpcr_create(C.byref(cfg), C.byref(self._h))
        if rc != 0:
            msg = self.lib.mpcr_last_error(self._h).decode() if self._h else ""
            if self._h:
                self.lib.mpcr_destroy(self._h)
                self._h = C.c_void_p()
            raise _capi.MpcrError(f"mpcr_create failed (rc={rc}) {msg}")
        self.lam_g_stride = self.lib.mpcr_lam_g_stride(self._h)

    def close(self):
        if getattr(self, "_h", None) and self._h:
            self.lib.mpcr_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------
    def launch_info(self):
        b, w, s = C.c_int(), C.c_int(), C.c_int()
        self.lib.mpcr_launch_info(self._h, C.byref(b), C.byref(w), C.byref(s))
        n = C.c_int()
        self.lib.mpcr_last_launch_count(self._h, C.byref(n))
        return dict(blocks=b.value, warps_per_block=w.value, smem_bytes=s.value, kernels_per_solve=n.value)

    def fp64_peak_tflops(self):
        v = C.c_double()
        self._check(self.lib.mpcr_fp64_peak_tflops(self._h, C.byref(v)))
        return v.value

    def set_dispatch_keys(self, keys):
        """Expected-cost keys (int32 CUDA tensor [B] or None) for the dispatch order of the next solves."""
        self._check(self.lib.mpcr_set_dispatch_keys(self._h, None if keys is None else keys.data_ptr()))

    def last_kernel_ms(self):
        ms = C.c_float()
        rc = self.lib.mpcr_last_kernel_ms(self._h, C.byref(ms))
        if rc != 0:
            raise _capi.MpcrError(self.lib.mpcr_last_error(self._h).decode())
        return ms.value

    # ------------------------------------------------------------------
    def model_eval(self, x, u, lam=None, model=None, integrator="euler", Ts=None, jac=True, hess=True):
        """Vehicle-model step x+ with analytic Jacobian [B,6,8] over (x, u) and lam-weighted Hessian
        [B,8,8] (host arrays).  integrator "euler" = the maps of control/MPC.py:186-260 that the
        reference's NLP uses; "rk4" = classical RK4 of the same vector field (extension)."""
        assert not self.on_device, "model_eval takes host arrays"
        x = np.ascontiguousarray(x, dtype=np.float64)
        u = np.ascontiguousarray(u, dtype=np.float64)
        B = x.shape[0]
        assert x.shape == (B, 6) and u.shape == (B, 2)
        lam = None if lam is None else np.ascontiguousarray(lam, dtype=np.float64)
        m = self.cfg.model if model is None else (MODEL_IDS[model] if isinstance(model, str) else int(model))
        integ = {"euler": 0, "rk4": 1}.get(integrator, integrator)
        f = np.empty((B, 6))
        J = np.empty((B, 6, 8)) if jac else None
        H = np.empty((B, 8, 8)) if (hess and lam is not None) else None
        ptr = lambda a: None if a is None else a.ctypes.data
        self._check(self.lib.mpcr_model_eval(self._h, m, int(integ), B, ptr(x), ptr(u),
                                             float(self.cfg.Ts if Ts is None else Ts), ptr(lam), ptr(f),
                                             ptr(J), ptr(H), None))
        return f, J, H

    # ------------------------------------------------------------------
    def solve(self, x0, u_prev, s0, cx, cy, max_err=None, params=None, u_warm=None,
              want_lam_g=True, trace=False, stream=None, out=None, sync=True) -> BatchResult:
        """``sync=False`` (host arrays only; device mode is always asynchronous): return once the
        copies and kernels are enqueued on ``stream`` (mpcr_solve_batch_async) - the inputs must
        stay untouched and the outputs are undefined until that stream is synchronised."""
        if self.on_device:
            return self._solve_device(x0, u_prev, s0, cx, cy, max_err, params, u_warm,
                                      want_lam_g, trace, stream, out)
        return self._solve_host(x0, u_prev, s0, cx, cy, max_err, params, u_warm, want_lam_g,
                                trace, stream, out, sync)

    def _check(self, rc):
        if rc != 0:
            raise _capi.MpcrError(f"mpcr_solve_batch failed (rc={rc}): "
                                  f"{self.lib.mpcr_last_error(self._h).decode()}")

    def _call(self, B, ins, outs, trace_ptr, stream, sync=True):
        args = [self._h, int(B)] + ins + outs
        if trace_ptr is not None:
            self._check(self.lib.mpcr_solve_batch_trace(*args, trace_ptr, stream))
        elif not sync:
            self._check(self.lib.mpcr_solve_batch_async(*args, stream))
        else:
            self._check(self.lib.mpcr_solve_batch(*args, stream))

    def alloc_host_outputs(self, B, want_lam_g=True, pinned=False):
        N, ls = self.N, self.lam_g_stride
        shapes = dict(U=(B, 2, N), X=(B, 6, N + 1), S=(B, N + 1), eC=(B, N), eL=(B, N),
                      lam_g=(B, ls), obj=(B,))
        if pinned:
            import torch
            mk = lambda s, dt: torch.empty(s, dtype=dt).pin_memory().numpy()
            o = {k: mk(s, torch.float64) for k, s in shapes.items()}
            o["status"] = mk((B,), torch.int32)
            o["iters"] = mk((B,), torch.int32)
        else:
            o = {k: np.empty(s, dtype=np.float64) for k, s in shapes.items()}
            o["status"] = np.empty(B, dtype=np.int32)
            o["iters"] = np.empty(B, dtype=np.int32)
        if not want_lam_g:
            o["lam_g"] = None
        return o

    def _solve_host(self, x0, u_prev, s0, cx, cy, max_err, params, u_warm, want_lam_g, trace,
                    stream, out, sync=True):
        f = lambda a: None if a is None else np.ascontiguousarray(a, dtype=np.float64)
        x0, u_prev, s0, cx, cy, max_err, params, u_warm = map(
            f, (x0, u_prev, s0, cx, cy, max_err, params, u_warm))
        B = x0.shape[0]
        N = self.N
        assert x0.shape == (B, 6) and u_prev.shape == (B, 2) and s0.shape == (B,)
        assert cx.shape == (B, 5) and cy.shape == (B, 5)
        assert u_warm is None or u_warm.shape == (B, 2, N)
        assert params is None or params.shape == (B, 5)
        o = out or self.alloc_host_outputs(B, want_lam_g)
        ptr = lambda a: None if a is None else a.ctypes.data
        tr = np.zeros((B, self.cfg.max_iter + 1, 8)) if trace else None
        self._call(B, [ptr(a) for a in (x0, u_prev, s0, cx, cy, max_err, params, u_warm)],
                   [ptr(o[k]) for k in ("U", "X", "S", "eC", "eL", "lam_g", "obj", "status", "iters")],
                   ptr(tr) if trace else None, stream, sync)
        res = BatchResult(trace=tr, **o)
        if not sync:
            res._keepalive = (x0, u_prev, s0, cx, cy, max_err, params, u_warm)   # read by the copies later
        return res

    def _solve_device(self, x0, u_prev, s0, cx, cy, max_err, params, u_warm, want_lam_g, trace,
                      stream, out):
        import torch
        dev = torch.device("cuda", self.cfg.device)
        for a in (x0, u_prev, s0, cx, cy, max_err, params, u_warm):
            assert a is None or (_is_torch(a) and a.is_cuda and a.dtype == torch.float64
                                 and a.is_contiguous()), "device mode needs contiguous cuda fp64"
        B, N, ls = x0.shape[0], self.N, self.lam_g_stride
        if out is None:
            mk = lambda *s: torch.empty(*s, dtype=torch.float64, device=dev)
            out = dict(U=mk(B, 2, N), X=mk(B, 6, N + 1), S=mk(B, N + 1), eC=mk(B, N), eL=mk(B, N),
                       lam_g=mk(B, ls) if want_lam_g else None, obj=mk(B),
                       status=torch.empty(B, dtype=torch.int32, device=dev),
                       iters=torch.empty(B, dtype=torch.int32, device=dev))
        tr = torch.zeros(B, self.cfg.max_iter + 1, 8, dtype=torch.float64, device=dev) if trace else None
        ptr = lambda a: None if a is None else a.data_ptr()
        if stream is None:
            stream = torch.cuda.current_stream(dev).cuda_stream
        self._call(B, [ptr(a) for a in (x0, u_prev, s0, cx, cy, max_err, params, u_warm)],
                   [ptr(out[k]) for k in ("U", "X", "S", "eC", "eL", "lam_g", "obj", "status", "iters")],
                   ptr(tr) if trace else None, stream)
        return BatchResult(trace=tr, **out)
