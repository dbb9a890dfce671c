"""Closed-loop receding-horizon rollouts, everything resident on the GPU.

Reproduces per step what the reference's agent does around each solve (agent.py:138-205,
:271-296) for R independent vehicles at once, with the reference's numpy plant models as the
simulator (there is no CARLA here):

  1. progress   = centerline.projection(X, Y, bounds=progress_bound())     agent.py:80-92,271
  2. cx, cy     = x/y_as_coeffs(progress - 5, lookahead, deg 4)            agent.py:156-165
  3. U          = MPC(state, progress, cx, cy, last_controls=U_prev).solve agent.py:171-205
  4. apply u_0  (negative throttle = brake)                                agent.py:281-296
  5. x          = <plant>.step(u_0, dt)                                    models/*.py

Four kernel launches per step (project, refit, solve, plant), device pointers throughout;
torch is only used for device memory and the elementwise bookkeeping.
"""
import ctypes as C

import numpy as np

from .solver import BatchedMPC, MODEL_IDS


class ClosedLoopRollouts:
    def __init__(self, centerline, N=15, Ts=0.05, nlp_model="dynamic", plant="blended",
                 lookahead=45.0, lookback=5.0, window=2.0, device=0, runtime_params=None):
        import torch
        self.torch = torch
        self.dev = torch.device("cuda", device)
        self.cl = centerline
        self.L = float(centerline.length)
        self.N, self.Ts = N, Ts
        self.lookahead, self.lookback, self.window = float(lookahead), float(lookback), float(window)
        self.plant_id = MODEL_IDS[plant]
        self.solver = BatchedMPC(N=N, Ts=Ts, model=nlp_model, device=device, on_device=True,
                                 runtime_params=runtime_params)
        self.lib = self.solver.lib
        self.track = centerline.device_track(self.solver)

    def close(self):
        if self.track:
            self.lib.mpcr_track_destroy(self.track)
            self.track = None
        self.solver.close()

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.mpcr_last_error(self.solver._h).decode())

    def step_prepare(self, x, progress, stream):
        """Steps 1-2: returns (progress_new, dist, sign, cx, cy)."""
        t = self.torch
        R = x.shape[0]
        lo = t.remainder(progress - self.window, self.L)
        hi = t.remainder(progress + self.window, self.L)
        lo2 = t.minimum(lo, hi)                      # agent.progress_bound swaps at the wrap:
        hi2 = t.maximum(lo, hi)                      # the window then spans ~L -> global search
        X = x[:, 0].contiguous()
        Y = x[:, 1].contiguous()
        s_new = t.empty(R, dtype=t.float64, device=self.dev)
        dist = t.empty(R, dtype=t.float64, device=self.dev)
        sign = t.empty(R, dtype=t.int32, device=self.dev)
        self._chk(self.lib.mpcr_track_project(self.track, R, X.data_ptr(), Y.data_ptr(), lo2.data_ptr(),
                                              hi2.data_ptr(), s_new.data_ptr(), dist.data_ptr(),
                                              sign.data_ptr(), stream))
        cx = t.empty(R, 5, dtype=t.float64, device=self.dev)
        cy = t.empty(R, 5, dtype=t.float64, device=self.dev)
        s_fit = (s_new - self.lookback).contiguous()
        self._chk(self.lib.mpcr_track_coeffs(self.track, R, s_fit.data_ptr(), self.lookahead,
                                             cx.data_ptr(), cy.data_ptr(), stream))
        return s_new, dist, sign, cx, cy

    def run(self, x0, s0, u_prev0=None, params=None, max_steps=2000, laps=1.0, seg_bounds=None,
            record=False):
        """Roll R vehicles until each has advanced ``laps`` * L of progress (or max_steps).

        x0 [R,6], s0 [R] numpy; params [R,5] per-vehicle runtime parameters (GA individuals).
        Returns numpy arrays: lap_time [R] (nan if not finished), steps [R], fail_steps [R],
        max_abs_error [R], seg_times [R,n_seg] if seg_bounds is given, and the state history
        if ``record``.
        """
        t = self.torch
        dev = self.dev
        R = x0.shape[0]
        stream = t.cuda.current_stream(dev).cuda_stream
        x = t.from_numpy(np.ascontiguousarray(x0, dtype=np.float64)).to(dev)
        progress = t.from_numpy(np.ascontiguousarray(s0, dtype=np.float64)).to(dev)
        u_prev = t.zeros(R, 2, dtype=t.float64, device=dev) if u_prev0 is None else \
            t.from_numpy(np.ascontiguousarray(u_prev0, dtype=np.float64)).to(dev)
        prm = None if params is None else t.from_numpy(np.ascontiguousarray(params, dtype=np.float64)).to(dev)
        u_warm = None
        out = None
        travelled = t.zeros(R, dtype=t.float64, device=dev)
        lap_time = t.full((R,), float("nan"), dtype=t.float64, device=dev)
        steps = t.zeros(R, dtype=t.int32, device=dev)
        fails = t.zeros(R, dtype=t.int32, device=dev)
        max_err = t.zeros(R, dtype=t.float64, device=dev)
        done = t.zeros(R, dtype=t.bool, device=dev)
        target = laps * self.L
        seg_t = None
        if seg_bounds is not None:
            sb = t.tensor(list(seg_bounds), dtype=t.float64, device=dev)       # [n_seg+1], sb[0] = 0
            seg_t = t.full((R, len(seg_bounds) - 1), float("nan"), dtype=t.float64, device=dev)
            start_abs = progress.clone()
        hist = [] if record else None
        x_new = t.empty_like(x)
        for k in range(max_steps):
            s_new, dist, sign, cx, cy = self.step_prepare(x, progress, stream)
            if k > 0:
                d = t.remainder(s_new - progress + 0.5 * self.L, self.L) - 0.5 * self.L
                travelled = t.where(done, travelled, travelled + d)
            progress = s_new
            max_err = t.where(done, max_err, t.maximum(max_err, dist))
            newly = (~done) & (travelled >= target)
            lap_time = t.where(newly, t.full_like(lap_time, k * self.Ts), lap_time)
            if seg_t is not None:
                for j in range(seg_t.shape[1]):
                    hit = t.isnan(seg_t[:, j]) & (travelled >= sb[j + 1] - 1e-12)
                    seg_t[:, j] = t.where(hit, t.full_like(lap_time, k * self.Ts), seg_t[:, j])
            done = done | newly
            if bool(done.all()):
                break
            res = self.solver.solve(x, u_prev, progress, cx, cy, params=prm, u_warm=u_warm,
                                    want_lam_g=False, out=out)
            out = dict(U=res.U, X=res.X, S=res.S, eC=res.eC, eL=res.eL, lam_g=None, obj=res.obj,
                       status=res.status, iters=res.iters)
            fails = fails + ((res.status < 0) & (~done)).to(t.int32)
            u0 = res.U[:, :, 0].contiguous()
            self._chk(self.lib.mpcr_plant_step(self.solver._h, self.plant_id, R, x.data_ptr(), u0.data_ptr(),
                                               self.Ts, x_new.data_ptr(), stream))
            if record:
                hist.append(dict(x=x.cpu().numpy().copy(), progress=progress.cpu().numpy().copy(),
                                 u0=u0.cpu().numpy().copy(), cx=cx.cpu().numpy().copy(),
                                 cy=cy.cpu().numpy().copy(), status=res.status.cpu().numpy().copy(),
                                 U=res.U.cpu().numpy().copy(), X=res.X.cpu().numpy().copy(),
                                 S=res.S.cpu().numpy().copy(), eC=res.eC.cpu().numpy().copy(),
                                 eL=res.eL.cpu().numpy().copy(), dist=dist.cpu().numpy().copy(),
                                 sign=sign.cpu().numpy().copy()))
            x = t.where(done[:, None], x, x_new)
            x_new = t.empty_like(x)
            u_prev = u0
            u_warm = res.U.clone()        # agent.py:205: last_controls = the full control sequence
            steps = steps + (~done).to(t.int32)
        ret = dict(lap_time=lap_time.cpu().numpy(), steps=steps.cpu().numpy(), fail_steps=fails.cpu().numpy(),
                   max_abs_error=max_err.cpu().numpy(), travelled=travelled.cpu().numpy(),
                   final_state=x.cpu().numpy())
        if seg_t is not None:
            st = seg_t.cpu().numpy()
            ret["seg_times"] = np.diff(np.concatenate([np.zeros((R, 1)), st], axis=1), axis=1)
        if record:
            ret["history"] = hist
        return ret


def start_states(centerline, R, seed=1234, v0=5.0):
    """s0 ~ U[0, L) on the centerline, tangent yaw, v_x = v0."""
    rng = np.random.default_rng(seed)
    s0 = rng.uniform(0, centerline.length, R)
    e = centerline.eval_all(s0)
    yaw = np.arctan2(e[:, 3], e[:, 2])
    x0 = np.stack([e[:, 0], e[:, 1], yaw, np.full(R, v0), np.zeros(R), np.zeros(R)], axis=1)
    return x0, s0
