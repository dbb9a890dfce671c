"""Closed-loop receding-horizon rollouts, everything resident on the GPU.

Reproduces per control step what the reference's agent does around each solve (agent.py:138-205,
:225-296) for R independent vehicles at once, with the reference's numpy plant models as the
simulator (there is no CARLA here):

  1. [state estimate: yaw rate = finite difference of the yaw, velocities world -> body] agent.py:240-254
  2. progress   = centerline.projection(X, Y, bounds=progress_bound())                   agent.py:80-92,271
  3. cx, cy     = x/y_as_coeffs(progress - 5, lookahead, deg 4)                          agent.py:156-165
  4. U          = MPC(state, progress, cx, cy, last_controls=U_prev).solve               agent.py:171-205
  5. apply u_0  ((0.5, 0) before start_control_at)                                       agent.py:278-296
  6. x          = <plant>.step(u_0, dt)                                                  models/*.py

Steps 1-3 + the lap bookkeeping are ONE kernel (mpcr_rollout_prepare), 5-6 another
(mpcr_rollout_apply), the solve in between writes into one of two plan buffers (the other one is
its warm start).  A pair of control steps is captured in a CUDA graph and replayed: no host work,
no host synchronisation and no allocation inside the loop; the host looks at the "all finished"
counter every ``check_every`` steps.  torch is only used for device memory and streams.
"""
import ctypes as C

import numpy as np

from . import _capi
from .solver import BatchedMPC, MODEL_IDS


class RolloutSession:
    """Device-resident state of R rollouts (mpcr_rollout_bufs + the two plan buffers)."""
    pass


class ClosedLoopRollouts:
    def __init__(self, centerline, N=15, Ts=0.05, nlp_model="dynamic", plant="blended",
                 lookahead=45.0, lookback=5.0, window=2.0, device=0, runtime_params=None,
                 restoration=False, estimate_state=False, start_control_at=0, use_graph=True,
                 solve_kernel="auto"):
        import torch
        self.torch = torch
        self.dev = torch.device("cuda", device)
        self.cl = centerline
        self.L = float(centerline.length)
        self.N, self.Ts = N, Ts
        self.lookahead, self.lookback, self.window = float(lookahead), float(lookback), float(window)
        self.plant_id = MODEL_IDS[plant]
        self.estimate_state = bool(estimate_state)       # agent.py:240-254 instead of the true plant state
        self.start_control_at = int(start_control_at)    # agent.py:66 (the agent waits 50 ticks)
        self.use_graph = bool(use_graph)
        self.keyed_dispatch = True
        # closed loop: every solve runs the build with the restoration phase ("always").  The two-stage
        # mode re-solves the problems that need the phase; in a warm-started loop a step takes ~1 ms per
        # problem and a re-solve 10-20 ms, which becomes the tail of every step (measured: 26.5 -> 32.0 ms).
        self.solver = BatchedMPC(N=N, Ts=Ts, model=nlp_model, device=device, on_device=True,
                                 runtime_params=runtime_params, solve_kernel=solve_kernel,
                                 restoration="always" if restoration is True else restoration)
        self.lib = self.solver.lib
        self.track = centerline.device_track(self.solver)
        self.stream = torch.cuda.Stream(device=self.dev)

    def close(self):
        if self.track:
            self.lib.mpcr_track_destroy(self.track)
            self.track = None
        self.solver.close()

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError(self.lib.mpcr_last_error(self.solver._h).decode())

    # ------------------------------------------------------------------
    def begin(self, x0, s0, u_prev0=None, params=None, laps=1.0, seg_bounds=None):
        """Upload the start states of R rollouts (x0 [R,6], s0 [R] numpy; params [R,5] per-vehicle
        runtime parameters = GA individuals) and allocate the device-resident state."""
        t = self.torch
        dev = self.dev
        R = x0.shape[0]
        up = lambda a: t.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
        f64 = lambda *s: t.zeros(*s, dtype=t.float64, device=dev)
        i32 = lambda *s: t.zeros(*s, dtype=t.int32, device=dev)
        S = RolloutSession()
        S.R = R
        S.x = up(x0)
        S.x_mpc = S.x.clone()
        S.progress = up(s0)
        S.yaw_prev = S.x[:, 2].clone()
        S.u_prev = f64(R, 2) if u_prev0 is None else up(u_prev0)
        S.params = None if params is None else up(params)
        S.cx, S.cy, S.dist, S.sign = f64(R, 5), f64(R, 5), f64(R), i32(R)
        S.travelled, S.max_err = f64(R), f64(R)
        S.lap_time = t.full((R,), float("nan"), dtype=t.float64, device=dev)
        n_seg = 0 if seg_bounds is None else len(seg_bounds) - 1
        S.seg_bounds = None if seg_bounds is None else up(np.asarray(list(seg_bounds), dtype=np.float64))
        S.seg_t = None if seg_bounds is None else t.full((R, n_seg), float("nan"), dtype=t.float64, device=dev)
        S.steps, S.fails, S.done, S.counters = i32(R), i32(R), i32(R), i32(4)
        # two plan buffers: the solve of tick k writes plan[k % 2] and warm-starts from plan[(k + 1) % 2]
        mk = lambda: dict(U=f64(R, 2, self.N), X=f64(R, 6, self.N + 1), S=f64(R, self.N + 1), eC=f64(R, self.N),
                          eL=f64(R, self.N), lam_g=None, obj=f64(R), status=i32(R), iters=i32(R))
        S.plan = [mk(), mk()]
        S.k = 0                      # host copy of the tick counter
        S.graph = None
        b = _capi.MpcrRolloutBufs()
        b.R, b.N, b.n_seg, b.plant_model = R, self.N, n_seg, self.plant_id
        b.estimate_state, b.start_control_at = int(self.estimate_state), self.start_control_at
        b.Ts, b.window, b.lookback, b.lookahead = self.Ts, self.window, self.lookback, self.lookahead
        b.target = float(laps) * self.L
        for name in ("x", "x_mpc", "progress", "yaw_prev", "u_prev", "cx", "cy", "dist", "sign", "travelled",
                     "lap_time", "max_err", "steps", "fails", "done", "counters"):
            setattr(b, name, getattr(S, name).data_ptr())
        b.seg_bounds = None if S.seg_bounds is None else S.seg_bounds.data_ptr()
        b.seg_t = None if S.seg_t is None else S.seg_t.data_ptr()
        S.bufs = b
        t.cuda.current_stream(dev).synchronize()
        return S

    def _tick(self, S, k, stream):
        """Launches of control step k on `stream` (no host synchronisation)."""
        st = stream.cuda_stream
        self._chk(self.lib.mpcr_rollout_prepare(self.track, C.byref(S.bufs), st))
        cur, prev = S.plan[k % 2], S.plan[(k + 1) % 2]
        if k >= self.start_control_at:
            warm = prev["U"] if k > self.start_control_at else None      # agent.py:205: last_controls
            # dispatch the vehicles whose previous solve took longest first
            self.solver.set_dispatch_keys(prev["iters"] if (warm is not None and self.keyed_dispatch) else None)
            self.solver.solve(S.x_mpc, S.u_prev, S.progress, S.cx, S.cy, params=S.params, u_warm=warm,
                              want_lam_g=False, out=cur, stream=st)
        self._chk(self.lib.mpcr_rollout_apply(self.solver._h, C.byref(S.bufs), cur["U"].data_ptr(),
                                              cur["status"].data_ptr(), st))

    def advance(self, S, K, sync=True):
        """K control steps of every rollout.  Returns converged solves of unfinished vehicles during
        these steps, the number of finished vehicles and the (device) travelled distances."""
        t = self.torch
        before = None
        if sync:
            before = S.counters.cpu().numpy().copy()
        with t.cuda.stream(self.stream):
            k_end = S.k + K
            while S.k < k_end:
                # the first ticks run ungraphed: cold start / warm-up phase, and they size the
                # solver's internal buffers (no allocation may happen during a capture)
                steady = S.k > self.start_control_at + 1
                if self.use_graph and steady and S.k % 2 == 0 and S.k + 2 <= k_end:
                    if S.graph is None:
                        g = t.cuda.CUDAGraph()
                        with t.cuda.graph(g, stream=self.stream):
                            self._tick(S, S.k, self.stream)
                            self._tick(S, S.k + 1, self.stream)
                        S.graph = g
                    S.graph.replay()
                    S.k += 2
                else:
                    self._tick(S, S.k, self.stream)
                    S.k += 1
        if not sync:
            return None
        self.stream.synchronize()
        after = S.counters.cpu().numpy()
        return dict(converged_solves=int(after[2] - before[2]), finished=int(after[1]), travelled=S.travelled,
                    tick=int(after[0]))

    def converged_so_far(self, S):
        """Converged solves of unfinished vehicles since begin() (one small device -> host read)."""
        self.stream.synchronize()
        return int(S.counters[2].item())

    def fetch_bytes(self, R):
        return R * (8 * (1 + 1 + 1 + 6) + 4 * 2)

    def fetch(self, S):
        """Results of the rollouts as numpy arrays (device -> host)."""
        self.stream.synchronize()
        ret = dict(lap_time=S.lap_time.cpu().numpy(), steps=S.steps.cpu().numpy(), fail_steps=S.fails.cpu().numpy(),
                   max_abs_error=S.max_err.cpu().numpy(), travelled=S.travelled.cpu().numpy(),
                   final_state=S.x.cpu().numpy())
        c = S.counters.cpu().numpy()
        ret["converged_solves"] = int(c[2])
        ret["finished"] = int(c[1])
        if S.seg_t is not None:
            st = S.seg_t.cpu().numpy()
            ret["seg_times"] = np.diff(np.concatenate([np.zeros((S.R, 1)), st], axis=1), axis=1)
        return ret

    # ------------------------------------------------------------------
    def run(self, x0, s0, u_prev0=None, params=None, max_steps=2000, laps=1.0, seg_bounds=None,
            record=False, check_every=32):
        """Roll R vehicles until each has advanced ``laps`` * L of progress (or max_steps).

        x0 [R,6], s0 [R] numpy; params [R,5] per-vehicle runtime parameters (GA individuals).
        Returns numpy arrays: lap_time [R] (nan if not finished), steps [R], fail_steps [R],
        max_abs_error [R], seg_times [R,n_seg] if seg_bounds is given, and the per-step history
        if ``record`` (one host round trip per step then)."""
        S = self.begin(x0, s0, u_prev0, params, laps, seg_bounds)
        R = S.R
        hist = [] if record else None
        while S.k < max_steps:
            if record:
                with self.torch.cuda.stream(self.stream):
                    k = S.k
                    self._tick(S, k, self.stream)
                    S.k += 1
                self.stream.synchronize()
                p = S.plan[k % 2]
                g = lambda a: a.cpu().numpy().copy()
                hist.append(dict(x=g(S.x_mpc), progress=g(S.progress), u0=g(S.u_prev), cx=g(S.cx), cy=g(S.cy),
                                 status=g(p["status"]), U=g(p["U"]), X=g(p["X"]), S=g(p["S"]), eC=g(p["eC"]),
                                 eL=g(p["eL"]), dist=g(S.dist), sign=g(S.sign), x_plant_next=g(S.x)))
                if int(S.counters[1].item()) >= R:
                    break
            else:
                r = self.advance(S, min(check_every, max_steps - S.k))
                if r["finished"] >= R:
                    break
        ret = self.fetch(S)
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
