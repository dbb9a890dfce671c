"""A pipeline of batched solves: ``depth`` solver handles on their own CUDA streams, used round-robin.

One ``mpcr_solve_batch`` launch is a persistent kernel that drains a work queue, one problem per warp.
Its last milliseconds are a TAIL: a few long problems (100+ iterations, not predictable from the
inputs) still run while most warps have nothing left - 2 to 10 % of the launch, depending on where
the long problems happened to sit in the dispatch order.  When batches arrive one after another
(GA sweeps, parameter studies, multi-GPU shards) that idle time is recoverable: enqueue batch i+1 on
a second stream while batch i is still running; its blocks take over the SM slots as the blocks of
batch i retire, and with host arrays the copies of one batch overlap the kernels of the other.

Results are bit-identical to serial solves (tests/test_gpu_schedule.py): every problem is still
solved by one warp, only the launch boundaries overlap.  No counterpart in the reference
(control/MPC.py:152-181: one blocking solve per call).
"""
from .solver import BatchedMPC


class SolvePipeline:
    def __init__(self, depth=2, device=0, host=False, **solver_kw):
        import torch
        self.torch = torch
        self.dev = torch.device("cuda", device)
        self.depth = int(depth)
        self.host = bool(host)
        self.solvers = [BatchedMPC(device=device, on_device=not host, **solver_kw) for _ in range(self.depth)]
        self.streams = [torch.cuda.Stream(device=self.dev) for _ in range(self.depth)]
        self.results = [None] * self.depth
        self.outs = [None] * self.depth
        self.n = 0

    def close(self):
        self.drain()
        for s in self.solvers:
            s.close()

    def next_stream(self):
        """The stream the next ``submit`` will use (to enqueue work that must precede that solve)."""
        return self.streams[self.n % self.depth]

    def submit(self, x0, u_prev, s0, cx, cy, max_err=None, params=None, u_warm=None, want_lam_g=True):
        """Enqueue one batch and return its ticket.  Does not wait for the batch; waits (host side) for
        the batch that used the same slot ``depth`` submits ago, whose result buffers are reused.
        Host mode: pass page-locked arrays and leave them untouched until ``result(ticket)``."""
        t = self.torch
        k = self.n % self.depth
        st = self.streams[k]
        if self.results[k] is not None:
            st.synchronize()
        solver = self.solvers[k]
        B = x0.shape[0]
        if self.outs[k] is None or self.outs[k]["obj"].shape[0] != B or \
                (self.outs[k]["lam_g"] is None) == bool(want_lam_g):
            self.outs[k] = self._alloc(solver, B, want_lam_g)
        if self.host:
            res = solver.solve(x0, u_prev, s0, cx, cy, max_err=max_err, params=params, u_warm=u_warm,
                               want_lam_g=want_lam_g, stream=st.cuda_stream, out=self.outs[k], sync=False)
        else:
            st.wait_stream(t.cuda.current_stream(self.dev))      # inputs produced on the caller's stream
            res = solver.solve(x0, u_prev, s0, cx, cy, max_err=max_err, params=params, u_warm=u_warm,
                               want_lam_g=want_lam_g, stream=st.cuda_stream, out=self.outs[k])
        self.results[k] = res
        self.n += 1
        return self.n - 1

    def _alloc(self, solver, B, want_lam_g):
        if self.host:
            return solver.alloc_host_outputs(B, want_lam_g=want_lam_g, pinned=True)
        t = self.torch
        N, ls = solver.N, solver.lam_g_stride
        mk = lambda *s: t.empty(*s, dtype=t.float64, device=self.dev)
        return dict(U=mk(B, 2, N), X=mk(B, 6, N + 1), S=mk(B, N + 1), eC=mk(B, N), eL=mk(B, N),
                    lam_g=mk(B, ls) if want_lam_g else None, obj=mk(B),
                    status=t.empty(B, dtype=t.int32, device=self.dev),
                    iters=t.empty(B, dtype=t.int32, device=self.dev))

    def result(self, ticket):
        """The BatchResult of a ticket, complete.  Its arrays belong to the slot: they are overwritten
        by the submit ``depth`` tickets later."""
        if not (self.n - self.depth <= ticket < self.n):
            raise ValueError("ticket is no longer (or not yet) in the pipeline")
        k = ticket % self.depth
        self.streams[k].synchronize()
        return self.results[k]

    def join(self, stream=None):
        """Make ``stream`` (default: the current stream) wait for everything submitted so far - a
        device-side join, no host synchronisation."""
        t = self.torch
        stream = stream or t.cuda.current_stream(self.dev)
        for st in self.streams:
            stream.wait_stream(st)

    def drain(self):
        for st in self.streams:
            st.synchronize()
