"""Fused, device-resident closed loop for many drones: follow_trajectory of the reference
(src/force_model/controller.py:8-56, src/jerk_model/controller.py:8-58) for `batch` instances at once.

Per control step ONE kernel does, per instance: yref windowing (set_up_ocp), x0 embedding, the SQP/HPIPM solve,
Converter.convert, the plant step (simulate_next_x) with the supplied noise draw, and the logged quantities.
Plant state, carried acceleration (jerk model), cost and |error| sums stay on the device between steps.
"""
import ctypes as C

import numpy as np
import torch

from ._lib import ClosedLoopArgs, check, lib
from .acados_shim import BatchedAcadosOcpSolver
from .params import ExperimentParameters

p = ExperimentParameters()


class CircleRef:
    """Per-instance circle reference generated on the device (no table in HBM): gen_circle_traj of the reference
    (src/generate_trajectory.py:7-28) with n samples per revolution; params [B, 4] = (radius, centre_x, centre_z, phase)."""

    def __init__(self, radius, center, phase, n=500):
        radius = torch.as_tensor(radius, dtype=torch.float64)
        center = torch.as_tensor(center, dtype=torch.float64)
        phase = torch.as_tensor(phase, dtype=torch.float64)
        self.params = torch.stack([radius, center[:, 0], center[:, 1], phase], 1).contiguous()
        self.n = int(n)


class PhiloxNoise:
    """Plant noise drawn on the device: eps(instance, step) = std * N(0, 1) from Philox4x32-10 keyed by `seed` with the
    counter (first_instance + i, step) - the scalar np.random.normal(0, noise) of the reference's simulate_next_x
    (src/force_model/ocp.py:114-115), without a [steps, batch] array and independent of how the batch is sharded."""

    def __init__(self, seed=2026, std=p.noise, first_instance=0):
        self.seed, self.std, self.first_instance = int(seed), float(std), int(first_instance)


class BatchedClosedLoop:
    def __init__(self, model='force', batch=1, device=0, precision='fp64', N_horizon=None, rti=False, **overrides):
        self.solver = BatchedAcadosOcpSolver(model, batch=batch, device=device, precision=precision,
                                             N_horizon=p.N_horizon if N_horizon is None else N_horizon, rti=rti,
                                             numpy_io=False, **overrides)
        self.batch, self.device, self.N = self.solver.batch, self.solver.device, self.solver.N
        self.step = 0
        self._logs = {}

    def _dev(self, t, shape=None, dtype=torch.float64):
        if t is None:
            return None
        t = torch.as_tensor(t, dtype=dtype).to(self.device).contiguous()
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError(f'expected shape {tuple(shape)}, got {tuple(t.shape)}')
        return t

    def init(self, x0, ref, noise=None, p_ctrl=None, p_plant=None, n_steps=None, log=True):
        """x0 [4, B]; ref [rows, 8] (one table shared by all drones), [rows, 8, B] (batch-minor), [B, rows, 8]
        (instance-major, preferred: a warp reads its window contiguously) or a CircleRef (no table); noise [n_steps, B] or None; p_ctrl / p_plant
        [2, B] or None (nominal mass 0.03277, g 9.81).  float64."""
        B = self.batch
        self.x0 = self._dev(x0, (4, B))
        if isinstance(ref, CircleRef):                      # generated on the fly from (radius, centre, phase)
            self.ref = self._dev(ref.params, (B, 4))
            self.ref_layout, self.ref_rows = 3, ref.n + self.N
        else:
            self.ref = self._dev(ref)
            if self.ref.dim() == 2:
                assert self.ref.shape[1] == 8
                self.ref_layout, self.ref_rows = 1, int(self.ref.shape[0])
            elif tuple(self.ref.shape[1:]) == (8, B):
                self.ref_layout, self.ref_rows = 0, int(self.ref.shape[0])
            else:
                assert self.ref.shape[0] == B and self.ref.shape[2] == 8, tuple(self.ref.shape)
                self.ref_layout, self.ref_rows = 2, int(self.ref.shape[1])
        self.n_steps = int(n_steps if n_steps is not None else self.ref_rows - self.N)
        assert self.ref_rows >= self.n_steps + self.N
        self.philox = noise if isinstance(noise, PhiloxNoise) else None
        self.noise = self._dev(noise)[:self.n_steps].contiguous() if noise is not None and self.philox is None else None
        assert self.noise is None or tuple(self.noise.shape) == (self.n_steps, B)
        self.p_ctrl = self._dev(p_ctrl, (2, B)) if p_ctrl is not None else None
        self.p_plant = self._dev(p_plant, (2, B)) if p_plant is not None else None
        ptr = lambda t: C.c_void_p(t.data_ptr()) if t is not None else None
        check(lib().bnmpc_closed_loop_init(self.solver.handle, ptr(self.x0), ptr(self.p_ctrl), ptr(self.p_plant)))
        self.step = 0
        self._logs = {}
        if log:
            S = self.n_steps
            z = lambda *sh, dt=torch.float64: torch.zeros(sh, dtype=dt, device=self.device)
            self._logs = dict(Xsim=z(S + 1, 4, B), U_plant=z(S, 2, B), U_ctrl=z(S, 2, B), a=z(S, 2, B),
                              status=z(S, B, dt=torch.int32), qp_iter=z(S, B, dt=torch.int32))
            self._logs['Xsim'][0] = self.x0
        return self

    def circle_table(self, rows=None):
        """The table [B, rows, 8] the on-the-fly circle reference corresponds to (same device arithmetic)."""
        assert self.ref_layout == 3
        rows = self.ref_rows if rows is None else int(rows)
        out = torch.empty((self.batch, rows, 8), dtype=torch.float64, device=self.device)
        check(lib().bnmpc_gen_circle_table(self.solver.handle, C.c_void_p(self.ref.data_ptr()), rows, C.c_void_p(out.data_ptr())))
        return out

    def noise_array(self):
        """The draws of the PhiloxNoise of this loop as an array [n_steps, B] (what the fused loop adds, bit for bit)."""
        assert self.philox is not None
        out = torch.empty((self.n_steps, self.batch), dtype=torch.float64, device=self.device)
        check(lib().bnmpc_philox_noise(self.solver.handle, self.philox.seed, self.philox.std, self.philox.first_instance, 0,
                                       self.n_steps, C.c_void_p(out.data_ptr())))
        return out

    def run(self, n_steps=None, steps_per_launch=1):
        """Advance `n_steps` control steps (default: the rest).  steps_per_launch = 1: one kernel launch per control step
        (every launch ends with all drones at the same step - the latency path).  steps_per_launch = k > 1: up to k control
        steps per launch; a drone keeps its working set on chip for a chunk of consecutive steps and drones advance
        independently inside the launch (Monte-Carlo throughput path, same results)."""
        n = self.n_steps - self.step if n_steps is None else int(n_steps)
        a = ClosedLoopArgs()
        a.n_steps, a.first_step, a.ref_rows = n, self.step, self.ref_rows
        a.ref_shared, a.log_stride = self.ref_layout, self.n_steps
        a.steps_per_launch = int(steps_per_launch)
        a.ref = self.ref.data_ptr()
        a.noise = self.noise.data_ptr() if self.noise is not None else None
        if self.philox is not None:
            a.noise_philox, a.noise_seed, a.noise_std, a.first_instance = 1, self.philox.seed, self.philox.std, self.philox.first_instance
        L = self._logs
        if L:
            a.Xsim, a.U_plant, a.U_ctrl, a.a_log = (L[k].data_ptr() for k in ('Xsim', 'U_plant', 'U_ctrl', 'a'))
            a.status, a.qp_iter = L['status'].data_ptr(), L['qp_iter'].data_ptr()
        check(lib().bnmpc_closed_loop_run(self.solver.handle, C.byref(a)))
        self.step += n
        return self

    def state(self):
        """(closedLoopCost [B], sum |pref - psim| [B], plant state [4, B], carried acceleration [2, B]) so far"""
        B = self.batch
        z = lambda *sh: torch.empty(sh, dtype=torch.float64, device=self.device)
        cost, err, x, acc = z(B), z(B), z(4, B), z(2, B)
        check(lib().bnmpc_closed_loop_state(self.solver.handle, *(C.c_void_p(t.data_ptr()) for t in (cost, err, x, acc))))
        return cost, err, x, acc

    def failures(self):
        """int32 [B]: control steps so far whose solve returned a non-zero status (the reference raises on the first one;
        the fused loop keeps stepping and counts)."""
        out = torch.empty(self.batch, dtype=torch.int32, device=self.device)
        check(lib().bnmpc_closed_loop_failures(self.solver.handle, C.c_void_p(out.data_ptr()), 1))
        return out

    def results(self):
        """Per-instance outputs in the reference's shapes: cost [B], AvgEucDist [B] (calc_aed over the steps run), and the
        logs Xsim [B, S+1, 4], a [B, S, 2], U_opt_plant [B, S, 2] (+ U_ctrl, status, qp_iter)."""
        cost, err, _, _ = self.state()
        out = dict(cost=cost, aed=err / (2.0 * max(self.step, 1)), failures=self.failures())
        for k, v in self._logs.items():
            out[k] = v.permute(2, 0, 1) if v.dim() == 3 else v.t()
        return out


def follow_trajectory_batched(model, ref, x0, noise=None, p_ctrl=None, p_plant=None, n_steps=None, device=0, precision='fp64',
                              steps_per_launch=1, **kw):
    """follow_trajectory for B drones in one call; returns the dict of BatchedClosedLoop.results()."""
    x0 = torch.as_tensor(x0, dtype=torch.float64)
    loop = BatchedClosedLoop(model, batch=x0.shape[1], device=device, precision=precision, **kw)
    loop.init(x0, ref, noise=noise, p_ctrl=p_ctrl, p_plant=p_plant, n_steps=n_steps)
    loop.run(steps_per_launch=steps_per_launch)
    return loop.results()
