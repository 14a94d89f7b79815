#!/usr/bin/env python3
"""bench.py - NMPC solves/s of the fused closed-loop step (BASELINE.json metric) on N B200s, one process per GPU.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--model force|jerk] [--batch B]

A "step" is one control step for the whole batch of drones: yref windowing + x0 embedding + SQP/HPIPM solve + converter
+ plant step + logs, ONE kernel launch (bnmpc_closed_loop_run).  Workload at every N: BASELINE config 2, "force_model
batched closed-loop, 4096 drones with randomised x0 and trajectories, FP64" per GPU (weak scaling: each rank owns its
own 4096 instances, no data-path collective).  Inputs are resident in HBM for `value`; `e2e` is the same metric through
the AcadosOcpSolver-style shim with pinned HOST buffers (yref window + x0 in, u0 + status out, every step).

--impl reference times the CPU restatement of the reference's path (oracle/nmpc_oracle.c, all host threads) - acados
itself is not installable here - on the same workload, a bounded sample of instances per step.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = 'nmpc_solves_per_sec'
UNIT = 'solves/s'


def flops_per_solve(nblk, n, m, N, erk_stages, qp_iters, sqp_iters=1.0):
    """Algorithmic flops (SURVEY 8d formulas, applied to the block structure the solver actually factorises):
    per stage factorisation s n^2 + s^2 n + s^3/3, one KKT solve 4n^2 + 4sn + 2s^2, residuals/barrier 2(2sn + s^2) + 10s;
    per IPM iteration N (F_f + 2 F_s + F_r); linearisation N (S_rk 2 n^2 s + 2 s^2)."""
    s = n + m
    f_f = s * n * n + s * s * n + s ** 3 / 3.0
    f_s = 4 * n * n + 4 * s * n + 2 * s * s
    f_r = 2 * (2 * s * n + s * s) + 10 * s
    f_it = nblk * N * (f_f + 2 * f_s + f_r)
    f_lin = nblk * N * (erk_stages * 2 * n * n * s + 2 * s * s)
    return sqp_iters * f_lin + qp_iters * f_it


def bytes_per_solve(nx, nu, N):
    """Algorithmic HBM bytes of one closed-loop step: x0 + yref window in, u0 + status/iters out, plus the
    rollout's state r/w, noise and p."""
    return nx * 8 + (N * (nx + nu) + nx) * 8 + nu * 8 + 8 + 64 + 8 + 16


class ClockSampler:
    """SM clock + throttle reasons during the timed region (pynvml; falls back to nvidia-smi)."""

    def __init__(self, index):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thr = None

    def _run(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            names = {getattr(nv, k): k for k in dir(nv) if k.startswith('nvmlClocksEventReason') or k.startswith('nvmlClocksThrottleReason')}
            while not self._stop.is_set():
                self.samples.append(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, name in names.items():
                    if isinstance(bit, int) and bit and (r & bit) == bit and bit & (bit - 1) == 0:
                        self.reasons.add(name.replace('nvmlClocksEventReason', '').replace('nvmlClocksThrottleReason', ''))
                time.sleep(0.05)
        except Exception as e:   # noqa: BLE001
            self.reasons.add(f'sampler_error:{type(e).__name__}')

    def __enter__(self):
        self._thr = threading.Thread(target=self._run, daemon=True)
        self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thr.join(timeout=2)

    def summary(self):
        med = float(np.median(self.samples)) if self.samples else None
        rs = sorted(r for r in self.reasons if r not in ('None', 'GpuIdle', 'ApplicationsClocksSetting'))
        return {'sm_mhz': med, 'sm_max_mhz': self.max_mhz, 'reasons': rs, 'samples': len(self.samples)}


def make_inputs(lo, hi, n_steps, rows, device, mass_sigma=0.0):
    """BASELINE config 2 / 4 / SURVEY 8d inputs for the global instances lo..hi-1 (per-instance seeds keyed by the global
    id: the numbers do not depend on how many ranks share the batch)."""
    from drone_attitude_control_b200.generate_trajectory import gen_circle_traj_batched
    from drone_attitude_control_b200.sharding import instance_inputs
    inp = instance_inputs(lo, hi, n_steps, mass_sigma=mass_sigma)
    ref = gen_circle_traj_batched(500, rows - 500, inp['radius'], inp['center'], inp['phase'], device=device)      # [rows, 8, b]
    x0 = ref[0, :4, :].clone() + inp['dx0'].to(device)
    return ref, x0, inp['noise'].to(device), inp


def run_ours(args):
    import torch
    import torch.distributed as dist
    import drone_attitude_control_b200 as pkg
    from drone_attitude_control_b200 import _lib
    import ctypes as C

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    B, K, W = args.batch, args.steps, args.warmup
    N = args.horizon
    rows = max(500 + N, W + K + N + 1)
    ref, x0, noise, inp = make_inputs(rank * B, (rank + 1) * B, W + K, rows, device=dev, mass_sigma=args.mass_sigma)   # weak scaling
    ref_im = ref.permute(2, 0, 1).contiguous()                                             # [B, rows, 8] instance-major
    loop = pkg.BatchedClosedLoop(args.model, batch=B, device=local, precision=args.precision, N_horizon=N, rti=args.rti)
    p_plant = None
    if args.mass_sigma > 0:                                                                # BASELINE config 4: model mismatch
        import torch as _t
        p_plant = _t.stack([0.03277 * inp['mass_scale'], _t.full((B,), 9.81, dtype=_t.float64)])
    ref_arg = pkg.CircleRef(inp['radius'], inp['center'], inp['phase'], n=rows - N) if args.ref == 'circle' else ref_im
    loop.init(x0, ref_arg, noise=noise, p_plant=p_plant, n_steps=W + K, log=True)
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev) if args.flush_l2 else None

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    loop.run(W)                                   # warm-up steps (untimed)
    barrier()
    l0 = loop.solver.launch_count()
    spl = max(1, args.steps_per_launch)
    nl = (K + spl - 1) // spl
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(nl)]
    with ClockSampler(local) as clk:
        t0 = time.perf_counter()
        for i in range(nl):
            if flush is not None:
                flush.fill_(i & 0xff)             # evict L2 between timed launches (not timed)
            evs[i][0].record(stream)
            loop.run(min(spl, K - i * spl), steps_per_launch=spl)
            evs[i][1].record(stream)
        barrier()
        wall = time.perf_counter() - t0
    launches = loop.solver.launch_count() - l0
    step_ms = np.array([a.elapsed_time(b) for a, b in evs])
    tot_ms = float(step_ms.sum())
    tmax = torch.tensor([tot_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
    tot_ms_all = float(tmax[0])
    res = loop.results()
    qp = res['qp_iter'][:, W:W + K].double()
    st = res['status'][:, W:W + K]
    stats = torch.tensor([float(qp.sum()), float((st != 0).sum()), float(res['cost'].sum()), float(res['aed'].sum()), float(B)],
                         dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats)                    # the only collective: final per-step metrics (< 1 KB)
    value = world * B * K / (tot_ms_all * 1e-3)

    # ---- e2e: the AcadosOcpSolver-style call sequence with pinned host buffers, every step ---------------------------
    e2e = None
    if not args.skip_e2e:
        s = pkg.BatchedAcadosOcpSolver(args.model, batch=B, device=local, precision=args.precision, N_horizon=N, rti=args.rti,
                                       numpy_io=False)
        ny, nx, nu = s.ny, s.nx, s.nu
        Ke = min(K, args.e2e_steps)
        ref_h = ref_im.cpu()                                      # [B, rows, 8]
        ycols = list(range(nx)) + [nx + j for j in range(nu)] if nx == 6 else [0, 1, 2, 3, 4, 5]
        yh = [torch.cat([ref_h[:, i:i + N, ycols].reshape(B, N * ny), ref_h[:, i + N, :nx]], 1).contiguous().pin_memory()
              for i in range(W + Ke)]
        xs_log = res['Xsim'].cpu()                                 # states the fused loop visited: realistic x0 stream
        acc0 = torch.tensor([0.0, 9.81], dtype=torch.float64).expand(B, 2)
        x0h = [(xs_log[:, i, :] if nx == 4 else torch.cat([xs_log[:, i, :], acc0], 1)).contiguous().pin_memory() for i in range(W + Ke)]
        u_host = torch.empty((B, nu), dtype=torch.float64).pin_memory()
        st_host = torch.empty(B, dtype=torch.int32).pin_memory()

        per = N * ny + nx
        ydev = [torch.empty((B, per), dtype=torch.float64, device=dev) for _ in range(2)]
        yev = [torch.cuda.Event(), torch.cuda.Event()]
        copy_stream = torch.cuda.Stream(device=dev)

        def prefetch(i):      # upload the reference window of step i on the copy stream (overlaps the solve of step i-1)
            with torch.cuda.stream(copy_stream):
                ydev[i % 2].copy_(yh[i], non_blocking=True)
                yev[i % 2].record(copy_stream)

        def e2e_step(i, pipelined, last):
            if pipelined:
                torch.cuda.current_stream().wait_event(yev[i % 2])
                s.set_yref_all(ydev[i % 2])
                s.solve_for_x0_into(x0h[i], u_host, st_host, wait=False)   # enqueue: x0 in, solve, u0 + status out (pinned host)
                if not last:
                    prefetch(i + 1)      # behind this step's x0 on the H2D engine, concurrent with the solve
                s.synchronize()          # u0 and status of this step are in host memory
            else:
                s.set_yref_all(yh[i])
                s.solve_for_x0_into(x0h[i], u_host, st_host)    # returns when u0 + status are in host memory

        def e2e_run(pipelined):
            s.reset()
            if pipelined:
                prefetch(0)
            for i in range(W):
                e2e_step(i, pipelined, False)
            barrier()
            t0 = time.perf_counter()
            for i in range(W, W + Ke):
                e2e_step(i, pipelined, i == W + Ke - 1)
            barrier()
            te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
            return world * B * Ke / float(te[0])

        e2e_serial = e2e_run(False)
        e2e_pipe = e2e_run(True)
        e2e = {'value': e2e_pipe, 'unit': UNIT, 'steps': Ke, 'serial_value': e2e_serial,
               'h2d_bytes_per_step': int(B * (N * ny + nx + 2 * nx) * 8), 'd2h_bytes_per_step': int(B * (nu * 8 + 4)),
               'api': 'BatchedAcadosOcpSolver.set_yref_all (OCP.set_up_ocp) + solve_for_x0 (x0 in, u0 and status out) with pinned host buffers, every step; '
                      'value: x0 upload, solve and u0/status download are enqueued asynchronously, then the upload of the next step\'s '
                      'reference window (known in advance) is enqueued on a copy stream and overlaps the solve, then the step waits; '
                      'serial_value: everything on one stream'}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---- roofline of the dominant kernel (k_loop_step = the whole step) ----------------------------------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
    except Exception:
        pass
    tf = C.c_double()
    _lib.check(_lib.lib().bnmpc_measure_fma_peak(local, _lib.FP32 if args.precision == 'fp32' else _lib.FP64, C.byref(tf)))
    dims = dict(force=(2, 2, 1, 4, 2, 4), jerk=(2, 3, 1, 6, 2, 1), force_dense=(1, 4, 2, 4, 2, 4), thrust=(1, 4, 2, 4, 2, 4))[args.model]
    nblk, n, m, nx, nu, erk = dims
    qp_local = float(qp.sum())
    fl = flops_per_solve(nblk, n, m, N, erk, qp_local / (B * K)) * B                      # per launch, this rank
    by = bytes_per_solve(nx, nu, N) * B
    ms_launch = tot_ms / K
    ach_tf = fl / (ms_launch * 1e-3) * 1e-12
    ach_gbs = by / (ms_launch * 1e-3) * 1e-9
    hbm_peak = peaks.get('hbm_gbs', 6650.0)
    traffic, ncu_view = None, None
    try:       # dram bytes per launch and pipe utilisation of the dominant kernel from the committed ncu capture of this config
        ncu_view = json.load(open(os.path.join(ROOT, 'profiles', 'traffic.json'))).get(f'{args.model}_{args.precision}_B{B}')
        traffic = ncu_view.get('traffic_bytes') if ncu_view else None
    except Exception:
        pass
    line = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
        'ms_per_step': tot_ms_all / K, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64' if args.precision == 'fp64' else 'f32', 'data': 'synthetic',
        'config': {'workload': f'{args.model}_model batched closed loop, {B} drones per GPU with randomised x0 and circle trajectories '
                               f'(BASELINE config 2), N_horizon {N}, {"SQP_RTI" if args.rti else "SQP to tol 1e-6"} + HPIPM-style IPM, '
                               f'noise sigma 0.01',
                   'batch_per_gpu': B, 'horizon': N, 'controller': args.model, 'plant_mass_sigma': args.mass_sigma,
                   'reference': 'per-instance table [B, rows, 8] in HBM' if args.ref == 'table' else 'generated in the kernel (CircleRef)',
                   'l2': 'flushed between timed steps (256 MiB write, untimed)' if args.flush_l2 else 'not flushed',
                   'timing': 'sum of per-step CUDA-event durations on the launch stream, max over ranks'},
        'p50_step_latency_ms': float(np.median(step_ms)), 'p99_step_latency_ms': float(np.percentile(step_ms, 99)),
        'qp_iter_mean': float(stats[0]) / (world * B * K), 'nonzero_status': int(stats[1]),
        'closed_loop_cost_mean': float(stats[2]) / float(stats[4]), 'aed_mean': float(stats[3]) / float(stats[4]),
        'gpu_launches': int(launches), 'wall_s_timed_region': wall,
        'clocks': clk.summary(),
        'e2e': e2e,
        'roofline': {'bound': 'fp64' if args.precision == 'fp64' else 'fp32', 'achieved': ach_tf, 'peak': tf.value, 'unit': 'TFLOP/s',
                     'frac': ach_tf / tf.value if tf.value else None, 'traffic': traffic,
                     'peak_source': 'measured live: bnmpc_measure_fma_peak (MEASURED_PEAKS.json has no vector-pipe figure)',
                     'kernel': 'k_loop_step', 'flops_per_launch': fl, 'ms_per_launch': ms_launch, 'ncu': ncu_view,
                     'hbm': {'bound': 'hbm', 'achieved': ach_gbs, 'peak': hbm_peak, 'unit': 'GB/s', 'frac': ach_gbs / hbm_peak,
                             'bytes_per_launch': by,
                             'peak_source': 'MEASURED_PEAKS.json hbm_gbs' if 'hbm_gbs' in peaks else 'fallback 6650'}},
    }
    if not args.skip_cpu and world == 1:
        line['cpu_baseline'] = cpu_baseline(args, sample_instances=args.cpu_instances, sample_steps=args.cpu_steps)
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def cpu_inputs(args, B, S):
    from oracle import nmpc_oracle as o
    rng = np.random.default_rng(2026)
    N = args.horizon
    refs = np.stack([o.gen_circle_traj(n_horizon=max(N, 30), radius=rng.uniform(0.5, 1.0), center=rng.uniform(-0.15, 0.15, 2),
                                       phase=rng.uniform(0, 2 * np.pi)) for _ in range(B)])
    x0 = refs[:, 0, :4] + rng.uniform(-0.05, 0.05, (B, 4))
    noise = rng.normal(0, 0.01, (S, B))
    pp = np.repeat(np.array([[o.MASS, o.GRAVITY_ACC]]), B, 0)
    return refs, x0, noise, pp


def cpu_baseline(args, sample_instances, sample_steps):
    """The oracle (CPU restatement, kind 'port': acados itself cannot be installed here) on the host cores."""
    from oracle import c_oracle as co
    model = co.MODEL_JERK if args.model.startswith('jerk') else co.MODEL_FORCE
    cores = co.lib().orc_num_cores()
    refs, x0, noise, pp = cpu_inputs(args, sample_instances, sample_steps)
    opts = co.default_opts(model, N=args.horizon, rti=args.rti)
    co.closed_loop(opts, refs[:cores], x0[:cores], noise[:2, :cores], pp[:cores], pp[:cores], 2, outputs=False)     # warm
    t0 = time.perf_counter()
    co.closed_loop(opts, refs, x0, noise, pp, pp, sample_steps, nthreads=cores, outputs=False)
    dt = time.perf_counter() - t0
    return {'value': sample_instances * sample_steps / dt, 'unit': UNIT, 'cores': cores, 'kind': 'port',
            'sample': f'{sample_instances} of the workload\'s instances x {sample_steps} closed-loop steps, oracle/nmpc_oracle.c, '
                      f'{cores} pthreads, {dt:.1f} s'}


def run_reference(args):
    """--impl reference: the CPU restatement of the reference's path on the host cores, same metric and config."""
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from oracle import c_oracle as co
    model = co.MODEL_JERK if args.model.startswith('jerk') else co.MODEL_FORCE
    cores = co.lib().orc_num_cores()
    K, W = args.steps, args.warmup
    Bs = min(args.cpu_instances, args.batch)
    refs, x0, noise, pp = cpu_inputs(args, Bs, W + K)
    opts = co.default_opts(model, N=args.horizon, rti=args.rti)
    # warm-up steps, then K timed steps continuing the same closed loop (the oracle API runs whole loops: time the
    # (W+K)-step run and the W-step run and subtract)
    t0 = time.perf_counter(); co.closed_loop(opts, refs, x0, noise, pp, pp, W, nthreads=cores, outputs=False); tw = time.perf_counter() - t0
    t0 = time.perf_counter(); co.closed_loop(opts, refs, x0, noise, pp, pp, W + K, nthreads=cores, outputs=False); tk = time.perf_counter() - t0
    dt = max(tk - tw, 1e-9)
    value = Bs * K / dt
    world = int(os.environ.get('WORLD_SIZE', '1'))
    sample = f'{Bs} of the workload\'s {args.batch} instances per step, oracle/nmpc_oracle.c, {cores} pthreads'
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': K, 'warmup': W,
        'ms_per_step': dt / K * 1e3, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': f'{args.model}_model batched closed loop (BASELINE config 2), N_horizon {args.horizon}; CPU restatement of the '
                               f'reference path (acados is not installable here), bounded sample', 'batch_per_gpu': args.batch,
                   'horizon': args.horizon, 'controller': args.model},
        'cpu_baseline': {'value': value, 'unit': UNIT, 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': value, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=100)
    ap.add_argument('--warmup', type=int, default=10)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--model', default='force', choices=['force', 'jerk', 'force_dense', 'thrust'])
    ap.add_argument('--batch', type=int, default=4096, help='instances per GPU')
    ap.add_argument('--horizon', type=int, default=30)
    ap.add_argument('--precision', default='fp64', choices=['fp64', 'fp32'])
    ap.add_argument('--rti', action='store_true')
    ap.add_argument('--ref', default='table', choices=['table', 'circle'], help='trajectory table in HBM, or generated in the kernel')
    ap.add_argument('--mass-sigma', type=float, default=0.0, help='BASELINE config 4: plant mass = 0.03277 (1 + N(0, sigma)) clipped to +-15 %%')
    ap.add_argument('--steps-per-launch', type=int, default=1)
    ap.add_argument('--no-flush-l2', dest='flush_l2', action='store_false')
    ap.add_argument('--skip-e2e', action='store_true')
    ap.add_argument('--skip-cpu', action='store_true')
    ap.add_argument('--e2e-steps', type=int, default=50)
    ap.add_argument('--cpu-instances', type=int, default=4096)
    ap.add_argument('--cpu-steps', type=int, default=300, help='closed-loop steps of the cpu_baseline sample (~10 s on 16 cores)')
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
