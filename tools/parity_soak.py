#!/usr/bin/env python3
"""Parity soak: EVERY instance of full-size closed-loop batches against the C oracle (oracle/nmpc_oracle.c, all host
threads), over several seeds, both models, nominal and perturbed plant mass.  Prints one line per run and a summary;
the committed output is profiles/r01_parity_soak.txt.  The oracle is the checker here, nothing it computes is shipped.
Usage (GPU box): python tools/parity_soak.py [--seeds 6] [--batch 4096] [--steps 40] [--horizon 30] [--multi]
(--horizon 50 / 100 or a small --batch select the kernels that run an instance on 2 / 4 warps - the factorisation scan and the
multi-warp stage scans; --multi runs all steps in one launch.)"""
import argparse, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import drone_attitude_control_b200 as pkg
from oracle import c_oracle as co
from test_gpu_parity import _fast_loop_inputs, MODEL_ID

ap = argparse.ArgumentParser()
ap.add_argument('--seeds', type=int, default=6); ap.add_argument('--batch', type=int, default=4096); ap.add_argument('--steps', type=int, default=40)
ap.add_argument('--horizon', type=int, default=30); ap.add_argument('--multi', action='store_true')
ap.add_argument('--spread', type=float, default=0.05, help='half-width of the uniform offset of x0 from the reference (default: the 0.05 of the parity tests; larger values drive inputs and states into their bounds)')
ap.add_argument('--models', default='force,jerk')
args = ap.parse_args()
B, S = args.batch, args.steps
tot = dict(solves=0, status_mismatch=0, iter_mismatch=0, nonzero_status=0, worst_dx=0.0, worst_du=0.0)
for model in args.models.split(','):
    for seed in range(args.seeds):
        sigma = 0.05 if seed % 2 else 0.0
        refs, x0, noise, pc, pp = _fast_loop_inputs(B, S, seed=1000 + seed, mass_sigma=sigma)
        x0 = refs[:, 0, :4] + (x0 - refs[:, 0, :4]) * (args.spread / 0.05)
        loop = pkg.BatchedClosedLoop(model, batch=B, device=0, N_horizon=args.horizon)
        loop.init(torch.tensor(x0.T.copy()), torch.tensor(np.ascontiguousarray(refs)), noise=torch.tensor(noise),
                  p_ctrl=torch.tensor(pc.T.copy()), p_plant=torch.tensor(pp.T.copy()), n_steps=S).run(steps_per_launch=S if args.multi else 1)
        got = {k: v.cpu().numpy() for k, v in loop.results().items()}
        t0 = time.perf_counter()
        want = co.closed_loop(co.default_opts(MODEL_ID[model], N=args.horizon), refs, x0, noise, pc, pp, S)
        t_or = time.perf_counter() - t0
        sm = int((got['status'] != want['status']).sum()); im = int((got['qp_iter'] != want['qp_iter']).sum())
        # instances whose status AND iteration counts agree everywhere are compared value by value
        same = ((got['status'] == want['status']) & (got['qp_iter'] == want['qp_iter'])).all(1)
        dx = float(np.abs(got['Xsim'][same] - want['Xsim'][same]).max()); du = float(np.abs(got['U_ctrl'][same] - want['U_ctrl'][same]).max())
        nz = int((want['status'] != 0).sum())
        print(f'{model:5s} N {args.horizon} B {B} {"multi-step" if args.multi else "per-step"} spread {args.spread} seed {1000 + seed} mass_sigma {sigma:.2f}: {B * S} solves, status mismatches {sm}, qp_iter mismatches {im}, '
              f'oracle non-zero statuses {nz}, max |dXsim| {dx:.2e}, max |du0| {du:.2e}  (oracle {t_or:.1f} s)', flush=True)
        tot['solves'] += B * S; tot['status_mismatch'] += sm; tot['iter_mismatch'] += im; tot['nonzero_status'] += nz
        tot['worst_dx'] = max(tot['worst_dx'], dx); tot['worst_du'] = max(tot['worst_du'], du)
print('TOTAL', tot)
sys.exit(0 if tot['status_mismatch'] == 0 and tot['iter_mismatch'] == 0 and tot['worst_dx'] < 1e-8 else 1)
