"""Diagnostic for a parity-soak mismatch: prints, for the instances whose status / iteration counts differ from the oracle's,
both sequences around the first difference.  usage: python tools/soak_diag.py model horizon batch steps seed [multi]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import torch
import drone_attitude_control_b200 as pkg
from oracle import c_oracle as co
from test_gpu_parity import _fast_loop_inputs, MODEL_ID

model, N, B, S, seed = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
multi = len(sys.argv) > 6
sigma = 0.05 if seed % 2 else 0.0
refs, x0, noise, pc, pp = _fast_loop_inputs(B, S, seed=seed, mass_sigma=sigma)
loop = pkg.BatchedClosedLoop(model, batch=B, device=0, N_horizon=N)
loop.init(torch.tensor(x0.T.copy()), torch.tensor(np.ascontiguousarray(refs)), noise=torch.tensor(noise),
          p_ctrl=torch.tensor(pc.T.copy()), p_plant=torch.tensor(pp.T.copy()), n_steps=S).run(steps_per_launch=S if multi else 1)
got = {k: v.cpu().numpy() for k, v in loop.results().items()}
want = co.closed_loop(co.default_opts(MODEL_ID[model], N=N), refs, x0, noise, pc, pp, S)
bad = np.where(((got['status'] != want['status']) | (got['qp_iter'] != want['qp_iter'])).any(1))[0]
print('instances with a difference:', bad.tolist(), '| oracle non-zero statuses at (instance, step):', np.argwhere(want['status'] != 0).tolist())
for i in bad:
    d = np.where((got['status'][i] != want['status'][i]) | (got['qp_iter'][i] != want['qp_iter'][i]))[0]
    k = d[0]
    print(f'instance {i}: first difference at step {k}')
    print('  oracle status ', want['status'][i, max(0, k - 2):k + 4].tolist(), 'qp_iter', want['qp_iter'][i, max(0, k - 2):k + 4].tolist())
    print('  gpu    status ', got['status'][i, max(0, k - 2):k + 4].tolist(), 'qp_iter', got['qp_iter'][i, max(0, k - 2):k + 4].tolist())
    print('  max |dXsim| before the difference', float(np.abs(got['Xsim'][i, :k + 1] - want['Xsim'][i, :k + 1]).max()),
          'after', float(np.abs(got['Xsim'][i] - want['Xsim'][i]).max()))
