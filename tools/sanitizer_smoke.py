import sys; sys.path.insert(0, '.')
import numpy as np, torch
from drone_attitude_control_b200 import BatchedClosedLoop, BatchedAcadosOcpSolver
from oracle import nmpc_oracle as o
rng = np.random.default_rng(1)
for model in ('force', 'jerk', 'thrust'):
    B, S = 9, 3
    ref = o.gen_circle_traj()
    if model == 'thrust':
        ref[:, 4] = 0.0; ref[:, 5] = o.GRAVITY
    x0 = ref[0, :4] + rng.uniform(-0.05, 0.05, (B, 4))
    loop = BatchedClosedLoop(model, batch=B, device=0)
    loop.init(torch.tensor(x0.T.copy()), torch.tensor(ref), noise=torch.tensor(rng.normal(0, 0.01, (S, B))), n_steps=S).run()
    r = loop.results(); torch.cuda.synchronize()
    print(model, 'ok', float(r['cost'].sum()), r['status'].max().item())
s = BatchedAcadosOcpSolver('force', batch=5, device=0, N_horizon=50)
s.set(0, 'lbx', np.tile([1.0, 0, 0, 0.6], (5, 1))); s.solve(); print('N50', s.get_stats('status').cpu().numpy())
