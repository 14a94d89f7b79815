"""Timing of the 3-D attitude model's closed loop (bnmpc_step_for_x0 per control step), device-resident: solves/s."""
import sys, os, time, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..', 'tests'))
import numpy as np
import torch
from test_att import att_inputs
from drone_attitude_control_b200.attitude_model import follow_trajectory_batched

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1184
S = int(sys.argv[2]) if len(sys.argv) > 2 else 10
prec = sys.argv[3] if len(sys.argv) > 3 else 'fp64'
refs, x0, pc, pp = att_inputs(min(B, 256), seed=1, rows=S + 30 + 5)
rep = (B + refs.shape[0] - 1) // refs.shape[0]
refs, x0, pc, pp = (np.tile(a, (rep,) + (1,) * (a.ndim - 1))[:B] for a in (refs, x0, pc, pp))
follow_trajectory_batched(refs, x0, 3, p_ctrl=pc, p_plant=pp, log=False, precision=prec)      # warm-up
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
r = follow_trajectory_batched(refs, x0, S, p_ctrl=pc, p_plant=pp, log=True, precision=prec)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
st = r['status'].cpu().numpy(); qi = r['qp_iter'].cpu().numpy(); si = r['sqp_iter'].cpu().numpy()
print(json.dumps(dict(model='att', batch=B, steps=S, precision=prec, ms_per_step=ms / S, solves_per_s=B * S / (ms * 1e-3),
                      nonzero_status=int((st != 0).sum()), qp_iter_mean=float(qi.mean()), sqp_iter_mean=float(si.mean()),
                      qp_iter_first_step=float(qi[:, 0].mean()))))
