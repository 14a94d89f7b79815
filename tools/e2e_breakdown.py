"""Where the time of an end-to-end step goes (bench.py's `e2e`): CUDA-event times of the pieces of
set_yref_all + solve_for_x0 on 4096 force-model instances, host-timed totals beside them.
Usage (GPU box): python tools/e2e_breakdown.py"""
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
import drone_attitude_control_b200 as pkg
from common import random_loop_inputs

B, N, S = 4096, 30, 40
refs, x0, noise, pc, pp = random_loop_inputs(256, S, seed=3)
rep = B // 256
refs = np.tile(refs, (rep, 1, 1)); x0 = np.tile(x0, (rep, 1)); noise = np.tile(noise, (1, rep)); pc = np.tile(pc, (rep, 1))
loop = pkg.BatchedClosedLoop('force', batch=B, device=0)
loop.init(torch.tensor(x0.T.copy()), torch.tensor(refs), noise=torch.tensor(noise), n_steps=S).run()
xs = loop.results()['Xsim'].cpu()
s = pkg.BatchedAcadosOcpSolver('force', batch=B, device=0, numpy_io=False)
ref_h = torch.tensor(refs)
yh = [torch.cat([ref_h[:, i:i + N, :6].reshape(B, N * 6), ref_h[:, i + N, :4]], 1).contiguous().pin_memory() for i in range(S)]
xh = [xs[:, i, :].contiguous().pin_memory() for i in range(S)]
xd = [t.cuda() for t in xh]; yd = [t.cuda() for t in yh]
uh = torch.empty((B, 2), dtype=torch.float64).pin_memory(); sh = torch.empty(B, dtype=torch.int32).pin_memory()
ud = torch.empty((B, 2), dtype=torch.float64, device='cuda'); sd = torch.empty(B, dtype=torch.int32, device='cuda')
cur = torch.cuda.current_stream()

def ev_time(fn, steps=range(10, S)):
    s.reset()
    for i in range(10): fn(i)
    cur.synchronize()
    tot = 0.0; t0 = time.perf_counter()
    for i in steps:
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(cur); fn(i); b.record(cur); cur.synchronize(); tot += a.elapsed_time(b)
    return tot / len(steps), (time.perf_counter() - t0) / len(steps) * 1e3

def dev_only(i): s.set_yref_all(yd[i]); s.solve_for_x0_device(xd[i], ud, sd)
def host_all(i): s.set_yref_all(yh[i]); s.solve_for_x0_into(xh[i], uh, sh)
def host_x0_dev_yref(i): s.set_yref_all(yd[i]); s.solve_for_x0_into(xh[i], uh, sh)
def yref_only_host(i): s.set_yref_all(yh[i])
def yref_only_dev(i): s.set_yref_all(yd[i])
for name, fn in (('yref D2D only', yref_only_dev), ('yref H2D only', yref_only_host), ('device buffers: yref D2D + x0 D2D + solve + u0/status D2D', dev_only),
                 ('yref D2D + host x0/u0/status', host_x0_dev_yref), ('all host (serial e2e step)', host_all)):
    g, h = ev_time(fn)
    print(f'{name:62s} gpu {g:7.3f} ms   host wall {h:7.3f} ms')
