import sys, time; sys.path.insert(0,'.')
import numpy as np, torch
import drone_attitude_control_b200 as pkg
B,N=4096,30
s=pkg.BatchedAcadosOcpSolver('force',batch=B,device=0,numpy_io=False)
sys.path.insert(0,'tests')
from common import random_solve_inputs
x0,yref=random_solve_inputs(0,256,seed=1)
x0=np.tile(x0,(16,1)); yref=np.tile(yref,(16,1))
yh=torch.tensor(yref).pin_memory(); xh=torch.tensor(x0).pin_memory()
uh=torch.empty((B,2),dtype=torch.float64).pin_memory(); sh=torch.empty(B,dtype=torch.int32).pin_memory()
def sync(): torch.cuda.current_stream().synchronize()
def T(f,n=30):
    for _ in range(3): f(); sync()
    t=time.perf_counter()
    for _ in range(n): f(); sync()
    return (time.perf_counter()-t)/n*1e3
print('set_yref_all host', T(lambda: s.set_yref_all(yh)))
yd=yh.cuda()
print('set_yref_all dev ', T(lambda: s.set_yref_all(yd)))
print('set lbx+ubx host ', T(lambda: (s.set(0,'lbx',xh), s.set(0,'ubx',xh))))
def solve_only():
    s.reset(); 
print('reset            ', T(lambda: s.reset()))
print('reset+solve      ', T(lambda: (s.reset(), s.solve())))
print('get u            ', T(lambda: s.get(0,'u')))
print('get u + d2h      ', T(lambda: uh.copy_(s.get(0,'u'),non_blocking=True)))
print('get_stats        ', T(lambda: s.get_stats('status')))
def full():
    s.set_yref_all(yh); s.set(0,'lbx',xh); s.set(0,'ubx',xh); st=s.solve(); uh.copy_(s.get(0,'u'),non_blocking=True); sh.copy_(st,non_blocking=True)
print('full step (warm iterate)', T(full))
