"""Experiment: the e2e loop of bench.py (set_yref_all + bnmpc_step_for_x0 per control step, pinned host buffers) with the
fleet split into G solver objects on G streams, software-pipelined: a sub-fleet is synchronised only right before its next
step is enqueued, so the tail of one sub-fleet's launch overlaps the others' launches.
usage: python tools/e2e_pipeline.py [B=4096] [steps=50] [G list ...]"""
import os, sys, time, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch
import bench
import drone_attitude_control_b200 as pkg

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
Ke = int(sys.argv[2]) if len(sys.argv) > 2 else 50
Gs = [int(a) for a in sys.argv[3:]] or [1, 2, 4]
W, N = 10, 30
dev = torch.device('cuda', 0)
torch.cuda.set_device(0)
inp = bench.workload(0, B, W + Ke, 500 + N, N)
ref_h = inp['ref'].permute(2, 0, 1).contiguous()          # [B, rows, 8]
ycols = [0, 1, 2, 3, 4, 5]
nx, nu, ny = 4, 2, 6
pin = lambda *sh, dt=torch.float64: torch.empty(sh, dtype=dt).pin_memory()


class Sub:
    def __init__(self, lo, hi):
        self.lo, self.hi, b = lo, hi, hi - lo
        self.stream = torch.cuda.Stream(device=dev)
        self.copy_stream = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(self.stream):
            self.s = pkg.BatchedAcadosOcpSolver('force', batch=b, device=0, N_horizon=N, numpy_io=False)
        self.yh = [torch.cat([ref_h[lo:hi, i:i + N, ycols].reshape(b, N * ny), ref_h[lo:hi, i + N, :nx]], 1).contiguous().pin_memory()
                   for i in range(W + Ke)]
        self.noise = inp['noise'][:W + Ke, lo:hi].contiguous().pin_memory()
        self.u, self.up, self.st = pin(b, nu), pin(b, 2), pin(b, dt=torch.int32)
        self.x0b = [pin(b, nx), pin(b, nx)]
        self.ydev = [torch.empty((b, N * ny + nx), dtype=torch.float64, device=dev) for _ in range(2)]
        self.yev = [torch.cuda.Event(), torch.cuda.Event()]

    def prefetch(self, i):
        pass

    def start(self):
        with torch.cuda.stream(self.stream):
            self.s.reset()
        self.x0b[0][:, :4] = inp['x0'][:, self.lo:self.hi].t()
        self.prefetch(0)

    def enqueue(self, i, last):
        if i == 0 or not os.environ.get('E2E_SKIP_YREF'):
            self.s.set_yref_all(self.yh[i])          # pinned host window -> H2D on the sub-fleet's own stream
        self.s.step_into(self.x0b[i % 2], self.noise[i], self.u, self.up, self.st, self.x0b[(i + 1) % 2], wait=False)


for G in Gs:
    cuts = [B * g // G for g in range(G + 1)]
    subs = [Sub(cuts[g], cuts[g + 1]) for g in range(G)]
    for sfl in subs:
        sfl.start()
    tot = W + Ke
    t0 = None
    for i in range(tot):
        if i == W:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        for sfl in subs:
            if i > 0:
                sfl.s.synchronize()          # step i-1 of THIS sub-fleet is back on the host; the others keep running
            sfl.enqueue(i, i == tot - 1)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    xs = torch.cat([sfl.x0b[tot % 2][:, :4] for sfl in subs], 0)
    bad = sum(int((sfl.st != 0).sum()) for sfl in subs)
    print(json.dumps({'G': G, 'solves_per_s': B * Ke / dt, 'ms_per_fleet_step': dt / Ke * 1e3, 'bad_last': bad,
                      'checksum': float(xs.double().abs().sum()), 'warps': [int(os.environ.get('BNMPC_WARPS_PER_SM', 0))]}))
    del subs
