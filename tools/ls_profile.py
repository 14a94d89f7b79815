"""Slot timing of the lockstep kernel (debug aid): CTA 0 records clock64 at every barrier of its schedule
(bnmpc_debug_profile); this prints where warp 0 of that CTA spends the half-rounds.  Usage (GPU box):
    python tools/ls_profile.py [--model force] [--batch 4096] [--steps 20] [--generation 0|1]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--model', default='force')
    ap.add_argument('--batch', type=int, default=4096)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--generation', type=int, default=0)
    ap.add_argument('--ref', default='table')
    args = ap.parse_args()
    os.environ['BNMPC_LS_GENERATION'] = str(args.generation)
    import drone_attitude_control_b200 as pkg
    from drone_attitude_control_b200 import _lib
    from bench import make_inputs
    dev = torch.device('cuda', 0)
    B, S = args.batch, args.steps
    ref, x0, noise, inp = make_inputs(0, B, 2 * S, 600, device=dev)
    loop = pkg.BatchedClosedLoop(args.model, batch=B, device=0)
    ref_arg = pkg.CircleRef(inp['radius'], inp['center'], inp['phase'], n=570) if args.ref == 'circle' else ref.permute(2, 0, 1).contiguous()
    loop.init(x0, ref_arg, noise=noise, n_steps=2 * S, log=False)
    loop.run(S, steps_per_launch=S)          # warm-up launch
    cap = 1 << 20
    buf = torch.zeros(cap, dtype=torch.int64, device=dev)
    _lib.check(_lib.lib().bnmpc_debug_profile(loop.solver.handle, C.c_void_p(buf.data_ptr()), cap))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); loop.run(S, steps_per_launch=S); e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    print(f'{B} instances x {S} steps in {ms:.3f} ms = {B * S / ms / 1e3:.2f} M solves/s')
    h = buf.cpu().numpy()
    n = int(h[0])
    t, tagw = h[1:n:2], h[2:n:2]
    tag = (tagw >> 56) & 0xff
    fac = ((tagw >> 24) & 0xffffffff)
    ipm = tagw & 0xffffff
    names = {(0, 1): 'P1 own', (1, 2): 'B1 wait', (2, 3): 'S factor', (3, 4): 'B2 wait', (2, 4): 'no factor', (4, 5): 'P2 own', (5, 6): 'B3 wait',
             (6, 7): 'S back', (7, 8): 'B4 wait', (6, 8): 'no back', (8, 9): 'P3 own', (9, 10): 'B5 wait', (10, 11): 'S fwd', (11, 12): 'B6 wait',
             (10, 12): 'no fwd', (12, 0): 'P4 own'}
    acc = {}
    for i in range(len(t) - 1):
        k = names.get((int(tag[i]), int(tag[i + 1])), f'{tag[i]}->{tag[i + 1]}')
        acc.setdefault(k, []).append(int(t[i + 1] - t[i]))
    tot = int(t[-1] - t[0])
    halves = int((tag == 0).sum())
    print(f'CTA 0: {halves} half-rounds, {tot} cycles, {tot / max(halves, 1):.0f} cycles per half-round')
    for k, v in acc.items():
        v = np.array(v)
        print(f'  {k:10s} n {len(v):6d}  mean {v.mean():8.0f}  p50 {np.median(v):8.0f}  p90 {np.percentile(v, 90):8.0f}  max {v.max():8d}  share {v.sum() / tot * 100:5.1f} %')
    m2 = tag == 2
    nf = np.array([bin(int(x)).count('1') for x in fac[m2]]); ni = np.array([bin(int(x)).count('1') for x in ipm[m2]])
    print(f'  warps in the IPM loop per half-round: mean {ni.mean():.2f}; in a factor sweep (when there is one): mean {nf[nf > 0].mean():.2f}, '
          f'half-rounds with a factor sweep {np.mean(nf > 0) * 100:.0f} %')


if __name__ == '__main__':
    main()
