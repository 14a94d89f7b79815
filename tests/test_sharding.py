"""Multi-GPU host logic on CPU: world_size-2 gloo processes own slices of the batch; inputs keyed by global instance id
make the union of the shards identical to the single-rank batch, and the only collective is the final metric reduce.
The per-instance work is done by the host emulation of the product kernels (tests/hostsim, test harness only)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, B, S, out_dir):
    sys.path.insert(0, ROOT); sys.path.insert(0, HERE)
    os.environ['MASTER_ADDR'] = '127.0.0.1'; os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from drone_attitude_control_b200 import sharding
    from drone_attitude_control_b200.generate_trajectory import gen_circle_traj_batched
    import hostsim as hs
    from oracle import c_oracle as co
    lo, hi = sharding.shard_range(B, rank, world)
    inp = sharding.instance_inputs(lo, hi, S, mass_sigma=0.05)
    ref = gen_circle_traj_batched(500, 30, inp['radius'], inp['center'], inp['phase'])          # [rows, 8, b]
    x0 = (ref[0, :4, :] + inp['dx0']).numpy().T.copy()
    refs = ref.permute(2, 0, 1).contiguous().numpy()
    b = hi - lo
    pc = np.repeat(np.array([[0.03277, 9.81]]), b, 0)
    pp = pc.copy(); pp[:, 0] *= inp['mass_scale'].numpy()
    oo = co.default_opts(co.MODEL_FORCE)
    r = hs.closed_loop(hs.MODEL_FORCE, hs.FP64, hs.opts_from_oracle(oo), refs, x0, inp['noise'].numpy(), pc, pp, S)
    tot = sharding.reduce_metrics([r['cost'].sum(), r['qp_iter'].sum(), float(b)])
    mx = sharding.reduce_metrics([r['cost'].max()], op='max')
    np.savez(os.path.join(out_dir, f'rank{rank}of{world}.npz'), lo=lo, hi=hi, Xsim=r['Xsim'], cost=r['cost'], tot=tot.numpy(), mx=mx.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_shard_ranges_cover_the_batch():
    from drone_attitude_control_b200.sharding import shard_range
    for B in (1, 7, 64, 4096, 262144):
        for G in (1, 2, 3, 4, 8):
            edges = [shard_range(B, r, G) for r in range(G)]
            assert edges[0][0] == 0 and edges[-1][1] == B
            assert all(edges[i][1] == edges[i + 1][0] for i in range(G - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


def test_inputs_do_not_depend_on_the_sharding():
    from drone_attitude_control_b200.sharding import instance_inputs, shard_range
    whole = instance_inputs(0, 12, 5, mass_sigma=0.05)
    for G in (2, 3):
        parts = [instance_inputs(*shard_range(12, r, G), 5, mass_sigma=0.05) for r in range(G)]
        for k, v in whole.items():
            cat = torch.cat([p[k] for p in parts], dim=-1 if v.dim() > 1 and k in ('dx0', 'noise') else 0)
            assert torch.equal(cat, v), k


def test_two_rank_gloo_run_matches_single_rank(tmp_path):
    B, S = 6, 4
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(1, port, B, S, str(tmp_path)), nprocs=1, join=True)
    mp.spawn(_worker, args=(2, port + 1, B, S, str(tmp_path)), nprocs=2, join=True)
    one = np.load(tmp_path / 'rank0of1.npz')
    two = [np.load(tmp_path / f'rank{r}of2.npz') for r in range(2)]
    assert (int(two[0]['lo']), int(two[0]['hi']), int(two[1]['lo']), int(two[1]['hi'])) == (0, 3, 3, 6)
    assert np.array_equal(np.concatenate([t['Xsim'] for t in two]), one['Xsim'])        # bit-identical per-instance results
    assert np.array_equal(np.concatenate([t['cost'] for t in two]), one['cost'])
    for t in two:                                                                       # every rank holds the reduced totals
        np.testing.assert_allclose(t['tot'], one['tot'], rtol=1e-13)
        assert t['tot'][2] == B
        assert t['mx'][0] == one['mx'][0]


def test_trajectory_families_shapes_and_formulas():
    """Batched generators of the reference's other trajectory families (src/jerk_model/gen_trajectory.py:8-71) against a
    direct evaluation of the reference's formulas at the MPC rate."""
    import numpy as np
    import torch
    from drone_attitude_control_b200 import generate_trajectory as gt
    T, dt, N, NH, g = 10.0, 0.02, 500, 30, 9.81
    initial = np.array([[0.1, -0.2], [0.0, 0.3]]); length = np.array([0.5, 1.0])
    st = gt.gen_static_point_traj_batched(N, NH, initial).numpy()
    assert st.shape == (2, N + NH, 8) and np.all(st[:, :, 0] == initial[:, None, 0]) and np.all(st[:, :, 5] == g) and np.all(st[:, :, 2:5] == 0)
    sl = gt.gen_straight_traj_batched(N, NH, initial, length).numpy()
    i = np.arange(N + NH)
    for b in range(2):
        jerk = 6 * length[b] / T ** 3
        np.testing.assert_allclose(sl[b, :, 0], initial[b, 0] + jerk * (i * dt) ** 3 / 6, rtol=1e-14, atol=1e-15)
        np.testing.assert_allclose(sl[b, :, 3], 0.5 * jerk * (i * dt) ** 2, rtol=1e-14, atol=1e-15)
        np.testing.assert_allclose(sl[b, :, 5], jerk * i * dt + g, rtol=1e-14)
    raw = gt.gen_straight_traj_batched(N, NH, initial, length, fill_acc=False).numpy()
    assert np.all(raw[:, :, 4] == 0) and np.all(raw[:, :, 5] == g) and np.all(raw[:, :, 6:] == 0)
    sq = gt.gen_square_traj_batched(N, NH, initial, length).numpy()
    side = N // 4
    for b in range(2):
        L = length[b]
        np.testing.assert_allclose(sq[b, side - 1, :2], initial[b] + [L * (side - 1) / side, 0], atol=1e-15)
        np.testing.assert_allclose(sq[b, 2 * side, :2], initial[b] + [L, L], atol=1e-15)
        np.testing.assert_allclose(sq[b, 3 * side + 1, :2], initial[b] + [0, L - L / side], atol=1e-15)
        np.testing.assert_allclose(sq[b, N:], sq[b, :NH], atol=0)


def test_fleet_group_bounds():
    """SolverFleet's slices: contiguous, cover the batch, sizes differ by at most one, never more groups than drones."""
    from drone_attitude_control_b200.fleet import group_bounds
    assert group_bounds(4096, 4) == [0, 1024, 2048, 3072, 4096]
    assert group_bounds(10, 3) == [0, 4, 7, 10]
    assert group_bounds(2, 5) == [0, 1, 2]
    for B, G in [(1, 1), (7, 7), (4097, 4), (50, 3)]:
        b = group_bounds(B, G)
        sizes = np.diff(b)
        assert b[0] == 0 and b[-1] == B and sizes.min() >= 1 and sizes.max() - sizes.min() <= 1
    with pytest.raises(ValueError):
        group_bounds(0, 2)


def test_fleet_pipeline_order_on_cpu():
    """SolverFleet's host logic with recording stand-ins for the solver objects: every sub-fleet gets its own rows of every
    buffer, is synchronised exactly once per step and only right before its own next step is enqueued (the others are left
    running), and the results hook sees each (sub-fleet, step) once, in step order, before that sub-fleet's next step."""
    from drone_attitude_control_b200.fleet import SolverFleet
    log = []

    class Rec:
        nx, nu, ny, ny_e, N = 4, 2, 6, 4, 30

        def __init__(self, b):
            self.b, self.id = b, len(made)
            made.append(self)

        def set_yref_all(self, y):
            log.append(('yref', self.id, tuple(y.shape), float(y[0, 0])))

        def step_into(self, x0, eps, u0, up, st, xn, p_plant_host=None, wait=True):
            assert not wait and x0.shape[0] == self.b and xn.shape[0] == self.b and (eps is None or eps.shape[0] == self.b)
            log.append(('step', self.id, float(x0[0, 0])))
            xn.copy_(x0 + 1.0)              # "plant": next state = state + 1
            st.fill_(self.id)

        def synchronize(self):
            log.append(('sync', self.id))

        def reset(self):
            log.append(('reset', self.id))

    made = []
    B, G, S = 10, 3, 4
    fleet = SolverFleet(batch=B, groups=G, solver_factory=Rec)
    assert fleet.slices() == [(0, 4), (4, 7), (7, 10)] and [r.b for r in made] == [4, 3, 3]
    xb = [torch.zeros(B, 4, dtype=torch.float64), torch.zeros(B, 4, dtype=torch.float64)]
    xb[0][:, 0] = torch.arange(B, dtype=torch.float64) * 100
    u0, up, st = torch.zeros(B, 2), torch.zeros(B, 2), torch.zeros(B, dtype=torch.int32)
    seen = []
    for i in range(S):
        y = torch.full((B, 30 * 6 + 4), float(i), dtype=torch.float64)
        fleet.step(y, xb[i % 2], None, u0, up, st, xb[(i + 1) % 2], on_results=lambda g, lo, hi: seen.append((g, lo, hi, len([e for e in log if e[0] == 'step' and e[1] == g]))))
    fleet.synchronize(on_results=lambda g, lo, hi: seen.append((g, lo, hi, S)))
    # each sub-fleet: yref, step per control step; a sync before every step but the first, and one at the end
    for g in range(G):
        mine = [e[0] for e in log if e[1] == g]
        assert mine == ['yref', 'step'] + ['sync', 'yref', 'step'] * (S - 1) + ['sync']
    # round-robin: the sync of sub-fleet g at step i comes after the enqueue of sub-fleet g-1 at step i (the others keep running)
    order = [(e[0], e[1]) for e in log if e[0] in ('sync', 'step')]
    assert order[:G] == [('step', 0), ('step', 1), ('step', 2)] and order[G:G + 4] == [('sync', 0), ('step', 0), ('sync', 1), ('step', 1)]
    # the hook saw every (sub-fleet, step) once, after exactly that many steps of the sub-fleet had been enqueued
    assert sorted(seen) == sorted((g, lo, hi, k) for g, (lo, hi) in enumerate(fleet.slices()) for k in range(1, S + 1))
    # the rows travelled through the right slices: S plant steps of +1 on every drone, statuses = sub-fleet id
    assert torch.equal(xb[S % 2][:, 0], torch.arange(B, dtype=torch.float64) * 100 + S)
    assert st.tolist() == [0] * 4 + [1] * 3 + [2] * 3
    fleet.reset()
    assert [e for e in log[-G:]] == [('reset', 0), ('reset', 1), ('reset', 2)]
