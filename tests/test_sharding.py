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
