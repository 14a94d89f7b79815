"""bench.py contract checks that need no GPU: the reference arm prints one JSON line with the agreed keys."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '4', '--warmup', '3',
                          '--cpu-instances', '64'], capture_output=True, text=True, check=True, cwd=ROOT).stdout
    lines = [l for l in out.splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'nmpc_solves_per_sec' and d['unit'] == 'solves/s'
    assert d['higher_is_better'] is True and d['n_gpus'] == 1 and d['steps'] == 4 and d['warmup'] == 3
    assert d['value'] > 0 and d['e2e']['value'] == d['value'] and d['e2e']['h2d_bytes_per_step'] == 0
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and 'sample' in cb
    assert d['config']['controller'] == 'force' and d['config']['horizon'] == 30 and d['vs_baseline'] is None
    assert 'workload' in d['config'] and 'model' not in d['config']


def test_flop_and_byte_model_matches_design():
    sys.path.insert(0, ROOT)
    import bench
    # per IPM iteration: 13.6 kflop force, 27.4 kflop jerk (block-structured count)
    f_it_force = bench.flops_per_solve(2, 2, 1, 30, 4, 1.0, 0.0)
    f_it_jerk = bench.flops_per_solve(2, 3, 1, 30, 1, 1.0, 0.0)
    assert abs(f_it_force - 13620) < 1 and abs(f_it_jerk - 27440) < 1
    assert bench.bytes_per_solve(4, 2, 30) == 1616 and bench.bytes_per_solve(6, 2, 30) == 2128
