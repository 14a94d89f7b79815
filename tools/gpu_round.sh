#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "fleet or one_call" 2>&1 | tail -4
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
e = d['e2e']
print('value %.3f M/s | per-step %.3f | e2e fleet(G=%d) %.3f M/s | single solver %.3f | ref call seq %.3f | diff %.1e | cpu %.3f M/s' % (
    d['value'] / 1e6, d['per_step_launch']['value'] / 1e6, e['groups'], e['value'] / 1e6, e['single_solver']['value'] / 1e6,
    e['reference_call_sequence']['value'] / 1e6, e['max_abs_state_difference_to_single_solver'], d['cpu_baseline']['value'] / 1e6))
PY
for g in 2 8; do python bench.py --skip-cpu --skip-extra --e2e-groups $g 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('G', d['e2e']['groups'], d['e2e']['value']/1e6)"; done
python bench.py --skip-cpu --skip-extra --model jerk --batch 16384 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('jerk G', e['groups'], e['value']/1e6, 'single', e['single_solver']['value']/1e6, e['max_abs_state_difference_to_single_solver'])"
