#!/bin/bash
# GPU box: what the driver runs at round end - the -m gpu suite, smoke(), both bench arms
cd /root/repo
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py --impl reference > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | grep real
( time python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err ) 2>&1 | grep real
tail -2 gpurun_out/bench.err
python - <<'PY'
import json
r = json.loads(open('gpurun_out/bench_ref.json').read().strip().splitlines()[-1])
d = json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
e = d['e2e']
print('reference arm %.3f M/s on %d cores | same config: %s' % (r['value'] / 1e6, r['cpu_baseline']['cores'], r['config'] == d['config']))
print('value %.3f M/s | per-step %.3f (p50 %.4f ms) | e2e fleet(G=%d) %.3f M/s | single solver %.3f | ref call seq %.3f | cpu %.3f M/s | frac %.4f | launches %s | clocks %s' % (
    d['value'] / 1e6, d['per_step_launch']['value'] / 1e6, d['p50_step_latency_ms'], e['groups'], e['value'] / 1e6, e['single_solver']['value'] / 1e6,
    e['reference_call_sequence']['value'] / 1e6, d['cpu_baseline']['value'] / 1e6, d['roofline']['frac'], d['gpu_launches'], d['clocks']))
for k, v in d['extra'].items():
    if isinstance(v, dict):
        print('  %-46s %7.3f M/s  failed %d  frac %.4f  %.1f ms' % (k, v['value'] / 1e6, v['failed_steps'], v['roofline_frac'], v['ms']))
PY
