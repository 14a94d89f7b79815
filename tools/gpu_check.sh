#!/bin/bash
# GPU box: the -m gpu suite, then the default bench line (what the driver runs at round end)
cd /root/repo
python -m pytest tests -x -q -m gpu 2>&1 | tail -8
python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1])
print('value %.3f M/s | per-step launches %.3f M/s (p50 %.4f ms) | e2e %.3f M/s | cpu %.3f M/s on %d cores | roofline frac %.4f | qp_iter %.4f / cpu %.4f' % (
    d['value'] / 1e6, d['per_step_launch']['value'] / 1e6, d['p50_step_latency_ms'], d['e2e']['value'] / 1e6, d['cpu_baseline']['value'] / 1e6,
    d['cpu_baseline']['cores'], d['roofline']['frac'], d['qp_iter_mean'], d['cpu_baseline']['qp_iter_mean']))
for k, v in d['extra'].items():
    if isinstance(v, dict):
        print('  %-46s %7.3f M/s  failed %d  frac %.4f  %.1f ms' % (k, v['value'] / 1e6, v['failed_steps'], v['roofline_frac'], v['ms']))
PY
