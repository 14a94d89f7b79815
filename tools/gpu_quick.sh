#!/bin/bash
# GPU box: a few parity tests, then the default bench line without the CPU arm
cd /root/repo
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "${KEXPR:-matches_oracle or launch_shape}" 2>&1 | tail -4
python bench.py --skip-cpu > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; tail -3 gpurun_out/bench_q.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_q.json').read().strip().splitlines()[-1])
print('value %.3f M/s | per-step launches %.3f M/s (p50 %.4f ms) | e2e %.3f M/s | roofline frac %.4f | qp_iter %.4f' % (
    d['value'] / 1e6, d['per_step_launch']['value'] / 1e6, d['p50_step_latency_ms'], d['e2e']['value'] / 1e6, d['roofline']['frac'], d['qp_iter_mean']))
for k, v in d['extra'].items():
    if isinstance(v, dict):
        print('  %-46s %7.3f M/s  failed %d  frac %.4f  %.1f ms' % (k, v['value'] / 1e6, v['failed_steps'], v['roofline_frac'], v['ms']))
PY
P='import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"]/1e6,4), "M/s multi;", round(d["per_step_launch"]["value"]/1e6,4), "M/s per-step; p50", round(d["p50_step_latency_ms"],4), d["nonzero_status"])'
for B in 1 256 1024; do echo "=== B=$B N=30"; timeout 300 python bench.py --skip-e2e --skip-cpu --skip-extra --steps 50 --warmup 5 --batch $B 2>&1 | tail -1 | python -c "$P"; done
