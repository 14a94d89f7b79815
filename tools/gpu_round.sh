#!/bin/bash
cd /root/repo
timeout 1200 python -m pytest tests -q -m gpu 2>&1 | tail -12
P='import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"]/1e6,4), "M/s multi;", round(d["per_step_launch"]["value"]/1e6,4), "M/s per-step; p50", round(d["p50_step_latency_ms"],4), d["nonzero_status"])'
for N in 50 100; do echo "=== N=$N 65536"; timeout 300 python bench.py --skip-e2e --skip-cpu --skip-extra --steps 4 --warmup 3 --batch 65536 --ref circle --horizon $N 2>&1 | tail -1 | python -c "$P"; done
for B in 1 256 1024; do echo "=== B=$B N=30"; timeout 300 python bench.py --skip-e2e --skip-cpu --skip-extra --steps 50 --warmup 5 --batch $B 2>&1 | tail -1 | python -c "$P"; done
echo "=== jerk N=100 16384"; timeout 300 python bench.py --skip-e2e --skip-cpu --skip-extra --steps 4 --warmup 3 --batch 16384 --model jerk --horizon 100 2>&1 | tail -1 | python -c "$P"
