#!/bin/bash
cd /root/repo
P='import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"]/1e6,4), "M/s multi;", round(d["per_step_launch"]["value"]/1e6,4), "M/s per-step; p50", round(d["p50_step_latency_ms"],4), d["nonzero_status"])'
for w in 16 8; do echo "=== force 4096 WARPS_PER_SM=$w"; BNMPC_WARPS_PER_SM=$w timeout 300 python bench.py --skip-e2e --skip-cpu --skip-extra 2>&1 | tail -1 | python -c "$P"; done
for w in 12 8; do echo "=== jerk 16384 WARPS_PER_SM=$w"; BNMPC_WARPS_PER_SM=$w timeout 300 python bench.py --skip-e2e --skip-cpu --skip-extra --model jerk --batch 16384 --steps 20 2>&1 | tail -1 | python -c "$P"; done
