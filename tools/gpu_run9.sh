cd /root/repo
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
P='import sys,json; d=json.loads(sys.stdin.read()); print(round(d["value"]/1e6,3), "M/s multi;", round(d["per_step_launch"]["value"]/1e6,3), "M/s per-step; p50", round(d["p50_step_latency_ms"],4), d["nonzero_status"])'
for N in 50 100; do for g in 1 4; do echo "=== N=$N wpg<=$g"; BNMPC_WARPS_PER_INSTANCE=$g timeout 300 python bench.py --skip-e2e --skip-cpu --skip-extra --steps 6 --warmup 3 --batch 65536 --ref circle --horizon $N 2>&1 | tail -1 | python -c "$P"; done; done
echo "=== B=1 latency N=30"; for g in 1 4; do BNMPC_WARPS_PER_INSTANCE=$g timeout 300 python bench.py --skip-e2e --skip-cpu --skip-extra --steps 50 --warmup 5 --batch 1 2>&1 | tail -1 | python -c "$P"; done
