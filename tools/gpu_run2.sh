cd /root/repo
for g in 0 1; do echo "=== generation $g"; python tools/ls_profile.py --generation $g; done
echo "=== gen1 bench spl"; BNMPC_LS_GENERATION=1 python bench.py --skip-e2e --skip-cpu --steps 60 --warmup 5 --steps-per-launch 60 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['qp_iter_mean'])"
echo "=== gen1 circle 65536"; BNMPC_LS_GENERATION=1 python bench.py --skip-e2e --skip-cpu --steps 12 --warmup 4 --batch 65536 --ref circle --steps-per-launch 12 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['qp_iter_mean'])"
