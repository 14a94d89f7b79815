set -x
cd /root/repo
python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "lockstep or philox or launch_shape or closed_loop_matches" 2>&1 | tail -15
B="python bench.py --skip-e2e --skip-cpu --steps 60 --warmup 5"
echo "=== warp per-step 4096"; BNMPC_LOOP_KERNEL=warp $B 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['qp_iter_mean'])"
echo "=== ls per-step 4096"; $B 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['qp_iter_mean'])"
for c in 0 2 5 10 60; do echo "=== ls spl=60 chunk=$c 4096"; BNMPC_CHUNK=$c $B --steps-per-launch 60 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['qp_iter_mean'])"; done
B2="python bench.py --skip-e2e --skip-cpu --steps 12 --warmup 4 --batch 65536 --ref circle"
echo "=== warp per-step 65536"; BNMPC_LOOP_KERNEL=warp $B2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['qp_iter_mean'])"
echo "=== ls per-step 65536"; $B2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['qp_iter_mean'])"
echo "=== ls spl=12 65536"; $B2 --steps-per-launch 12 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['qp_iter_mean'])"
echo "=== jerk warp 16384"; BNMPC_LOOP_KERNEL=warp python bench.py --skip-e2e --skip-cpu --steps 20 --warmup 4 --batch 16384 --model jerk 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['qp_iter_mean'])"
echo "=== jerk ls spl 16384"; python bench.py --skip-e2e --skip-cpu --steps 20 --warmup 4 --batch 16384 --model jerk --steps-per-launch 20 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['qp_iter_mean'])"
