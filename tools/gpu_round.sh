#!/bin/bash
cd /root/repo
{
echo "# parity soak on the final round-2 build: every instance of each closed-loop batch against the C oracle"
echo "# (1) the one-warp kernel, one launch per step"; timeout 900 python tools/parity_soak.py --seeds 4 2>&1 | grep -v Warning
echo "# (2) the one-warp kernel, all steps in one launch"; timeout 900 python tools/parity_soak.py --seeds 2 --multi 2>&1 | grep -v Warning
echo "# (3) N = 100, four warps per instance (factorisation scan + multi-warp stage scans), all steps in one launch"; timeout 900 python tools/parity_soak.py --seeds 2 --horizon 100 --batch 2048 --steps 20 --multi 2>&1 | grep -v Warning
echo "# (4) N = 50, two warps per instance, one launch per step"; timeout 900 python tools/parity_soak.py --seeds 2 --horizon 50 --batch 2048 --steps 20 2>&1 | grep -v Warning
echo "# (5) N = 30, small batch (512 drones): four warps per instance"; timeout 900 python tools/parity_soak.py --seeds 2 --batch 512 --steps 40 2>&1 | grep -v Warning
} > gpurun_out/r02_parity_soak.txt 2>&1
grep -E "^#|TOTAL" gpurun_out/r02_parity_soak.txt
