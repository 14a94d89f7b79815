#!/bin/bash
cd /root/repo
{
echo "# (6) stress: x0 up to 0.4 off the reference (inputs and states run into their bounds), force model, N = 100 four warps / N = 50 two warps / N = 30 one warp"
timeout 900 python tools/parity_soak.py --models force --seeds 2 --horizon 100 --batch 2048 --steps 20 --multi --spread 0.4 2>&1 | grep -v Warning
timeout 900 python tools/parity_soak.py --models force --seeds 2 --horizon 50 --batch 2048 --steps 20 --spread 0.4 2>&1 | grep -v Warning
timeout 900 python tools/parity_soak.py --models force --seeds 2 --batch 4096 --steps 20 --spread 0.4 2>&1 | grep -v Warning
} > gpurun_out/r02_parity_soak_stress.txt 2>&1
cat gpurun_out/r02_parity_soak_stress.txt | cut -c1-260
