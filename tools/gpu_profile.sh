#!/bin/bash
# GPU box: ncu captures of the closed-loop kernel for profiles/ (each only after the same command exited 0 without ncu)
cd /root/repo
F="python bench.py --skip-e2e --skip-cpu --skip-extra --steps 20 --warmup 3"
J="python bench.py --skip-e2e --skip-cpu --skip-extra --steps 3 --warmup 3 --model jerk --batch 16384"
$F > gpurun_out/plain_force.log 2>&1 && {
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches.csv $F > gpurun_out/ncu_l.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_loop_step -s 3 -c 1 -o gpurun_out/r02_force_multistep $F > gpurun_out/ncu_m.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:k_loop_step -s 4 -c 1 -o gpurun_out/r02_force_perstep $F > gpurun_out/ncu_p.log 2>&1
}
$J > gpurun_out/plain_jerk.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_loop_step -s 4 -c 1 -o gpurun_out/r02_jerk_perstep $J > gpurun_out/ncu_j.log 2>&1
tail -2 gpurun_out/ncu_m.log gpurun_out/ncu_p.log gpurun_out/ncu_j.log
ls -la gpurun_out/*.ncu-rep
