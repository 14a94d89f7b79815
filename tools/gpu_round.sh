#!/bin/bash
cd /root/repo
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
echo "=== e2e pipeline"; timeout 600 python tools/e2e_pipeline.py 4096 50 1 2 3 4 6 8 2>&1 | tail -8
echo "=== att"; timeout 300 python tools/att_bench.py 296 8 2>&1 | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_solve -s 4 -c 1 -o gpurun_out/r02_att_solve python tools/att_bench.py 296 3 > gpurun_out/ncu_att.log 2>&1
tail -2 gpurun_out/ncu_att.log
