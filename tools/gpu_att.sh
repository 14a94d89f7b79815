#!/bin/bash
# GPU box: the 3-D attitude model - parity tests, then closed-loop timing at one and four waves of instances
cd /root/repo
timeout 900 python -m pytest tests/test_att.py -x -q -m gpu 2>&1 | tail -15
timeout 300 python tools/att_bench.py 296 8 2>&1 | tail -1
timeout 300 python tools/att_bench.py 1184 8 2>&1 | tail -1
timeout 300 python tools/att_bench.py 4736 8 2>&1 | tail -1
if [ -n "$ATT_NCU" ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_solve -s 4 -c 1 -o gpurun_out/r02_att_solve python tools/att_bench.py 296 3 > gpurun_out/ncu_att.log 2>&1
  tail -2 gpurun_out/ncu_att.log
fi
