#!/bin/bash
cd /root/repo
for w in 3 4 6; do echo "WARPS_PER_SM=$w"; BNMPC_WARPS_PER_SM=$w timeout 300 python tools/att_bench.py 4736 8 2>&1 | tail -1 | cut -c1-120; done
