cd /root/repo
timeout 900 python -m pytest tests/test_att.py -x -q -m gpu 2>&1 | tail -15
timeout 300 python tools/att_bench.py 296 8 2>&1 | tail -2
timeout 300 python tools/att_bench.py 1184 8 2>&1 | tail -2
timeout 1200 python -m pytest tests -x -q -m gpu --deselect tests/test_att.py 2>&1 | tail -6
