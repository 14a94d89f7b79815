#!/bin/bash
cd /root/repo
python bench.py --skip-cpu --skip-e2e > gpurun_out/bench_q.json 2> gpurun_out/bench_q.err; tail -3 gpurun_out/bench_q.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_q.json').read().strip().splitlines()[-1])
for k, v in d['extra'].items():
    if isinstance(v, dict):
        print('  %-46s %7.3f M/s  failed %d  frac %.4f  %.1f ms %s' % (k, v['value'] / 1e6, v['failed_steps'], v['roofline_frac'], v['ms'], v.get('mean_position_error_m', '')), v.get('qp_iter_mean', ''))
PY
F="python bench.py --skip-e2e --skip-cpu --skip-extra --steps 4 --warmup 3 --batch 65536 --ref circle --horizon 100"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_loop_step -s 3 -c 1 -o gpurun_out/r02_force_N100_v10 $F > gpurun_out/ncu_n100.log 2>&1
tail -1 gpurun_out/ncu_n100.log
