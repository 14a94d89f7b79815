#!/usr/bin/env python3
"""Summarise an ncu report of bench.py into profiles/: the raw page as CSV, a small JSON of the metrics quoted in
DESIGN.md / profiles/README.md, and the entry of profiles/traffic.json that bench.py reports as roofline.traffic.

Usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep r01_v3 force_fp64_B4096"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = {
    'gpu__time_duration.sum': 'duration',
    'dram__bytes_read.sum': 'dram_read',
    'dram__bytes_write.sum': 'dram_write',
    'smsp__inst_executed.sum': 'warp_instructions',
    'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active': 'fp64_pipe_pct',
    'smsp__issue_active.avg.pct_of_peak_sustained_active': 'issue_active_pct',
    'sm__warps_active.avg.pct_of_peak_sustained_active': 'warps_active_pct',
    'launch__registers_per_thread': 'registers_per_thread',
    'launch__grid_size': 'grid_size',
    'smsp__thread_inst_executed_per_inst_executed.ratio': 'active_threads_per_instruction',
    'smsp__average_warps_issue_stalled_wait_per_issue_active.ratio': 'stall_wait',
    'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio': 'stall_long_scoreboard',
    'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio': 'stall_short_scoreboard',
    'smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio': 'stall_no_instruction',
    'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio': 'stall_barrier',
    'smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio': 'stall_branch_resolving',
    'smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio': 'stall_math_pipe_throttle',
}
UNIT = {'Mbyte': 1e6, 'Kbyte': 1e3, 'Gbyte': 1e9, 'byte': 1.0, 'ms': 1e-3, 'us': 1e-6, 'ns': 1e-9, 's': 1.0}


def main():
    rep, tag, key = sys.argv[1], sys.argv[2], sys.argv[3]
    raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True, check=True).stdout
    os.makedirs(os.path.join(ROOT, 'profiles'), exist_ok=True)
    open(os.path.join(ROOT, 'profiles', f'{tag}_k_loop_step_full_raw.csv'), 'w').write(raw)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out = {'kernel': data[0][hdr.index('Kernel Name')] if 'Kernel Name' in hdr else None, 'launches_profiled': len(data)}
    for name, short in WANT.items():
        if name in hdr:
            i = hdr.index(name)
            vals = [float(r[i].replace(',', '')) * UNIT.get(units[i], 1.0) for r in data]
            out[short] = sum(vals) / len(vals)
    out['traffic_bytes'] = out.get('dram_read', 0.0) + out.get('dram_write', 0.0)
    out['source'] = f'profiles/{tag}_k_loop_step_full_raw.csv (ncu --set full --clock-control none, per launch)'
    json.dump(out, open(os.path.join(ROOT, 'profiles', f'{tag}_summary.json'), 'w'), indent=1)
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    t = json.load(open(tpath)) if os.path.exists(tpath) else {}
    t[key] = {k: out[k] for k in ('traffic_bytes', 'fp64_pipe_pct', 'issue_active_pct', 'warps_active_pct', 'duration', 'source') if k in out}
    json.dump(t, open(tpath, 'w'), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == '__main__':
    main()
