"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by solver pass (line ranges of bnmpc_core.cuh).
Usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > src.csv; python tools/ncu_source_regions.py src.csv"""
import bisect, collections, csv, re, sys

def regions_of(path):
    pat = re.compile(r'^\s*(?:static\s+)?BN_HD\s+[\w:<>,\s\*&]+?\s+(\w+)\s*\(')
    out = [(1, 'prelude')]
    for i, line in enumerate(open(path), 1):
        m = pat.match(line)
        if m and line.startswith('    ') and not line.startswith('        '): out.append((i, m.group(1)))
        elif m and not line.startswith(' '): out.append((i, m.group(1)))
    return out

def main(src_csv, core='drone_attitude_control_b200/csrc/bnmpc_core.cuh'):
    regions = regions_of(core); starts = [r[0] for r in regions]
    rows = csv.reader(open(src_csv)); hdr = None; cur = None; idx = None
    S = collections.Counter(); I = collections.Counter(); TI = collections.Counter(); ST = collections.defaultdict(collections.Counter)
    for r in rows:
        if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
        if r and r[0] == 'Line No':
            hdr = r; idx = {h: i for i, h in reversed(list(enumerate(hdr)))}; continue
        if hdr is None or not r or not r[0].strip().isdigit() or len(r) != len(hdr): continue
        try:
            ln = int(r[0]); s = int(r[idx['# Samples']] or 0); ie = int(r[idx['Instructions Executed']] or 0)
            ti = int(r[idx['Thread Instructions Executed']] or 0)
        except ValueError: continue
        name = regions[bisect.bisect_right(starts, ln) - 1][1] if cur == core.split('/')[-1] else cur
        S[name] += s; I[name] += ie; TI[name] += ti
        for c, i in idx.items():
            if c.startswith('stall_') and 'Not Issued' not in c and r[i]:
                try: ST[name][c] += int(r[i])
                except ValueError: pass
    tot = sum(S.values()); ti = sum(I.values())
    print('total samples', tot, 'warp instructions', ti)
    print(f"{'region':22s} {'samp%':>6} {'instr%':>6} {'relCPI':>6} {'thr/inst':>8}  top stalls (% of region samples)")
    for k, v in S.most_common():
        if v < tot * 0.002: continue
        top = ', '.join(f"{c[6:]}:{100 * n / max(v, 1):.0f}" for c, n in ST[k].most_common(5))
        print(f"{k:22s} {100 * v / tot:6.1f} {100 * I[k] / ti:6.1f} {(v / tot) / (I[k] / ti + 1e-12):6.2f} {TI[k] / max(I[k], 1):8.1f}  {top}")

if __name__ == '__main__': main(*sys.argv[1:])
