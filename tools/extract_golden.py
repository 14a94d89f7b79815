#!/usr/bin/env python3
"""Recover the closed-loop series of the reference's one committed run from its vector plots.

The reference ships no numeric fixtures; the only artefacts of a real acados run are
  /root/reference/experiment_data/img/example_acc_trajectory_component.pdf   (force model)
  /root/reference/experiment_data/img/example_jerk_trajectory_component.pdf  (jerk model)
produced by create_componentwise (reference src/store_results.py:140-212) from main.py as committed
(seed 42, noise on; reference src/main.py:43-46).  Every plotted line is a polyline in PDF points; the
abscissa is arange(0, T, dt) (store_results.py:153), so vertex k <-> control step k.  Matplotlib's path
simplification drops near-collinear vertices, so only a subset of the 500 steps survives per series.

This script decodes the polylines, calibrates each axis from its own tick gridlines + tick labels, and
writes tests/golden/acados_{force,jerk}.npz with, per series, the surviving step indices and values.
It needs /root/reference and therefore runs only in the build container; the .npz files are committed.

Usage: python tools/extract_golden.py [--ref /root/reference] [--out tests/golden]
"""
import argparse
import os
import re
import zlib

import numpy as np

T_END, DT, NSTEP = 10.0, 0.02, 500


def page_stream(pdf_path):
    raw = open(pdf_path, 'rb').read()
    streams = []
    for m in re.finditer(rb'stream\r?\n(.*?)endstream', raw, re.S):
        try:
            streams.append(zlib.decompress(m.group(1)))
        except Exception:
            pass
    return max(streams, key=len).decode('latin1')


def parse_axes(page):
    """Return list of axes: dict(clip=(x,y,w,h), yticks=[(ypos, value)], lines=[(colour, dashed, pts)])."""
    axes = {}
    blocks = re.split(r'\bQ\b', page)
    for bi, block in enumerate(blocks):
        clip = re.search(r'([\d.]+) ([\d.]+) ([\d.]+) ([\d.]+) re\s+W n', block)
        if not clip:
            continue
        c = tuple(float(v) for v in clip.groups())
        ax = axes.setdefault(c, dict(clip=c, yticks=[], lines=[]))
        pts = np.array(re.findall(r'(-?[\d.]+) (-?[\d.]+) [ml]\b', block), float)
        if len(pts) > 100:
            col = re.findall(r'([\d.]+) ([\d.]+) ([\d.]+) RG', block)
            dashed = re.search(r'\[\s*[\d.]+ [\d.]+\s*\]\s*\d+\s*d', block) is not None
            ax['lines'].append((tuple(float(v) for v in col[-1]), dashed, pts))
        elif len(pts) == 2 and abs(pts[0, 1] - pts[1, 1]) < 1e-9 and abs(pts[0, 0] - c[0]) < 1e-4:
            # horizontal grid line spanning the axes; its label is in the next block's BT..ET
            nxt = blocks[bi + 1] if bi + 1 < len(blocks) else ''
            bt = re.search(r'BT(.*?)ET', nxt, re.S)
            if not bt:
                continue
            toks = re.findall(r'/(F\d+) [\d.]+ Tf|\(((?:[^()\\]|\\.)*)\) Tj', bt.group(1))
            text, neg, font = '', False, None
            for f, s in toks:
                if f:
                    font = f
                    continue
                # usetex fonts: the decimal point is ':' of the math-italic font (cmmi), the minus sign
                # is a non-printing glyph of the symbol font (cmsy); digits come from cmr
                if re.fullmatch(r'[\d.]+', s):
                    text += s
                elif s == ':':
                    text += '.'
                else:
                    neg = True
            if text:
                ax['yticks'].append((pts[0, 1], -float(text) if neg else float(text)))
    return list(axes.values())


def calibrate(ax):
    ty = np.array(sorted(set(ax['yticks'])))
    assert len(ty) >= 2, ax['yticks']
    # least-squares line through all labelled ticks (they are exactly linear)
    A = np.stack([ty[:, 0], np.ones(len(ty))], 1)
    k, b = np.linalg.lstsq(A, ty[:, 1], rcond=None)[0]
    assert np.max(np.abs(A @ np.array([k, b]) - ty[:, 1])) < 1e-6
    return k, b


def series(ax, pts):
    k, b = calibrate(ax)
    x0, wd = ax['clip'][0], ax['clip'][2]
    # the x-limits are 0 .. T (six ticks 0,2,..,10 across the clip box)
    step_f = (pts[:, 0] - x0) / wd * (T_END / DT)
    step = np.rint(step_f).astype(int)
    assert np.max(np.abs(step_f - step)) < 1e-3
    return step, k * pts[:, 1] + b


SOLID_SIM = (0.5294117647, 0.8078431373, 0.9215686275)
SOLID_U = (0.0, 0.3921568627, 0.0)
SOLID_A = (0.0, 0.7490196078, 1.0)


def pick(ax, colour, dashed=False):
    return [pts for (c, d, pts) in ax['lines'] if np.allclose(c, colour, atol=1e-6) and d == dashed]


def extract(pdf_path, jerk):
    axes = parse_axes(page_stream(pdf_path))
    axes = [a for a in axes if a['lines']]
    axes.sort(key=lambda a: -a['clip'][1])  # top to bottom: p, v, (a), theta, F_d
    out = {}
    names = ['p', 'v'] + (['a'] if jerk else []) + ['theta', 'Fd']
    assert len(axes) == len(names), (len(axes), names)
    for name, ax in zip(names, axes):
        if name in ('p', 'v'):
            lx, lz = pick(ax, SOLID_SIM)
            for comp, pts in (('x', lx), ('z', lz)):
                out[f'{name}{comp}_step'], out[f'{name}{comp}'] = series(ax, pts)
        elif name == 'a':
            lx, lz = pick(ax, SOLID_A, dashed=False)
            for comp, pts in (('x', lx), ('z', lz)):
                out[f'a{comp}_step'], out[f'a{comp}'] = series(ax, pts)
        else:
            (pts,) = pick(ax, SOLID_U)
            out[f'{name}_step'], out[name] = series(ax, pts)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--ref', default='/root/reference')
    ap.add_argument('--out', default=os.path.join(os.path.dirname(__file__), '..', 'tests', 'golden'))
    args = ap.parse_args()
    img = os.path.join(args.ref, 'experiment_data', 'img')
    for tag, fn, jerk in (('force', 'example_acc_trajectory_component.pdf', False),
                          ('jerk', 'example_jerk_trajectory_component.pdf', True)):
        d = extract(os.path.join(img, fn), jerk)
        path = os.path.join(args.out, f'acados_{tag}.npz')
        np.savez_compressed(path, **d)
        print(tag, {k: (len(v), float(v[0])) for k, v in d.items() if not k.endswith('_step')})


if __name__ == '__main__':
    main()
