#!/usr/bin/env python3
"""Model code generator: symbolic ODE -> straight-line CUDA device code.

Plays the role CasADi's C export plays for the reference (expl_ode_fun / expl_vde_forw): each
model is written down symbolically exactly as in the reference

    force   reference src/force_model/dynamics.py:32-37     xdot = [vx, vz, Fx/m, Fz/m - g]
    jerk    reference src/jerk_model/dynamics.py:35-42      xdot = [vx, vz, ax, az - g, hx, hz]
    plant   reference src/plant.py:27-33                    xdot = [vx, vz, Fd sin(th)/m, Fd cos(th)/m - g]

(mass and g become the per-instance parameter vector p = (m, g) instead of baked-in constants), and this script
emits, per model, a struct with
  * f / jac               the ODE right-hand side and its Jacobians w.r.t. x and u, CSE'd straight-line code
  * f_blk / jac_blk       the same per *independent block*: the generator finds the connected components of the
                          dependency graph of the ODE (state i <- variables its derivative reads).  When all
                          components have the same shape the OCP splits into NBLK identical-shape sub-problems that
                          share only the interior-point scalars; the solver then maps one thread per (instance, block)
  * JAC_CONST             true when no Jacobian entry depends on (x, u): sensitivities are computed once per solve
The output csrc/generated/models_gen.cuh is committed, so neither the GPU box nor a user needs sympy.

Usage: python -m drone_attitude_control_b200.codegen.gen_models
"""
import os

import sympy as sp
from sympy.printing.c import C99CodePrinter


class _Printer(C99CodePrinter):
    """Literals are wrapped in T(...) so the same code instantiates for double and float."""

    def _print_Float(self, e):
        return 'T(%s)' % super()._print_Float(e)

    def _print_Integer(self, e):
        return 'T(%d)' % int(e)

    def _print_Rational(self, e):
        return '(T(%d)/T(%d))' % (e.p, e.q)

    def _print_Pow(self, e):
        if e.exp == -1:
            return '(T(1)/(%s))' % self._print(e.base)
        if e.exp.is_Integer and 0 < int(e.exp) <= 3:
            return '(' + '*'.join(['(%s)' % self._print(e.base)] * int(e.exp)) + ')'
        return super()._print_Pow(e)


_pr = _Printer()


def _components(xs, us, f):
    """Connected components of the graph on (states + inputs): state i -- every variable xdot_i reads."""
    allv = list(xs) + list(us)
    parent = {v: v for v in allv}

    def find(a):
        while parent[a] != a:
            parent[a] = parent[parent[a]]
            a = parent[a]
        return a

    for xi, fi in zip(xs, f):
        for v in fi.free_symbols:
            if v in parent:
                parent[find(v)] = find(xi)
    comps = {}
    for v in allv:
        comps.setdefault(find(v), []).append(v)
    out = []
    for c in comps.values():
        out.append(([i for i, x in enumerate(xs) if x in c], [i for i, u in enumerate(us) if u in c]))
    out.sort(key=lambda c: c[0][0] if c[0] else 10 ** 6)
    return out


def _emit_assign(outs, exprs, indent):
    """CSE + straight-line assignments; outs = list of lvalue strings."""
    repl, red = sp.cse(list(exprs), symbols=sp.numbered_symbols('t_'), optimizations='basic')
    lines = ['%sconst T %s = %s;' % (indent, _pr.doprint(s), _pr.doprint(e)) for s, e in repl]
    lines += ['%s%s = %s;' % (indent, o, _pr.doprint(e)) for o, e in zip(outs, red)]
    return lines


def _subs_arrays(exprs, xs, us, ps, xmap=None, umap=None):
    """Replace symbols by array references x[i], u[i], p[i] (optionally through local index maps)."""
    d = {}
    for i, x in enumerate(xs):
        if xmap is None or i in xmap:
            d[x] = sp.Symbol('x[%d]' % (i if xmap is None else xmap[i]))
    for i, u in enumerate(us):
        if umap is None or i in umap:
            d[u] = sp.Symbol('u[%d]' % (i if umap is None else umap[i]))
    for i, p in enumerate(ps):
        d[p] = sp.Symbol('p[%d]' % i)
    return [sp.sympify(e).subs(d) for e in exprs]


_ERK_TABLEAUS = {
    1: ([[0]], [1]),
    2: ([[0, 0], [sp.Rational(1, 2), 0]], [0, 1]),
    3: ([[0, 0, 0], [sp.Rational(1, 2), 0, 0], [-1, 2, 0]], [sp.Rational(1, 6), sp.Rational(2, 3), sp.Rational(1, 6)]),
    4: ([[0, 0, 0, 0], [sp.Rational(1, 2), 0, 0, 0], [0, sp.Rational(1, 2), 0, 0], [0, 0, 1, 0]],
        [sp.Rational(1, 6), sp.Rational(1, 3), sp.Rational(1, 3), sp.Rational(1, 6)]),
}


def _sens_structure(xs, us, f, comps):
    """Structural zeros / ones of the discrete-time sensitivities A = d x+/d x, B = d x+/d u of one explicit RK step,
    valid for every stage count 1..4, every step size and every block (the intersection is emitted, so the flags do not
    depend on run-time options).  Only used for constant-Jacobian models."""
    h = sp.Symbol('h', positive=True)
    nxb, nub = len(comps[0][0]), len(comps[0][1])
    a_zero = [[True] * nxb for _ in range(nxb)]; a_one = [[True] * nxb for _ in range(nxb)]
    b_zero = [[True] * nub for _ in range(nxb)]
    for cx, cu in comps:
        xb = [xs[i] for i in cx]; ub = [us[i] for i in cu]; fb = [f[i] for i in cx]
        for ns, (At, bt) in _ERK_TABLEAUS.items():
            K = []
            for i in range(ns):
                xi = [xb[r] + h * sum(At[i][j] * K[j][r] for j in range(i)) for r in range(nxb)]
                K.append([sp.expand(e.subs(dict(zip(xb, xi)), simultaneous=True)) for e in fb])
            xn = [sp.expand(xb[r] + h * sum(bt[i] * K[i][r] for i in range(ns))) for r in range(nxb)]
            A = sp.Matrix(xn).jacobian(sp.Matrix(xb)); B = sp.Matrix(xn).jacobian(sp.Matrix(ub))
            for r in range(nxb):
                for c in range(nxb):
                    e = sp.simplify(A[r, c])
                    a_zero[r][c] &= (e == 0); a_one[r][c] &= (e == 1)
                for c in range(nub):
                    b_zero[r][c] &= (sp.simplify(B[r, c]) == 0)
    return a_zero, a_one, b_zero


def gen_model(name, xs, us, ps, f, force_single_block=False):
    nx, nu, npar = len(xs), len(us), len(ps)
    f = [sp.sympify(e) for e in f]
    Jx = sp.Matrix(f).jacobian(sp.Matrix(xs))
    Ju = sp.Matrix(f).jacobian(sp.Matrix(us))
    jac_const = not any((set(xs) | set(us)) & e.free_symbols for e in list(Jx) + list(Ju))
    comps = _components(xs, us, f)
    shapes = {(len(c[0]), len(c[1])) for c in comps}
    if force_single_block or len(comps) == 1 or len(shapes) != 1 or any(len(c[0]) == 0 for c in comps):
        comps = [(list(range(nx)), list(range(nu)))]
    nblk = len(comps)
    nxb, nub = len(comps[0][0]), len(comps[0][1])
    L = []
    L.append('// model "%s": nx=%d nu=%d np=%d, %d block(s) of (%d states, %d inputs), Jacobian %s' % (
        name, nx, nu, npar, nblk, nxb, nub, 'constant' if jac_const else 'state/input dependent'))
    L.append('struct Model_%s {' % name)
    L.append('    static constexpr int NX = %d, NU = %d, NP = %d;' % (nx, nu, npar))
    L.append('    static constexpr int NBLK = %d, NXB = %d, NUB = %d;' % (nblk, nxb, nub))
    L.append('    static constexpr bool JAC_CONST = %s;' % ('true' if jac_const else 'false'))
    L.append('    static constexpr const char* name() { return "%s"; }' % name)
    if jac_const:
        a_zero, a_one, b_zero = _sens_structure(xs, us, f, comps)
    else:
        a_zero = [[False] * nxb for _ in range(nxb)]; a_one = [[False] * nxb for _ in range(nxb)]
        b_zero = [[False] * max(nub, 1) for _ in range(nxb)]
    L.append('    // structure of the discrete-time sensitivities A (NXB x NXB), B (NXB x NUB) of one ERK step, found symbolically:')
    L.append('    // entries that are identically 0 / 1 for every stage count, step size and block (all false if not constant)')
    for nm, tab, nc in (('a_zero', a_zero, nxb), ('a_one', a_one, nxb), ('b_zero', b_zero, max(nub, 1))):
        rows = ', '.join('{' + ', '.join('true' if v else 'false' for v in row) + '}' for row in tab)
        L.append('    __host__ __device__ static constexpr bool %s(int r, int c) {' % nm)
        L.append('        constexpr bool tab[%d][%d] = {%s};' % (nxb, nc, rows))
        L.append('        return tab[r][c];')
        L.append('    }')
    # structural zeros of the continuous-time Jacobians of a block (entries that are identically zero in every block): the
    # sensitivity propagation of a large block skips those products (erk_step)
    jx_zero = [[all(Jx[cx[r], cx[c]] == 0 for cx, _ in comps) for c in range(nxb)] for r in range(nxb)]
    ju_zero = [[all(Ju[cx[r], cu[c]] == 0 for cx, cu in comps) for c in range(nub)] for r in range(nxb)]
    for nm, tab, nc in (('jx_zero', jx_zero, nxb), ('ju_zero', ju_zero, max(nub, 1))):
        rows = ', '.join('{' + ', '.join('true' if v else 'false' for v in row) + '}' for row in tab)
        L.append('    __host__ __device__ static constexpr bool %s(int r, int c) {' % nm)
        L.append('        constexpr bool tab[%d][%d] = {%s};' % (nxb, nc, rows))
        L.append('        return tab[r][c];')
        L.append('    }')
    for nm, idx in (('xg', 0), ('ug', 1)):
        n_loc = nxb if idx == 0 else nub
        rows = ', '.join('{' + ', '.join(str(i) for i in c[idx]) + '}' for c in comps)
        L.append('    // global index of block-local %s j of block b' % ('state' if idx == 0 else 'input'))
        L.append('    __host__ __device__ static constexpr int %s(int b, int j) {' % nm)
        L.append('        constexpr int tab[%d][%d] = {%s};' % (nblk, max(n_loc, 1), rows))
        L.append('        return tab[b][j];')
        L.append('    }')
    # full model
    fe = _subs_arrays(f, xs, us, ps)
    L.append('    template <class T> __host__ __device__ __forceinline__ static void f(const T* x, const T* u, const T* p, T* xd) {')
    L.append('        (void)x; (void)u; (void)p;')
    L += _emit_assign(['xd[%d]' % i for i in range(nx)], fe, '        ')
    L.append('    }')
    je = _subs_arrays(list(Jx) + list(Ju), xs, us, ps)
    L.append('    // fx: NX x NX row-major, fu: NX x NU row-major')
    L.append('    template <class T> __host__ __device__ __forceinline__ static void jac(const T* x, const T* u, const T* p, T* fx, T* fu) {')
    L.append('        (void)x; (void)u; (void)p;')
    L += _emit_assign(['fx[%d]' % i for i in range(nx * nx)] + ['fu[%d]' % i for i in range(nx * nu)], je, '        ')
    L.append('    }')
    # per block
    L.append('    // block-local versions: x has NXB entries, u has NUB entries (order given by xg/ug)')
    L.append('    template <class T> __host__ __device__ __forceinline__ static void f_blk(int b, const T* x, const T* u, const T* p, T* xd) {')
    L.append('        (void)x; (void)u; (void)p;')
    L.append('        switch (b) {')
    for b, (cx, cu) in enumerate(comps):
        xmap = {g: j for j, g in enumerate(cx)}
        umap = {g: j for j, g in enumerate(cu)}
        fb = _subs_arrays([f[g] for g in cx], xs, us, ps, xmap, umap)
        L.append('        %s: {' % ('default' if b == nblk - 1 else 'case %d' % b))
        L += _emit_assign(['xd[%d]' % j for j in range(nxb)], fb, '            ')
        L.append('        } break;')
    L.append('        }')
    L.append('    }')
    L.append('    template <class T> __host__ __device__ __forceinline__ static void jac_blk(int b, const T* x, const T* u, const T* p, T* fx, T* fu) {')
    L.append('        (void)x; (void)u; (void)p;')
    L.append('        switch (b) {')
    for b, (cx, cu) in enumerate(comps):
        xmap = {g: j for j, g in enumerate(cx)}
        umap = {g: j for j, g in enumerate(cu)}
        jb = [Jx[r, c] for r in cx for c in cx] + [Ju[r, c] for r in cx for c in cu]
        jb = _subs_arrays(jb, xs, us, ps, xmap, umap)
        L.append('        %s: {' % ('default' if b == nblk - 1 else 'case %d' % b))
        L += _emit_assign(['fx[%d]' % i for i in range(nxb * nxb)] + ['fu[%d]' % i for i in range(nxb * nub)], jb, '            ')
        L.append('        } break;')
    L.append('        }')
    L.append('    }')
    L.append('};')
    return '\n'.join(L), dict(name=name, nx=nx, nu=nu, nblk=nblk, nxb=nxb, nub=nub, jac_const=jac_const, comps=comps)


def models():
    m, g = sp.symbols('m g')
    px, pz, vx, vz, ax, az = sp.symbols('px pz vx vz ax az')
    Fx, Fz, hx, hz, th, Fd = sp.symbols('Fx Fz hx hz theta Fd')
    out = []
    out.append(gen_model('force', [px, pz, vx, vz], [Fx, Fz], [m, g], [vx, vz, Fx / m, Fz / m - g]))
    out.append(gen_model('jerk', [px, pz, vx, vz, ax, az], [hx, hz], [m, g], [vx, vz, ax, az - g, hx, hz]))
    out.append(gen_model('plant', [px, pz, vx, vz], [th, Fd], [m, g],
                         [vx, vz, Fd * sp.sin(th) / m, Fd * sp.cos(th) / m - g]))
    # dense (single-block) variants of the controller models: same ODEs, block detection disabled.  They exercise the
    # generic coupled path and serve as an in-product cross-check of the block-split path.
    out.append(gen_model('force_dense', [px, pz, vx, vz], [Fx, Fz], [m, g], [vx, vz, Fx / m, Fz / m - g], True))
    out.append(gen_model('jerk_dense', [px, pz, vx, vz, ax, az], [hx, hz], [m, g], [vx, vz, ax, az - g, hx, hz], True))
    # NOT in the reference (north-star extension): the 3-D attitude-and-total-thrust model - position,
    # velocity and the attitude quaternion (body -> world) as states, total thrust and the body rates as inputs:
    #   pdot = v,  vdot = (T / m) R(q) e3 - g e3,  qdot = 1/2 q (x) (0, w)
    # (the planar plant of src/plant.py:27-33 is its restriction to the x-z plane: theta = pitch, Fd = T)
    py, vy, qw, qx, qy, qz, Tt, wx, wy, wz = sp.symbols('py vy qw qx qy qz Tt wx wy wz')
    half = sp.Rational(1, 2)
    out.append(gen_model('att', [px, py, pz, vx, vy, vz, qw, qx, qy, qz], [Tt, wx, wy, wz], [m, g], [
        vx, vy, vz,
        2 * (qx * qz + qw * qy) * Tt / m, 2 * (qy * qz - qw * qx) * Tt / m, (1 - 2 * (qx * qx + qy * qy)) * Tt / m - g,
        half * (-qx * wx - qy * wy - qz * wz), half * (qw * wx + qy * wz - qz * wy),
        half * (qw * wy - qx * wz + qz * wx), half * (qw * wz + qx * wy - qy * wx)]))
    return out


def main():
    here = os.path.dirname(os.path.abspath(__file__))
    dst = os.path.join(here, '..', 'csrc', 'generated', 'models_gen.cuh')
    parts = ['// GENERATED by drone_attitude_control_b200/codegen/gen_models.py - do not edit.', '#pragma once', '',
             'namespace bnmpc {', '']
    for code, info in models():
        parts += [code, '']
        print(info)
    parts += ['}  // namespace bnmpc', '']
    with open(dst, 'w') as fh:
        fh.write('\n'.join(parts))
    print('wrote', os.path.normpath(dst))


if __name__ == '__main__':
    main()
