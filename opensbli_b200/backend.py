"""`B200(alg)` -- the drop-in replacement of the reference's `OPSC(alg)` back end.

Same call shape as `OPSC(algorithm, operation_count=False, OPS_diagnostics=1)` (opensbli/code_generation/opsc.py:253).
Where OPSC walks the algorithm tree and prints OPS-C text, this class walks the same tree
(`alg.prg.components`, as `OPSC.loop_alg` does, opsc.py:595-609), recognises the loop families of the solver hot path
 and distils them into an execution *plan* for the hand-written sm_100a kernels:

    CRu_i / CRp / CRa / CRT ...................... constituent relations (verified numerically against the canonical forms)
    LLF{Teno,Weno}_reconstruction_d_direction .... characteristic flux sweep: scheme, order, formulation, averaging
    LLF... Residual .............................. flux difference
    Convective ... / Viscous ... / Derivative ... . Central(4) / StoreSome(4) convective + viscous terms
    Temporal solution / Sub stage / Save ......... RungeKuttaLS / RungeKutta, coefficients read from the IR
    ExchangeSelf, '<BC> boundary dir d side s' ... periodic exchanges and boundary kernels
    Grid_based_initialisation .................... cold kernel, turned into numpy statements evaluated by the runner

It writes, in the current directory (like OPSC writes opensbli.cpp & co, opsc.py:278,420,492):
    opensbli_b200.plan.json  the symbolic plan (constants still named)
    opensbli.cpp             a parameter stub holding the `name=Input;` lines, so that the unchanged
                             `substitute_simulation_parameters` (utilities/helperfunctions.py:130-149) and
                             `print_iteration_ops` (172-190) keep working; `python -m opensbli_b200.run` reads it back.
Anything outside the recognised path raises NotImplementedError naming the offending loop -- there is no silent fallback.
"""
import json
import re

import numpy as np

PLAN_FILE = 'opensbli_b200.plan.json'
STUB_FILE = 'opensbli.cpp'


class UnsupportedByB200(NotImplementedError):
    pass


def _walk(components, out, path=()):
    for c in components:
        if hasattr(c, 'components'):
            _walk(c.components, out, path + (c,))
        else:
            out.append((path, c))


def _name(c):
    return getattr(c, 'computation_name', None) or getattr(c, 'name', None) or type(c).__name__


def _strip(ds):
    s = str(ds)
    return s[:-3] if s.endswith('_B0') else s


class _Printer(object):
    """numpy-syntax printer for cold-kernel equations (initialisation, Dirichlet states)."""

    def __init__(self):
        from sympy.printing.numpy import NumPyPrinter

        class P(NumPyPrinter):
            def _print_DataSet(s, e):
                return _strip(e.base)

            def _print_Grididx(s, e):
                return 'idx%d' % int(e.number)

            def _print_Indexed(s, e):
                if type(e).__name__ == 'Grididx':
                    return 'idx%d' % int(e.number)
                if type(e).__name__ == 'DataSet':
                    return _strip(e.base)
                return str(e)

            def _print_Piecewise(s, e):
                # nested numpy.where: broadcasts over the grid and accepts the trailing (value, True) pair
                out = None
                for val, cond in reversed(e.args):
                    v = s._print(val)
                    out = v if (cond is True or cond == True) and out is None else 'numpy.where(%s, %s, %s)' % (s._print(cond), v, out if out is not None else 'numpy.nan')  # noqa: E712
                return out

            def _print_GroupedPiecewise(s, e):
                return s._print_Piecewise(e)

            def _print_GridVariable(s, e):
                return str(e)

            def _print_ConstantObject(s, e):
                return str(e)

            def _print_Symbol(s, e):
                return str(e)
        self.p = P({'fully_qualified_modules': True, 'precision': 17})

    def __call__(self, expr):
        return self.p.doprint(expr)


def _cold_statements(kernel):
    pr = _Printer()
    out = []
    for e in kernel.equations:
        if not hasattr(e, 'lhs'):
            raise UnsupportedByB200('cold kernel %s: unsupported equation %r' % (_name(kernel), e))
        out.append([pr(e.lhs), pr(e.rhs)])
    return out


def _lambdify_check(eq, canonical, names, ntry=8, tol=1e-12):
    """numerically compare eq.rhs with canonical(**values) on random positive inputs."""
    from sympy import lambdify, Symbol
    from opensbli.core.opensbliobjects import DataSet, ConstantObject
    atoms = list(eq.rhs.atoms(DataSet)) + list(eq.rhs.atoms(ConstantObject))
    rep = {a: Symbol('v_' + (_strip(a.base) if type(a).__name__ == 'DataSet' else str(a))) for a in atoms}
    syms = sorted(set(rep.values()), key=str)
    f = lambdify(syms, eq.rhs.xreplace(rep), 'math')
    rng = np.random.default_rng(0)
    for _ in range(ntry):
        vals = {str(s)[2:]: float(0.5 + rng.random()) for s in syms}
        got = f(*[vals[str(s)[2:]] for s in syms])
        want = canonical(vals)
        if abs(got - want) > tol * max(1.0, abs(want)):
            return False
    return True


def _check_constituent(kernels, ndim):
    """The hand-written kernels hard-wire the ideal-gas constituent relations; make sure the app's are those."""
    canon = {
        'p': lambda v: (v['gama'] - 1.0) * (v['rhoE'] - 0.5 * v['rho'] * sum(v['u%d' % d] ** 2 for d in range(ndim))),
        'a': lambda v: (v['gama'] * v['p'] / v['rho']) ** 0.5,
        'T': lambda v: v['p'] * v['gama'] * v['Minf'] ** 2 / v['rho'],
    }
    for d in range(ndim):
        canon['u%d' % d] = (lambda d: (lambda v: v['rhou%d' % d] / v['rho']))(d)
    seen = []
    for k in kernels:
        for e in k.equations:
            lhs = _strip(e.lhs.base) if hasattr(e.lhs, 'base') else str(e.lhs)
            if lhs not in canon:
                raise UnsupportedByB200("constituent relation for '%s' (kernel %s) is outside the canonical ideal-gas set "
                                        "{u_i, p, a, T} the B200 kernels implement" % (lhs, _name(k)))
            if not _lambdify_check(e, canon[lhs], None):
                raise UnsupportedByB200("constituent relation %s = %s differs from the canonical form" % (lhs, e.rhs))
            seen.append(lhs)
    return seen


def _datasets_used(kernels):
    from opensbli.core.opensbliobjects import DataSet
    used = set()
    for k in kernels:
        for e in k.equations:
            if hasattr(e, 'rhs'):
                used |= set(_strip(ds.base) for ds in e.rhs.atoms(DataSet))
    return used


def _recon_info(k):
    names = [str(e.lhs) for e in k.equations if hasattr(e, 'lhs')]
    m = re.match(r'LLF(Teno|Weno)_reconstruction_(\d)_direction', _name(k))
    kind = m.group(1).lower()
    info = {'conv': kind, 'direction': int(m.group(2))}
    if kind == 'teno':
        nst = len(set(n for n in names if re.match(r'delta_\d+$', n)))
        info['order'] = {3: 5, 4: 6}.get(nst)
        if info['order'] is None:
            raise UnsupportedByB200('TENO with %d candidate stencils is not implemented' % nst)
        if 'TENO_CT' in names:
            raise UnsupportedByB200('adaptive TENO (shock-sensor controlled C_T) is not implemented yet')
        info['weno_formulation'] = 'JS'
    else:
        nst = len(set(n for n in names if re.match(r'omega_\d+$', n)))
        if nst != 3:
            raise UnsupportedByB200('WENO order %d is not implemented (only 5)' % (2 * nst - 1))
        info['order'] = 5
        a0 = k.equations[names.index('alpha_0')].rhs
        from sympy import Abs
        info['weno_formulation'] = 'Z' if a0.has(Abs) else 'JS'
    d = info['direction']
    info['averaging'] = 'roe' if ('AVG_%d_inv_rho' % d) in names else 'simple'
    used = _datasets_used([k])
    extra = [u for u in used if re.match(r'(D\d\d|detJ|SD\d+)$', u)]
    if extra:
        raise UnsupportedByB200('curvilinear metric terms %s in %s are not implemented yet' % (sorted(extra), _name(k)))
    return info


def _const_value(c):
    from sympy.printing.c import ccode
    v = c.value
    if isinstance(v, str):
        return v
    return ccode(v)


def extract_plan(algorithm):
    """Distil the plan from an OpenSBLI algorithm object (see module docstring)."""
    from opensbli.core.kernel import ConstantsToDeclare
    if getattr(algorithm, 'MultiBlock', False) or len(algorithm.block_descriptions) != 1:
        raise UnsupportedByB200('multi-block algorithms are not implemented')
    if str(algorithm.dtype).lower() not in ('double', 'dtype.double'):
        pass
    ndim = algorithm.block_descriptions[0].ndim
    flat = []
    _walk(algorithm.prg.components, flat)
    before, in_iter, in_stage, after = [], [], [], []
    for path, c in flat:
        loops = [type(p).__name__ for p in path]
        nloops = loops.count('DoLoop')
        (before if nloops == 0 and 'Timers' not in loops else in_iter if nloops == 1 else in_stage if nloops == 2 else after).append(c)
        if nloops == 0 and 'Timers' in loops:
            after.append(c)
    plan = {'ndim': ndim, 'viscous': False, 'averaging': 'roe', 'weno_formulation': 'JS'}
    q_names = ['rho'] + ['rhou%d' % d for d in range(ndim)] + ['rhoE']

    # ---- stage loop: classify every kernel
    cr, recon, central_conv, viscous, rk_kernels, stage_bcs, unknown = [], [], [], [], [], [], []
    for c in in_stage:
        t, n = type(c).__name__, _name(c)
        if t == 'ExchangeSelf' or ' boundary dir' in n:
            stage_bcs.append(c)
        elif t != 'Kernel':
            unknown.append(c)
        elif n.startswith('CR'):
            cr.append(c)
        elif re.match(r'LLF(Teno|Weno)_reconstruction_\d_direction', n):
            recon.append(c)
        elif re.match(r'LLF(Teno|Weno) Residual', n):
            pass
        elif n.startswith('Convective'):
            central_conv.append(c)
        elif n.startswith('Viscous') or n.startswith('Derivative evaluation'):
            viscous.append(c)
        elif n in ('Temporal solution advancement', 'Sub stage advancement'):
            rk_kernels.append(c)
        else:
            unknown.append(c)
    if unknown:
        raise UnsupportedByB200('loops outside the accelerated hot path: %s' % sorted(set(_name(c) for c in unknown)))
    _check_constituent(cr, ndim)
    if recon and central_conv:
        raise UnsupportedByB200('mixed shock-capturing and central convective terms are not implemented')
    if recon:
        infos = [_recon_info(k) for k in recon]
        if sorted(i['direction'] for i in infos) != list(range(ndim)):
            raise UnsupportedByB200('reconstruction kernels do not cover every direction once')
        for key in ('conv', 'order', 'weno_formulation', 'averaging'):
            if len(set(i[key] for i in infos)) != 1:
                raise UnsupportedByB200('direction-dependent %s is not implemented' % key)
            plan[key] = infos[0][key]
    elif central_conv:
        plan.update(conv='central', order=4)
        if len([k for k in central_conv if 'CD' in _name(k)]) != {1: 6, 2: 18, 3: 39}.get(ndim, -1) and ndim == 3:
            raise UnsupportedByB200('unexpected set of central convective derivative loops (only the skew-symmetric '
                                    'Navier-Stokes form of apps/taylor_green_vortex is implemented)')
    else:
        raise UnsupportedByB200('no convective discretisation found in the stage loop')
    if viscous:
        plan['viscous'] = True
        used = _datasets_used(viscous)
        if 'mu' in used:
            raise UnsupportedByB200('variable viscosity (mu as a constituent relation) is not implemented yet')
        if any(re.match(r'(D\d\d|detJ|SD\d+)$', u) for u in used):
            raise UnsupportedByB200('curvilinear metric terms in the viscous loops are not implemented yet')

    # ---- Runge-Kutta kind and coefficients (rk_LS.py:70-102, rk_sbli.py:58-61)
    consts = {}
    for k in rk_kernels:
        for ic in k.IndexedConstants:
            consts[str(ic.base.label)] = [float(v) for v in ic.value]
    if 'rkA' in consts and 'rkB' in consts:
        plan.update(rk='ls', rk_a=consts['rkA'], rk_b=consts['rkB'])
    elif 'rkold' in consts and 'rknew' in consts:
        plan.update(rk='sbli', rk_a=consts['rkold'], rk_b=consts['rknew'])
    else:
        raise UnsupportedByB200('unrecognised Runge-Kutta update kernels (constants %s)' % sorted(consts))

    # ---- boundary conditions, from the iteration-start list (algorithm.py:442)
    bc = [[None, None] for _ in range(ndim)]
    cold_pr = _Printer()
    for c in [c for c in in_iter if type(c).__name__ == 'ExchangeSelf' or ' boundary dir' in _name(c)]:
        if type(c).__name__ == 'ExchangeSelf':
            arrays = [_strip(a) for a in c.transfer_arrays]
            if arrays != q_names:
                raise UnsupportedByB200('periodic exchange of %s (expected the conserved arrays)' % arrays)
            side = {'left': 0, 'right': 1}.get(c.side, c.side)
            bc[int(c.direction)][int(side)] = {'type': 'periodic'}
            continue
        m = re.match(r'(\w+) boundary dir(\d) side(\d)', _name(c))
        kind, d, s = m.group(1), int(m.group(2)), int(m.group(3))
        if kind != 'Dirichlet':
            raise UnsupportedByB200("boundary condition '%s' is not implemented yet" % kind)
        bc[d][s] = {'type': 'dirichlet', 'statements': _cold_statements(c)}
    if any(b is None for pair in bc for b in pair):
        raise UnsupportedByB200('a block face has no recognised boundary condition')
    plan['bc'] = bc

    # ---- cold kernels before the time loop
    init = []
    for c in before:
        if type(c).__name__ == 'Kernel':
            if not _name(c).startswith('Grid_based_initialisation'):
                raise UnsupportedByB200('cold kernel %s is not implemented yet' % _name(c))
            init += _cold_statements(c)
    plan['init'] = init
    plan['q_names'] = q_names

    # ---- constants, in declaration order (opsc.py:625-654)
    plan['constant_decls'] = []
    for c in ConstantsToDeclare.constants:
        if type(c).__name__ == 'ConstantObject':
            plan['constant_decls'].append([str(c), 'int' if 'int' in str(c.datatype.opsc()).lower() else 'double', _const_value(c)])
    return plan


def write_stub(plan, path=STUB_FILE):
    """opensbli.cpp-named parameter stub: the `name=Input;` lines of OPSC's main program (opsc.py:625-632)."""
    L = ['// OpenSBLI B200 back end: simulation parameters (filled in by substitute_simulation_parameters)',
         '// run with:  python -m opensbli_b200.run', 'int main(int argc, char **argv)', '{']
    for name, dtype, value in plan['constant_decls']:
        L.append('%s=%s;' % (name, value) if value == 'Input' else '%s = %s;' % (name, value))
    L += ['int iter=0;', '', '}']
    open(path, 'w').write('\n'.join(L) + '\n')


class B200(object):
    def __init__(self, algorithm, operation_count=False, OPS_diagnostics=1, workdir='.'):
        import os
        self.operation_count = operation_count
        self.OPS_diagnostics = OPS_diagnostics
        self.plan = extract_plan(algorithm)
        with open(os.path.join(workdir, PLAN_FILE), 'w') as f:
            json.dump(self.plan, f, indent=1)
        write_stub(self.plan, os.path.join(workdir, STUB_FILE))
        print("Successfully generated the B200 execution plan.")
