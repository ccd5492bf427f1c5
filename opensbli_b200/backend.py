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
import os
import re

import numpy as np

PLAN_FILE = 'opensbli_b200.plan.json'
STUB_FILE = 'opensbli.cpp'


class UnsupportedByB200(NotImplementedError):
    pass


def _walk(components, out, path=()):
    for c in components:
        if hasattr(c, 'components') and type(c).__name__ not in ('SimulationMonitor',):
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
                off = tuple(int(i) for i in e.indices)
                if any(off):
                    return "_A('%s', %r)" % (_strip(e.base), off)
                return "_A('%s')" % _strip(e.base)

            def _print_Grididx(s, e):
                return 'idx%d' % int(e.number)

            def _print_Indexed(s, e):
                if type(e).__name__ == 'Grididx':
                    return 'idx%d' % int(e.number)
                if type(e).__name__ == 'DataSet':
                    return s._print_DataSet(e)
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


def _user_kernel(kernel, when, stencil=False, thin_direction=None):
    """A user kernel compiled at run time (NVRTC): an ordered list of assignments [lhs, is_dataset, rhs in C, lhs index] over
    the kernel's range -- app-specific arithmetic outside the solver's hot loops cannot be hand-written.
    stencil=False: point-wise kernels (statistics accumulation, e.g. channel_flow/*/stats.py; `User kernel` in the algorithm).
    stencil=True: datasets may be read and written at relative offsets (boundary-condition kernels of classes without a
    hand-written kernel, `when` = 'bc_<dir>_<side>'); every point of the range runs in its own thread, so the kernel must not
    write where another point of the range reads or writes -- checked here from the offsets and the range's unit extents."""
    from sympy.printing.c import C99CodePrinter, ccode
    from opensbli.core.opensbliobjects import DataSet

    def index(e):
        off = [int(i) for i in e.indices]
        if any(off) and not stencil:
            raise UnsupportedByB200('user kernel %s accesses %s at an offset: only point-wise user kernels are implemented' % (_name(kernel), e))
        terms = ['X'] + ['(%d)%s' % (o, ('', '*s1', '*s2')[d]) for d, o in enumerate(off) if o]
        return ' + '.join(terms), tuple(off + [0] * (3 - len(off)))

    class P(C99CodePrinter):
        def _print_DataSet(s, e):
            return '%s[%s]' % (_strip(e.base), index(e)[0])

        def _print_Indexed(s, e):
            if type(e).__name__ == 'DataSet':
                return s._print_DataSet(e)
            if type(e).__name__ == 'Grididx':
                return 'idx%d' % int(e.indices[0])
            return C99CodePrinter._print_Indexed(s, e)

        def _print_Grididx(s, e):
            return 'idx%d' % int(e.indices[0])

        def _print_Rational(s, e):
            return '(%d.0/%d.0)' % (e.p, e.q)

        def _print_GroupedPiecewise(s, e):      # used as an EXPRESSION (teno.py:456-459: delta_r = 0 if ... else 1): a plain Piecewise;
            return s._print_Piecewise(e)        # SymPy >= 1.6 no longer finds the base class's printer for Function subclasses

    pr = P({'precision': 17})
    out, reads, writes, local = [], [], [], []
    roff, woff = {}, {}
    indexed = {}                  # IndexedConstants of the kernel (rkA[stage], ...): name -> values

    def one(e):
        if not hasattr(e, 'lhs'):
            raise UnsupportedByB200('user kernel %s: unsupported equation %r' % (_name(kernel), e))
        for ds in e.rhs.atoms(DataSet):
            n = _strip(ds.base)
            roff.setdefault(n, set()).add(index(ds)[1])
            if n not in reads:
                reads.append(n)
        if type(e.lhs).__name__ == 'DataSet':
            n = _strip(e.lhs.base)
            idx, off = index(e.lhs)
            woff.setdefault(n, set()).add(off)
            if n not in writes:
                writes.append(n)
            out.append([n, True, pr.doprint(e.rhs), idx])
        else:
            local.append(str(e.lhs))
            out.append([str(e.lhs), False, pr.doprint(e.rhs), None])

    def condition(c):
        for ds in c.atoms(DataSet):
            n = _strip(ds.base)
            roff.setdefault(n, set()).add(index(ds)[1])
            if n not in reads:
                reads.append(n)
        return pr.doprint(c)

    for e in kernel.equations:
        if type(e).__name__ == 'GroupedPiecewise':
            # groups of equations under if / else if / else, printed as opsc.py:372-397 does
            for i, (eqs, cond) in enumerate(e.args):
                out.append(['#if' if i == 0 else '#else' if cond == True else '#elif', None, None if (i and cond == True) else condition(cond), None])   # noqa: E712
                for q in (eqs if isinstance(eqs, (list, tuple)) or type(eqs).__name__ == 'Tuple' else [eqs]):
                    one(q)
            out.append(['#end', None, None, None])
        else:
            one(e)
    for ic in getattr(kernel, 'IndexedConstants', []):
        try:
            indexed[str(ic.base.label)] = [float(v) for v in ic.value]
        except Exception:
            raise UnsupportedByB200('kernel %s: indexed constant %s has no values' % (_name(kernel), ic))
    rng = kernel.total_range()
    if stencil:
        # two different points p, p' of the range touch the same element iff p - p' = (read or write offset) - (write offset);
        # that is impossible when the difference is non-zero along a direction in which the range is one point thick
        if thin_direction is None:
            thin = [str(rng[2 * d + 1] - rng[2 * d]) == '1' for d in range(len(rng) // 2)] + [True] * 3
        else:       # ranges given at run time (SplitBC): a boundary kernel covers ONE plane normal to its direction (checked when resolved)
            thin = [d == thin_direction for d in range(len(rng) // 2)] + [True] * 3
        for n, ws in woff.items():
            for w in ws:
                for o in ws | roff.get(n, set()):
                    diff = [a - b for a, b in zip(o, w)]
                    if any(diff) and not any(diff[d] and thin[d] for d in range(3)):
                        raise UnsupportedByB200('kernel %s writes %s where a neighbouring point of its range reads or writes it (offsets %s / %s): '
                                                'it cannot run one thread per point' % (_name(kernel), n, w, o))
    consts = sorted(set(str(s) for e in kernel.equations for s in (e.rhs if hasattr(e, 'rhs') else e).free_symbols
                        if type(s).__name__ == 'ConstantObject'))
    out_k = {'name': _name(kernel), 'when': when, 'range': [r if isinstance(r, str) else ccode(r) for r in rng], 'reads': reads,
             'writes': writes, 'locals': local, 'constants': consts, 'statements': out}
    if thin_direction is not None:
        out_k['one_plane_along'] = thin_direction
    if indexed:
        out_k['indexed_constants'] = indexed
    return out_k


def _cold_kernel(kernel):
    """A cold (one-off or boundary-value) kernel as data: its iteration range (C expressions in block0np{d}) and an
    ordered list of assignments  [lhs name, lhs offset or None for a kernel-local variable, rhs in numpy syntax]."""
    from sympy.printing.c import ccode
    pr = _Printer()
    out = []
    for e in kernel.equations:
        if not hasattr(e, 'lhs'):
            raise UnsupportedByB200('cold kernel %s: unsupported equation %r' % (_name(kernel), e))
        if type(e.lhs).__name__ == 'DataSet':
            lhs, off = _strip(e.lhs.base), [int(i) for i in e.lhs.indices]
        else:
            lhs, off = str(e.lhs), None
        out.append([lhs, off, pr(e.rhs)])
    rng = [ccode(r) for r in kernel.total_range()]
    return {'name': _name(kernel), 'range': rng, 'statements': out}


def _lambdify_check(eq, canonical, names, ntry=8, tol=1e-12):
    """numerically compare eq.rhs with canonical(**values) on random positive inputs."""
    from sympy import lambdify, Symbol
    from opensbli.core.opensbliobjects import DataSet
    rep = {a: Symbol('v_' + _strip(a.base)) for a in eq.rhs.atoms(DataSet)}
    expr = eq.rhs.xreplace(rep)
    # constants may be ConstantObjects or plain Symbols (apps build BC equations with parse_expr): go by name
    rep2 = {a: Symbol('v_' + str(a)) for a in expr.free_symbols if not str(a).startswith('v_')}
    expr = expr.xreplace(rep2)
    syms = sorted(expr.free_symbols, key=str)
    f = lambdify(syms, expr, 'math')
    rng = np.random.default_rng(0)
    for _ in range(ntry):
        vals = {str(s)[2:]: float(0.5 + rng.random()) for s in syms}
        got = f(*[vals[str(s)[2:]] for s in syms])
        try:
            want = canonical(vals)
        except KeyError:          # the canonical form needs a symbol the equation does not have
            return False
        if abs(got - want) > tol * max(1.0, abs(want)):
            return False
    return True


def _check_adiabatic_wall(kernel, direction, side, ndim):
    """adiabatic_wall.py:28-79: the wall-energy equation must be rho_wall T_wall / (gama (gama-1) Minf^2) with
    T_wall = 6/11 (3 T_1 - 3/2 T_2 + 1/3 T_3), T_h the temperature h points above the wall; checked numerically."""
    from opensbli.core.opensbliobjects import DataSet
    import random
    rnd = random.Random(7)
    walls = [e for e in kernel.equations if hasattr(e.lhs, 'base') and _strip(e.lhs.base) == 'rhoE' and not any(int(i) != 0 for i in e.lhs.indices[:ndim])]
    if not walls:
        return False
    ex = walls[0].rhs
    inward = 1 if side == 0 else -1
    for _ in range(4):
        vals, mapping = {}, {}
        for ds in ex.atoms(DataSet):
            key = (_strip(ds.base), int(ds.indices[direction]) * inward)
            if any(int(ds.indices[e]) != 0 for e in range(ndim) if e != direction):
                return False
            mapping[ds] = vals.setdefault(key, rnd.uniform(1.0, 2.0))
        cst = {str(s): rnd.uniform(1.2, 1.6) for s in ex.free_symbols}
        try:
            got = float(ex.xreplace(mapping).subs({s: cst[str(s)] for s in ex.free_symbols}))
            gm, M2 = cst['gama'], cst['Minf'] ** 2
            T = {}
            for h in (1, 2, 3):
                ke = sum(0.5 * vals[('rhou%d' % d, h)] ** 2 for d in range(ndim))
                T[h] = gm * M2 * (gm - 1.0) * (vals[('rhoE', h)] - ke / vals[('rho', h)]) / vals[('rho', h)]
            want = vals[('rho', 0)] * (6.0 / 11.0) * (3.0 * T[1] + T[3] / 3.0 - 1.5 * T[2]) / (gm * (gm - 1.0) * M2)
        except (KeyError, TypeError):
            return False
        if abs(got - want) > 1e-11 * abs(want):
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
        # contravariant velocity of the curvilinear form (euler_wave.py:29)
        canon['U%d' % d] = (lambda d: (lambda v: sum(v['D%d%d' % (d, j)] * v['u%d' % j] for j in range(ndim))))(d)
    info = {'viscosity': {'type': 'constant'}, 'sensor': False}
    for k in kernels:
        lhs_names = [(_strip(e.lhs.base) if hasattr(e.lhs, 'base') else str(e.lhs)) for e in k.equations if hasattr(e, 'lhs')]
        if 'theta' in lhs_names:          # modified Ducros sensor (shock_sensors.py:12-49), recognised by its output array
            from sympy import tanh
            if not any(e.rhs.has(tanh) for e in k.equations if hasattr(e, 'rhs')):
                raise UnsupportedByB200('shock sensor %s is not the modified Ducros sensor' % _name(k))
            info['sensor'] = True
            continue
        for e in k.equations:
            lhs = _strip(e.lhs.base) if hasattr(e.lhs, 'base') else str(e.lhs)
            if lhs == 'mu':
                suth = lambda v: v['T'] ** 1.5 * (1.0 + v['SuthT'] / v['RefT']) / (v['T'] + v['SuthT'] / v['RefT'])
                if _lambdify_check(e, suth, None):
                    info['viscosity'] = {'type': 'sutherland'}
                    continue
                from sympy import Pow
                pw = [a for a in e.rhs.atoms(Pow)]
                if len(pw) == 1 and e.rhs == pw[0] and pw[0].exp.is_number:
                    info['viscosity'] = {'type': 'power', 'exponent': float(pw[0].exp)}
                    continue
                raise UnsupportedByB200('viscosity law mu = %s is not implemented (Sutherland and T**n are)' % e.rhs)
            if lhs not in canon:
                raise UnsupportedByB200("constituent relation for '%s' (kernel %s) is outside the set {u_i, p, a, T, mu, theta} "
                                        "the B200 kernels implement" % (lhs, _name(k)))
            if not _lambdify_check(e, canon[lhs], None):
                raise UnsupportedByB200("constituent relation %s = %s differs from the canonical form" % (lhs, e.rhs))
    return info


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
        info['teno_adaptive'] = 'TENO_CT' in names
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
    info['curvilinear'] = bool(extra)
    if extra:   # metric direction cosines in the eigensystem (euler_eigensystem.py:18-54): the 2-D strong-conservation form
        want = sorted(['D%d0' % d, 'D%d1' % d, 'detJ'])
        if sorted(extra) != want or 'U%d' % d not in used:
            raise UnsupportedByB200('curvilinear metric terms %s in %s: only the 2-D strong-conservation form (D_%d0, D_%d1, detJ, '
                                    'contravariant velocity U%d) is implemented' % (sorted(extra), _name(k), d, d, d))
    info.setdefault('teno_adaptive', False)
    return info


def _check_curvilinear_residual(recon, resid, ndim):
    """Strong-conservation form: Residual_m = - sum_d (F^d_m[i] - F^d_m[i-1]) / Delta_d / detJ, with F^d_m the m-th array
    written by the reconstruction loop of direction d (shock_capturing.py:21-34; euler_wave.py:12-18).  Checked numerically
    with every constant (the 1/Delta factors) set to one."""
    import random
    from opensbli.core.opensbliobjects import DataSet
    rnd = random.Random(99)
    flux = {}
    for k in recon:
        d = _recon_info(k)['direction']
        outs = [_strip(e.lhs.base) for e in k.equations if hasattr(e, 'lhs') and hasattr(e.lhs, 'base')]
        flux[d] = outs
    vals = {}
    for k in resid:
        for e in k.equations:
            m = re.match(r'Residual(\d)$', _strip(e.lhs.base) if hasattr(e.lhs, 'base') else '')
            if not m:
                continue
            eq = int(m.group(1))
            mapping = {}
            for ds in e.rhs.atoms(DataSet):
                key = (_strip(ds.base), tuple(int(i) for i in ds.indices[:ndim]))
                mapping[ds] = vals.setdefault(key, rnd.uniform(1.0, 2.0))
            ex = e.rhs.xreplace(mapping)
            got = float(ex.subs({s: 1.0 for s in ex.free_symbols}))
            zero = (0,) * ndim
            want = 0.0
            try:
                for d in range(ndim):
                    back = tuple(-1 if e_ == d else 0 for e_ in range(ndim))
                    want -= vals[(flux[d][eq], zero)] - vals[(flux[d][eq], back)]
                want /= vals[('detJ', zero)]
            except (KeyError, IndexError):
                raise UnsupportedByB200('curvilinear residual equation of Residual%d does not difference the reconstructed fluxes' % eq)
            if abs(got - want) > 1e-11 * max(1.0, abs(want)):
                raise UnsupportedByB200('curvilinear residual equation of Residual%d is not -(dF/dxi)/detJ' % eq)


# boundary classes routed through the run-time compiled path although a hand-written kernel exists (tests of that path)
_SPLIT_BC = re.compile(r'(\w+) bc direction-(\d) side-(\d) split-(\d+)')
GENERIC_BCS = set(k for k in os.environ.get('OSB_GENERIC_BC', '').split(',') if k)

NATIVE_BCS = ('Dirichlet', 'Extrapolation', 'InletPressureExtrapolate', 'Symmetry', 'AdiabaticWall', 'IsothermalWall', 'ZeroGradientOutlet', 'PressureOutlet')

KNOWN_CONSTANTS = re.compile(r'^(gama|gamma_m1|Minf|Re|Pr|mu|dt|eps|TENO_CT|teno_a1|teno_a2|epsilon|SuthT|RefT|Twall|back_pressure|niter|c[0-2]|'
                             r'block0np\d|Delta\dblock0|inv_\d+|rc\d+|rcinv\d+|inv_rfact\d*_block0)$')


def _check_constants_used(components):
    """Every runtime constant the hot loops read must be one the hand-written kernels account for; anything else
    (body-force constants c_j of the channel apps, user constants of custom terms) means the equations are not the
    canonical system and must not be silently dropped."""
    from opensbli.core.opensbliobjects import ConstantObject
    bad = set()
    for c in components:
        if type(c).__name__ != 'Kernel':
            continue
        for e in c.equations:
            if hasattr(e, 'rhs'):
                for a in e.rhs.atoms(ConstantObject):
                    if not KNOWN_CONSTANTS.match(str(a)):
                        bad.add('%s (in %s)' % (a, _name(c)))
    if bad:
        raise UnsupportedByB200('the hot loops use constants outside the implemented canonical system: %s' % sorted(bad))


def _dirichlet_free(kernel, ndim, q_names):
    """Dirichlet faces that do not impose every conserved variable (transitional_SBLI.py:134-141 leaves the spanwise
    momentum alone and adds its kinetic energy to the imposed energy): which variables are free, and whether the energy
    equation is  E_imposed + 1/2 sum(free momentum^2)/rho  -- checked numerically."""
    import random
    from opensbli.core.opensbliobjects import DataSet
    rnd = random.Random(11)
    centre = {}
    for e in kernel.equations:
        if hasattr(e.lhs, 'base') and not any(int(i) != 0 for i in e.lhs.indices[:ndim]):
            centre[_strip(e.lhs.base)] = e
    free = [m for m, n in enumerate(q_names) if n not in centre]
    if not free:
        return {}
    if any(not (1 <= m <= ndim) for m in free) or 'rhoE' not in centre:
        raise UnsupportedByB200('Dirichlet face that leaves %s free is not implemented' % [q_names[m] for m in free])
    ex = centre['rhoE'].rhs
    dss = [a for a in ex.atoms(DataSet) if _strip(a.base) in q_names]
    if not dss:
        return {'free': free}
    others = {s: rnd.uniform(1.0, 2.0) for s in ex.free_symbols}
    def value(vals):
        return float(ex.xreplace({a: vals[_strip(a.base)] for a in dss}).subs(others))
    for _ in range(3):
        vals = {n: rnd.uniform(1.0, 2.0) for n in q_names}
        zero = dict(vals, **{q_names[m]: 0.0 for m in free})
        ke = 0.5 * sum(vals[q_names[m]] ** 2 for m in free) / vals['rho']
        if any(_strip(a.base) not in [q_names[m] for m in free] + ['rho'] for a in dss) or abs(value(vals) - value(zero) - ke) > 1e-12:
            raise UnsupportedByB200('Dirichlet energy equation depends on the solution in a way the B200 kernels do not implement')
    return {'free': free, 'ke_free': True}


def _mass_source(cr, hot, ndim):
    """Time-periodic mass source of apps/transitional_SBLI (transitional_SBLI.py:77-89): a constituent relation
    S = f(x, constants) * sin(w * iter) that enters the continuity residual as +S.  Returns (plan entry, cold kernel that
    evaluates f, the CR kernel) or (None, None, None)."""
    from sympy import sin, Symbol, diff
    from sympy.printing.c import ccode
    from opensbli.core.opensbliobjects import DataSet
    for k in cr:
        for e in k.equations:
            its = [s for s in e.rhs.free_symbols if str(s) == 'iter'] if hasattr(e, 'rhs') else []
            if not its:
                continue
            it = its[0]
            sname = _strip(e.lhs.base)
            spatial, temporal = e.rhs.as_independent(it, as_Add=False)
            if len(k.equations) != 1 or temporal.func != sin or (temporal.args[0] / it).has(it):
                raise UnsupportedByB200('time-dependent constituent relation %s is not of the form f(x) sin(w iter)' % sname)
            rate = temporal.args[0] / it
            uses = 0
            for h in hot:
                for he in h.equations:
                    if not hasattr(he, 'rhs'):
                        continue
                    ds = [a for a in he.rhs.atoms(DataSet) if _strip(a.base) == sname]
                    if not ds:
                        continue
                    lhs = _strip(he.lhs.base) if hasattr(he.lhs, 'base') else str(he.lhs)
                    if lhs != 'Residual0' or len(ds) != 1 or any(int(i) != 0 for i in ds[0].indices[:ndim]) or diff(he.rhs, ds[0]) != 1:
                        raise UnsupportedByB200('source term %s enters %s in a form the B200 kernels do not implement' % (sname, lhs))
                    uses += 1
            if uses != 1:
                raise UnsupportedByB200('source term %s is not added to the continuity residual exactly once' % sname)
            cold = {'name': 'Mass source amplitude', 'range': [ccode(r) for r in k.total_range()],
                    'statements': [['BF_amp', [0] * ndim, _Printer()(spatial)]]}
            return {'field': 'BF_amp', 'rate': ccode(rate)}, cold, k
    return None, None, None


def _check_forcing(kernels, ndim):
    """Constant body force of the channel apps (turbulent_channel.py:15-16): momentum_i gets -c_i, the energy equation
    -c_j u_j.  Verified on the residual equations: d Residual_i / d c_i = -1 and d Residual_E / d c_j = -u_j."""
    from sympy import diff, simplify, Symbol
    from opensbli.core.opensbliobjects import ConstantObject, DataSet
    found = False
    for k in kernels:
        for e in k.equations:
            cs = [a for a in e.rhs.atoms(ConstantObject) if re.match(r'c[0-2]$', str(a))] if hasattr(e, 'rhs') else []
            if not cs:
                continue
            found = True
            m = re.match(r'Residual(\d)$', _strip(e.lhs.base) if hasattr(e.lhs, 'base') else '')
            if not m:
                raise UnsupportedByB200('body-force constant outside a residual equation (%s)' % _name(k))
            eq = int(m.group(1))
            for cst in cs:
                j = int(str(cst)[1])
                dd = diff(e.rhs, cst)
                if 1 <= eq <= ndim:
                    ok = (eq - 1 == j) and dd == -1
                else:
                    us = [a for a in dd.atoms(DataSet)]
                    ok = eq == ndim + 1 and len(us) == 1 and _strip(us[0].base) == 'u%d' % j and simplify(dd + us[0]) == 0
                if not ok:
                    raise UnsupportedByB200('forcing term of %s in Residual%d is not the canonical constant body force' % (cst, eq))
    return found


def _central_form(kernels, ndim, q_names):
    """Central(4) convective terms: which splitting do the loops evaluate?  The convective loops (in program order: product
    work arrays, first-derivative loops or local evaluations, residual equations; scheme.py:187-271, StoreSome.py:71-161)
    are interpreted numerically on random data:
      * every derivative loop must be the 4th-order central difference of one of the functions the kernels implement
        (q_m, u_a, p, q_m u_d, p u_d, rhoE/rho) along one direction;
      * the residual equations must equal the Blaisdell skew form (parsing.py:75-111; taylor_green_vortex.py:8-11,
        laminar_channel.py:7-9) or the Feiereisen quadratic split (compressible_TCF_Central/turbulent_channel.py:12-20),
        with diagonal metric factors D_dd and the constant body force where present.
    Returns 'blaisdell' or 'feiereisen'."""
    import random
    from sympy import Piecewise
    from opensbli.core.opensbliobjects import DataSet
    rnd = random.Random(20240607)
    W = {-2: 1.0 / 12, -1: -8.0 / 12, 1: 8.0 / 12, 2: -1.0 / 12}
    unames = ['u%d' % d for d in range(ndim)]
    vals, tvals, Dv, cv = {}, {}, {}, {}

    def v(name, s):
        return vals.setdefault((name, s), rnd.uniform(1.0, 2.0))

    def default_branch(ex):
        for _ in range(4):
            pws = list(ex.atoms(Piecewise))
            if not pws:
                break
            ex = ex.xreplace({pw: pw.args[-1][0] for pw in pws})
        return ex

    def numeric(ex, mapping):
        ex = ex.xreplace(mapping)
        return float(ex.subs({s: 1.0 for s in ex.free_symbols}))

    point_defs, terms = {}, {}

    def field_value(name, s):
        """value of dataset `name` at shift s along the current direction: a field, or a pointwise product work array"""
        if name in point_defs:
            ex = point_defs[name]
            return numeric(ex, {ds: field_value(_strip(ds.base), s) for ds in ex.atoms(DataSet)})
        if name in q_names or name in unames or name in ('p', 'T', 'mu'):
            return v(name, s)
        raise UnsupportedByB200('central convective loops read %s, which is outside the canonical system' % name)

    def expected(form, m, forced):
        D = lambda d: Dv.get(d, 1.0)
        T = lambda f, d: tvals[(f, d)]
        r = 0.0
        for d in range(ndim):
            ud, q = v('u%d' % d, 0), v(q_names[m], 0)
            if form == 'blaisdell':
                r -= 0.5 * D(d) * (T('%s*u%d' % (q_names[m], d), d) + ud * T(q_names[m], d) + q * T('u%d' % d, d))
            elif m == 0:
                r -= D(d) * T('rhou%d' % d, d)
            elif m <= ndim:
                r -= 0.5 * D(d) * (T('rhou%d*u%d' % (m - 1, d), d) + v('rhou%d' % d, 0) * T('u%d' % (m - 1), d) + v('u%d' % (m - 1), 0) * T('rhou%d' % d, d))
            else:
                r -= 0.5 * D(d) * (T('rhoE*u%d' % d, d) + v('rhou%d' % d, 0) * T('rhoE/rho', d) + (v('rhoE', 0) / v('rho', 0)) * T('rhou%d' % d, d))
            if m == ndim + 1:
                r -= D(d) * T('p*u%d' % d, d)
        if 1 <= m <= ndim:
            r -= D(m - 1) * T('p', m - 1)
            if forced:
                r -= cv[m - 1]
        if m == ndim + 1 and forced:
            r -= sum(cv[j] * v('u%d' % j, 0) for j in range(ndim))
        return r

    forms = set()
    for k in kernels:
        for e in k.equations:
            if not hasattr(e, 'rhs'):
                continue
            lname = _strip(e.lhs.base) if hasattr(e.lhs, 'base') else str(e.lhs)
            ex = default_branch(e.rhs)
            dss = list(ex.atoms(DataSet))
            m = re.match(r'Residual(\d)$', lname)
            if m:
                # ---- a residual equation: evaluate it on random term / field / metric values
                mapping, forced = {}, False
                for ds in dss:
                    n = _strip(ds.base)
                    if any(int(i) != 0 for i in ds.indices[:ndim]):
                        raise UnsupportedByB200('stencil access inside the convective residual equation of %s' % _name(k))
                    if n in terms:
                        mapping[ds] = tvals.setdefault(terms[n], rnd.uniform(1.0, 2.0))
                    elif re.match(r'D(\d)\1$', n):
                        mapping[ds] = Dv.setdefault(int(n[1]), rnd.uniform(1.0, 2.0))
                    else:
                        mapping[ds] = field_value(n, 0)
                for s in ex.free_symbols:
                    n = str(s)
                    if n in terms:
                        mapping[s] = tvals.setdefault(terms[n], rnd.uniform(1.0, 2.0))
                    elif re.match(r'c[0-2]$', n):
                        mapping[s] = cv.setdefault(int(n[1]), rnd.uniform(1.0, 2.0)); forced = True
                for j in range(ndim):
                    cv.setdefault(j, rnd.uniform(1.0, 2.0))
                got = numeric(ex, mapping)
                ok = []
                for form in ('blaisdell', 'feiereisen'):
                    try:
                        if abs(expected(form, int(m.group(1)), forced) - got) < 1e-11 * max(1.0, abs(got)):
                            ok.append(form)
                    except KeyError:
                        pass
                if not ok:
                    raise UnsupportedByB200('central convective terms of %s are neither the Blaisdell skew form nor the Feiereisen split '
                                            'the B200 kernels implement' % lname)
                forms.add(frozenset(ok))
                continue
            dirs = set(d for ds in dss for d in range(ndim) if int(ds.indices[d]) != 0)
            if not dirs:
                point_defs[lname] = ex                   # pointwise product work array ("Convective terms group d")
                terms.pop(lname, None)
                continue
            if len(dirs) != 1:
                raise UnsupportedByB200('derivative loop %s differentiates along more than one direction' % _name(k))
            d = dirs.pop()
            got = numeric(ex, {ds: field_value(_strip(ds.base), int(ds.indices[d])) for ds in dss})
            cands = {n: (lambda s, n=n: v(n, s)) for n in q_names + unames + ['p', 'T', 'mu']}
            cands.update({'%s*u%d' % (n, d): (lambda s, n=n: v(n, s) * v('u%d' % d, s)) for n in q_names + ['p']})
            cands['rhoE/rho'] = lambda s: v('rhoE', s) / v('rho', s)
            match = [c for c, f in cands.items() if abs(sum(W[s] * f(s) for s in W) - got) < 1e-12]
            if len(match) != 1:
                raise UnsupportedByB200('derivative loop %s (%s) is not a 4th-order central difference of a function the B200 '
                                        'kernels implement' % (_name(k), lname))
            terms[lname] = (match[0], d)
            point_defs.pop(lname, None)
    common = set(['blaisdell', 'feiereisen'])
    for f in forms:
        common &= f
    if not forms or not common:
        raise UnsupportedByB200('central convective residual equations do not agree on one splitting')
    return 'blaisdell' if 'blaisdell' in common else 'feiereisen'


def _check_viscous_form(kernels, ndim):
    """The viscous loops (Derivative evaluation / Viscous CD / Viscous terms / Viscous residual, in program order;
    StoreSome.py:71-161, scheme.py:256-271) are interpreted numerically on random data, like the central convective loops:
    every derivative loop must be a 4th-order first, second or mixed central difference of u_a, T or mu, and what the
    residual equations ADD must equal the Navier-Stokes terms the kernels implement (k_viscous*, osb_kernels.cuh):
        tau_ij = mu/Re (d_j u_i + d_i u_j - 2/3 delta_ij d_k u_k) ,  q_j = mu/((gama-1) Minf^2 Pr Re) d_j T
        momentum_i += d_j tau_ij (- c_i) ,  energy += d_j q_j + d_j (u_i tau_ij) (- c_j u_j)
    with diagonal metrics  d_j f = D_jj delta_j f ,  d_jj f = D_jj^2 delta_jj f + D_jj SD_jjj delta_j f."""
    import random
    from sympy import Piecewise
    from opensbli.core.opensbliobjects import DataSet
    rnd = random.Random(4711)
    W1 = {-2: 1.0 / 12, -1: -8.0 / 12, 1: 8.0 / 12, 2: -1.0 / 12}
    W2 = {-2: -1.0 / 12, -1: 16.0 / 12, 0: -30.0 / 12, 1: 16.0 / 12, 2: -1.0 / 12}
    fields = ['u0', 'u1', 'u2', 'T', 'mu']
    vals, Dv, SDv, cst = {}, {}, {}, {}

    def v(key, s):
        return vals.setdefault((key, s), rnd.uniform(1.0, 2.0))

    def default_branch(ex):
        for _ in range(4):
            pws = list(ex.atoms(Piecewise))
            if not pws:
                break
            ex = ex.xreplace({pw: pw.args[-1][0] for pw in pws})
        return ex

    def constant(name):
        return cst.setdefault(name, rnd.uniform(1.2, 1.8))

    def numeric(ex, mapping, derivative):
        ex = ex.xreplace(mapping)
        # derivative loops: strip the 1/Delta factors; residual equations: physical constants get random values
        return float(ex.subs({s: (1.0 if derivative else constant(str(s))) for s in ex.free_symbols}))

    terms, checked = {}, 0
    for k in kernels:
        for e in k.equations:
            if not hasattr(e, 'rhs'):
                continue
            lname = _strip(e.lhs.base) if hasattr(e.lhs, 'base') else str(e.lhs)
            ex = default_branch(e.rhs)
            dss = list(ex.atoms(DataSet))
            m = re.match(r'Residual(\d)$', lname)
            if m:
                eq = int(m.group(1))
                mapping, r0 = {}, rnd.uniform(1.0, 2.0)
                for ds in dss:
                    n = _strip(ds.base)
                    if any(int(i) != 0 for i in ds.indices[:ndim]):
                        raise UnsupportedByB200('stencil access inside the viscous residual equation of %s' % _name(k))
                    if n == lname:
                        mapping[ds] = r0
                    elif n in terms:
                        mapping[ds] = v(terms[n], 0)
                    elif re.match(r'D(\d)\1$', n):
                        mapping[ds] = Dv.setdefault(int(n[1]), rnd.uniform(1.0, 2.0))
                    elif re.match(r'SD(\d)\1\1$', n):
                        mapping[ds] = SDv.setdefault(int(n[2]), rnd.uniform(1.0, 2.0))
                    elif n in fields:
                        mapping[ds] = v(n, 0)
                    else:
                        raise UnsupportedByB200('viscous residual equation reads %s, which the B200 kernels do not implement' % n)
                for s in ex.free_symbols:
                    if str(s) in terms:
                        mapping[s] = v(terms[str(s)], 0)
                got = numeric(ex, mapping, False) - r0
                # ---- the terms the kernels implement, from the same random values
                has_mu_field = any(_strip(ds.base) == 'mu' for kk in kernels for ee in kk.equations if hasattr(ee, 'rhs') for ds in ee.rhs.atoms(DataSet))
                mu = v('mu', 0) if has_mu_field else (cst['mu'] if 'mu' in cst else 1.0)
                Re, gama, Minf, Pr = (constant(n) for n in ('Re', 'gama', 'Minf', 'Pr'))
                D = lambda d: Dv.get(d, 1.0)
                SD = lambda d: SDv.get(d, 0.0)
                # a derivative the loops do not evaluate can only be one whose coefficient vanishes (e.g. d T without metrics and
                # with constant viscosity); if it does not, the comparison below fails
                T = lambda *key: v(key, 0) if key in terms.values() else 0.0
                try:
                    du = [[D(b) * T('d1', 'u%d' % a, b) for b in range(ndim)] for a in range(ndim)]
                    dT = [D(b) * T('d1', 'T', b) for b in range(ndim)]
                    dmu = [D(b) * T('d1', 'mu', b) for b in range(ndim)]
                    div = sum(du[a][a] for a in range(ndim))
                    lap = lambda f, b: D(b) ** 2 * T('d2', f, b) + D(b) * SD(b) * T('d1', f, b)
                    S = lambda a, b: du[a][b] + du[b][a] - ((2.0 / 3.0) * div if a == b else 0.0)
                    vis = []
                    for a in range(ndim):
                        s1 = sum(dmu[b] * S(a, b) for b in range(ndim))
                        s2 = 0.0
                        for b in range(ndim):
                            if a == b:
                                s2 += (4.0 / 3.0) * lap('u%d' % a, a)
                            else:
                                s2 += lap('u%d' % a, b) + (1.0 / 3.0) * D(a) * D(b) * T('mix', 'u%d' % b, min(a, b), max(a, b))
                        vis.append((s1 + mu * s2) / Re)
                    forced = [cst.get('c%d' % a) for a in range(ndim)]
                    if 1 <= eq <= ndim:
                        want = vis[eq - 1] - (forced[eq - 1] or 0.0)
                    elif eq == ndim + 1:
                        kq = 1.0 / ((gama - 1.0) * Minf ** 2 * Pr * Re)
                        want = kq * (sum(dmu[d] * dT[d] for d in range(ndim)) + mu * sum(lap('T', d) for d in range(ndim)))
                        want += sum(vis[a] * v('u%d' % a, 0) for a in range(ndim))
                        want += (mu / Re) * sum(S(a, b) * du[a][b] for a in range(ndim) for b in range(ndim))
                        want -= sum((forced[a] or 0.0) * v('u%d' % a, 0) for a in range(ndim))
                    else:
                        want = 0.0
                except KeyError as err:
                    raise UnsupportedByB200('viscous terms: derivative %s needed by the implemented Navier-Stokes terms is not evaluated by the loops' % (err.args[0],))
                if abs(got - want) > 1e-10 * max(1.0, abs(want)):
                    raise UnsupportedByB200('viscous terms added to %s differ from the Navier-Stokes terms the B200 kernels implement '
                                            '(loops give %.12g, kernels %.12g on the same random data)' % (lname, got, want))
                checked += 1
                continue
            dirs = set(d for ds in dss for d in range(ndim) if int(ds.indices[d]) != 0)
            if len(dirs) != 1:
                raise UnsupportedByB200('viscous derivative loop %s does not differentiate along exactly one direction' % _name(k))
            d = dirs.pop()
            mapping = {}
            for ds in dss:
                n, s = _strip(ds.base), int(ds.indices[d])
                if n in terms:
                    mapping[ds] = v(terms[n], s)
                elif n in fields:
                    mapping[ds] = v(n, s)
                else:
                    raise UnsupportedByB200('viscous derivative loop %s reads %s' % (_name(k), n))
            got = numeric(ex, mapping, True)
            cands = {}
            for f in fields:
                cands[('d1', f, d)] = sum(W1[s] * v(f, s) for s in W1)
                cands[('d2', f, d)] = sum(W2[s] * v(f, s) for s in W2)
            for key in set(terms.values()):
                if key[0] == 'd1' and key[2] != d:
                    cands[('mix', key[1], min(key[2], d), max(key[2], d))] = sum(W1[s] * v(key, s) for s in W1)
            match = [key for key, val in cands.items() if abs(val - got) < 1e-11]
            if len(match) != 1:
                raise UnsupportedByB200('viscous derivative loop %s (%s) is not a 4th-order central difference the B200 kernels implement' % (_name(k), lname))
            terms[lname] = match[0]
    if not checked:
        raise UnsupportedByB200('no viscous residual equations found in the viscous loops')


def _metric_directions(kernels, ndim):
    """Stretched directions from the metric arrays the residual / viscous loops read: only diagonal metrics D_dd
    (+ SD_ddd) are implemented, i.e. grids stretched along their own coordinate (metric.py:137-147)."""
    used = _datasets_used(kernels)
    out = [None] * ndim
    for u in used:
        m = re.match(r'S?D(\d)(\d)(\d?)$', u)
        if m:
            idx = [int(x) for x in m.groups() if x != '']
            if len(set(idx)) != 1:
                raise UnsupportedByB200('off-diagonal metric term %s: general curvilinear grids are not implemented yet' % u)
            out[idx[0]] = 'D%d%d' % (idx[0], idx[0])
        elif u == 'detJ':
            raise UnsupportedByB200('detJ in the hot loops (strong-conservation curvilinear form) is not implemented yet')
    return out


def _closure_tables(kernels, ndim):
    """One-sided closure rows from the idx-conditional formulas of the first-derivative loops
    (opensblifunctions.py:523-534).  Returns ({(dir, side): nrows}, d1 table) with the table read off the equations:
    row r = weights of the boundary-absolute points 0..np-1."""
    from sympy import Piecewise, Eq
    from opensbli.core.opensbliobjects import DataSet, Grididx
    faces, table = {}, None
    for k in kernels:
        for e in k.equations:
            for pw in e.rhs.atoms(Piecewise):
                rows0 = {}
                for expr, cond in pw.args:
                    if cond is True or cond == True or not isinstance(cond, Eq):  # noqa: E712
                        continue
                    idxs = list(cond.lhs.atoms(Grididx))
                    if len(idxs) != 1:
                        continue
                    d = int(idxs[0].number)
                    # Eq(idx, r) | Eq(np - idx, r + 1) | Eq(idx, np - 1 - r), whichever way SymPy canonicalised it
                    diff = cond.lhs - cond.rhs
                    at = -diff.subs(idxs[0], 0) / diff.coeff(idxs[0])          # the grid index the row applies to
                    if at.is_number:
                        side, row = 0, int(at)
                    else:
                        nps = [s for s in at.free_symbols]
                        if len(nps) != 1 or not (nps[0] - at).is_number:
                            continue
                        side, row = 1, int(nps[0] - at) - 1
                    faces[(d, side)] = max(faces.get((d, side), 0), row + 1)
                    if side == 0:
                        rows0[row] = (expr, d)
                first = not re.search(r' CD .*x\w*\d .*x\w*\d', _name(k))       # not a second / mixed derivative loop
                if first and rows0 and table is None and len(e.rhs.atoms(DataSet)) and all(len(x.atoms(DataSet)) >= 4 for x, _ in rows0.values()):
                    # only first-derivative formulas of a plain dataset are used to read the weights
                    bases = set(ds.base for x, _ in rows0.values() for ds in x.atoms(DataSet))
                    if len(bases) != 1:
                        continue
                    tab = []
                    for r in sorted(rows0):
                        expr, d = rows0[r]
                        w = {}
                        ex = expr.expand()
                        for ds in ex.atoms(DataSet):
                            cf = ex.coeff(ds)
                            syms = [a for a in cf.free_symbols]
                            cf = cf.subs({a: 1 for a in syms})      # strip the 1/Delta factor
                            w[int(ds.indices[d]) + r] = float(cf)
                        tab.append([w.get(p, 0.0) for p in range(max(w) + 1)])
                    n = max(len(t) for t in tab)
                    table = [t + [0.0] * (n - len(t)) for t in tab]
    return faces, table


def _const_value(c):
    from sympy.printing.c import ccode
    v = c.value
    if isinstance(v, str):
        return v
    return ccode(v)


class _NotForGenericPath(UnsupportedByB200):
    """Raised for programs neither path can run (precision, multi-block, input files): not retried on the generic path."""


def extract_plan(algorithm):
    """Distil the plan from an OpenSBLI algorithm object (see module docstring).  Programs the hand-written kernels do not cover
    (other scheme orders, fully curvilinear 3-D / viscous terms, loops of classes unknown here ...) fall back to the GENERIC
    path: every loop of the time step is printed from its equations as CUDA C and compiled at run time, launched in program
    order -- one thread per point, the reference's own arithmetic, at the speed of a plain loop-per-kernel code."""
    import os
    if os.environ.get('OSB_FORCE_GENERIC_PATH'):          # testing / measuring: the same program through the printed kernels
        _extract_specialised_checks_only(algorithm)
        return _extract_generic(algorithm, 'forced by OSB_FORCE_GENERIC_PATH')
    try:
        return _extract_specialised(algorithm)
    except _NotForGenericPath:
        raise
    except UnsupportedByB200 as e:
        if os.environ.get('OSB_NO_GENERIC_PATH'):
            raise
        try:
            plan = _extract_generic(algorithm, str(e))
        except UnsupportedByB200 as e2:
            raise UnsupportedByB200('%s -- and the generic path cannot run the program either: %s' % (e, e2))
        print('B200: %s' % e)
        print('B200: the program runs on the GENERIC path (every loop compiled at run time from its equations); expect the speed of '
              'a loop-per-kernel code, not of the hand-written kernels')
        return plan


def _extract_specialised_checks_only(algorithm):
    """the refusals that hold for both paths (multi-block, precision)"""
    try:
        _extract_specialised(algorithm)
    except _NotForGenericPath:
        raise
    except UnsupportedByB200:
        pass


def _generic_exchange(c, when):
    """A periodic self-exchange (exchange.py:9-57) as a printed copy kernel: over the destination box, a[X] = a[X + from - to]"""
    from sympy.printing.c import ccode
    arrays = [_strip(a) for a in c.transfer_arrays]
    size, frm, to = [ccode(v) for v in c.transfer_size], [ccode(v) for v in c.transfer_from], [ccode(v) for v in c.transfer_to]
    nd = len(size)
    shift = ' + '.join('((%s) - (%s))%s' % (frm[d], to[d], ('', '*s1', '*s2')[d]) for d in range(nd))
    rng = []
    for d in range(nd):
        rng += [to[d], '(%s) + (%s)' % (to[d], size[d])]
    return {'name': 'exchange %s %s' % (getattr(c, 'direction', ''), getattr(c, 'side', '')), 'when': when, 'range': rng, 'reads': list(arrays),
            'writes': list(arrays), 'locals': [], 'constants': [], 'statements': [[a, True, '%s[X + %s]' % (a, shift), 'X'] for a in arrays]}


def _extract_generic(algorithm, reason):
    from opensbli.core.kernel import ConstantsToDeclare
    ndim = algorithm.block_descriptions[0].ndim
    flat = []
    _walk(algorithm.prg.components, flat)
    plan = {'ndim': ndim, 'generic': {'reason': reason}, 'conv': 'generic', 'order': 0, 'viscous': False, 'averaging': 'roe', 'weno_formulation': 'JS'}
    q_names = ['rho'] + ['rhou%d' % d for d in range(ndim)] + ['rhoE']
    cold, user, io_specs, monitor = [], [], [], None
    seen_loop = seen_stage = False
    nstages = 0
    halo_depths = {}
    not_executed = set()
    for path, c in flat:
        loops = [type(p).__name__ for p in path]
        t, n = type(c).__name__, _name(c)
        nloops = loops.count('DoLoop')
        if t == 'iohdf5':
            spec = {'arrays': [_strip(a) for a in c.arrays], 'iotype': c.kwargs.get('iotype', 'write'), 'name': c.kwargs.get('name'),
                    'filename': c.kwargs.get('filename')}
            if spec['iotype'] == 'read':
                raise UnsupportedByB200('initial data read from an HDF5 file: pass the file to the runner with --restart instead')
            cond = [p_ for p_ in path if type(p_).__name__ == 'Condition']
            if cond:
                m = re.search(r'Mod\([^,]+,\s*(\d+)\)', str(cond[-1].condition))
                if not m:
                    raise UnsupportedByB200('output condition %s' % cond[-1].condition)
                spec.update(when='in_loop', every=int(m.group(1)))
            else:
                spec['when'] = 'after' if (seen_loop or 'DoLoop' in loops) else 'before'
            io_specs.append(spec)
            seen_loop = seen_loop or 'Timers' in loops or nloops > 0
            continue
        seen_loop = seen_loop or 'Timers' in loops or nloops > 0
        if t == 'SimulationMonitor':
            monitor = {'arrays': [_strip(m.flow_var) for m in c.monitors],
                       'probes': [[str(x) for x in (m.probe_loc if isinstance(m.probe_loc, (tuple, list)) else (m.probe_loc,))] for m in c.monitors],
                       'frequency': int(c.frequency), 'precision': int(c.fp_precision), 'output_file': c.output_file}
            continue
        if t not in ('Kernel', 'ExchangeSelf'):
            if t not in ('DoLoop', 'Timers', 'Condition'):
                not_executed.add(t)
            continue
        if nloops == 0 and 'Timers' not in loops:
            if seen_loop:                                           # after the time loop
                if t != 'Kernel':
                    raise UnsupportedByB200('exchange after the time loop')
                user.append(_user_kernel(c, 'after_loop', stencil=True))
            elif t == 'ExchangeSelf':
                from sympy.printing.c import ccode
                cold.append({'name': 'exchange', 'exchange': True, 'arrays': [_strip(a) for a in c.transfer_arrays],
                             'size': [ccode(v) for v in c.transfer_size], 'from': [ccode(v) for v in c.transfer_from],
                             'to': [ccode(v) for v in c.transfer_to]})
            else:
                cold.append(_cold_kernel(c))
            continue
        if nloops == 2:
            seen_stage = True
            stage_loop = [p_ for p_ in path if type(p_).__name__ == 'DoLoop'][-1]
            nstages = int(stage_loop.loop.upper) - int(stage_loop.loop.lower) + 1
            when = 'stage'
        elif nloops == 1:
            when = 'iteration_end' if seen_stage else 'iteration_start'
        else:
            raise UnsupportedByB200('loop nest deeper than iteration / stage: %s' % n)
        if t == 'ExchangeSelf':
            side = {'left': 0, 'right': 1}.get(c.side, c.side)
            halo_depths.setdefault(int(side), set()).add(int(c.transfer_size[int(c.direction)]))
            user.append(_generic_exchange(c, when))
        else:
            user.append(_user_kernel(c, when, stencil=True))
    if not nstages:
        raise UnsupportedByB200('no Runge-Kutta stage loop found in the program')
    plan['generic']['nstages'] = nstages
    plan.update(rk='ls', rk_a=[0.0] * nstages, rk_b=[0.0] * nstages)          # placeholders: the printed RK kernels carry their own coefficients
    for k in user:
        ic = k.get('indexed_constants', {})
        if 'rkold' in ic:
            plan['rk'] = 'sbli'
    # halos the boundary kernels / exchanges fill (the specialised driver code is not used; the depth only sizes nothing here but is
    # kept for the runner's dataset files)
    depth = (3, 4)
    if halo_depths and all(len(v) == 1 for v in halo_depths.values()) and set(halo_depths) == {0, 1}:
        depth = (min(5, max(2, next(iter(halo_depths[0])))), min(5, max(2, next(iter(halo_depths[1])))))
    plan['halos'] = list(depth)
    plan['bc'] = [[{'type': 'open'}, {'type': 'open'}] for _ in range(ndim)]   # the boundary kernels are in the kernel lists
    plan['cold'] = cold
    plan['user_kernels'] = user
    plan['io'] = io_specs
    if monitor:
        plan['monitor'] = monitor
    plan['q_names'] = q_names
    plan['not_executed'] = sorted(not_executed)
    plan['constant_decls'] = []
    for c in ConstantsToDeclare.constants:
        if type(c).__name__ == 'ConstantObject':
            plan['constant_decls'].append([str(c), 'int' if 'int' in str(c.datatype.opsc()).lower() else 'double', _const_value(c)])
    arrays = [[str(c.base), 'int', 2 * ndim] for c in ConstantsToDeclare.constants
              if type(c).__name__ == 'ConstantIndexed' and str(c.base).startswith(('split_range_', 'split_halo_range_'))]
    if arrays:
        plan['constant_array_decls'] = arrays
    return plan


def _extract_specialised(algorithm):
    from opensbli.core.kernel import ConstantsToDeclare
    if getattr(algorithm, 'MultiBlock', False) or len(algorithm.block_descriptions) != 1:
        raise _NotForGenericPath('multi-block algorithms are not implemented')
    # the kernels are fp64 only (the reference's `SimulationDataType.set_datatype(Double)`, datatypes.py:2-30)
    try:
        from opensbli.core.datatypes import SimulationDataType
        ctype = SimulationDataType.opsc()
    except Exception:
        ctype = 'double'
    dt = algorithm.dtype if isinstance(algorithm.dtype, str) else getattr(algorithm.dtype, 'opsc', lambda: 'double')()
    if ctype != 'double' or str(dt).lower() != 'double':
        raise _NotForGenericPath('datatype %s/%s: the B200 back end computes in double precision only' % (ctype, dt))
    ndim = algorithm.block_descriptions[0].ndim
    flat = []
    _walk(algorithm.prg.components, flat)
    before, in_iter, in_stage, after = [], [], [], []
    seen_loop = False
    io_specs = []
    for path, c in flat:
        loops = [type(p).__name__ for p in path]
        if type(c).__name__ == 'iohdf5':
            # dataset files (core/io_hdf5.py:18-127): written after the time loop, every N iterations inside it
            # (`iohdf5(save_every=N)`: Condition((iter + 1) %% N == 0), algorithm.py:456-463), or read before it
            spec = {'arrays': [_strip(a) for a in c.arrays], 'iotype': c.kwargs.get('iotype', 'write'), 'name': c.kwargs.get('name'),
                    'filename': c.kwargs.get('filename')}
            cond = [p for p in path if type(p).__name__ == 'Condition']
            if cond:
                m = re.search(r'Mod\([^,]+,\s*(\d+)\)', str(cond[-1].condition))
                if not m:
                    raise UnsupportedByB200('output condition %s' % cond[-1].condition)
                spec.update(when='in_loop', every=int(m.group(1)))
            else:
                spec['when'] = 'before' if spec['iotype'] == 'read' else ('after' if (seen_loop or 'DoLoop' in loops) else 'before')
            io_specs.append(spec)
            seen_loop = seen_loop or 'Timers' in loops or loops.count('DoLoop') > 0
            continue
        nloops = loops.count('DoLoop')
        seen_loop = seen_loop or 'Timers' in loops or nloops > 0
        if 'Condition' in loops and type(c).__name__ in ('Kernel', 'ExchangeSelf'):
            # InTheSimulation(frequency=N) on a loop (algorithm.py:456-463): only dataset files are honoured every N iterations
            raise _NotForGenericPath('loop %s runs under a condition (every N iterations): conditional loops are not implemented' % _name(c))
        if nloops == 0 and 'Timers' not in loops:
            (after if seen_loop else before).append(c)          # top level: before / after the timed time loop
        elif nloops == 1:
            in_iter.append(c)
        elif nloops == 2:
            in_stage.append(c)
        else:
            after.append(c)
    plan = {'ndim': ndim, 'viscous': False, 'averaging': 'roe', 'weno_formulation': 'JS'}
    monitor = None
    for c in in_iter:
        if type(c).__name__ == 'SimulationMonitor':
            # probes of dataset values printed every `frequency` iterations (simulation_monitors.py:17-160, algorithm.py:433-437)
            monitor = {'arrays': [_strip(m.flow_var) for m in c.monitors],
                       'probes': [[str(x) for x in (m.probe_loc if isinstance(m.probe_loc, (tuple, list)) else (m.probe_loc,))] for m in c.monitors],
                       'frequency': int(c.frequency), 'precision': int(c.fp_precision), 'output_file': c.output_file}
    # loops of UserDefinedEquations (utilities/user_defined_kernels.py): `User kernel: <name>` and, when their equations hold
    # derivatives, the `UserDefinedEquations CD ...` / `UserDefinedEquations evaluation` loops that evaluate those first
    # (statistics, SFD, the WENO filter's sensor): compiled at run time, launched in program order at the end of the iteration
    user = []
    is_user = lambda c: type(c).__name__ == 'Kernel' and _name(c).startswith(('User kernel', 'UserDefinedEquations'))
    for c in in_iter:
        if is_user(c):
            user.append(_user_kernel(c, 'iteration_end', stencil=True))
        elif type(c).__name__ == 'Kernel' and not (' boundary dir' in _name(c) or _SPLIT_BC.match(_name(c)) or _name(c) == 'Save equations'):
            raise UnsupportedByB200('loop outside the accelerated hot path at iteration level: %s' % _name(c))
    for c in after:
        if is_user(c):
            user.append(_user_kernel(c, 'after_loop', stencil=True))
        elif type(c).__name__ == 'Kernel':
            raise UnsupportedByB200('loop after the time loop is not implemented: %s' % _name(c))
    # components of the program that are not part of the per-step hot path (file output, monitors, timers): not executed
    # by the B200 run-time; listed in the plan and printed so that nothing is dropped silently
    plan['not_executed'] = sorted(set(type(c).__name__ for c in in_iter + after + before
                                      if type(c).__name__ not in ('Kernel', 'ExchangeSelf', 'DoLoop', 'Timers', 'SimulationMonitor', 'Condition')))
    if any(sp_['iotype'] == 'read' for sp_ in io_specs):
        raise _NotForGenericPath('initial data read from an HDF5 file by ops_decl_dat_hdf5 (iohdf5(iotype="read")): pass the file to the '
                                 'runner with --restart instead')
    plan['io'] = io_specs
    q_names = ['rho'] + ['rhou%d' % d for d in range(ndim)] + ['rhoE']

    # ---- stage loop: classify every kernel
    cr, recon, resid, central_conv, viscous, rk_kernels, stage_bcs, unknown = [], [], [], [], [], [], [], []
    for c in in_stage:
        t, n = type(c).__name__, _name(c)
        if t == 'ExchangeSelf' or ' boundary dir' in n or _SPLIT_BC.match(n or ''):
            stage_bcs.append(c)
        elif t != 'Kernel':
            unknown.append(c)
        elif n.startswith('CR') or n == 'ConstituentRelations evaluation':
            cr.append(c)
        elif re.match(r'LLF(Teno|Weno)_reconstruction_\d_direction', n):
            recon.append(c)
        elif re.match(r'LLF(Teno|Weno) Residual', n):
            resid.append(c)
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
    mass_source, source_cold, source_kernel = _mass_source(cr, [c for c in in_stage if type(c).__name__ == 'Kernel' and c not in cr], ndim)
    if source_kernel is not None:
        cr.remove(source_kernel)
    def _generic_bc(c):           # boundary kernels that will be compiled at run time carry their own constants
        if _SPLIT_BC.match(_name(c) or '') or (_name(c) or '').startswith(('User kernel', 'UserDefinedEquations')):
            return True
        m = re.match(r'(\w+) boundary dir(\d) side(\d)', _name(c) or '')
        return bool(m) and (m.group(1) not in NATIVE_BCS or m.group(1) in GENERIC_BCS)
    _check_constants_used([c for c in in_stage + in_iter if c is not source_kernel and not _generic_bc(c)])
    crinfo = _check_constituent(cr, ndim)
    plan['viscosity'] = crinfo['viscosity']
    if recon and central_conv:
        raise UnsupportedByB200('mixed shock-capturing and central convective terms are not implemented')
    if recon:
        infos = [_recon_info(k) for k in recon]
        if sorted(i['direction'] for i in infos) != list(range(ndim)):
            raise UnsupportedByB200('reconstruction kernels do not cover every direction once')
        for key in ('conv', 'order', 'weno_formulation', 'averaging', 'teno_adaptive', 'curvilinear'):
            if len(set(i[key] for i in infos)) != 1:
                raise UnsupportedByB200('direction-dependent %s is not implemented' % key)
            plan[key] = infos[0][key]
        if plan['teno_adaptive'] and not crinfo['sensor']:
            raise UnsupportedByB200('adaptive TENO without the Ducros sensor relation is not implemented')
    elif central_conv:
        plan.update(conv='central', order=4, teno_adaptive=False)
        plan['central_form'] = _central_form([k for k in in_stage if k in central_conv or _name(k).startswith('Derivative evaluation')], ndim, q_names)
    else:
        raise UnsupportedByB200('no convective discretisation found in the stage loop')
    plan['viscous'] = bool(viscous)
    plan['curvilinear'] = bool(plan.get('curvilinear'))
    if plan['curvilinear']:
        if ndim != 2 or viscous or plan.get('teno_adaptive'):
            raise UnsupportedByB200('curvilinear grids are implemented for 2-D inviscid shock-capturing schemes only')
        _check_curvilinear_residual(recon, resid, ndim)
    if viscous:
        _check_viscous_form([k for k in in_stage if k in viscous], ndim)
    plan['forcing'] = _check_forcing(resid + viscous + central_conv, ndim)
    if plan['viscosity']['type'] != 'constant' and not viscous:
        plan['viscosity'] = {'type': 'constant'}
    plan['metric_fields'] = [None] * ndim if plan['curvilinear'] else _metric_directions(resid + viscous + central_conv + [k for k in cr if _name(k) == 'ConstituentRelations evaluation'], ndim)
    faces, d1tab = _closure_tables([k for k in viscous + cr + central_conv if _name(k).startswith(('Derivative evaluation', 'Viscous CD', 'Convective CD'))
                                    or _name(k) == 'ConstituentRelations evaluation'], ndim)
    closure_name = None
    if faces:
        nrows = set(faces.values())
        if len(nrows) != 1 or d1tab is None:
            raise UnsupportedByB200('could not read the one-sided derivative closure from the derivative loops')
        closure_name = {2: 'reduced_access', 4: 'carpenter'}.get(nrows.pop(), 'custom')
        # second-derivative rows are the same in both of the reference's schemes (reduced_access_scheme.py:76-83,
        # Carpenter_scheme.py:69-76); first-derivative rows were read off the equations
        plan['closures'] = {closure_name: {'d1': d1tab, 'd2': [[35.0 / 12, -104.0 / 12, 114.0 / 12, -56.0 / 12, 11.0 / 12],
                                                                 [11.0 / 12, -20.0 / 12, 6.0 / 12, 4.0 / 12, -1.0 / 12]]}}

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
    halo_depths = {}
    for c in [c for c in in_iter if type(c).__name__ == 'ExchangeSelf' or ' boundary dir' in _name(c) or _SPLIT_BC.match(_name(c) or '')]:
        ms = _SPLIT_BC.match(_name(c) or '')
        if ms:
            # SplitBC (bc_core.py:200-217): several boundary classes share a face, each over its own part of the plane.  The
            # parts are run-time integer arrays (split_range_<d><s><n> + split_halo_range_<d><s><n>, bc_core.py:110-127) that
            # the user fills in; every part becomes a run-time compiled kernel on the face, applied in the order given
            d, sd = int(ms.group(2)), int(ms.group(3))
            user.append(_user_kernel(c, 'bc_%d_%d' % (d, sd), stencil=True, thin_direction=d))
            entry = bc[d][sd] or {'type': 'generic', 'class': 'Split', 'parts': []}
            entry['parts'].append(ms.group(1))
            if (d, sd) in faces:
                entry['closure'] = closure_name
            bc[d][sd] = entry
            continue
        if type(c).__name__ == 'ExchangeSelf':
            arrays = [_strip(a) for a in c.transfer_arrays]
            if arrays != q_names:
                raise UnsupportedByB200('periodic exchange of %s (expected the conserved arrays)' % arrays)
            side = {'left': 0, 'right': 1}.get(c.side, c.side)
            bc[int(c.direction)][int(side)] = {'type': 'periodic'}
            halo_depths.setdefault(int(side), set()).add(int(c.transfer_size[int(c.direction)]))   # periodic.py:42-56: hm planes go up, hp down
            continue
        m = re.match(r'(\w+) boundary dir(\d) side(\d)', _name(c))
        kind, d, sd = m.group(1), int(m.group(2)), int(m.group(3))
        entry = None
        if kind in GENERIC_BCS:
            pass
        elif kind == 'Dirichlet':
            # imposed state = whatever the BC equations evaluate to on the face (constants or functions of the position)
            entry = {'type': 'dirichlet_field', 'kernel': _cold_kernel(c)}
            entry.update(_dirichlet_free(c, ndim, q_names))
        elif kind == 'Extrapolation':
            # order 0 copies one interior value into boundary + halos; order 1 extrapolates linearly (extrapolation.py:37-55)
            lin = any(e.rhs.is_Add for e in c.equations if hasattr(e, 'rhs'))
            entry = {'type': 'extrapolation', 'order': 1 if lin else 0}
        elif kind == 'InletPressureExtrapolate':
            entry = {'type': 'inlet_pressure_extrapolate'}
        elif kind == 'Symmetry':
            # SymmetryBC (symmetry.py:23-50) and InviscidWallBC (inviscid_wall.py:24-52) both name their kernel 'Symmetry'; the
            # inviscid wall additionally assigns the boundary point itself (state one point inside, normal momentum removed)
            writes_boundary = any(hasattr(e.lhs, 'indices') and int(e.lhs.indices[d]) == 0 for e in c.equations if hasattr(e, 'lhs'))
            if writes_boundary and (plan.get('curvilinear') or any(plan.get('metric_fields') or [])):
                raise UnsupportedByB200('inviscid wall on a stretched / curvilinear block (metric-dependent normal) is not implemented')
            entry = {'type': 'inviscid_wall' if writes_boundary else 'symmetry'}
        elif kind == 'ZeroGradientOutlet':
            entry = {'type': 'zero_gradient_outlet'}
        elif kind == 'PressureOutlet':
            entry = {'type': 'pressure_outlet'}
        elif kind == 'AdiabaticWall':
            if not _check_adiabatic_wall(c, d, sd, ndim):
                raise UnsupportedByB200('adiabatic wall with a non-canonical wall-energy equation is not implemented')
            entry = {'type': 'adiabatic_wall'}
        elif kind == 'IsothermalWall':
            # the wall-energy equation must be the canonical rhoE = rho Twall/(gama (gama-1) Minf^2) (isothermal_wall.py:40-45)
            walls = [e for e in c.equations if hasattr(e.lhs, 'base') and _strip(e.lhs.base) == 'rhoE' and not any(e.lhs.indices)]
            canon = lambda v: v['rho'] * v['Twall'] / (v['gama'] * (v['gama'] - 1.0) * v['Minf'] ** 2)
            if not walls or not _lambdify_check(walls[0], canon, None):
                raise UnsupportedByB200('isothermal wall with a non-canonical wall-energy equation is not implemented')
            entry = {'type': 'isothermal_wall'}
        if entry is None or kind in GENERIC_BCS:
            # no hand-written kernel for this boundary class (ForcingStripWall, InletLawal, InletTransfer, InviscidWall2D, ...
            # or forced through OSB_GENERIC_BC for testing): its equations are printed as CUDA C and compiled at run time --
            # boundary kernels touch a plane of points per application, nothing to gain from hand-writing them
            entry = {'type': 'generic', 'class': kind}
            user.append(_user_kernel(c, 'bc_%d_%d' % (d, sd), stencil=True))
        if (d, sd) in faces:
            entry['closure'] = closure_name
        bc[d][sd] = entry
    if any(b is None for pair in bc for b in pair):
        raise UnsupportedByB200('a block face has no recognised boundary condition')
    plan['bc'] = bc
    # depth of the halos the boundary conditions fill: the scheme's own (2/2 central, 3/4 WENO/TENO) unless the block has further
    # consumers (block.shock_filter: a WENO filter on a central scheme makes the exchanges 3/4 deep)
    if halo_depths:
        if any(len(v) != 1 for v in halo_depths.values()) or set(halo_depths) != {0, 1}:
            raise UnsupportedByB200('periodic exchanges of different depths: %s' % halo_depths)
        depth = (halo_depths[0].pop(), halo_depths[1].pop())
        if depth != ((2, 2) if plan['conv'] == 'central' else (3, 4)):
            if not (2 <= depth[0] <= 5 and 2 <= depth[1] <= 5):
                raise UnsupportedByB200('halo depth %s' % (depth,))
            plan['halos'] = list(depth)

    # ---- cold kernels before the time loop (initialisation, metric evaluation, metric boundaries), in program order
    cold = []
    for c in before:
        if type(c).__name__ == 'ExchangeSelf':
            from sympy.printing.c import ccode
            cold.append({'name': 'exchange', 'exchange': True, 'arrays': [_strip(a) for a in c.transfer_arrays],
                         'size': [ccode(s) for s in c.transfer_size], 'from': [ccode(s) for s in c.transfer_from],
                         'to': [ccode(s) for s in c.transfer_to]})
        if type(c).__name__ == 'Kernel':
            n = _name(c)
            # 'User kernel: ...' placed BeforeSimulationStarts (e.g. the SFD filter's `Initialize the filter`, filters/SFD.py:50-62):
            # evaluated once by the cold path like the initialisation, its datasets uploaded with the plan
            if not (n.startswith('Grid_based_initialisation') or n.startswith('MetricsEquation') or n.startswith('Metric boundary')
                    or n.startswith('User kernel')):
                raise UnsupportedByB200('cold kernel %s is not implemented yet' % n)
            cold.append(_cold_kernel(c))
    if mass_source:
        if not viscous:
            raise UnsupportedByB200('mass source without viscous terms is not implemented')
        cold.append(source_cold)
        plan['mass_source'] = mass_source
    plan['cold'] = cold
    plan['user_kernels'] = user
    if monitor:
        plan['monitor'] = monitor
    plan['q_names'] = q_names

    # ---- constants, in declaration order (opsc.py:625-654)
    plan['constant_decls'] = []
    for c in ConstantsToDeclare.constants:
        if type(c).__name__ == 'ConstantObject':
            plan['constant_decls'].append([str(c), 'int' if 'int' in str(c.datatype.opsc()).lower() else 'double', _const_value(c)])
    # run-time integer arrays of SplitBC: declared like OPSC does, `int name[] = {Input, ...};`, for the user to fill in
    arrays = [[str(c.base), 'int', 2 * ndim] for c in ConstantsToDeclare.constants
              if type(c).__name__ == 'ConstantIndexed' and str(c.base).startswith(('split_range_', 'split_halo_range_'))]
    if arrays:
        plan['constant_array_decls'] = arrays
    return plan


def write_stub(plan, path=STUB_FILE):
    """opensbli.cpp-named parameter stub: the `name=Input;` lines of OPSC's main program (opsc.py:625-632)."""
    L = ['// OpenSBLI B200 back end: simulation parameters (filled in by substitute_simulation_parameters)',
         '// run with:  python -m opensbli_b200.run', 'int main(int argc, char **argv)', '{']
    for name, dtype, value in plan['constant_decls']:
        L.append('%s=%s;' % (name, value) if value == 'Input' else '%s = %s;' % (name, value))
    for name, dtype, count in plan.get('constant_array_decls', []):
        L.append('%s %s[] = {%s};' % (dtype, name, ', '.join(['Input'] * count)))
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
        if self.plan.get('not_executed'):
            print("B200: components outside the accelerated time loop are not executed: %s" % ', '.join(self.plan['not_executed']))
        print("Successfully generated the B200 execution plan.")
