"""Execution plan of the B200 back end.

A *plan* is what `B200(alg)` (backend.py) distils from an OpenSBLI algorithm object -- the same
information the reference's OPSC back end turns into OPS-C text (opensbli/code_generation/opsc.py:250-282):
dimensions, grid sizes, which spatial/temporal schemes act on the canonical compressible
Euler / Navier-Stokes system, boundary conditions per side, and the numerical constants.  It is a plain
JSON-able dict; `to_text` serialises it to the line format parsed by the C ABI (`osb_create`).

    ndim            1 | 2 | 3
    np              [n0, n1, n2][:ndim]           block0np{d}
    delta           [d0, ...]                     Delta{d}block0
    conv            'central' | 'weno' | 'teno'   Central / LLFWeno / LLFTeno
    order           4 | 5 | 6
    weno_formulation 'JS' | 'Z'
    averaging       'roe' | 'simple'              RoeAverage / SimpleAverage
    viscous         bool                          constant-viscosity Navier-Stokes terms (Central / StoreSome)
    halos           optional [hm, hp]             depth of the boundary / periodic halos when the block has consumers beyond the
                                                  scheme itself (a WENO filter on a central scheme: 3, 4); default 2/2 central, 3/4 WENO/TENO
    rk              'ls' | 'sbli'                 RungeKuttaLS / RungeKutta
    rk_a, rk_b      stage coefficients            LS: A, B ; SBLI: rkold, rknew
    constants       {name: float}                 gama, Minf, Re, Pr, dt, eps, TENO_CT, ...
    bc              [[side0, side1] per direction] each {'type': 'periodic'} | {'type': 'dirichlet', 'q': [...]}
                    | {'type': 'exchange'} (halo owned by the neighbouring rank of a slab decomposition)
                    | {'type': 'open'} (cut face of a window of a larger block: the halo is left as uploaded, hostpipe.py)
                    | {'type': 'isothermal_wall'} | {'type': 'adiabatic_wall'} | {'type': 'extrapolation', 'order': 0|1} | {'type': 'symmetry'}
                    | {'type': 'inlet_pressure_extrapolate'} | {'type': 'dirichlet_field', 'table': ndarray [nv, tangential]}
                    | {'type': 'zero_gradient_outlet'} | {'type': 'pressure_outlet'} (side 1; constant back_pressure) | {'type': 'inviscid_wall'}
                    | {'type': 'split', 'parts': [{'type': t, 'range': [lo0, hi0, lo1, hi1, ...], ('order': n)}, ...]}: SplitBC
                      (bc_core.py:200-217) -- several boundary classes share the face, each over its own part of the plane; `range`
                      is the part's evaluation range per direction (halo extension included), one plane thick along the face normal
                    | {'type': 'generic'}: the face's kernel is a run-time compiled user kernel with when = 'bc_<dir>_<side>' (user_kernels)
                    every non-periodic face may carry 'closure': 'reduced_access' | 'carpenter' (one-sided derivative rows)
    viscosity       {'type': 'constant'} | {'type': 'sutherland'} | {'type': 'power', 'exponent': e}
                    (constant: an optional constant 'mu' scales 1/Re, e.g. viscous_shock_tube.py:14-16)
    metric_fields   per direction None | 'D11'...: stretched direction; fields['D11'], fields['SD111'] hold the metric arrays
    teno_adaptive   bool: C_T from the Ducros sensor (constants teno_a1, teno_a2, epsilon)
    curvilinear     bool: full metric tensor; fields D00, D01, D10, D11, detJ (2-D, inviscid shock-capturing schemes;
                    strong-conservation form of apps/euler_wave_curvilinear)
    mass_source     {'field': name, 'rate': w}: Residual_rho += fields[name] * sin(w * iteration)  (transitional_SBLI.py:77-89);
                    iteration0: iteration number of the first step (restart)
                    dirichlet_field faces may carry 'free': [m, ...] (variables left untouched) and 'ke_free': True (imposed
                    energy = table + 1/2 sum(free momentum^2)/rho)
    forcing         bool: constant body force c0, c1, c2 (constants): momentum_i -= c_i, energy -= c_j u_j
    central_form    'blaisdell' (default; Skew() split of taylor_green_vortex / laminar_2D) | 'feiereisen' (quadratic split of
                    compressible_TCF_Central / turbulent_3D): how Central(4) writes the convective terms
    init            optional list of [lhs, rhs] assignment strings (numpy syntax) for the cold initialisation
    niter           optional int
"""
import copy
import json

CONV = ('central', 'weno', 'teno', 'generic')
BC_TYPES = ('periodic', 'dirichlet', 'exchange', 'isothermal_wall', 'extrapolation', 'inlet_pressure_extrapolate', 'symmetry',
            'dirichlet_field', 'adiabatic_wall', 'zero_gradient_outlet', 'pressure_outlet', 'inviscid_wall', 'generic', 'open', 'split')
SPLIT_PART_TYPES = ('dirichlet', 'isothermal_wall', 'adiabatic_wall', 'extrapolation', 'inlet_pressure_extrapolate', 'symmetry',
                    'zero_gradient_outlet', 'pressure_outlet', 'inviscid_wall')

# one-sided derivative closures: rows idx = 0.. next to the face x weights of the boundary-absolute points 0..np-1
# (reduced_access_scheme.py:36-43,76-83; Carpenter's first-derivative rows are taken from the scheme object by the back end)
CLOSURES = {
    'reduced_access': {'d1': [[-25.0 / 12, 48.0 / 12, -36.0 / 12, 16.0 / 12, -3.0 / 12], [-3.0 / 12, -10.0 / 12, 18.0 / 12, -6.0 / 12, 1.0 / 12]],
                       'd2': [[35.0 / 12, -104.0 / 12, 114.0 / 12, -56.0 / 12, 11.0 / 12], [11.0 / 12, -20.0 / 12, 6.0 / 12, 4.0 / 12, -1.0 / 12]]},
}


class PlanError(ValueError):
    pass


def validate(plan):
    nd = plan.get('ndim')
    if nd not in (1, 2, 3):
        raise PlanError('ndim must be 1, 2 or 3')
    for k in ('np', 'delta'):
        if len(plan.get(k, ())) < nd:
            raise PlanError('%s needs %d entries' % (k, nd))
    if plan.get('conv') not in CONV:
        raise PlanError('conv must be one of %s' % (CONV,))
    order = plan.get('order')
    ok = {'central': (4,), 'weno': (5,), 'teno': (5, 6), 'generic': (order,)}[plan['conv']]
    if plan['conv'] == 'generic' and not (plan.get('generic') or {}).get('nstages'):
        raise PlanError("conv 'generic' needs generic.nstages (every loop of the step is a run-time compiled kernel, see backend.extract_plan)")
    if order not in ok:
        raise PlanError('%s scheme: order %r is not implemented by the B200 back end (supported: %s)' % (plan['conv'], order, ok))
    if plan.get('rk') not in ('ls', 'sbli'):
        raise PlanError("rk must be 'ls' or 'sbli'")
    if not plan.get('rk_a') or len(plan['rk_a']) != len(plan.get('rk_b', ())):
        raise PlanError('rk_a / rk_b missing or of different length')
    if len(plan.get('bc', ())) < nd:
        raise PlanError('bc needs one [side0, side1] pair per direction')
    for d in range(nd):
        for s in range(2):
            b = plan['bc'][d][s]
            if b['type'] not in BC_TYPES:
                raise PlanError("boundary condition '%s' is not implemented by the B200 back end" % b['type'])
            if b['type'] == 'dirichlet' and len(b.get('q', ())) != nd + 2:
                raise PlanError('dirichlet bc needs %d conservative values' % (nd + 2))
            if b['type'] == 'pressure_outlet' and (s != 1 or 'back_pressure' not in plan.get('constants', {})):
                raise PlanError('pressure_outlet is defined for side 1 and needs the constant back_pressure (pressure_outlet.py:22-31)')
            if b['type'] == 'split':
                if not b.get('parts') or len(b['parts']) > 8:
                    raise PlanError('split bc needs 1..8 parts')
                plane = 0 if s == 0 else plan['np'][d] - 1
                for part in b['parts']:
                    if part.get('type') not in SPLIT_PART_TYPES:
                        raise PlanError("split bc: part type '%s' is not implemented (supported: %s)" % (part.get('type'), SPLIT_PART_TYPES))
                    r = part.get('range', ())
                    if len(r) != 2 * nd or any(r[2 * e] >= r[2 * e + 1] for e in range(nd)):
                        raise PlanError('split bc: every part needs a non-empty range [lo, hi) per direction')
                    if r[2 * d] != plane or r[2 * d + 1] != plane + 1:
                        raise PlanError('split bc: a part must cover the boundary plane of its face only (direction %d: [%d, %d))' % (d, plane, plane + 1))
                    if any(r[2 * e] < -5 or r[2 * e + 1] > plan['np'][e] + 5 for e in range(nd)):
                        raise PlanError('split bc: part range outside the padded block')
                    if part['type'] == 'dirichlet' and len(part.get('q', ())) != nd + 2:
                        raise PlanError('dirichlet bc needs %d conservative values' % (nd + 2))
                    if part['type'] == 'pressure_outlet' and (s != 1 or 'back_pressure' not in plan.get('constants', {})):
                        raise PlanError('pressure_outlet is defined for side 1 and needs the constant back_pressure (pressure_outlet.py:22-31)')
    c = plan.get('constants', {})
    need = [] if plan['conv'] == 'generic' else ['gama', 'dt'] + (['Re', 'Pr', 'Minf'] if plan.get('viscous') else [])
    for k in need:
        if k not in c:
            raise PlanError('missing constant %s' % k)
    return plan


def _f(x):
    return repr(float(x))


def to_text(plan):
    """Serialise for osb_create (see include/osbli_b200.h)."""
    validate(plan)
    nd = plan['ndim']
    L = ['osbli_plan 1', 'ndim %d' % nd,
         'np ' + ' '.join(str(int(n)) for n in plan['np'][:nd]),
         'delta ' + ' '.join(_f(d) for d in plan['delta'][:nd]),
         'conv %s' % plan['conv'], 'order %d' % plan['order'],
         'weno_formulation %s' % plan.get('weno_formulation', 'JS'),
         'averaging %s' % plan.get('averaging', 'roe'),
         'viscous %d' % (1 if plan.get('viscous') else 0)] + (['halos %d %d' % tuple(plan['halos'])] if plan.get('halos') else []) + [
         'rk %s' % plan['rk']] + (['generic_stages %d' % int(plan['generic']['nstages'])] if plan['conv'] == 'generic' else []) + [
         'rk_a ' + ' '.join(_f(v) for v in plan['rk_a']),
         'rk_b ' + ' '.join(_f(v) for v in plan['rk_b'])]
    for k, v in sorted(plan['constants'].items()):
        L.append('const %s %s' % (k, _f(v)))
    closures = set()
    for d in range(nd):
        for s in range(2):
            b = plan['bc'][d][s]
            cl = ' closure' if b.get('closure') else ''
            if b.get('closure'):
                closures.add(b['closure'])
            if b['type'] == 'dirichlet':
                L.append('bc %d %d dirichlet %s%s' % (d, s, ' '.join(_f(v) for v in b['q']), cl))
            elif b['type'] == 'extrapolation':
                L.append('bc %d %d extrapolation %d%s' % (d, s, int(b.get('order', 0)), cl))
            elif b['type'] == 'split':
                L.append('bc %d %d split%s' % (d, s, cl))
                for part in b['parts']:
                    extra = ' '.join(_f(v) for v in part['q']) if part['type'] == 'dirichlet' else str(int(part.get('order', 0)))
                    L.append('bc_part %d %d %s %s %s' % (d, s, part['type'], ' '.join(str(int(v)) for v in part['range']), extra))
            elif b['type'] == 'dirichlet_field':
                fr = ''.join(' free %d' % m for m in b.get('free', [])) + (' ke_free' if b.get('ke_free') else '')
                L.append('bc %d %d dirichlet_field%s%s' % (d, s, cl, fr))
            else:
                L.append('bc %d %d %s%s' % (d, s, b['type'], cl))
    if len(closures) > 1:
        raise PlanError('only one closure scheme per block is implemented (got %s)' % sorted(closures))
    for name in closures:
        tab = plan.get('closures', CLOSURES).get(name) or CLOSURES.get(name)
        if tab is None:
            raise PlanError("no coefficient table for closure '%s'" % name)
        for key in ('d1', 'd2'):
            rows = tab[key]
            L.append('closure_%s %d %d %s' % (key, len(rows), len(rows[0]), ' '.join(_f(v) for r in rows for v in r)))
    visc = plan.get('viscosity', {'type': 'constant'})
    L.append('viscosity %s%s' % (visc['type'], ' ' + _f(visc['exponent']) if visc['type'] == 'power' else ''))
    for d, name in enumerate(plan.get('metric_fields', [None] * nd)):
        if name:
            L.append('metric %d 1' % d)
    if plan.get('teno_adaptive'):
        L.append('teno_adaptive 1')
    if plan.get('forcing'):
        L.append('forcing 1')
    if plan.get('curvilinear'):
        L.append('curvilinear 1')
    if plan.get('mass_source'):
        L.append('mass_source %s %d' % (_f(plan['mass_source']['rate']), int(plan.get('iteration0', 0))))
    if plan.get('central_form', 'blaisdell') != 'blaisdell':
        if plan['central_form'] != 'feiereisen':
            raise PlanError("central_form must be 'blaisdell' or 'feiereisen'")
        L.append('central_form feiereisen')
    return '\n'.join(L) + '\n'


def save(plan, path):
    with open(path, 'w') as f:
        json.dump(plan, f, indent=1, sort_keys=True)


def load(path):
    with open(path) as f:
        return validate(json.load(f))


def with_size(plan, np_, delta=None, **constants):
    """Copy of a plan on another grid / with other constants."""
    p = copy.deepcopy(plan)
    p['np'] = list(np_)
    if delta is not None:
        p['delta'] = list(delta)
    p['constants'].update(constants)
    return validate(p)
