"""`python -m opensbli_b200.run [dir]` -- run what `B200(alg)` + `substitute_simulation_parameters` left in a directory.

The analogue of `make && ./opensbli_seq` in the reference workflow: reads opensbli_b200.plan.json (symbolic plan)
and the parameter stub opensbli.cpp (the `name = value;` lines, values being C expressions exactly as the reference's
apps pass them, e.g. '2*M_PI/block0np0', 'ceil(0.2/0.0002)'), resolves them with C arithmetic semantics, evaluates the
cold initialisation kernel with numpy on the host, runs `niter` steps on the GPU through the C ABI and writes
opensbli_output.npz (conserved fields, reference layout incl. halos)."""
import ast
import json
import re
import math
import os
import sys

import numpy as np

from . import plan as _plan
from .backend import PLAN_FILE, STUB_FILE

_FUNCS = {n: getattr(math, n) for n in ('ceil', 'floor', 'sqrt', 'sin', 'cos', 'tan', 'exp', 'log', 'tanh', 'sinh', 'cosh', 'pow', 'fabs', 'atan', 'asin', 'acos')}


def c_eval(expr, env):
    """Evaluate a C arithmetic expression: int/int truncates, ceil/floor return doubles, M_PI known."""
    def ev(n):
        if isinstance(n, ast.Expression):
            return ev(n.body)
        if isinstance(n, ast.Constant):
            return n.value
        if isinstance(n, ast.Name):
            if n.id == 'M_PI':
                return math.pi
            return env[n.id]
        if isinstance(n, ast.UnaryOp):
            v = ev(n.operand)
            return -v if isinstance(n.op, ast.USub) else +v
        if isinstance(n, ast.BinOp):
            a, b = ev(n.left), ev(n.right)
            if isinstance(n.op, ast.Add):
                return a + b
            if isinstance(n.op, ast.Sub):
                return a - b
            if isinstance(n.op, ast.Mult):
                return a * b
            if isinstance(n.op, ast.Div):
                if isinstance(a, int) and isinstance(b, int):
                    return int(a / b)          # C integer division truncates toward zero
                return a / b
            raise ValueError('operator %s not supported in %r' % (type(n.op).__name__, expr))
        if isinstance(n, ast.Call) and isinstance(n.func, ast.Name) and n.func.id in _FUNCS:
            return float(_FUNCS[n.func.id](*[ev(a) for a in n.args]))
        raise ValueError('cannot evaluate %r' % expr)
    return ev(ast.parse(expr.strip(), mode='eval'))


def read_stub(text, decls, overrides=None):
    """-> ordered {name: value} from the stub's `name = expr;` lines (declared types from the plan).  `overrides` replaces
    the value of a parameter (e.g. block0np0) before the expressions that depend on it are evaluated."""
    types = {n: t for n, t, _ in decls}
    env = {}
    overrides = overrides or {}
    for line in text.splitlines():
        line = line.strip()
        if not line.endswith(';') or '=' not in line or line.startswith(('//', 'int iter')):
            continue
        name, expr = line[:-1].split('=', 1)
        name, expr = name.strip(), expr.strip()
        if name not in types:
            continue
        if expr == 'Input':
            raise ValueError("simulation parameter '%s' has no value: call substitute_simulation_parameters" % name)
        v = overrides[name] if name in overrides else c_eval(expr, env)
        env[name] = int(v) if types[name] == 'int' else float(v)
    return env


class ColdRunner(object):
    """numpy interpreter of the cold kernels distilled by the back end (initialisation, metric evaluation, metric
    boundaries, Dirichlet states): every kernel is a range plus ordered assignments whose right-hand sides read
    datasets at relative offsets through `_A(name, offset)`.  Vectorised over the kernel's range, which is valid
    because no cold kernel reads, at another point, a value it has itself written."""

    def __init__(self, nd, np_, env, halo=5):
        self.nd, self.np_, self.env, self.h = nd, list(np_), dict(env), halo
        self.shape = tuple(n + 2 * halo for n in reversed(self.np_))
        self.arrays = {}

    def array(self, name):
        if name not in self.arrays:
            self.arrays[name] = np.zeros(self.shape)       # OPS semantics: datasets start zeroed
        return self.arrays[name]

    def _slices(self, lo, hi, off):
        return tuple(slice(lo[d] + off[d] + self.h, hi[d] + off[d] + self.h) for d in reversed(range(self.nd)))

    def exchange(self, op):
        """Periodic copy of a cold phase (ExchangeSelf, periodic.py:42-56; e.g. the metric arrays of a curvilinear grid):
        size / from / to are the reference's own transfer description (C expressions in block0np{d})."""
        size = [int(c_eval(s, self.env)) for s in op['size']]
        src = [int(c_eval(s, self.env)) for s in op['from']]
        dst = [int(c_eval(s, self.env)) for s in op['to']]
        s_from = tuple(slice(src[d] + self.h, src[d] + size[d] + self.h) for d in reversed(range(self.nd)))
        s_to = tuple(slice(dst[d] + self.h, dst[d] + size[d] + self.h) for d in reversed(range(self.nd)))
        for name in op['arrays']:
            a = self.array(name)
            a[s_to] = a[s_from].copy()

    def run(self, kernel):
        if kernel.get('exchange'):
            return self.exchange(kernel)
        rng = [int(c_eval(r, self.env)) for r in kernel['range']]
        lo, hi = rng[0::2], rng[1::2]
        ax = [np.arange(lo[d], hi[d]) for d in range(self.nd)]
        grids = np.meshgrid(*reversed(ax), indexing='ij')[::-1]
        zero = (0,) * self.nd

        def _A(name, off=zero):
            return self.array(name)[self._slices(lo, hi, off)]
        ns = dict(self.env)
        ns.update({'idx%d' % d: grids[d] for d in range(self.nd)})
        g = {'numpy': np, 'math': math, '_A': _A}
        for lhs, off, rhs in kernel['statements']:
            val = eval(rhs, g, ns)
            if off is None:
                ns[lhs] = val
            else:
                self.array(lhs)[self._slices(lo, hi, off)] = val
        return lo, hi


def user_kernel_source(k, index, env, nd):
    """CUDA C of one point-wise user kernel (entry signature: include/osbli_b200.h, osb_add_user_kernel)."""
    fields = list(k['reads']) + [w for w in k['writes'] if w not in k['reads']]
    entry = 'osb_user_kernel_%d' % index
    text = ' '.join(s[2] for s in k['statements'])
    L = ['struct UserFields { double *p[48]; };']
    for name, val in env.items():                       # the constants the statements mention, as the reference's C globals
        if re.search(r'\b%s\b' % re.escape(name), text):
            L.append('#define %s (%s)' % (name, repr(float(val))))
    L += ['extern "C" __global__ void %s(long long off, int n0, int n1, int n2, int lo0, int lo1, int lo2, long long s1, long long s2, UserFields f) {' % entry,
          '  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, k = blockIdx.z;',
          '  if (i >= n0 || j >= n1 || k >= n2) return;',
          '  const long long X = off + (lo0 + i) + (lo1 + j) * s1 + (lo2 + k) * s2;']
    for n, name in enumerate(fields):
        L.append('  double *%s = f.p[%d];' % (name, n))
    for name in dict.fromkeys(k['locals']):
        L.append('  double %s;' % name)
    for lhs, is_field, rhs in k['statements']:
        L.append('  %s%s = %s;' % (lhs, '[X]' if is_field else '', rhs))
    L.append('}')
    rng = [int(c_eval(r, env)) for r in k['range']]
    return {'name': k['name'], 'entry': entry, 'source': '\n'.join(L) + '\n', 'fields': fields, 'range': rng, 'when': k['when'],
            'writes': list(k['writes'])}


def resolve(plan_sym, env):
    """symbolic plan + parameter values -> numeric plan accepted by opensbli_b200.Simulation (cold kernels evaluated:
    metric fields, tabulated Dirichlet states); returns (plan, ColdRunner holding every cold dataset)."""
    nd = plan_sym['ndim']
    p = {k: plan_sym[k] for k in ('ndim', 'conv', 'order', 'weno_formulation', 'averaging', 'viscous', 'rk', 'rk_a', 'rk_b')}
    for k in ('viscosity', 'metric_fields', 'teno_adaptive', 'closures', 'forcing', 'central_form', 'curvilinear'):   # copied verbatim
        if k in plan_sym:
            p[k] = plan_sym[k]
    p['np'] = [int(env['block0np%d' % d]) for d in range(nd)]
    p['delta'] = [float(env['Delta%dblock0' % d]) for d in range(nd)]
    p['constants'] = {k: float(v) for k, v in env.items() if not k.startswith(('block0np', 'Delta', 'niter'))}
    p['niter'] = int(env.get('niter', 0))
    cold = ColdRunner(nd, p['np'], env)
    for k in plan_sym.get('cold', []):
        cold.run(k)
    p['fields'] = {}
    for d, name in enumerate(p.get('metric_fields') or []):
        if name:
            for f in (name, 'S' + name + str(d)):
                p['fields'][f] = cold.array(f).copy()
    p['user_kernels'] = [user_kernel_source(k, n, env, nd) for n, k in enumerate(plan_sym.get('user_kernels', []))]
    # datasets the user kernels only read and the cold path has evaluated (coordinates, metric terms ...): shipped with the
    # plan; the runtime declares and uploads those the solver does not hold itself
    written = set(w for k in p['user_kernels'] for w in k['writes'])
    p['user_fields'] = {f: cold.array(f).copy() for k in p['user_kernels'] for f in k['fields']
                        if f not in written and f in cold.arrays and f not in plan_sym['q_names']}
    if plan_sym.get('monitor'):
        m = dict(plan_sym['monitor'])
        m['probes'] = [[int(c_eval(x, env)) for x in pr] for pr in m['probes']]
        p['monitor'] = m
    if plan_sym.get('curvilinear'):
        for f in ['D%d%d' % (i, j) for i in range(nd) for j in range(nd)] + ['detJ']:
            p['fields'][f] = cold.array(f).copy()
    if plan_sym.get('mass_source'):
        ms = plan_sym['mass_source']
        p['mass_source'] = {'field': ms['field'], 'rate': float(c_eval(ms['rate'], env))}
        p['fields'][ms['field']] = cold.array(ms['field']).copy()
    bc = []
    for d in range(nd):
        pair = []
        for s in range(2):
            b = dict(plan_sym['bc'][d][s])
            if b['type'] == 'dirichlet_field':
                # evaluate the BC equations on a scratch copy of the datasets, read the imposed state off the boundary plane
                scratch = ColdRunner(nd, p['np'], env)
                scratch.arrays = {k: v.copy() for k, v in cold.arrays.items()}
                for m in b.get('free', []):                   # the kinetic energy of the free momenta is added at run time
                    scratch.array(plan_sym['q_names'][m])[...] = 0.0
                scratch.run(b.pop('kernel'))
                plane = p['np'][d] - 1 + scratch.h if s == 1 else scratch.h
                tabs = []
                for n in plan_sym['q_names']:
                    a = np.moveaxis(scratch.array(n), nd - 1 - d, 0)
                    tabs.append(a[plane].reshape(-1))
                table = np.stack(tabs)
                if all(np.all(t == t[0]) for t in table):        # constant state: plain Dirichlet
                    b = {'type': 'dirichlet', 'q': [float(t[0]) for t in table], **({'closure': b['closure']} if b.get('closure') else {})}
                else:
                    b['table'] = table
            pair.append(b)
        bc.append(pair)
    p['bc'] = bc
    return _plan.validate(p), cold


def initial_state(plan_sym, cold):
    return [np.ascontiguousarray(cold.array(n)) for n in plan_sym['q_names']]


def time_loop(sim, plan, niter, workdir='.'):
    """The reference's time loop as the runner drives it: niter steps on the GPU; with a SimulationMonitor the run is cut at
    the iterations that print (iter == 0 or (iter+1) %% frequency == 0, algorithm.py:433-437) and the probe values are written
    in the reference's format (simulation_monitors.py:128-160).  Returns the device time of the steps in ms."""
    mon = plan.get('monitor')
    if not mon or niter <= 0:
        return sim.step_timed(niter) if niter > 0 else 0.0
    dt = plan['constants']['dt']
    fmt = '%%.%df' % mon['precision']
    out = open(os.path.join(workdir, mon['output_file']), 'w') if mon.get('output_file') else sys.stdout
    ms, done = 0.0, 0
    stops = sorted(set([1] + list(range(mon['frequency'], niter + 1, mon['frequency']))))
    try:
        for stop in stops:
            ms += sim.step_timed(stop - done)
            done = stop
            if stop == 1:
                out.write(', '.join(['Iteration', 'Time'] + ['%s_B0(%s)' % (a, ', '.join(str(x) for x in pr)) for a, pr in zip(mon['arrays'], mon['probes'])]) + '\n')
            vals = []
            for a, pr in zip(mon['arrays'], mon['probes']):
                v = sim.read_point(a, *pr)
                vals.append(v / stop if 'mean' in a else v)      # running sums are reported as means
            out.write(', '.join(['%d' % stop, fmt % (stop * dt)] + [fmt % v for v in vals]) + '\n')
        if done < niter:
            ms += sim.step_timed(niter - done)
    finally:
        if out is not sys.stdout:
            out.close()
    return ms


def load_case(workdir='.', overrides=None):
    plan_sym = json.load(open(os.path.join(workdir, PLAN_FILE)))
    env = read_stub(open(os.path.join(workdir, STUB_FILE)).read(), plan_sym['constant_decls'], overrides)
    plan_num, cold = resolve(plan_sym, env)
    return plan_sym, env, plan_num, cold


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    workdir = argv[0] if argv else '.'
    from .runtime import Simulation
    plan_sym, env, plan_num, cold = load_case(workdir)
    q0 = initial_state(plan_sym, cold)
    niter = plan_num.get('niter', 0)
    with Simulation(plan_num) as sim:
        sim.set_state(q0)
        ms = time_loop(sim, plan_num, niter, workdir)
        if any(k['when'] == 'after_loop' for k in plan_num.get('user_kernels', [])):
            sim.run_user_kernels('after_loop')                  # loops after the time loop (e.g. statistics / niter)
        q = sim.get_state()
        extra = {w: sim.download(w) for k in plan_num.get('user_kernels', []) for w in k['writes']}
    print('Total Wall time %f' % (ms * 1e-3))      # same span as the reference's Timers (algorithm.py:301-327)
    np.savez(os.path.join(workdir, 'opensbli_output.npz'), **{n: a for n, a in zip(plan_sym['q_names'], q)}, **extra)
    return 0


if __name__ == '__main__':
    sys.exit(main())
