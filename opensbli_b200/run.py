"""`python -m opensbli_b200.run [dir]` -- run what `B200(alg)` + `substitute_simulation_parameters` left in a directory.

The analogue of `make && ./opensbli_seq` in the reference workflow: reads opensbli_b200.plan.json (symbolic plan)
and the parameter stub opensbli.cpp (the `name = value;` lines, values being C expressions exactly as the reference's
apps pass them, e.g. '2*M_PI/block0np0', 'ceil(0.2/0.0002)'), resolves them with C arithmetic semantics, evaluates the
cold initialisation kernel with numpy on the host, runs `niter` steps on the GPU through the C ABI and writes
opensbli_output.npz (conserved fields, reference layout incl. halos)."""
import ast
import json
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


def read_stub(text, decls):
    """-> ordered {name: value} from the stub's `name = expr;` lines (declared types from the plan)."""
    types = {n: t for n, t, _ in decls}
    env = {}
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
        v = c_eval(expr, env)
        env[name] = int(v) if types[name] == 'int' else float(v)
    return env


def _exec_statements(statements, ns):
    for lhs, rhs in statements:
        ns[lhs] = eval(rhs, {'numpy': np, 'math': math}, ns)
    return ns


def resolve(plan_sym, env):
    """symbolic plan + parameter values -> numeric plan accepted by opensbli_b200.Simulation."""
    nd = plan_sym['ndim']
    p = {k: plan_sym[k] for k in ('ndim', 'conv', 'order', 'weno_formulation', 'averaging', 'viscous', 'rk', 'rk_a', 'rk_b')}
    p['np'] = [int(env['block0np%d' % d]) for d in range(nd)]
    p['delta'] = [float(env['Delta%dblock0' % d]) for d in range(nd)]
    p['constants'] = {k: float(v) for k, v in env.items() if not k.startswith(('block0np', 'Delta', 'niter'))}
    p['niter'] = int(env.get('niter', 0))
    bc = []
    for d in range(nd):
        pair = []
        for s in range(2):
            b = plan_sym['bc'][d][s]
            if b['type'] == 'dirichlet':
                ns = _exec_statements(b['statements'], dict(env))
                pair.append({'type': 'dirichlet', 'q': [float(ns[n]) for n in plan_sym['q_names']]})
            else:
                pair.append({'type': b['type']})
        bc.append(pair)
    p['bc'] = bc
    return _plan.validate(p)


def initial_state(plan_sym, plan_num, env, halo=5):
    """Evaluate the Grid_based_initialisation statements over the padded block (gridbasedinit.py:46-57)."""
    nd = plan_num['ndim']
    ax = [np.arange(-halo, n + halo) for n in plan_num['np']]
    grids = np.meshgrid(*reversed(ax), indexing='ij')[::-1]
    ns = dict(env)
    for d in range(nd):
        ns['idx%d' % d] = grids[d]
    _exec_statements(plan_sym['init'], ns)
    shape = grids[0].shape
    return [np.ascontiguousarray(np.broadcast_to(np.asarray(ns[n], dtype=np.float64), shape)) for n in plan_sym['q_names']]


def load_case(workdir='.'):
    plan_sym = json.load(open(os.path.join(workdir, PLAN_FILE)))
    env = read_stub(open(os.path.join(workdir, STUB_FILE)).read(), plan_sym['constant_decls'])
    return plan_sym, env, resolve(plan_sym, env)


def main(argv=None):
    argv = sys.argv[1:] if argv is None else argv
    workdir = argv[0] if argv else '.'
    from .runtime import Simulation
    plan_sym, env, plan_num = load_case(workdir)
    q0 = initial_state(plan_sym, plan_num, env)
    niter = plan_num.get('niter', 0)
    with Simulation(plan_num) as sim:
        sim.set_state(q0)
        ms = sim.step_timed(niter) if niter > 0 else 0.0
        q = sim.get_state()
    print('Total Wall time %f' % (ms * 1e-3))      # same span as the reference's Timers (algorithm.py:301-327)
    np.savez(os.path.join(workdir, 'opensbli_output.npz'), **{n: a for n, a in zip(plan_sym['q_names'], q)})
    return 0


if __name__ == '__main__':
    sys.exit(main())
