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
        if isinstance(n, ast.Subscript):        # element of a run-time integer array (SplitBC ranges)
            return ev(n.value)[int(ev(n.slice))]
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


def read_stub(text, decls, overrides=None, array_decls=()):
    """-> ordered {name: value} from the stub's `name = expr;` lines (declared types from the plan).  `overrides` replaces
    the value of a parameter (e.g. block0np0) before the expressions that depend on it are evaluated.  Integer arrays
    (`int name[] = {a, b, ...};`, the SplitBC ranges the user edits by hand as in the reference's opensbli.cpp) become lists."""
    types = {n: t for n, t, _ in decls}
    arrays = {n: int(cnt) for n, _, cnt in array_decls}
    env = {}
    overrides = overrides or {}
    for line in text.splitlines():
        line = line.strip()
        m = re.match(r'int\s+(\w+)\[\]\s*=\s*\{(.*)\};$', line)
        if m and m.group(1) in arrays:
            if m.group(1) in overrides:
                env[m.group(1)] = [int(v) for v in overrides[m.group(1)]]
                continue
            items = [x.strip() for x in m.group(2).split(',')]
            if 'Input' in items or len(items) != arrays[m.group(1)]:
                raise ValueError("range array '%s' needs %d integer values in the parameter file (opensbli.cpp): {first, last+1} per direction "
                                 "for split_range_*, {extension below, extension above} for split_halo_range_*" % (m.group(1), arrays[m.group(1)]))
            env[m.group(1)] = [int(c_eval(x, env)) for x in items]
            continue
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


def user_kernel_source(k, index, env, nd, stage=None):
    """CUDA C of one run-time compiled kernel (entry signature: include/osbli_b200.h, osb_add_user_kernel).  `stage`: the RK stage
    the copy is compiled for when its statements index constants by the stage counter (rkA[stage], generic path)."""
    fields = list(k['reads']) + [w for w in k['writes'] if w not in k['reads']]
    if len(fields) > 96:
        raise ValueError('kernel %s touches %d datasets (limit 96, OSB_MAX_USER_FIELDS)' % (k['name'], len(fields)))
    entry = 'osb_user_kernel_%d' % index
    text = ' '.join(s[2] for s in k['statements'] if s[2])
    L = ['struct UserFields { double *p[96]; long long iter; };']          # OSB_MAX_USER_FIELDS (include/osbli_b200.h)
    for name, val in env.items():                       # the constants the statements mention, as the reference's C globals
        if not isinstance(val, list) and re.search(r'\b%s\b' % re.escape(name), text):
            # integer parameters (block0np<d>, niter) stay integers, as the reference's C globals are: index arithmetic and
            # integer division behave as in its program
            L.append('#define %s (%s)' % (name, str(val) if isinstance(val, int) and not isinstance(val, bool) else repr(float(val))))
    # OSB_GOFF<d>: global index of the rank's first point along direction d (slab-decomposed runs; decomp.local_plan defines it)
    L += ['#ifndef OSB_GOFF0', '#define OSB_GOFF0 0', '#endif', '#ifndef OSB_GOFF1', '#define OSB_GOFF1 0', '#endif', '#ifndef OSB_GOFF2', '#define OSB_GOFF2 0', '#endif']
    L += ['extern "C" __global__ void %s(long long off, int n0, int n1, int n2, int lo0, int lo1, int lo2, long long s1, long long s2, UserFields f) {' % entry,
          '  const int i = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y, k = blockIdx.z;',
          '  if (i >= n0 || j >= n1 || k >= n2) return;',
          '  const long long X = off + (lo0 + i) + (lo1 + j) * s1 + (lo2 + k) * s2;',
          '  const int idx0 = lo0 + i + OSB_GOFF0, idx1 = lo1 + j + OSB_GOFF1, idx2 = lo2 + k + OSB_GOFF2;   // grid indices (block.grid_indexes)',
          '  (void)idx0; (void)idx1; (void)idx2;']
    for name, vals in sorted(k.get('indexed_constants', {}).items()):      # rkA[stage] ...: the table and the stage this copy runs in
        L.append('  const double %s[%d] = {%s};' % (name, len(vals), ', '.join(repr(float(v)) for v in vals)))
    if stage is not None:
        L.append('  const int stage = %d; (void)stage;' % stage)
    if re.search(r'\biter\b', text):                      # the loop counter, an ops_arg_gbl in the reference (e.g. sin(omega dt iter))
        L.append('  const int iter = (int)f.iter;')
    for n, name in enumerate(fields):
        L.append('  double *%s = f.p[%d];' % (name, n))
    # a kernel-local variable may carry the name of a dataset the kernel also touches (the reference tells `x0` from `x0_B0[...]`;
    # here datasets lost the suffix): such a local is renamed -- a dataset name is always followed by its index, a local never
    clash = {name: re.compile(r'\b%s\b(?!\s*\[)' % re.escape(name)) for name in dict.fromkeys(k['locals']) if name in fields}
    def loc(text):
        for name, rx in clash.items():
            text = rx.sub(name + '_local', text)
        return text
    for name in dict.fromkeys(k['locals']):
        L.append('  double %s = 0.0;' % loc(name))       # kernel locals start at zero, as in the reference's generated C (opsc.py:340-343)
    for st in k['statements']:
        lhs, is_field, rhs = st[0], st[1], st[2]
        if lhs in ('#if', '#elif', '#else', '#end'):              # GroupedPiecewise: equations under if / else if / else (opsc.py:372-397)
            L.append({'#if': '  if (%s) {' % loc(rhs or ''), '#elif': '  } else if (%s) {' % loc(rhs or ''), '#else': '  } else {', '#end': '  }'}[lhs])
            continue
        at = (st[3] if len(st) > 3 and st[3] else 'X')            # relative write (boundary kernels): index printed by the back end
        L.append('  %s%s = %s;' % (lhs if is_field else loc(lhs), '[%s]' % at if is_field else '', loc(rhs)))
    L.append('}')
    rng = [int(c_eval(r, env)) for r in k['range']]
    if 'one_plane_along' in k:              # SplitBC part: the race check of the back end relied on it
        d = k['one_plane_along']
        if rng[2 * d + 1] - rng[2 * d] != 1:
            raise ValueError('boundary kernel %s: its range must be ONE plane along direction %d, got [%d, %d)' % (k['name'], d, rng[2 * d], rng[2 * d + 1]))
    when = k['when'] if stage is None else 'stage_%d' % stage
    return {'name': k['name'], 'entry': entry, 'source': '\n'.join(L) + '\n', 'fields': fields, 'range': rng, 'when': when,
            'writes': list(k['writes'])}


def user_kernel_sources(plan_sym, env, nd):
    """All run-time compiled kernels of a plan, in program order.  On the generic path a kernel of the stage loop that reads the
    stage counter is compiled once per stage; registration order = launch order within each list."""
    out = []
    nstages = (plan_sym.get('generic') or {}).get('nstages', 0)
    for k in plan_sym.get('user_kernels', []):
        uses_stage = k['when'] == 'stage' and any(s[2] and re.search(r'\bstage\b', s[2]) for s in k['statements'])
        if uses_stage:
            for s in range(nstages):
                out.append(user_kernel_source(k, len(out), env, nd, stage=s))
        else:
            out.append(user_kernel_source(k, len(out), env, nd))
    return out


def resolve(plan_sym, env):
    """symbolic plan + parameter values -> numeric plan accepted by opensbli_b200.Simulation (cold kernels evaluated:
    metric fields, tabulated Dirichlet states); returns (plan, ColdRunner holding every cold dataset)."""
    nd = plan_sym['ndim']
    p = {k: plan_sym[k] for k in ('ndim', 'conv', 'order', 'weno_formulation', 'averaging', 'viscous', 'rk', 'rk_a', 'rk_b')}
    for k in ('viscosity', 'metric_fields', 'teno_adaptive', 'closures', 'forcing', 'central_form', 'curvilinear', 'halos', 'generic'):   # copied verbatim
        if k in plan_sym:
            p[k] = plan_sym[k]
    p['np'] = [int(env['block0np%d' % d]) for d in range(nd)]
    p['delta'] = [float(env['Delta%dblock0' % d]) for d in range(nd)]
    p['constants'] = {k: float(v) for k, v in env.items() if not k.startswith(('block0np', 'Delta', 'niter')) and not isinstance(v, list)}
    p['niter'] = int(env.get('niter', 0))
    cold = ColdRunner(nd, p['np'], env)
    for k in plan_sym.get('cold', []):
        cold.run(k)
    p['fields'] = {}
    for d, name in enumerate(p.get('metric_fields') or []):
        if name:
            for f in (name, 'S' + name + str(d)):
                p['fields'][f] = cold.array(f).copy()
    p['user_kernels'] = user_kernel_sources(plan_sym, env, nd)
    # datasets the user kernels only read and the cold path has evaluated (coordinates, metric terms ...): shipped with the
    # plan; the runtime declares and uploads those the solver does not hold itself
    written = set(w for k in p['user_kernels'] for w in k['writes'])
    p['user_fields'] = {f: cold.array(f).copy() for k in p['user_kernels'] for f in k['fields']
                        if f in cold.arrays and f not in plan_sym['q_names']}
    if plan_sym.get('generic'):
        # generic path: a dataset some loop reads but no loop writes and the cold path never evaluated holds the zeros OPS declares
        # it with (e.g. `u2` of StoreSome(4, 'u0 u1 u2 T') in a 2-D app, katzer_SBLI.py:60-75) -- declared explicitly, and listed
        known = set(plan_sym['q_names']) | set('u%d' % d for d in range(nd)) | {'p', 'a', 'T'} | \
            set('Residual%d' % m for m in range(nd + 2)) | set(pre + n for n in plan_sym['q_names'] for pre in ('tempRK_',)) | \
            set(n + '_RKold' for n in plan_sym['q_names'])
        zero = sorted(set(f for k in p['user_kernels'] for f in k['fields']) - written - set(p['user_fields']) - known)
        for f in zero:
            p['user_fields'][f] = np.zeros(cold.shape)
        p['zero_datasets'] = zero
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


def read_iteration_ops(stub_text):
    """What `print_iteration_ops(every=N, NaN_check='rho_B0')` (utilities/helperfunctions.py:172-190) injected after the
    `int iter=0;` line of the program text: -> (every or None, dataset to NaN-check or None)."""
    m = re.search(r'if\(fmod\(iter\+1,\s*(\d+)\)\s*==\s*0\)', stub_text)
    n = re.search(r'ops_NaNcheck\((\w+)\)', stub_text)
    name = n.group(1) if n else None
    if name and name.endswith('_B0'):
        name = name[:-3]
    return (int(m.group(1)) if m else None), name


class Runner(object):
    """One rank of a run: a Simulation (one GPU) or a DistributedSimulation (torchrun, one process per GPU; the block is cut
    into slabs along its slowest axis).  Field access is by GLOBAL padded arrays on rank 0."""

    def __init__(self, plan, dist=None, device=-1):
        from .runtime import Simulation
        from . import decomp
        self.plan_global = plan
        self.dist = dist
        self.world = dist.get_world_size() if dist is not None else 1
        self.rank = dist.get_rank() if dist is not None else 0
        if self.world > 1:
            self.dsim = decomp.DistributedSimulation(plan, dist, device)
            self.sim = self.dsim.sim
            self.k0, self.nk = self.dsim.offset, self.dsim.nloc
        else:
            self.dsim = None
            self.sim = Simulation(plan, device=device)
            self.k0, self.nk = 0, plan['np'][plan['ndim'] - 1]

    def close(self):
        self.sim.close()

    def local_part(self, a):
        """slab of a global padded array (slowest axis first in numpy order) incl. 5 halo planes on both sides"""
        return np.ascontiguousarray(a[self.k0:self.k0 + self.nk + 10]) if self.world > 1 else a

    def set_state(self, q_global):
        q = [self.local_part(a) for a in q_global]
        if self.dsim is not None:
            self.dsim.set_state(q)
        else:
            self.sim.set_state(q)

    def step_timed(self, n):
        if n <= 0:
            return 0.0
        if self.dsim is None:
            return self.sim.step_timed(n)
        self.dsim.barrier()
        self.sim.timer_start()
        self.dsim.step(n)
        ms = self.sim.timer_stop()
        self.dsim.barrier()
        return ms

    def gather(self, name):
        """global padded array of a dataset on rank 0 (None elsewhere): interior planes of every slab, outer halo planes of
        the first and the last one"""
        a = self.sim.download(name)
        if self.world == 1:
            return a
        parts = [None] * self.world
        self.dist.all_gather_object(parts, (self.k0, self.nk, a))
        if self.rank != 0:
            return None
        parts.sort(key=lambda p: p[0])
        pieces = [parts[0][2][:5]] + [p[2][5:5 + p[1]] for p in parts] + [parts[-1][2][-5:]]
        return np.concatenate(pieces, axis=0)

    def read_point(self, name, idx):
        """value of a dataset at a GLOBAL grid index (every rank gets it)"""
        idx = list(idx) + [0] * (3 - len(idx))
        ax = self.plan_global['ndim'] - 1
        if self.world == 1:
            return self.sim.read_point(name, *idx)
        mine = self.k0 <= idx[ax] < self.k0 + self.nk
        loc = list(idx)
        loc[ax] -= self.k0
        v = self.sim.read_point(name, *loc) if mine else None
        vals = [None] * self.world
        self.dist.all_gather_object(vals, v)
        return [x for x in vals if x is not None][0]

    def nan_count(self, name):
        n = self.sim.nan_check(name)
        if self.world > 1:
            vals = [None] * self.world
            self.dist.all_gather_object(vals, n)
            n = sum(vals)
        return n

    def diagnostics(self):
        """volume averages over the whole block: kinetic energy, enstrophy, mass, total energy; max Mach number"""
        d = self.sim.diagnostics()
        if self.world > 1:
            vals = [None] * self.world
            self.dist.all_gather_object(vals, d)
            d = {k: (max(v[k] for v in vals) if k == 'max_mach' else sum(v[k] for v in vals)) for k in d}
        npts = float(np.prod(self.plan_global['np']))
        return {'ke': d['sum_ke'] / npts, 'enstrophy': d['sum_enstrophy'] / npts, 'mass': d['sum_rho'] / npts,
                'energy': d['sum_rhoE'] / npts, 'max_mach': d['max_mach'], 'nonfinite': d['nonfinite']}


def write_output(runner, plan_sym, plan, cold, spec, workdir, iteration=None):
    """One dataset file of an `iohdf5` component in the reference's layout (iodata.py): `opensbli_output.h5` after the loop,
    `opensbli_output_%06d.h5` for the in-loop dumps of save_every (io_hdf5.py:99-127; npz stand-in without h5py)."""
    from . import iodata
    known = set(runner.sim.field_names())
    arrays = {}
    for name in spec['arrays']:
        if name in known:
            a = runner.gather(name)
        elif name in cold.arrays:                    # coordinates, metric terms: evaluated once by the cold path
            a = cold.arrays[name]
        else:
            raise KeyError("dataset '%s' requested by the output component exists neither on the device nor in the cold data" % name)
        arrays[name] = a
    if runner.rank != 0:
        return None
    base = spec.get('name') or 'opensbli_output'
    base = base[:-3] if base.endswith('.h5') else base
    if iteration is not None:
        base = '%s_%06d' % (base, iteration)
    return iodata.write_datasets(os.path.join(workdir, base), arrays, plan['np'])


def time_loop(runner, plan, niter, workdir='.', plan_sym=None, cold=None, iteration_ops=(None, None), log=None):
    """The reference's time loop as the runner drives it: niter steps on the GPU(s), cut at the iterations where the
    generated program does something besides stepping (algorithm.py:433-463):
      * SimulationMonitor: iter == 0 or (iter+1) %% frequency == 0 -> probe values in the reference's format (simulation_monitors.py:128-160)
      * print_iteration_ops(every, NaN_check): (iter+1) %% every == 0 -> "Iteration is N" (+ NaN check of the dataset; a
        non-finite value stops the run as ops_NaNcheck does)
      * iohdf5(save_every=N): (iter+1) %% N == 0 -> opensbli_output_%%06d file
    Accepts a Runner or a bare Simulation.  Returns the device time of the steps in ms."""
    if not isinstance(runner, Runner):
        sim = runner
        runner = Runner.__new__(Runner)
        runner.plan_global, runner.dist, runner.world, runner.rank, runner.dsim, runner.sim = plan, None, 1, 0, None, sim
        runner.k0, runner.nk = 0, plan['np'][plan['ndim'] - 1]
    mon = plan.get('monitor')
    every, nan_name = iteration_ops
    dumps = [sp for sp in (plan_sym or {}).get('io', []) if sp.get('when') == 'in_loop']
    log = log if log is not None else (lambda s: print(s, flush=True))
    if niter <= 0:
        return 0.0
    stops = set()
    if mon:
        stops |= set([1] + list(range(mon['frequency'], niter + 1, mon['frequency'])))
    if every:
        stops |= set(range(every, niter + 1, every))
    for sp in dumps:
        stops |= set(range(sp['every'], niter + 1, sp['every']))
    if not stops:
        return runner.step_timed(niter)
    dt = plan['constants']['dt']
    out = None
    if mon and runner.rank == 0:
        out = open(os.path.join(workdir, mon['output_file']), 'w') if mon.get('output_file') else sys.stdout
        fmt = '%%.%df' % mon['precision']
    ms, done = 0.0, 0
    try:
        for stop in sorted(stops):
            ms += runner.step_timed(stop - done)
            done = stop
            if every and stop % every == 0:
                if runner.rank == 0:
                    log('Iteration is %d' % stop)
                if nan_name:
                    bad = runner.nan_count(nan_name)
                    if bad:
                        raise RuntimeError('NaN check: %d non-finite values in %s at iteration %d' % (bad, nan_name, stop))
            if mon and (stop == 1 or stop % mon['frequency'] == 0):
                vals = []
                for a, pr in zip(mon['arrays'], mon['probes']):
                    v = runner.read_point(a, pr)
                    vals.append(v / stop if 'mean' in a else v)      # running sums are reported as means
                if runner.rank == 0:
                    if stop == 1:
                        out.write(', '.join(['Iteration', 'Time'] + ['%s_B0(%s)' % (a, ', '.join(str(x) for x in pr)) for a, pr in zip(mon['arrays'], mon['probes'])]) + '\n')
                    out.write(', '.join(['%d' % stop, fmt % (stop * dt)] + [fmt % v for v in vals]) + '\n')
            for sp in dumps:
                if stop % sp['every'] == 0:
                    write_output(runner, plan_sym, plan, cold, sp, workdir, iteration=stop)
        if done < niter:
            ms += runner.step_timed(niter - done)
    finally:
        if out is not None and out is not sys.stdout:
            out.close()
    return ms


def load_case(workdir='.', overrides=None):
    plan_sym = json.load(open(os.path.join(workdir, PLAN_FILE)))
    env = read_stub(open(os.path.join(workdir, STUB_FILE)).read(), plan_sym['constant_decls'], overrides, plan_sym.get('constant_array_decls', ()))
    plan_num, cold = resolve(plan_sym, env)
    return plan_sym, env, plan_num, cold


def main(argv=None):
    """python -m opensbli_b200.run [dir] [--restart FILE [--iteration N]] [--niter N]
    Under `python -m torch.distributed.run --nproc-per-node N ...` the block is decomposed over N GPUs (one rank each)."""
    import argparse
    ap = argparse.ArgumentParser(prog='python -m opensbli_b200.run')
    ap.add_argument('workdir', nargs='?', default='.')
    ap.add_argument('--restart', default=None, help='dataset file (.h5 / .npz, reference layout) holding the conserved arrays to start from')
    ap.add_argument('--iteration', type=int, default=0, help='iteration number the restart file was written at (time-dependent source terms)')
    ap.add_argument('--niter', type=int, default=None, help='override the number of iterations of the program')
    args = ap.parse_args(sys.argv[1:] if argv is None else argv)
    workdir = args.workdir
    plan_sym, env, plan_num, cold = load_case(workdir)
    stub = open(os.path.join(workdir, STUB_FILE)).read()
    q0 = initial_state(plan_sym, cold)
    if plan_num.get('generic'):
        print('B200: generic path (%s): %d run-time compiled loops per step' % (plan_num['generic'].get('reason'), len(plan_num['user_kernels'])))
        if plan_num.get('zero_datasets'):
            print('B200: datasets read but never written hold zeros, as OPS declares them: %s' % ', '.join(plan_num['zero_datasets']))
    if args.restart:                                  # the reference's ops_decl_dat_hdf5 route (opsc.py:702-705, generate_restart.py)
        from . import iodata
        data, _ = iodata.read_datasets(args.restart if os.path.isabs(args.restart) else os.path.join(workdir, args.restart))
        for m, n in enumerate(plan_sym['q_names']):
            if n not in data or data[n].shape != q0[m].shape:
                raise ValueError('restart file %s: dataset %s missing or of another shape' % (args.restart, n))
            q0[m] = np.ascontiguousarray(data[n], dtype=np.float64)
        plan_num['iteration0'] = args.iteration
    niter = plan_num.get('niter', 0) if args.niter is None else args.niter
    world = int(os.environ.get('WORLD_SIZE', '1'))
    dist, device = None, -1
    if world > 1:
        import torch
        import torch.distributed as dist
        device = int(os.environ.get('LOCAL_RANK', '0'))
        torch.cuda.set_device(device)
        dist.init_process_group('nccl', device_id=torch.device('cuda', device))
    runner = Runner(plan_num, dist, device)
    try:
        runner.set_state(q0)
        if args.restart:
            runner.sim.set_iteration(args.iteration)
        ms = time_loop(runner, plan_num, niter, workdir, plan_sym=plan_sym, cold=cold, iteration_ops=read_iteration_ops(stub))
        if any(k['when'] == 'after_loop' for k in plan_num.get('user_kernels', [])):
            runner.sim.run_user_kernels('after_loop')             # loops after the time loop (e.g. statistics / niter)
        if runner.rank == 0:
            print('Total Wall time %f' % (ms * 1e-3))             # same span as the reference's Timers (algorithm.py:301-327)
        # dataset files the program writes after the loop; without an iohdf5 component the conserved arrays are still kept
        specs = [sp for sp in plan_sym.get('io', []) if sp.get('when') == 'after'] or [{'arrays': list(plan_sym['q_names']), 'name': None}]
        extra = [w for k in plan_num.get('user_kernels', []) for w in k['writes']]
        for sp in specs:
            sp = dict(sp, arrays=list(dict.fromkeys(list(sp['arrays']) + extra)))
            path = write_output(runner, plan_sym, plan_num, cold, sp, workdir)
            if runner.rank == 0:
                print('wrote %s' % path)
    finally:
        runner.close()
        if world > 1:
            dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())
